#! /usr/bin/env python
"""Supervised U2GNN training CLI on the B200 engine — same flags, seeds, log lines and output files as the
reference script (U2GNN_pytorch/train_pytorch_U2GNN_Sup.py:24-39,191-212), plus --attn_axis / --precision /
--dataset_root.  The train step is the fused CUDA step (u2gnn_b200.trainer.SupTrainer); batches are built on the
host with the global numpy RNG in the reference's call order (permutation per train batch, choice per node)."""
import os
import sys
import time
from argparse import ArgumentParser, ArgumentDefaultsHelpFormatter

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import u2gnn_b200 as U                                    # noqa: E402
from u2gnn_b200 import engine as E                        # noqa: E402
from u2gnn_b200.data import DeviceBatchBuilder, build_batch, load_data, separate_data   # noqa: E402
from u2gnn_b200.parallel import balanced_graph_ranges, shard_graph_batch   # noqa: E402
from u2gnn_b200.evaluate import ConditionalStepLR, sup_accuracy   # noqa: E402
from u2gnn_b200.trainer import SupTrainer                 # noqa: E402


def parse_args(argv=None):
    p = ArgumentParser("U2GNN", formatter_class=ArgumentDefaultsHelpFormatter, conflict_handler="resolve")
    p.add_argument("--run_folder", default="../", help="")
    p.add_argument("--dataset", default="PTC", help="Name of the dataset.")
    p.add_argument("--learning_rate", default=0.0005, type=float, help="Learning rate")
    p.add_argument("--batch_size", default=4, type=int, help="Batch Size")
    p.add_argument("--num_epochs", default=50, type=int, help="Number of training epochs")
    p.add_argument("--model_name", default="PTC", help="")
    p.add_argument("--sampled_num", "--num_sampled", default=512, type=int, help="")
    p.add_argument("--dropout", default=0.5, type=float, help="")
    p.add_argument("--num_hidden_layers", default=1, type=int, help="")
    p.add_argument("--num_timesteps", default=1, type=int, help="Timestep T ~ Number of self-attention layers within each U2GNN layer")
    p.add_argument("--ff_hidden_size", default=1024, type=int, help="The hidden size for the feedforward layer")
    p.add_argument("--num_neighbors", default=4, type=int, help="")
    p.add_argument("--fold_idx", type=int, default=1, help="The fold index. 0-9.")
    p.add_argument("--degree_as_tag", action="store_true", help="use node degrees as tags (README/TF-era flag)")
    p.add_argument("--attn_axis", default="nodes", choices=["nodes", "neighbors"],
                   help="'nodes' = reference as written, 'neighbors' = intended layout (SURVEY.md F1)")
    p.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    p.add_argument("--label_smoothing_style", default="reference", choices=["reference", "tf"],
                   help="reference: 0.9 / 0.1/(C-1) targets of the PyTorch file; tf: tf.losses.softmax_cross_entropy(label_smoothing=0.1)")
    p.add_argument("--tie_timesteps", action="store_true",
                   help="share ONE encoder weight set across the T timesteps (the published Universal-Transformer U2GNN); "
                        "default = T independent sets like the reference PyTorch file")
    p.add_argument("--dataset_root", default=None)
    p.add_argument("--batch_builder", default="host", choices=["host", "device"],
                   help="host: the reference's numpy loop (reproduces its index stream bit for bit); device: dataset resident in HBM, "
                        "neighbours sampled by the CUDA batch builder (counter-based stream, no host loop)")
    p.add_argument("--world_size", default=1, type=int,
                   help="data-parallel ranks (launch with torchrun; every rank draws the same graph batch and trains on its balanced share, "
                        "gradients are all-reduced over NCCL)")
    return p.parse_args(argv)


def run(args, log=print):
    torch.manual_seed(123)
    np.random.seed(123)
    torch.cuda.manual_seed_all(123)
    U.require_device()
    world, rank = max(1, args.world_size), 0
    if world > 1:
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        if dist.get_world_size() != world:
            raise ValueError("--world_size %d but torchrun started %d ranks" % (world, dist.get_world_size()))
        rank = dist.get_rank()
    dev = torch.device("cuda", torch.cuda.current_device())
    if world > 1 and args.attn_axis == "nodes":
        raise ValueError("attn_axis='nodes' couples every node of a batch (SURVEY.md F1): a batch cannot be sharded across ranks; "
                         "use --attn_axis neighbors with --world_size > 1")
    if rank != 0:
        log = lambda *a, **k: None
    log(args)
    degree_as_tag = args.degree_as_tag or args.dataset in ("COLLAB", "IMDBBINARY", "IMDBMULTI")
    graphs, num_classes = load_data(args.dataset, degree_as_tag, args.dataset_root)
    train_graphs, test_graphs = separate_data(graphs, args.fold_idx)
    d = graphs[0].node_features.shape[1]
    reddit = 4 if "REDDIT" in args.dataset else None
    if reddit:
        d = 4
    model = U.TransformerU2GNN(feature_dim_size=d, ff_hidden_size=args.ff_hidden_size, num_classes=num_classes,
                               dropout=args.dropout, num_self_att_layers=args.num_timesteps,
                               num_U2GNN_layers=args.num_hidden_layers, attn_axis=args.attn_axis,
                               tie_timesteps=args.tie_timesteps).to(dev)
    trainer = SupTrainer(model, lr=args.learning_rate, precision=args.precision, seed=123,
                         smoothing_style=args.label_smoothing_style)
    steps_per_epoch = int((len(train_graphs) - 1) / args.batch_size) + 1

    def to_dev(batch):
        ix, rp, X, y = batch
        return (torch.from_numpy(ix).to(dev), torch.from_numpy(rp).to(dev), torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev))

    builder = DeviceBatchBuilder(train_graphs, args.num_neighbors, device=dev, seed=123) if args.batch_builder == "device" else None
    if builder is not None and reddit:
        builder.X = (builder.X.repeat(1, reddit) * 0.01).contiguous()       # the reference's REDDIT feature rule (train_pytorch_U2GNN_Sup.py:93-95)

    def train_batch(sel):
        """One global batch `sel` (the same on every rank: the numpy stream is seeded identically) -> this rank's share."""
        G = len(sel)
        if builder is not None:
            sizes = np.array([train_graphs[i].n for i in sel])
            g0, g1 = balanced_graph_ranges(torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)])), world)[rank]
            ix, rp, X, y, _ = builder.build(np.asarray(sel)[g0:g1])
            return ix, rp, X, y, G
        ix, rp, X, y = to_dev(build_batch([train_graphs[i] for i in sel], args.num_neighbors, np.random, reddit))
        if world > 1:
            sh = shard_graph_batch(ix, rp, X, y, rank, world)
            ix, rp, X, y = sh["input_x"], sh["rowptr"], sh["X"], sh["labels"]
        return ix, rp, X, y, G

    def train_epoch():
        model.train()
        total = 0.0
        for _ in range(steps_per_epoch):
            sel = np.random.permutation(len(train_graphs))[:args.batch_size]
            ix, rp, X, y, G = train_batch(sel)
            total += float(trainer.step(ix, rp, X, y, G_total=G).item())      # loss.item(): the reference's per-step device sync
            E.check_device_errors(dev)
        return total

    def evaluate():
        return sup_accuracy(model, test_graphs, args.batch_size, args.num_neighbors, np.random, reddit, dev)

    out_dir = os.path.abspath(os.path.join(args.run_folder, "../runs_pytorch_U2GNN_Sup", args.model_name))
    log("Writing to {}\n".format(out_dir))
    ckpt = os.path.join(out_dir, "checkpoints")
    os.makedirs(ckpt, exist_ok=True)
    accs, losses = [], []
    sched = ConditionalStepLR(args.learning_rate, steps_per_epoch)
    with open(os.path.join(ckpt, "model_acc.txt") if rank == 0 else os.devnull, "w") as w:
        for epoch in range(1, args.num_epochs + 1):
            t0 = time.time()
            loss = train_epoch()
            if world > 1:                            # every rank holds its share of the (globally normalised) loss
                t = torch.tensor([loss], device=dev, dtype=torch.float64)
                torch.distributed.all_reduce(t)
                loss = float(t.item())
            losses.append(loss)
            acc = evaluate()
            accs.append(acc)
            log("| epoch {:3d} | time: {:5.2f}s | loss {:5.2f} | test acc {:5.2f} | ".format(epoch, time.time() - t0, loss, acc * 100))
            trainer.lr = sched.epoch_end(loss)       # the reference's conditional StepLR (train_pytorch_U2GNN_Sup.py:147,209-210)
            w.write("epoch " + str(epoch) + " fold " + str(args.fold_idx) + " acc " + str(acc * 100) + "%\n")
    return accs, losses


if __name__ == "__main__":
    run(parse_args())
