#! /usr/bin/env python
"""Unsupervised U2GNN training CLI on the B200 engine — flags, seeds, transductive protocol and log lines of
U2GNN_pytorch/train_pytorch_U2GNN_UnSup.py:27-42,149-213 (the reference script itself does not run as shipped,
SURVEY.md F3; its calling convention is the contract).  Per epoch: fused train steps over random graph batches
(vocab = every node of the dataset, sampled softmax with device-drawn log-uniform negatives), then graph
embeddings = sum-pool of the class table -> 10-fold LogisticRegression accuracy."""
import os
import sys
import time
from argparse import ArgumentParser, ArgumentDefaultsHelpFormatter

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import u2gnn_b200 as U                                    # noqa: E402
from u2gnn_b200 import engine as E                        # noqa: E402
from u2gnn_b200.data import DeviceBatchBuilder, build_batch, global_node_ids, load_data   # noqa: E402
from u2gnn_b200.evaluate import ConditionalStepLR, unsup_accuracy   # noqa: E402
from u2gnn_b200.trainer import UnSupTrainer               # noqa: E402


def parse_args(argv=None):
    p = ArgumentParser("U2GNN", formatter_class=ArgumentDefaultsHelpFormatter, conflict_handler="resolve")
    p.add_argument("--run_folder", default="../", help="")
    p.add_argument("--dataset", default="PTC", help="Name of the dataset.")
    p.add_argument("--learning_rate", default=0.005, type=float, help="Learning rate")
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
    p.add_argument("--attn_axis", default="nodes", choices=["nodes", "neighbors"])
    p.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    p.add_argument("--tie_timesteps", action="store_true",
                   help="share ONE encoder weight set across the T timesteps (the published Universal-Transformer U2GNN); "
                        "default = T independent sets like the reference PyTorch file")
    p.add_argument("--dataset_root", default=None)
    p.add_argument("--batch_builder", default="host", choices=["host", "device"],
                   help="host: the reference's numpy loop (reproduces its index stream bit for bit); device: dataset resident in HBM, "
                        "neighbours sampled by the CUDA batch builder")
    return p.parse_args(argv)


def run(args, log=print):
    torch.manual_seed(123)
    np.random.seed(123)
    torch.cuda.manual_seed_all(123)
    U.require_device()
    dev = torch.device("cuda")
    log(args)
    degree_as_tag = args.degree_as_tag or args.dataset in ("COLLAB", "IMDBBINARY", "IMDBMULTI")
    graphs, _ = load_data(args.dataset, degree_as_tag, args.dataset_root)
    labels = np.array([g.label for g in graphs])
    d = graphs[0].node_features.shape[1]
    vocab = int(sum(g.n for g in graphs))
    pool_rowptr = torch.from_numpy(np.concatenate([[0], np.cumsum([g.n for g in graphs])]).astype(np.int64)).to(dev)
    model = U.TransformerU2GNNUnSup(feature_dim_size=d, ff_hidden_size=args.ff_hidden_size, dropout=args.dropout,
                                    num_self_att_layers=args.num_timesteps, vocab_size=vocab, sampled_num=args.sampled_num,
                                    num_U2GNN_layers=args.num_hidden_layers, device=dev, attn_axis=args.attn_axis,
                                    precision=args.precision, tie_timesteps=args.tie_timesteps).to(dev)
    trainer = UnSupTrainer(model, lr=args.learning_rate, seed=123)
    steps_per_epoch = int((len(graphs) - 1) / args.batch_size) + 1

    builder = DeviceBatchBuilder(graphs, args.num_neighbors, device=dev, seed=123) if args.batch_builder == "device" else None

    def train_epoch():
        model.train()
        total = 0.0
        for _ in range(steps_per_epoch):
            sel = np.random.permutation(len(graphs))[:args.batch_size]
            if builder is not None:                   # input_y = the dataset-wide node ids the builder returns (train_pytorch_U2GNN_UnSup.py:96-99)
                ix, _, X, _, iy = builder.build(sel)
                loss = trainer.step(X, ix, iy)
            else:
                ix, _, X, _ = build_batch([graphs[i] for i in sel], args.num_neighbors, np.random)
                iy = global_node_ids(graphs, sel)
                loss = trainer.step(torch.from_numpy(X).to(dev), torch.from_numpy(ix).to(dev), torch.from_numpy(iy).to(dev))
            total += float(loss.sum().item())         # the reference's per-step device sync (loss.item())
            E.check_device_errors(dev)
        return total

    def evaluate():
        return unsup_accuracy(model.ss.weight.data, pool_rowptr, labels)

    out_dir = os.path.abspath(os.path.join(args.run_folder, "../runs_pytorch_U2GNN_UnSup", args.model_name))
    log("Writing to {}\n".format(out_dir))
    ckpt = os.path.join(out_dir, "checkpoints")
    os.makedirs(ckpt, exist_ok=True)
    hist = []
    sched = ConditionalStepLR(args.learning_rate, steps_per_epoch)
    with open(os.path.join(ckpt, "model_acc.txt"), "w") as w:
        for epoch in range(1, args.num_epochs + 1):
            t0 = time.time()
            loss = train_epoch()
            mean, std = evaluate()
            hist.append((mean, std, loss))
            trainer.lr = sched.epoch_end(loss)       # the reference's conditional StepLR (train_pytorch_U2GNN_UnSup.py:147,210-211)
            log("| epoch {:3d} | time: {:5.2f}s | loss {:5.2f} | mean {:5.2f} | std {:5.2f} | ".format(epoch, time.time() - t0, loss, mean, std))
            w.write("epoch " + str(epoch) + " mean: " + str(mean) + " std: " + str(std) + "\n")
    return hist


if __name__ == "__main__":
    run(parse_args())
