#!/usr/bin/env python
"""bench.py - nodes/sec of the U2GNN train step (forward + loss + backward + clip + Adam).

    python bench.py --gpus N --steps K --warmup W [--workload cfg5]     # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W               # the reference's CPU path (torch.nn port)

Workloads (config.workload) are the five BASELINE.json configs as synthetic batches of the named shapes:
  cfg1  supervised, MUTAG shape      d 7,  k 8,  T 3, ff 1024, 4 graphs (~72 nodes) per step, attention as the reference runs it
  cfg2  unsupervised, PTC-degree shape d 4, k 4,  T 2, ff 1024, ns 512, V 8 792, 4 graphs (~100 nodes) per step
  cfg3  supervised, IMDBBINARY shape d 65, k 16, T 4, ff 1024, 4 096 graphs (~82 K nodes) per rank per step
  cfg4  unsupervised, REDDITMULTI5K shape: 4 999 graphs / ~2.54 M nodes, d 4, k 4, T 2, ff 1024, ns 512, class table row-sharded
        over the ranks, 512 graphs (~260 K nodes) per rank per step
  cfg5  supervised, 64 M-node-graph shape: d 64, k 16, T 4, ff 2048, 262 144 nodes per rank per step   (DEFAULT, the metric's config)
Supervised multi-GPU runs shard ONE global batch with parallel.shard_graph_batch (weak scaling: the global batch grows with N;
the only collective is the gradient all-reduce).  Prints ONE JSON line (contract in the task statement); the oracle / torch
port is used only for the cpu_baseline leg and the reference arm, which never imports the CUDA package.
"""
import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "graph-transformer_b200")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "nodes/sec U2GNN train step"
WORKLOADS = {
    "cfg1": dict(kind="sup", d=7, k=8, T=3, L=1, ff=1024, C=2, axis="nodes", nodes=72, avg_graph=18, precision="fp32",
                 desc="cfg1-shape (MUTAG) synthetic batch: supervised U2GNN, d 7, num_neighbors 8, T 3, L 1, ff 1024, 4 graphs per step, "
                      "attn_axis=nodes (reference as written) (BASELINE.json configs[0])"),
    "cfg2": dict(kind="unsup", d=4, k=4, T=2, L=1, ff=1024, ns=512, V=8792, axis="nodes", nodes=100, avg_graph=25, precision="fp32",
                 desc="cfg2-shape (PTC degree-as-tag) synthetic batch: unsupervised U2GNN, d 4, num_neighbors 4, T 2, L 1, ff 1024, "
                      "sampled softmax ns 512 over V 8792, 4 graphs per step, attn_axis=nodes (BASELINE.json configs[1])"),
    "cfg3": dict(kind="sup", d=65, k=16, T=4, L=1, ff=1024, C=2, axis="neighbors", nodes=4096 * 20, avg_graph=20, precision="bf16",
                 desc="cfg3-shape (IMDBBINARY) synthetic batch: supervised U2GNN, d 65, num_neighbors 16, T 4, L 1, ff 1024, 4096 graphs "
                      "per rank per step, attn_axis=neighbors (BASELINE.json configs[2])"),
    "cfg4": dict(kind="unsup", d=4, k=4, T=2, L=1, ff=1024, ns=512, V=2540000, graphs=4999, graphs_per_step=512, axis="neighbors",
                 precision="bf16",
                 desc="cfg4-shape (REDDITMULTI5K) synthetic dataset: 4999 graphs / ~2.54 M nodes, unsupervised U2GNN, d 4, num_neighbors 4, "
                      "T 2, L 1, ff 1024, sampled softmax ns 512, class table row-sharded over the ranks, 512 graphs per rank per step, "
                      "attn_axis=neighbors (BASELINE.json configs[3])"),
    "cfg5": dict(kind="sup", d=64, k=16, T=4, L=1, ff=2048, C=2, axis="neighbors", nodes=262144, avg_graph=61, precision="bf16",
                 desc="cfg5-shape synthetic graph batch: supervised U2GNN, d 64, num_neighbors 16, T 4, L 1, ff 2048, "
                      "attn_axis=neighbors (BASELINE.json configs[4])"),
    # one PASS over a rank's share of the 64 M-node / 1 B-edge graph (8 M nodes, ~125 M directed edges per rank: the full graph at 8
    # GPUs), dataset CSR + features resident in HBM, every micro-batch built on the device INSIDE the timed region
    "cfg5-full": dict(kind="epoch", d=64, k=16, T=4, L=1, ff=2048, C=2, axis="neighbors", nodes=262144, avg_graph=61, precision="bf16",
                      nodes_per_rank=8 * 1024 * 1024, mean_degree=15.6,
                      desc="cfg5 full pass: 8 Mi nodes / ~125 M directed edges per rank (64 Mi nodes / 1 B edges at 8 GPUs), power-law degrees, "
                           "graphs of ~61 nodes, supervised U2GNN d 64, num_neighbors 16, T 4, L 1, ff 2048; micro-batches of <= 262144 nodes "
                           "built on the device (CSR neighbour sampling + feature gather) inside the timed region (BASELINE.json configs[4])"),
}
# entry points timed per launch for the HBM-side roofline list (north_star: gather / attention / pooling kernels)
HBM_KERNELS = ("u2gnn_gather_rows", "u2gnn_inproj_seqattn_tc_fwd", "u2gnn_seqattn_tc_bwd_ex", "u2gnn_segment_sum",
               "u2gnn_add_dropout_ln_bwd_ex", "u2gnn_gemm_tc_rows_ln", "u2gnn_gemm_tc_dgrad_wgrad", "u2gnn_clip_adam")


def _load_synthetic():
    """u2gnn_b200/synthetic.py loaded by PATH: the reference arm must not import the package (its __init__ dlopens the CUDA library)."""
    spec = importlib.util.spec_from_file_location("u2gnn_synthetic", os.path.join(PKG, "u2gnn_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def algorithmic_per_node(d, S, T, ff, L=1):
    """fwd flops F = T(8Sd^2 + 4S^2 d + 4Sd ff); fwd+bwd = 3F.  HBM bytes fwd
    B = 8S + 4Sd + 4d + 4d; fwd+bwd = 2B."""
    F = T * (8 * S * d * d + 4 * S * S * d + 4 * S * d * ff) * L
    B = (8 * S + 4 * S * d + 4 * d + 4 * d) * L
    return 3 * F, 2 * B


# ----------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi, during the timed region)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# reference CPU path (oracle/torch_port.py: the reference's own torch.nn modules)
# ----------------------------------------------------------------------------------------------
def _host_graphs(b, k):
    """Graph objects (neighbour lists) of a synthetic batch for the host batch builder: every node's candidates are the other
    nodes of its graph, which is what make_batch samples from."""
    import numpy as np
    rp = b["rowptr"].cpu().numpy()
    X = b["X"].cpu().numpy()
    labels = b["labels"].cpu().numpy() if "labels" in b else np.zeros(len(rp) - 1, dtype=np.int64)
    graphs = []
    for g in range(len(rp) - 1):
        n = int(rp[g + 1] - rp[g])
        nb = [np.delete(np.arange(n, dtype=np.int64), i) for i in range(n)]
        graphs.append(dict(n=n, neighbors=nb, node_features=X[rp[g]:rp[g + 1]], label=int(labels[g])))
    return graphs


def _host_build_batch(graphs, k):
    """The reference's host batch builder restated (train_pytorch_U2GNN_Sup.py:91-119: per node np.random.choice with
    replacement, isolated nodes repeat themselves) - timed for the `with_host_batch_builder` CPU number."""
    import numpy as np
    rows, start = [], 0
    for g in graphs:
        for i, nb in enumerate(g["neighbors"]):
            node = start + i
            rows.append([node] + list(start + np.random.choice(nb, k, replace=True)) if len(nb) else [node] * (k + 1))
        start += g["n"]
    return np.array(rows, dtype=np.int64), np.concatenate([g["node_features"] for g in graphs], 0)


def cpu_train_rate(w, nodes, steps, warmup, threads, with_builder=False):
    """nodes/s of the reference CPU train step (train mode, real p=0.5 dropouts, clip 0.5, Adam) on a bounded sample of the
    workload: `nodes` nodes of the same synthetic batch generator."""
    from oracle import torch_port as TP
    syn = _load_synthetic()
    torch.set_num_threads(threads)
    torch.manual_seed(123)
    vocab = w.get("V")
    b = syn.make_batch(nodes, w["k"], w["d"], w.get("C", 2), avg_graph=w.get("avg_graph", 61), seed=2024, device="cpu",
                       vocab=max(vocab, nodes) if vocab else None)
    N, G = b["X"].shape[0], b["G"]
    if w["kind"] == "sup":
        model = TP.SupPort(w["d"], w["ff"], w["C"], w["T"], 0.5, w["L"], attn_axis=w["axis"])
        idx = torch.stack([torch.repeat_interleave(torch.arange(G), b["rowptr"][1:] - b["rowptr"][:-1]), torch.arange(N)])
        gp = torch.sparse_coo_tensor(idx, torch.ones(N), (G, N))
        soft = TP.smooth_labels(b["labels"], w["C"])
        loss_of = lambda ix, X: TP.soft_ce(model(ix, gp, X), soft)
    else:
        from oracle.sampler import OracleSampler
        model = TP.UnSupPort(max(vocab, nodes), w["d"], w["ff"], w["T"], w["L"], 0.5, attn_axis=w["axis"])
        sampler = OracleSampler(max(vocab, nodes))
        loss_of = lambda ix, X: torch.sum(model(X, ix, b["input_y"], sampler.sample_with_tries(w["ns"])[0]))
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=5e-4)
    graphs = _host_graphs(b, w["k"]) if with_builder else None

    def one():
        if with_builder:
            ix, X = _host_build_batch(graphs, w["k"])
            ix, X = torch.from_numpy(ix), torch.from_numpy(X)
        else:
            ix, X = b["input_x"], b["X"]
        TP.train_step(model, opt, lambda: loss_of(ix, X))

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return N * steps / dt, dt / steps


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # bounded sample: calibrate on a small batch, then size the step so the whole run takes ~2 minutes
    probe = min(128, w.get("nodes", 128))
    rate, _ = cpu_train_rate(w, probe, 1, 1, threads)
    budget = 120.0
    cap = w.get("nodes", 8192) if w["axis"] == "nodes" else 8192
    nodes = int(max(min(64, cap), min(cap, rate * budget / max(1, args.steps + args.warmup))))
    rate, sec = cpu_train_rate(w, nodes, args.steps, args.warmup, threads)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "nodes/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, w, nodes),
            "cpu_baseline": {"value": rate, "unit": "nodes/s", "cores": threads, "kind": "port",
                             "sample": "%d nodes/step of the %s-shape batch (node subsample of the GPU arm's step, SURVEY.md 8(d)), torch.nn "
                                       "port of the reference model (oracle/torch_port.py), fp32, train mode with p=0.5 dropouts"
                                       % (nodes, args.workload)},
            "e2e": {"value": rate, "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, w, nodes, extra=None):
    c = {"workload": w["desc"], "workload_id": args.workload, "nodes_per_rank_per_step": nodes,
         "l2_policy": ("inputs larger than L2 (X + input_x + activations >> 126 MB)" if nodes >= 65536 else
                       "L2 flushed between timed steps (256 MB write): the batch itself fits in L2")}
    if extra:
        c.update(extra)
    return c


# ----------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------
def _tree_dataset(w, seed, device):
    """cfg4: REDDITMULTI5K-shaped dataset resident on the device: `graphs` graphs, log-normal sizes summing to V, tree-like
    (every node but the first of its graph links to a random earlier node of the graph: ~1 undirected edge per node), features
    one-hot(degree mod 4) * 0.01-scaled like the reference's REDDIT rule.  -> (rowptr, col, gstart, X)."""
    g = torch.Generator(device=device).manual_seed(seed)
    G, V = w["graphs"], w["V"]
    raw = torch.exp(torch.randn(G, generator=g, device=device) * 0.75)
    sizes = torch.clamp((raw / raw.sum() * V).long(), min=2)
    sizes[-1] += V - int(sizes.sum().item())
    gstart = torch.zeros(G + 1, dtype=torch.int64, device=device)
    gstart[1:] = torch.cumsum(sizes, 0)
    gid = torch.repeat_interleave(torch.arange(G, device=device), sizes)
    local = torch.arange(V, device=device) - gstart[gid]
    parent = gstart[gid] + (torch.rand(V, generator=g, device=device) * local.clamp(min=1)).long().clamp(max=(local - 1).clamp(min=0))
    child = torch.arange(V, device=device)
    keep = local > 0
    src = torch.cat([child[keep], parent[keep]])
    dst = torch.cat([parent[keep], child[keep]])
    order = torch.argsort(src * V + dst)
    src, dst = src[order], dst[order]
    deg = torch.bincount(src, minlength=V)
    rowptr = torch.zeros(V + 1, dtype=torch.int64, device=device)
    rowptr[1:] = torch.cumsum(deg, 0)
    X = torch.zeros((V, w["d"]), dtype=torch.float32, device=device)
    X[torch.arange(V, device=device), deg % w["d"]] = 1.0
    X = X * 0.99 + 0.01
    return rowptr, dst.contiguous(), gstart, X


def _powerlaw_dataset(w, n_nodes, seed, device):
    """Graph-sharded share of the cfg5 graph, generated in HBM: graphs of ~avg_graph nodes, per-node degree ~ Pareto (alpha 2.1)
    clipped to [1, graph size - 1] and scaled to the target mean, neighbours uniform inside the node's graph (no cross-rank
    edges).  -> (rowptr int64 [V+1], col int64 [E], gstart int64 [G+1], X f32 [V, d], labels int64 [G])."""
    g = torch.Generator(device=device).manual_seed(seed)
    avg = w["avg_graph"]
    G = n_nodes // avg
    sizes = torch.randint(avg // 2, avg + avg // 2 + 1, (G,), generator=g, device=device, dtype=torch.int64)
    sizes[-1] += n_nodes - int(sizes.sum().item())
    gstart = torch.zeros(G + 1, dtype=torch.int64, device=device)
    gstart[1:] = torch.cumsum(sizes, 0)
    gid = torch.repeat_interleave(torch.arange(G, device=device), sizes)
    size = sizes[gid]
    u = torch.rand(n_nodes, generator=g, device=device).clamp_(min=1e-7)
    pareto = u.pow_(-1.0 / 1.1)                                   # density ~ x^-2.1 on [1, inf)
    lo, hi = 0.5, 32.0
    for _ in range(24):                                           # scale so the clipped mean hits the target
        mid = 0.5 * (lo + hi)
        mean = torch.minimum((pareto * mid).floor().clamp_(min=1), (size - 1).float()).mean().item()
        lo, hi = (mid, hi) if mean < w["mean_degree"] else (lo, mid)
    deg = torch.minimum((pareto * hi).floor().clamp_(min=1), (size - 1).float()).long()
    del pareto, u
    rowptr = torch.zeros(n_nodes + 1, dtype=torch.int64, device=device)
    rowptr[1:] = torch.cumsum(deg, 0)
    E_ = int(rowptr[-1].item())
    src = torch.repeat_interleave(torch.arange(n_nodes, device=device), deg)
    col = gstart[gid[src]] + (torch.rand(E_, generator=g, device=device) * sizes[gid[src]].float()).long().clamp_(max=10 ** 9)
    col = torch.minimum(col, gstart[gid[src] + 1] - 1)
    del src
    X = torch.randn((n_nodes, w["d"]), generator=g, device=device)
    labels = torch.randint(0, w["C"], (G,), generator=g, device=device, dtype=torch.int64)
    return rowptr, col.contiguous(), gstart, X, labels, float(deg.float().mean().item())


def run_full_epoch(args, w):
    """cfg5-full: ONE pass over this rank's share of the big graph.  The dataset (CSR adjacency + features) lives in HBM; every
    micro-batch is built by the device batch builder (u2gnn_build_batch + u2gnn_gather_rows) and trained inside the timed region,
    so input generation, neighbour sampling and load balance are part of the number.  Every rank runs the same number of
    micro-batches (balanced by node count) because each step ends in the gradient all-reduce."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sys.path.insert(0, PKG)
    import u2gnn_b200 as U
    from u2gnn_b200 import engine as E
    from u2gnn_b200 import parallel as P
    from u2gnn_b200.data import DeviceBatchBuilder
    from u2gnn_b200.trainer import SupTrainer
    U.require_device()
    dev = torch.device("cuda", local)
    V = args.full_nodes or w["nodes_per_rank"]
    rowptr, col, gstart, X, labels, mean_deg = _powerlaw_dataset(w, V, 2024 + rank, dev)
    edges = int(col.numel())
    ds = DeviceBatchBuilder.from_device_tensors(rowptr, col, gstart, X, w["k"], seed=7, labels=labels.cpu().numpy())
    torch.manual_seed(123)
    model = U.TransformerU2GNN(w["d"], w["ff"], w["C"], w["T"], 0.5, w["L"], attn_axis=w["axis"]).cuda()
    trainer = SupTrainer(model, lr=5e-4, precision=args.precision or w["precision"])
    nb = (V + w["nodes"] - 1) // w["nodes"]
    ranges = P.balanced_graph_ranges(gstart, nb)                  # nb micro-batches of whole graphs, balanced by node count
    G_total = sum(r[1] - r[0] for r in ranges) * world            # per-step loss normalisation is per micro-batch below
    import numpy as np

    def epoch(limit=None):
        done = 0
        loss = None
        for i, (g0, g1) in enumerate(ranges[:limit]):
            ix, rp, Xc, y, _ = ds.build(np.arange(g0, g1), stream_id=i)
            loss = trainer.step(ix, rp, Xc, y, G_total=(g1 - g0) * world)
            done += int(Xc.shape[0])
        return done, loss

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    epoch(limit=3)                                                # warm-up: 3 micro-batches
    sync()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    U.LIB.launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    done, loss = epoch()
    e1.record()
    sync()
    E.check_device_errors()
    launches = U.LIB.launches
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    n = torch.tensor([float(done)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    ms, total = float(t.item()), float(n.item())
    if rank == 0:
        flops_node, bytes_node = algorithmic_per_node(w["d"], w["k"] + 1, w["T"], w["ff"], w["L"])
        line = {"metric": METRIC, "value": total / (ms / 1e3), "unit": "nodes/s", "n_gpus": world, "steps": nb, "warmup": 3,
                "ms_per_step": ms / nb, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(args, w, int(total / nb / world),
                                          {"precision": args.precision or w["precision"], "parallelism": "dp%d" % world,
                                           "nodes_per_rank": V, "directed_edges_per_rank": edges, "mean_degree": mean_deg,
                                           "micro_batches_per_rank": nb, "epoch_s": ms / 1e3, "loss": float(loss.item()),
                                           "dataset_bytes_in_hbm_per_rank": int(rowptr.numel() * 8 + col.numel() * 8 + X.numel() * 4),
                                           "algorithmic_mflop_per_node": flops_node / 1e6, "algorithmic_bytes_per_node": bytes_node,
                                           "timed_region": "whole pass: device batch build (neighbour sampling, feature gather) + train step per micro-batch"}),
                "clocks": clk, "gpu_launches": launches,
                "e2e": {"value": total / (ms / 1e3), "unit": "nodes/s", "h2d_bytes_per_step": int(16 * (len(ranges[0]) + 2) * 0 + 2 * 8 * (ranges[0][1] - ranges[0][0] + 1)),
                        "d2h_bytes_per_step": 0, "note": "the dataset is resident in HBM; per micro-batch only the graph offsets travel host->device"}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_ours(args, w):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sys.path.insert(0, PKG)
    import u2gnn_b200 as U
    from u2gnn_b200 import engine as E
    from u2gnn_b200 import parallel as P
    from u2gnn_b200.synthetic import make_batch
    from u2gnn_b200.trainer import SupTrainer, UnSupTrainer
    U.require_device()
    S = w["k"] + 1
    precision = args.precision or w["precision"]
    note = None
    if precision == "bf16" and not (E.ffn_tc_supported(w["d"], w["ff"]) or E.ffn_wide_supported(w["d"], w["ff"])):
        precision, note = "fp32", "no tcgen05 FFN path for this feature size: this workload ran on the fp32 path"
    elif precision == "bf16" and E.ffn_wide_supported(w["d"], w["ff"]):
        note = ("bf16 FFN as named by the config: 64 < d <= 128 runs the FFN as tcgen05 GEMMs with the hidden materialised in bf16 "
                "(engine.ffn_wide_*); the attention block of this feature size stays on the fp32 kernels")
    torch.manual_seed(123)
    dev = torch.device("cuda", local)
    batches = []                                   # resident batches, cycled over the timed steps
    if w["kind"] == "sup":
        nodes = args.nodes or w["nodes"]
        if precision == "fp32" and not args.nodes and nodes > 65536:
            # the fp32 mode keeps the [rows, ff] hidden of every timestep in HBM (8 KB per row and timestep at ff 2048): 65 536
            # nodes per step is what fits next to the other saved activations; throughput is per node, the batch is the same shape
            nodes = 65536
            note = "fp32 mode: micro-batch of 65 536 nodes (the [rows, ff] fp32 hidden is materialised)"
        model = U.TransformerU2GNN(w["d"], w["ff"], w["C"], w["T"], 0.5, w["L"], attn_axis=w["axis"]).cuda()
        trainer = SupTrainer(model, lr=5e-4, precision=precision)
        # ONE global batch (the same on every rank: seeded device generator), sharded by balanced graph ranges
        gb = make_batch(nodes * world, w["k"], w["d"], w["C"], avg_graph=w["avg_graph"], seed=2024, device="cuda")
        sh = P.shard_graph_batch(gb["input_x"], gb["rowptr"], gb["X"], gb["labels"], rank, world)
        G_total = gb["G"]
        del gb
        torch.cuda.empty_cache()
        batches.append(dict(input_x=sh["input_x"], rowptr=sh["rowptr"], X=sh["X"], labels=sh["labels"]))
        my_nodes = sh["X"].shape[0]

        def step_on(b):
            return trainer.step(b["input_x"], b["rowptr"], b["X"], b["labels"], G_total=G_total)
    else:
        if "graphs" in w:                          # cfg4: dataset resident in HBM, device batch builder, row-sharded table
            rowptr, col, gstart, X = _tree_dataset(w, 2024, dev)
            V = w["V"]
            ranges = P.balanced_graph_ranges(gstart, world)
            g0, g1 = ranges[rank]
            bounds = [int(gstart[r[0]]) for r in ranges] + [V]
            shard = P.RowShard(V, world, rank, bounds) if world > 1 else None
            from u2gnn_b200.data import DeviceBatchBuilder
            ds = DeviceBatchBuilder.from_device_tensors(rowptr, col, gstart, X, w["k"], seed=7)
            gen = torch.Generator().manual_seed(100 + rank)
            gps = min(args.graphs_per_step or w["graphs_per_step"], g1 - g0)
            for i in range(4):
                sel = (g0 + torch.randperm(g1 - g0, generator=gen)[:gps]).sort().values.numpy()
                ix, rp, Xc, _, node_global = ds.build(sel, stream_id=i)
                batches.append(dict(input_x=ix, X=Xc, input_y=node_global))
            local_rows = bounds[rank + 1] - bounds[rank] if shard is not None else V
            model = U.TransformerU2GNNUnSup(local_rows, w["d"], w["ff"], w["ns"], w["T"], w["L"], 0.5, dev, attn_axis=w["axis"],
                                            precision=precision).cuda()
            trainer = UnSupTrainer(model, lr=5e-3, row_shard=shard, global_vocab=V)
            my_nodes = sum(b["X"].shape[0] for b in batches) / len(batches)
        else:                                      # cfg2: one tiny batch, whole table on every rank
            nodes = args.nodes or w["nodes"]
            b = make_batch(nodes, w["k"], w["d"], 2, avg_graph=w["avg_graph"], seed=2024 + rank, device="cuda", vocab=w["V"])
            batches.append(dict(input_x=b["input_x"], X=b["X"], input_y=b["input_y"]))
            model = U.TransformerU2GNNUnSup(w["V"], w["d"], w["ff"], w["ns"], w["T"], w["L"], 0.5, dev, attn_axis=w["axis"],
                                            precision=precision).cuda()
            trainer = UnSupTrainer(model, lr=5e-3)
            my_nodes = nodes

        def step_on(b):
            return trainer.step(b["X"], b["input_x"], b["input_y"])
    dominant = trainer.dominant_kernel()
    small = my_nodes < 65536
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if small else None

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step_on(batches[i % len(batches)])
    sync()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    U.LIB.launches = 0
    names = (dominant,) + tuple(n for n in HBM_KERNELS if n != dominant)
    U.LIB.timed = {n: [] for n in names}
    E.FLOPS.clear(); E.BYTES.clear()
    nodes_done = 0
    ms = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    if not small:
        e0.record()
        for i in range(args.steps):
            b = batches[i % len(batches)]
            loss = step_on(b)
            nodes_done += b["X"].shape[0]
        e1.record()
        sync()
        ms = e0.elapsed_time(e1)
    else:
        # tiny batches fit in L2: flush it between the timed steps and time each step on its own
        for i in range(args.steps):
            b = batches[i % len(batches)]
            flush.fill_(i & 0xFF)
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            loss = step_on(b)
            b_.record()
            torch.cuda.synchronize()
            ms += a_.elapsed_time(b_)
            nodes_done += b["X"].shape[0]
    launches = U.LIB.launches
    timed = U.LIB.timed
    U.LIB.timed = None
    kern = {n: (sum(a.elapsed_time(b_) for a, b_ in v), len(v)) for n, v in timed.items()}
    flops, nbytes = dict(E.FLOPS), dict(E.BYTES)
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms, float(nodes_done)], device="cuda", dtype=torch.float64)
    if world > 1:
        tn = t[1:].clone()
        dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(tn, op=dist.ReduceOp.SUM)
        t[1] = tn[0]
    ms, total_nodes = float(t[0].item()), float(t[1].item())
    value = total_nodes / (ms / 1e3)

    # ---- end-to-end through the public API with HOST buffers (H2D of the batch + D2H of the loss per step)
    hosts = [{k: v.cpu().pin_memory() for k, v in b.items() if torch.is_tensor(v)} for b in batches]
    h2d = sum(v.numel() * v.element_size() for v in hosts[0].values())

    # The input feed is pipelined the way a training loop's loader is: the host->device copy of step i+1's batch is issued on
    # a copy stream while step i computes; every timed step still copies its own inputs from pinned host memory (K copies
    # in the region) and ends with the device->host read of its loss.
    copy_stream = torch.cuda.Stream()

    def issue_copy(i):
        with torch.cuda.stream(copy_stream):
            devb = {k: v.cuda(non_blocking=True) for k, v in hosts[i % len(hosts)].items()}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return devb, ev

    def run_host_steps(n):
        nxt = issue_copy(0)
        last, done = 0.0, 0
        for i in range(n):
            devb, ev = nxt
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            for v in devb.values():
                v.record_stream(cur)
            l = step_on(devb)
            done += devb["X"].shape[0]
            if i + 1 < n:
                nxt = issue_copy(i + 1)
            last = float(l.sum().item())      # device->host read of the step's loss
        return last, done

    run_host_steps(min(2, args.warmup))
    sync()
    e0.record()
    _, done = run_host_steps(args.steps)
    e1.record()
    sync()
    t = torch.tensor([e0.elapsed_time(e1), float(done)], device="cuda", dtype=torch.float64)
    if world > 1:
        tn = t[1:].clone()
        dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(tn, op=dist.ReduceOp.SUM)
        t[1] = tn[0]
    e2e = float(t[1].item()) / (float(t[0].item()) / 1e3)
    E.check_device_errors()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        flops_node, bytes_node = algorithmic_per_node(w["d"], S, w["T"], w["ff"], w["L"])
        ncu = None
        try:
            ncu = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_ffn_summary.json")))
        except Exception:
            pass
        roof = trainer.roofline(dominant, kern[dominant][0], kern[dominant][1], peaks, flops, ncu)
        hbm_peak = peaks.get("hbm_gbs", 6550.0)
        hbm = []
        for n in HBM_KERNELS:
            tms, cnt = kern.get(n, (0.0, 0))
            if cnt and nbytes.get(n):
                gbs = nbytes[n] / (tms * 1e-3) / 1e9
                hbm.append({"kernel": n, "achieved_gbs": gbs, "frac": gbs / hbm_peak, "launches": cnt, "ms_per_step": tms / args.steps,
                            "algorithmic_bytes_per_launch": nbytes[n] / cnt})
        line = {"metric": METRIC, "value": value, "unit": "nodes/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
                "config": workload_config(args, w, int(round(total_nodes / args.steps / world)),
                                          {"precision": precision, "parallelism": "dp%d" % world,
                                           "algorithmic_mflop_per_node": flops_node / 1e6, "algorithmic_bytes_per_node": bytes_node,
                                           "loss": float(loss.sum().item())}),
                "roofline": roof, "roofline_hbm": hbm, "hbm_peak_gbs": hbm_peak, "clocks": clk,
                "e2e": {"value": e2e, "unit": "nodes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
                "gpu_launches": launches}
        if note:
            line["config"]["note"] = note
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cn = min(args.cpu_nodes, w.get("nodes", args.cpu_nodes)) if w["axis"] == "nodes" else args.cpu_nodes
            rate, sec = cpu_train_rate(w, cn, 2, 1, threads)
            rate_b, _ = cpu_train_rate(w, cn, 2, 1, threads, with_builder=True)
            rate_1, _ = cpu_train_rate(w, cn, 1, 1, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "nodes/s", "cores": threads, "kind": "port",
                                    "with_host_batch_builder": rate_b, "one_thread": rate_1,
                                    "sample": "%d nodes/step x 2 steps (+1 warm-up) of the same %s-shape batch through the torch.nn "
                                              "port of the reference model, fp32, train mode; with_host_batch_builder adds the "
                                              "reference's per-node np.random.choice loop per step; one_thread = 1 step on 1 thread"
                                              % (cn, args.workload)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("U2GNN_WORKLOAD", "cfg5"), choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=os.environ.get("U2GNN_PRECISION"), choices=["fp32", "bf16"])
    ap.add_argument("--nodes", type=int, default=int(os.environ.get("U2GNN_BENCH_NODES", 0)),
                    help="nodes per rank per step (default: the workload's)")
    ap.add_argument("--graphs-per-step", type=int, default=0, help="cfg4: graphs per rank per step")
    ap.add_argument("--full-nodes", type=int, default=0, help="cfg5-full: nodes per rank (default 8 Mi)")
    ap.add_argument("--cpu-nodes", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "ours":
        args.warmup = max(args.warmup, 3)
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    elif w["kind"] == "epoch":
        run_full_epoch(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
