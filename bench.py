#!/usr/bin/env python
"""bench.py — nodes/sec of the U2GNN train step (forward + loss + backward + clip + Adam).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU path (torch.nn port)

Workload (config.workload): BASELINE.json configs[4] shape — supervised U2GNN on a synthetic graph
batch, d 64, num_neighbors 16 (S 17), T 4, L 1, ff 2048, attn_axis="neighbors", `--nodes` nodes per
rank per step (weak scaling: every rank owns its own graphs; the only collective is the gradient
all-reduce).  Prints ONE JSON line (contract in the task statement); the oracle / torch port is used
only for the cpu_baseline leg and the reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "graph-transformer_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "nodes/sec U2GNN train step"
CFG = dict(d=64, k=16, T=4, L=1, ff=2048, C=2)


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
def cpu_train_rate(nodes, steps, warmup, threads):
    """nodes/s of the reference CPU train step (train mode, real p=0.5 dropouts, clip 0.5, Adam) on a
    bounded sample of the workload: `nodes` nodes of the same synthetic graph batch."""
    from oracle import torch_port as TP
    from u2gnn_b200.synthetic import make_batch
    torch.set_num_threads(threads)
    torch.manual_seed(123)
    b = make_batch(nodes, CFG["k"], CFG["d"], CFG["C"], seed=2024, device="cpu")
    model = TP.SupPort(CFG["d"], CFG["ff"], CFG["C"], CFG["T"], 0.5, CFG["L"], attn_axis="neighbors")
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=5e-4)
    N, G = b["X"].shape[0], b["G"]
    idx = torch.stack([torch.repeat_interleave(torch.arange(G), b["rowptr"][1:] - b["rowptr"][:-1]), torch.arange(N)])
    gp = torch.sparse_coo_tensor(idx, torch.ones(N), (G, N))
    soft = TP.smooth_labels(b["labels"], CFG["C"])

    def loss_fn():
        return TP.soft_ce(model(b["input_x"], gp, b["X"]), soft)

    for _ in range(warmup):
        TP.train_step(model, opt, loss_fn)
    t0 = time.perf_counter()
    for _ in range(steps):
        TP.train_step(model, opt, loss_fn)
    dt = time.perf_counter() - t0
    return N * steps / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # bounded sample: calibrate on 128 nodes, then size the step so the whole run takes ~2 minutes
    rate, _ = cpu_train_rate(128, 1, 1, threads)
    budget = 120.0
    nodes = int(max(64, min(8192, rate * budget / max(1, args.steps + args.warmup))))
    rate, sec = cpu_train_rate(nodes, args.steps, args.warmup, threads)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "nodes/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(nodes),
            "cpu_baseline": {"value": rate, "unit": "nodes/s", "cores": threads, "kind": "port",
                             "sample": "%d nodes/step of the cfg5-shape batch, torch.nn port of the reference model "
                                       "(oracle/torch_port.py), train mode with p=0.5 dropouts" % nodes},
            "e2e": {"value": rate, "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(nodes, extra=None):
    c = {"workload": "cfg5-shape synthetic graph batch: supervised U2GNN, d 64, num_neighbors 16, T 4, L 1, ff 2048, "
                     "attn_axis=neighbors (BASELINE.json configs[4])",
         "nodes_per_rank_per_step": nodes, "l2_policy": "inputs larger than L2 (X + input_x + activations >> 126 MB)"}
    if extra:
        c.update(extra)
    return c


# ----------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import u2gnn_b200 as U
    from u2gnn_b200 import engine as E
    from u2gnn_b200.synthetic import make_batch
    from u2gnn_b200.trainer import SupTrainer
    U.require_device()
    S = CFG["k"] + 1
    nodes = args.nodes
    torch.manual_seed(123)
    model = U.TransformerU2GNN(CFG["d"], CFG["ff"], CFG["C"], CFG["T"], 0.5, CFG["L"], attn_axis="neighbors").cuda()
    trainer = SupTrainer(model, lr=5e-4, precision=args.precision)
    b = make_batch(nodes, CFG["k"], CFG["d"], CFG["C"], seed=2024 + rank, device="cuda")
    G_total = b["G"] * world
    dominant = trainer.dominant_kernel()

    def step():
        return trainer.step(b["input_x"], b["rowptr"], b["X"], b["labels"], G_total=G_total)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    U.LIB.launches = 0
    U.LIB.timed = {dominant: []}
    E.FLOPS.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    launches = U.LIB.launches
    timed = U.LIB.timed[dominant]
    U.LIB.timed = None
    kern_ms = sum(a.elapsed_time(b_) for a, b_ in timed)
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = nodes * world * args.steps / (ms / 1e3)

    # ---- end-to-end through the public API with HOST buffers (H2D of the batch + D2H of the loss per step)
    host = {k: v.cpu().pin_memory() for k, v in b.items() if torch.is_tensor(v)}
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    # The input feed is pipelined the way a training loop's loader is: the host->device copy of step i+1's batch is issued on
    # a copy stream while step i computes; every timed step still copies its own inputs from pinned host memory (K copies
    # in the region) and ends with the device->host read of its loss.
    copy_stream = torch.cuda.Stream()

    def issue_copy():
        with torch.cuda.stream(copy_stream):
            dev = {k: v.cuda(non_blocking=True) for k, v in host.items()}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return dev, ev

    def run_host_steps(n):
        nxt = issue_copy()
        last = 0.0
        for i in range(n):
            dev, ev = nxt
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            for v in dev.values():
                v.record_stream(cur)
            l = trainer.step(dev["input_x"], dev["rowptr"], dev["X"], dev["labels"], G_total=G_total)
            if i + 1 < n:
                nxt = issue_copy()
            last = float(l.item())      # device->host read of the step's loss
        return last

    run_host_steps(min(2, args.warmup))
    sync()
    e0.record()
    run_host_steps(args.steps)
    e1.record()
    sync()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = nodes * world * args.steps / (float(t.item()) / 1e3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        flops_node, bytes_node = algorithmic_per_node(CFG["d"], S, CFG["T"], CFG["ff"], CFG["L"])
        ncu = None
        try:
            ncu = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_ffn_summary.json")))
        except Exception:
            pass
        roof = trainer.roofline(dominant, kern_ms, len(timed), peaks, E.FLOPS, ncu)
        line = {"metric": METRIC, "value": value, "unit": "nodes/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
                "config": workload_config(nodes, {"precision": args.precision, "parallelism": "dp%d" % world,
                                                   "algorithmic_mflop_per_node": flops_node / 1e6,
                                                   "algorithmic_bytes_per_node": bytes_node,
                                                   "loss": float(loss.item())}),
                "roofline": roof, "clocks": clk,
                "e2e": {"value": e2e, "unit": "nodes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
                "gpu_launches": launches}
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            rate, sec = cpu_train_rate(args.cpu_nodes, 2, 1, threads)
            line["cpu_baseline"] = {"value": rate, "unit": "nodes/s", "cores": threads, "kind": "port",
                                    "sample": "%d nodes/step x 2 steps (+1 warm-up) of the same cfg5-shape batch through the "
                                              "torch.nn port of the reference model, train mode" % args.cpu_nodes}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("U2GNN_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--nodes", type=int, default=int(os.environ.get("U2GNN_BENCH_NODES", 262144)),
                    help="nodes per rank per step")
    ap.add_argument("--cpu-nodes", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
