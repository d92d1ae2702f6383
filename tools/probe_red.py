"""L2 reduction (red.global.add) throughput for the split-K pattern of the FFN backward: 16 CTAs -> one [128 x 64] tile."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-transformer_b200"))
import torch
import u2gnn_b200 as U
from u2gnn_b200 import engine as E
from u2gnn_b200._lib import probe_lib
PROBE = probe_lib()     # libu2gnn_b200_probe.so: the product library exports no probe / trace entry points
n_tiles = 34816                      # 4.456 M rows
buf = torch.zeros(n_tiles * 8192, device="cuda")
for groups in (16, 8, 1):
    for mode in (0, 3, 1, 2, 4, 5, 6):
        def run(): PROBE.call("u2gnn_red_probe", buf.data_ptr(), n_tiles, groups, mode, E._stream())
        run(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        gb = n_tiles * 32768 * groups / 1e9
        print("groups=%2d mode=%d (%s): %.3f ms for %.1f GB of updates = %.2f TB/s" % (groups, mode, ("red.v4", "atomicAdd x4", "st.v4", "red.v4 row-per-thread", "bulk red 256 B x 128 threads", "bulk red 32 KB x 1 thread", "bulk red 64-row halves, one warp, read-wait")[mode], ms, gb, gb / ms))
