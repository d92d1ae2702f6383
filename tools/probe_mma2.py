"""tcgen05.mma cta_group::2 (M = 256 across a CTA pair) issue / execute rate against cta_group::1 (M = 128), by N and operand mode."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-transformer_b200"))
import torch
import u2gnn_b200 as U
from u2gnn_b200 import engine as E
from u2gnn_b200._lib import probe_lib
PROBE = probe_lib()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for N in (64, 96, 128, 256):
    for ts in (0, 1):
        for count in (64, 512):
            PROBE.call("u2gnn_tc_probe", N, ts, 2, count, out.data_ptr(), E._stream())
            torch.cuda.synchronize()
            a1, b1 = out.tolist()
            PROBE.call("u2gnn_tc_probe2", N, ts, count, out.data_ptr(), E._stream())
            torch.cuda.synchronize()
            a2, b2 = out.tolist()
            print("N=%3d %s count=%4d | cta_group::1 M=128: %6.1f cyc/mma (ideal %3d) | cta_group::2 M=256: issue %6.1f total %6.1f cyc/mma (ideal %3d per pair) -> %.2fx the MACs per cycle per SM"
                  % (N, ("SS", "TS")[ts], count, b1 / count, 128 * N // 256, a2 / count, b2 / count, 128 * N // 256, (b1 / count) / (b2 / count)))
