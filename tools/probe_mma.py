import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-transformer_b200"))
import torch
import u2gnn_b200 as U
from u2gnn_b200 import engine as E
from u2gnn_b200._lib import probe_lib
PROBE = probe_lib()     # libu2gnn_b200_probe.so: the product library exports no probe / trace entry points
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for N in (64, 80, 128, 256):
    for ts in (0, 1, 2):
        for rot in ((2,) if ts == 2 else (0, 2)):
            for count in (64, 512):
                PROBE.call("u2gnn_tc_probe", N, ts, rot, count, out.data_ptr(), E._stream())
                torch.cuda.synchronize()
                a, b = out.tolist()
                print("N=%3d %s rotate=%d count=%4d  issue %6.1f cyc/mma  total %6.1f cyc/mma  (ideal %d)" % (N, ("SS", "TS", "SS-MN")[ts], rot, count, a / count, b / count, 128 * N // 256))
