import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-transformer_b200"))
import torch
import u2gnn_b200 as U
from u2gnn_b200 import engine as E
from u2gnn_b200._lib import probe_lib
PROBE = probe_lib()     # libu2gnn_b200_probe.so: the product library exports no probe / trace entry points
out = torch.zeros(4, dtype=torch.int64, device="cuda")
for warps in (1, 4, 8, 16):
    for batch in (1, 2):
        PROBE.call("u2gnn_tmem_bw_probe", warps, 2000, batch, out.data_ptr(), E._stream())
        torch.cuda.synchronize()
        cyc, byt = out.tolist()[:2]
        print("warps=%2d batch=%d  %8d cycles  %6.1f B/cycle/SM  %6.1f cycles per LDTM.x32 per warp" % (warps, batch, cyc, byt / cyc, cyc / (2000 * batch)))
