"""Times the bf16-split tcgen05 GEMMs of the fp32 mode at one full-size timestep (rows = nodes * 17, d 64, ff 2048) with CUDA events:
linear1 (+ReLU+dropout), linear2, dPre (aux mask), dy1 (beta 1), dW2, dW1, in_proj, dx.  GB/s = the bytes each product must move once."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-transformer_b200"))
import torch
import u2gnn_b200 as U
from u2gnn_b200 import engine as E

nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
M, d, ff = nodes * 17, 64, 2048
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
y1, df = rnd(M, d), rnd(M, d)
W1, b1, W2, b2 = rnd(ff, d) / 8, rnd(ff), rnd(d, ff) / 45, rnd(d)
Win, bin_ = rnd(3 * d, d) / 8, rnd(3 * d)
hd = torch.empty(M, ff, device="cuda")
dpre = torch.empty(M, ff, device="cuda")
f = torch.empty(M, d, device="cuda")
qkv = torch.empty(M, 3 * d, device="cuda")
dW1, db1, dW2, dWin, dbin = torch.zeros(ff, d, device="cuda"), torch.zeros(ff, device="cuda"), torch.zeros(d, ff, device="cuda"), torch.zeros(3 * d, d, device="cuda"), torch.zeros(3 * d, device="cuda")
cases = [
    ("linear1+relu+drop", lambda: E.linear_fp32(y1, M, d, W1, 0, ff, hd, bias=b1, relu=True, drop=(1, 2, 128)), M * (4 * d + 4 * ff)),
    ("linear2", lambda: E.linear_fp32(hd, M, ff, W2, 0, d, f, bias=b2), M * (4 * d + 4 * ff)),
    ("dPre (aux)", lambda: E.linear_fp32(df, M, d, W2, 1, ff, dpre, aux=hd, aux_scale=2.0), M * (4 * d + 8 * ff)),
    ("dy1 (beta)", lambda: E.linear_fp32(dpre, M, ff, W1, 1, d, f, beta=1.0), M * (8 * d + 4 * ff)),
    ("dW2", lambda: E.wgrad_fp32(df, M, d, hd, ff, dW2, None), M * (4 * d + 4 * ff)),
    ("dW1+db1", lambda: E.wgrad_fp32(dpre, M, ff, y1, d, dW1, db1), M * (4 * d + 4 * ff)),
    ("in_proj", lambda: E.linear_fp32(y1, M, d, Win, 0, 3 * d, qkv, bias=bin_), M * (4 * d + 12 * d)),
    ("dx (beta)", lambda: E.linear_fp32(qkv, M, 3 * d, Win, 1, d, f, beta=1.0), M * (12 * d + 8 * d)),
    ("dWin+db", lambda: E.wgrad_fp32(qkv, M, 3 * d, y1, d, dWin, dbin), M * (4 * d + 12 * d)),
]
for name, fn, nbytes in cases:
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(5):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print("%-20s %8.3f ms  %7.0f GB/s (%.2f of 6552)" % (name, ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / 6552))
