"""One full-size encoder layer (argv[3]: bf16 (default) or fp32) (forward + backward) on synthetic rows: the launches `tools/ncu_capture_r02.sh` profiles.
Rows = nodes * S (default 65536 * 17 = 1 114 112, the per-timestep row count of `bench.py --nodes 65536`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-transformer_b200"))
import torch
import u2gnn_b200 as U
from u2gnn_b200 import engine as E

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
PREC = sys.argv[3] if len(sys.argv) > 3 else "bf16"
S, d, ff, thr = 17, 64, 2048, 128
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
p = {"self_attn.in_proj_weight": rnd(3 * d, d) / 8, "self_attn.in_proj_bias": 0.1 * rnd(3 * d), "self_attn.out_proj.weight": rnd(d, d) / 8,
     "self_attn.out_proj.bias": 0.1 * rnd(d), "linear1.weight": rnd(ff, d) / 8, "linear1.bias": 0.1 * rnd(ff), "linear2.weight": rnd(d, ff) / 45,
     "linear2.bias": 0.1 * rnd(d), "norm1.weight": 1 + 0.1 * rnd(d), "norm1.bias": 0.1 * rnd(d), "norm2.weight": 1 + 0.1 * rnd(d), "norm2.bias": 0.1 * rnd(d)}
gr = {n: torch.zeros_like(v) for n, v in p.items()}
x = rnd(B * S, d)
dy = rnd(B * S, d)
for _ in range(iters):
    y, sv = E.encoder_layer_fwd(x, B, S, S, p, d, ff, [16, 17, 18, 19], 123, thr, False, PREC)
    dx = E.encoder_layer_bwd(dy, sv, p, gr, d, ff, [16, 17, 18, 19], 123, thr, False, need_dx=True)
torch.cuda.synchronize()
print("rows", B * S, "ok", bool(torch.isfinite(dx).all()))
