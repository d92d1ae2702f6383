"""Pipeline trace of the fused FFN backward kernels (dgrad, wgrad): clock stamps of CTA 0 -> per-phase cycle medians."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-transformer_b200"))
import numpy as np
import torch
import u2gnn_b200 as U
from u2gnn_b200 import engine as E
from u2gnn_b200._lib import probe_lib
PROBE = probe_lib()     # libu2gnn_b200_probe.so: the product library exports no probe / trace entry points

d, ff = 64, 2048
pairs = 6
M = 148 * 256 * pairs
thr = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g = torch.Generator(device="cuda").manual_seed(0)
y1 = torch.randn(M, d, device="cuda", generator=g); df = torch.randn(M, d, device="cuda", generator=g); dz = torch.randn(M, d, device="cuda", generator=g)
W1 = torch.randn(ff, d, device="cuda", generator=g) / 8; W2 = torch.randn(d, ff, device="cuda", generator=g) / 45
b1 = torch.zeros(ff, device="cuda"); b2 = torch.zeros(d, device="cuda")
nb = PROBE.call("u2gnn_ffn_tc_packed_bytes", d, ff)
packed = torch.zeros(nb, dtype=torch.uint8, device="cuda")
PROBE.call("u2gnn_ffn_tc_prepare", W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), d, ff, 2.0, packed.data_ptr(), nb, E._stream())
dy = torch.empty_like(y1); dW1 = torch.zeros_like(W1); db1 = torch.zeros_like(b1); dW2 = torch.zeros_like(W2)
ws = torch.empty(PROBE.call("u2gnn_ffn_tc_bwd_workspace_bytes", M), dtype=torch.uint8, device="cuda")
def run():
    PROBE.call("u2gnn_ffn_tc_bwd", y1.data_ptr(), df.data_ptr(), 0, 0, 0, dz.data_ptr(), M, d, ff, packed.data_ptr(), 2.0, 1, 2, thr,
               dy.data_ptr(), dW1.data_ptr(), db1.data_ptr(), dW2.data_ptr(), ws.data_ptr(), ws.numel(), E._stream())
CAP, SLOTS = 1024, 64
tr = torch.zeros(SLOTS * CAP, dtype=torch.int32, device="cuda")
run(); torch.cuda.synchronize()
PROBE.call("u2gnn_ffn_tc_set_trace", tr.data_ptr())
run(); torch.cuda.synchronize()
PROBE.call("u2gnn_ffn_tc_set_trace", 0)
t = tr.cpu().numpy().astype(np.int64).reshape(SLOTS, CAP)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", "ffn_bwd_trace.npy"), t)
NC = ff // 128
med = lambda a: float(np.median(a))
# ---------------- dgrad (slots 32..48): MMA warp per (chunk, tile): [before p_full wait, after]
m = t[32][: pairs * NC * 4].reshape(pairs, NC, 2, 2)
print("== dgrad: CTA0 last MMA stamp %d cycles = %.0f per chunk-pair incl. pair boundaries" % (m[-1, -1, 1, 1], m[-1, -1, 1, 1] / (pairs * NC)))
print("chunk-pair period median %.0f | MMA waits for dPre: tile0 %.0f tile1 %.0f | gaps between pairs %s" % (
    med(np.diff(m[1:, :, 0, 0], axis=1)), med(m[1:, :, 0, 1] - m[1:, :, 0, 0]), med(m[1:, :, 1, 1] - m[1:, :, 1, 0]),
    (m[1:, 0, 0, 0] - m[:-1, -1, 1, 1]).tolist()))
for sl in (33, 37, 41, 45):      # epilogue warps: per pair [x_free wait start, img landed, row in TMEM] + per chunk [before d_full wait, after, dPre stored] + [chunk loop end]
    e = t[sl][: pairs * (4 + NC * 3)].reshape(pairs, 4 + NC * 3)
    body = e[1:, 3:3 + NC * 3].reshape(-1, NC, 3)
    dd = np.diff(body, axis=2)
    print("epi slot %d: wait x_free + image %.0f | row -> TMEM %.0f | per chunk: wait D %.0f | D -> dPre %.0f | chunk period %.0f | drain -> next pair %.0f" % (
        sl, med(e[1:, 1] - e[1:, 0]), med(e[1:, 2] - e[1:, 1]), med(dd[..., 0]), med(dd[..., 1]), med(np.diff(body[:, :, 0], axis=1)),
        med(e[2:, 0] - e[1:-1, -1])))
# ---------------- wgrad (slots 0..16): MMA per iteration n >= 1: [a_done wait start, end, p_full(n-1) wait start, end, W1 issued]
n_tiles = (M + 127) // 128
my = (n_tiles + 8) // 9
mm = t[0][2: 2 + (my - 2) * 5].reshape(-1, 5)[4:]
print("== wgrad: tiles of CTA0 %d; per-tile period median %.0f mean %.0f" % (my, med(np.diff(mm[:, 0])), np.diff(mm[:, 0]).mean()))
print("MMA: wait a_done %.0f | issue W2(n) D(n), wait b_done, S(n+1) %.0f | wait p_full %.0f | issue W1(n-1) %.0f | loop %.0f" % (
    med(mm[:, 1] - mm[:, 0]), med(mm[:, 2] - mm[:, 1]), med(mm[:, 3] - mm[:, 2]), med(mm[:, 4] - mm[:, 3]), med(mm[1:, 0] - mm[:-1, 4])))
# epilogue warps per iteration n >= 1: A: [before s wait, after, a_done arrived], B(n-1): [before d wait, after, dPre computed, p_free seen, stored in TMEM]
for sl in (1, 6, 11, 16):
    k = min(my - 2, (1024 - 3) // 8)
    e = t[sl][3: 3 + k * 8].reshape(k, 8)[4:]
    dd = np.diff(e, axis=1)
    print("epi slot %2d: wait S %.0f | A %.0f | gap %.0f | wait D %.0f | B ld+cvt %.0f | wait p_free %.0f | st+arrive %.0f | loop %.0f | period %.0f" % (
        sl, med(dd[:, 0]), med(dd[:, 1]), med(dd[:, 2]), med(dd[:, 3]), med(dd[:, 4]), med(dd[:, 5]), med(dd[:, 6]), med(e[1:, 0] - e[:-1, 7]), med(np.diff(e[:, 0]))))
