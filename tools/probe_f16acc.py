"""Hardware probe: tcgen05.mma kind::f16 with an fp16 accumulator (layout of D in tensor memory) and with an fp16 A
operand in tensor memory against a bf16 B."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-transformer_b200"))
import numpy as np, torch
import u2gnn_b200 as U
from u2gnn_b200 import engine as E
from u2gnn_b200._lib import probe_lib
PROBE = probe_lib()     # libu2gnn_b200_probe.so: the product library exports no probe / trace entry points
torch.manual_seed(0)
K, N = 64, 128
A = torch.randn(128, K, device="cuda"); B = torch.randn(N, K, device="cuda")
bf = lambda x: x.to(torch.bfloat16).float()
ref = bf(A) @ bf(B).t()
scratch = torch.zeros(65536, dtype=torch.uint8, device="cuda")
C = torch.zeros(128, N, device="cuda")
PROBE.call("u2gnn_tc_selftest", 4, A.data_ptr(), B.data_ptr(), C.data_ptr(), K, N, scratch.data_ptr(), E._stream())
torch.cuda.synchronize()
raw = C.cpu().numpy().view(np.uint32)
lo = (raw & 0xFFFF).astype(np.uint16).view(np.float16).astype(np.float32)
hi = (raw >> 16).astype(np.uint16).view(np.float16).astype(np.float32)
r = ref.cpu().numpy()
print("row0 ref[:8]   ", r[0, :8])
print("row0 lo16[:8]  ", lo[0, :8])
print("row0 hi16[:8]  ", hi[0, :8])
print("hyp unpacked (col j = elem j in low half): max err", np.abs(lo - r).max(), " hi halves max", np.abs(hi).max())
pk = np.stack([lo[:, : N // 2], hi[:, : N // 2]], axis=2).reshape(128, N)
print("hyp packed (col j = elems 2j,2j+1): max err", np.abs(pk - r).max())
print("cols N/2.. raw nonzero:", int((raw[:, N // 2:] != 0).sum()))
C2 = torch.zeros(128, N, device="cuda")
PROBE.call("u2gnn_tc_selftest", 5, A.data_ptr(), B.data_ptr(), C2.data_ptr(), K, N, scratch.data_ptr(), E._stream())
torch.cuda.synchronize()
ref5 = A.half().float() @ bf(B).t()
print("fp16 A (TMEM) x bf16 B: rel err", ((C2 - ref5).abs().max() / ref5.abs().max()).item())
