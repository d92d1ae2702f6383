"""Microbenchmark of the fused tcgen05 FFN kernels (CUDA events, inputs larger than L2)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-transformer_b200"))
import torch
import u2gnn_b200 as U
from u2gnn_b200 import engine as E

def main():
    d, ff = 64, 2048
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * 1024 * 1024
    thr = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    emit = int(sys.argv[3]) if len(sys.argv) > 3 else 1      # 1 = write the mask words as the product step does
    g = torch.Generator(device="cuda").manual_seed(0)
    y1 = torch.randn(M, d, device="cuda", generator=g)
    W1 = torch.randn(ff, d, device="cuda", generator=g) / 8
    W2 = torch.randn(d, ff, device="cuda", generator=g) / 45
    b1 = torch.zeros(ff, device="cuda"); b2 = torch.zeros(d, device="cuda")
    gamma = torch.ones(d, device="cuda"); beta = torch.zeros(d, device="cuda")
    nb = U.LIB.call("u2gnn_ffn_tc_packed_bytes", d, ff)
    packed = torch.zeros(nb, dtype=torch.uint8, device="cuda")
    U.LIB.call("u2gnn_ffn_tc_prepare", W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), d, ff, 2.0, packed.data_ptr(), nb, E._stream())
    z = torch.empty_like(y1); xn = torch.empty_like(y1); st = torch.empty(M, 2, device="cuda")
    mask = torch.empty(U.LIB.call("u2gnn_ffn_tc_mask_bytes", M, ff), dtype=torch.uint8, device="cuda") if emit else None
    def run():
        U.LIB.call("u2gnn_ffn_tc_fwd", y1.data_ptr(), M, d, ff, packed.data_ptr(), 1, 2, 3, thr, gamma.data_ptr(), beta.data_ptr(),
                   z.data_ptr(), st.data_ptr(), xn.data_ptr(), mask.data_ptr() if emit else 0, E._stream())
    for _ in range(3): run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): run()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    fl = 4.0 * M * d * ff
    print(json.dumps({"kernel": "ffn_tc_fwd", "M": M, "thr": thr, "emit": emit, "ms": ms, "tflops": fl / ms / 1e9}))

if __name__ == "__main__":
    main()
