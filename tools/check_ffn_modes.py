"""The experiment switches of the fused FFN forward (u2gnn_ffn_tc_debug) must not change results: compare every mode
bit-for-bit with mode 0, then time each."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-transformer_b200"))
import torch
import u2gnn_b200 as U
from u2gnn_b200 import engine as E

def main():
    modes = [int(a) for a in sys.argv[1:]] or [0, 8]
    d, ff = 64, 2048
    M = 4 * 1024 * 1024 + 77
    g = torch.Generator(device="cuda").manual_seed(0)
    y1 = torch.randn(M, d, device="cuda", generator=g)
    W1 = torch.randn(ff, d, device="cuda", generator=g) / 8
    W2 = torch.randn(d, ff, device="cuda", generator=g) / 45
    b1 = torch.randn(ff, device="cuda", generator=g) * 0.1; b2 = torch.randn(d, device="cuda", generator=g) * 0.1
    gamma = torch.ones(d, device="cuda"); beta = torch.zeros(d, device="cuda")
    nb = U.LIB.call("u2gnn_ffn_tc_packed_bytes", d, ff)
    packed = torch.zeros(nb, dtype=torch.uint8, device="cuda")
    U.LIB.call("u2gnn_ffn_tc_prepare", W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), d, ff, 2.0, packed.data_ptr(), nb, E._stream())
    ref = None
    for thr in (128, 0):
        for dbg in modes:
            U.LIB.call("u2gnn_ffn_tc_debug", dbg)
            z = torch.empty_like(y1); xn = torch.empty_like(y1); st = torch.empty(M, 2, device="cuda")
            def run():
                U.LIB.call("u2gnn_ffn_tc_fwd", y1.data_ptr(), M, d, ff, packed.data_ptr(), 1, 2, 3, thr, gamma.data_ptr(), beta.data_ptr(),
                           z.data_ptr(), st.data_ptr(), xn.data_ptr(), E._stream())
            for _ in range(3): run()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5): run()
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            if dbg == modes[0]:
                ref = (z.clone(), xn.clone(), st.clone())
                same = True
            else:
                same = bool(torch.equal(z, ref[0]) and torch.equal(xn, ref[1]) and torch.equal(st, ref[2]))
            print(json.dumps({"thr": thr, "dbg": dbg, "ms": round(ms, 4), "tflops": round(4.0 * M * d * ff / ms / 1e9, 1), "identical_to_first_mode": same}))
    U.LIB.call("u2gnn_ffn_tc_debug", 0)

if __name__ == "__main__":
    main()
