"""Pipeline trace of the fused FFN forward kernel: clock stamps of CTA 0's MMA warp and epilogue warps
(u2gnn_ffn_tc_set_trace) -> per-phase cycle averages in steady state."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-transformer_b200"))
import numpy as np
import torch
import u2gnn_b200 as U
from u2gnn_b200 import engine as E
from u2gnn_b200._lib import probe_lib
PROBE = probe_lib()     # libu2gnn_b200_probe.so: the product library exports no probe / trace entry points

def main():
    d, ff = 64, 2048
    pairs = 6
    M = 148 * 256 * pairs
    thr = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    g = torch.Generator(device="cuda").manual_seed(0)
    y1 = torch.randn(M, d, device="cuda", generator=g)
    W1 = torch.randn(ff, d, device="cuda", generator=g) / 8
    W2 = torch.randn(d, ff, device="cuda", generator=g) / 45
    b1 = torch.zeros(ff, device="cuda"); b2 = torch.zeros(d, device="cuda")
    gamma = torch.ones(d, device="cuda"); beta = torch.zeros(d, device="cuda")
    nb = PROBE.call("u2gnn_ffn_tc_packed_bytes", d, ff)
    packed = torch.zeros(nb, dtype=torch.uint8, device="cuda")
    PROBE.call("u2gnn_ffn_tc_prepare", W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), d, ff, 2.0, packed.data_ptr(), nb, E._stream())
    z = torch.empty_like(y1); xn = torch.empty_like(y1); st = torch.empty(M, 2, device="cuda")
    emit = int(sys.argv[2]) if len(sys.argv) > 2 else 1       # 1 = write the mask words as the product step does
    mask = torch.empty(PROBE.call("u2gnn_ffn_tc_mask_bytes", M, ff), dtype=torch.uint8, device="cuda") if emit else None
    CAP = 1024
    SLOTS = 21
    tr = torch.zeros(SLOTS * CAP, dtype=torch.int32, device="cuda")
    def run():
        PROBE.call("u2gnn_ffn_tc_fwd", y1.data_ptr(), M, d, ff, packed.data_ptr(), 1, 2, 3, thr, gamma.data_ptr(), beta.data_ptr(),
                   z.data_ptr(), st.data_ptr(), xn.data_ptr(), mask.data_ptr() if emit else 0, E._stream())
    run(); torch.cuda.synchronize()
    PROBE.call("u2gnn_ffn_tc_set_trace", tr.data_ptr())
    run(); torch.cuda.synchronize()
    PROBE.call("u2gnn_ffn_tc_set_trace", 0)
    t = tr.cpu().numpy().astype(np.int64).reshape(SLOTS, CAP)
    NC = ff // 128
    # MMA warp: per (c, i): [before h wait, after h wait]
    m = t[0][: pairs * NC * 4].reshape(pairs, NC, 2, 2)
    print("kernel cycles (CTA0 last MMA stamp): %d = %.0f per chunk-pair incl. pair boundaries" % (m[-1, -1, -1, -1], m[-1, -1, -1, -1] / (pairs * NC)))
    print("pair starts:", m[:, 0, 0, 0].tolist())
    print("gap last stamp of pair -> first of next:", (m[1:, 0, 0, 0] - m[:-1, -1, 1, 1]).tolist())
    per_chunk = np.diff(m[:, :, 0, 0], axis=1)
    print("MMA warp: cycles per chunk-pair  median %.0f  mean %.0f (pairs 1..)" % (np.median(per_chunk[1:]), per_chunk[1:].mean()))
    hw = (m[..., 1] - m[..., 0])[1:]
    print("MMA warp: wait for H  tile0 median %.0f  tile1 median %.0f" % (np.median(hw[:, :, 0]), np.median(hw[:, :, 1])))
    iss = (m[1:, :, 1, 0] - m[1:, :, 0, 1])
    print("MMA warp: tile0 h-ready -> tile1 wait start (issue G2+G1+commits) median %.0f" % np.median(iss))
    # epilogue warps: per (c, i) 5 stamps: before s wait, after s wait, after ld wait, before st, after arrive
    for w in (1, 6, 11, 16):
        e = t[w][: pairs * NC * 10].reshape(pairs, NC, 2, 5)[1:]
        dd = np.diff(e, axis=3)
        for i in (0, 1):
            print("epi warp slot %2d tile %d: wait S %.0f | ld %.0f | compute %.0f | st+fence+arrive %.0f" % (
                w, i, np.median(dd[:, :, i, 0]), np.median(dd[:, :, i, 1]), np.median(dd[:, :, i, 2]), np.median(dd[:, :, i, 3])))
    e_all = np.stack([t[w][: pairs * NC * 10].reshape(pairs, NC, 2, 5) for w in range(1, 17)])
    for i in (0, 1):
        last_arrive = e_all[:, :, :, i, 4].max(axis=0)
        first_arrive = e_all[:, :, :, i, 4].min(axis=0)
        s_rel = e_all[:, :, :, i, 1].min(axis=0)
        print("tile%d: spread of the 16 warps' arrive %.0f; last arrive -> MMA warp released %.0f; S seen -> last arrive %.0f" % (
            i, np.median((last_arrive - first_arrive)[1:]), np.median((m[:, :, i, 1] - last_arrive)[1:]), np.median((last_arrive - s_rel)[1:])))
        print("tile%d: MMA h-ready(c) -> epilogue sees S(c+1) %.0f" % (i, np.median((s_rel[1:, 1:] - m[1:, :-1, i, 1]))))
    io = t[17][: pairs * 4].reshape(pairs, 2, 2)
    print("I/O warp: y_full wait start / end per pair, tile 0:", io[:, 0].tolist())
    print("I/O warp: y_full wait start / end per pair, tile 1:", io[:, 1].tolist())
    np.save(os.path.join(ROOT, "gpurun_out", "ffn_trace.npy"), t)

if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    main()
