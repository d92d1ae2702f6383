"""tcgen05 / TMEM path: hardware layout self-tests and the fused bf16 FFN kernels against the fp32
CUDA path and the oracle (tolerance 2e-2, BASELINE.json north_star "bf16-FFN mode")."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def U():
    import u2gnn_b200
    u2gnn_b200.require_device()
    return u2gnn_b200


def bf16_round(x):
    return x.to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("mode,K,N", [(0, 64, 64), (0, 64, 128), (0, 128, 128), (0, 128, 64), (1, 128, 64),
                                      (1, 128, 128), (1, 32, 64), (2, 64, 128), (2, 128, 64), (3, 64, 128), (3, 128, 64)])
def test_tcgen05_operand_paths(U, mode, K, N):
    from u2gnn_b200 import engine as E
    g = torch.Generator(device="cuda").manual_seed(mode * 1000 + K + N)
    if mode == 1:
        A = torch.randn(K, 128, device="cuda", generator=g)      # At[K, M]
        B = torch.randn(K, N, device="cuda", generator=g)        # Bt[K, N]
        ref = bf16_round(A).t() @ bf16_round(B)
    else:
        A = torch.randn(128, K, device="cuda", generator=g)
        B = torch.randn(N, K, device="cuda", generator=g)
        ref = bf16_round(A) @ bf16_round(B).t()
    C = torch.zeros(128, N, device="cuda")
    scratch = torch.zeros(65536, dtype=torch.uint8, device="cuda")
    U.LIB.call("u2gnn_tc_selftest", mode, A.data_ptr(), B.data_ptr(), C.data_ptr(), K, N, scratch.data_ptr(), E._stream())
    torch.cuda.synchronize()
    err = (C - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-5, (mode, K, N, err)


# ------------------------------------------------------------------ fused FFN forward
def _bf(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(torch.bfloat16).to(torch.float32).numpy()


def _ffn_case(U, M, d, ff, p, seed=11):
    from u2gnn_b200 import engine as E
    from oracle import u2gnn_oracle as O
    rng = np.random.default_rng(seed + M + d + ff)
    y1 = rng.standard_normal((M, d)).astype(np.float32)
    W1 = (rng.standard_normal((ff, d)) / np.sqrt(d)).astype(np.float32)
    b1 = (0.1 * rng.standard_normal(ff)).astype(np.float32)
    W2 = (rng.standard_normal((d, ff)) / np.sqrt(ff)).astype(np.float32)
    b2 = (0.1 * rng.standard_normal(d)).astype(np.float32)
    gamma = (1 + 0.1 * rng.standard_normal(d)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(d)).astype(np.float32)
    thr = E.dropout_threshold(p)
    scale = 256.0 / (256.0 - thr) if thr else 1.0
    dev = lambda a: torch.from_numpy(a).cuda()
    nbytes = U.LIB.call("u2gnn_ffn_tc_packed_bytes", d, ff)
    packed = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    t = {k: dev(v) for k, v in dict(y1=y1, W1=W1, b1=b1, W2=W2, b2=b2, gamma=gamma, beta=beta).items()}
    U.LIB.call("u2gnn_ffn_tc_prepare", t["W1"].data_ptr(), t["b1"].data_ptr(), t["W2"].data_ptr(), t["b2"].data_ptr(), d, ff,
               scale, packed.data_ptr(), nbytes, E._stream())
    z = torch.full((M, d), float("nan"), device="cuda")
    stats = torch.zeros((M, 2), device="cuda")
    xn = torch.zeros((M, d), device="cuda")
    SEED, S2, S3 = 0xABCDEF0123, 18, 19
    U.LIB.call("u2gnn_ffn_tc_fwd", t["y1"].data_ptr(), M, d, ff, packed.data_ptr(), SEED, S2, S3, thr, t["gamma"].data_ptr(),
               t["beta"].data_ptr(), z.data_ptr(), stats.data_ptr(), xn.data_ptr(), E._stream())
    torch.cuda.synchronize()
    # oracle (exact fp32 semantics) and a bf16-operand emulation of what the tensor cores compute
    if thr:
        k2, _ = O.dropout_keep_mask(SEED, S2, M * ff, p)
        k3, _ = O.dropout_keep_mask(SEED, S3, M * d, p)
        m2 = k2.reshape(M, ff).astype(np.float64) * scale
        m3 = k3.reshape(M, d).astype(np.float64) * scale
    else:
        m2, m3 = np.ones((M, ff)), np.ones((M, d))
    y64 = y1.astype(np.float64)
    h = np.maximum(y64 @ W1.T.astype(np.float64) + b1, 0) * m2
    z_ref = y64 + (h @ W2.T.astype(np.float64) + b2) * m3
    hb = _bf(np.maximum(_bf(y1).astype(np.float64) @ _bf(W1).T.astype(np.float64) + b1, 0) * (m2 > 0))
    z_emu = y64 + (hb.astype(np.float64) @ _bf(W2 * scale).T.astype(np.float64) + b2) * m3
    mean = z_ref.mean(1)
    rstd = 1.0 / np.sqrt(z_ref.var(1) + 1e-5)
    xn_ref = (z_ref - mean[:, None]) * rstd[:, None] * gamma + beta
    return z.cpu().numpy(), stats.cpu().numpy(), xn.cpu().numpy(), z_ref, z_emu, mean, rstd, xn_ref


@pytest.mark.parametrize("M,d,ff,p", [(256, 64, 256, 0.0), (1000, 64, 2048, 0.0), (37, 64, 128, 0.0), (257, 64, 1024, 0.5),
                                      (300, 7, 256, 0.5), (129, 12, 128, 0.0), (300, 64, 2048, 0.25),
                                      (148 * 512 + 5, 64, 256, 0.5)])
def test_ffn_tc_forward(U, M, d, ff, p):
    z, stats, xn, z_ref, z_emu, mean, rstd, xn_ref = _ffn_case(U, M, d, ff, p)
    scale = np.abs(z_ref).max()
    assert np.isfinite(z).all()
    assert np.abs(z - z_emu).max() / scale < 2e-3          # same bf16 operands, fp32 accumulation
    assert np.abs(z - z_ref).max() / scale < 2e-2          # tolerance of the bf16-FFN mode (north_star)
    assert np.abs(stats[:, 0] - mean).max() < 2e-2 * max(1.0, np.abs(mean).max())
    assert np.abs(stats[:, 1] / rstd - 1).max() < 2e-2
    assert np.abs(xn - xn_ref).max() / np.abs(xn_ref).max() < 2e-2
