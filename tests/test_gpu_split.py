"""fp32 mode on the tensor cores: the three-product bf16-split GEMMs (csrc/gemm_split.cu) against fp64 numpy.

The entry points replace u2gnn_sgemm on F.linear and its autograd (linear1 / linear2 / in_proj / out_proj of
torch/nn/modules/transformer.py:944-982) in precision="fp32"; the bar is the fp32 tolerance of BASELINE.json's north_star
(1e-4 relative) - the split itself is good for ~2e-5, which is what the kernel-level checks ask for.  The model-level fixtures
(tests/test_gpu_parity.py) run through these kernels because engine.FP32_TC is on by default; the last test here repeats two
of them on the CUDA-core SGEMM so that path stays covered as well.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import u2gnn_oracle as O

pytestmark = pytest.mark.gpu
TOL = 3e-5


@pytest.fixture(scope="module")
def U():
    import u2gnn_b200
    u2gnn_b200.require_device()
    return u2gnn_b200


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def call_rows(U, A, W, w_kn, N, bias=None, epi=0, seed=0, stream=0, thr=0, aux=None, aux_scale=1.0, beta=0.0, C0=None, lda=None,
              ldw=None, ldc=None, packed=None):
    from u2gnn_b200 import engine as E
    M, K = A.shape
    lda = lda or A.stride(0)
    ldw = ldw or W.stride(0)
    C = torch.zeros((M, ldc or N), dtype=torch.float32, device="cuda") if C0 is None else C0
    U.LIB.call("u2gnn_gemm_split_rows", A.data_ptr(), M, K, lda, W.data_ptr(), w_kn, ldw, N, E._ptr(bias), epi, seed, stream, thr, 0,
               E._ptr(aux), 0 if aux is None else aux.stride(0), aux_scale, beta, C.data_ptr(), C.stride(0), E._ptr(packed), E._stream())
    torch.cuda.synchronize()
    return C


@pytest.mark.parametrize("M,K,N,w_kn", [(1000, 64, 192, 0), (300, 7, 21, 0), (257, 2048, 64, 0), (513, 2048, 64, 1), (200, 65, 195, 0),
                                         (129, 195, 65, 1), (4097, 64, 2048, 0), (128, 1, 1, 0), (640, 192, 64, 1), (77, 100, 300, 1)])
def test_split_rows_matches_fp64(U, M, K, N, w_kn):
    rng = np.random.default_rng(M + K + N)
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((K, N) if w_kn else (N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    ref = A.astype(np.float64) @ (W.astype(np.float64) if w_kn else W.astype(np.float64).T) + b
    C = call_rows(U, dev(A), dev(W), w_kn, N, bias=dev(b), epi=1)
    assert rel_err(C.cpu().numpy(), ref) < TOL
    # the same product with the weights pre-packed into the kernel's operand images (one bulk copy per step): bit-identical
    from u2gnn_b200 import engine as E
    Wd = dev(W)
    nb = U.LIB.call("u2gnn_gemm_split_packed_bytes", N, K)
    pk = torch.empty(nb, dtype=torch.uint8, device="cuda")
    U.LIB.call("u2gnn_gemm_split_pack", Wd.data_ptr(), w_kn, Wd.stride(0), N, K, pk.data_ptr(), nb, E._stream())
    Cp = call_rows(U, dev(A), Wd, w_kn, N, bias=dev(b), epi=1, packed=pk)
    assert torch.equal(Cp, C)
    # beta = 1 accumulates over the old C
    C0 = rng.standard_normal((M, N)).astype(np.float32)
    C2 = call_rows(U, dev(A), dev(W), w_kn, N, beta=1.0, C0=dev(C0))
    assert rel_err(C2.cpu().numpy(), ref - b + C0) < TOL


def test_split_rows_is_more_accurate_than_one_bf16_product(U):
    """The point of the split: a plain bf16 product is ~4e-3 off, the three-product split ~1e-5."""
    rng = np.random.default_rng(5)
    A = rng.standard_normal((512, 256)).astype(np.float32)
    W = rng.standard_normal((64, 256)).astype(np.float32)
    ref = A.astype(np.float64) @ W.astype(np.float64).T
    C = call_rows(U, dev(A), dev(W), 0, 64).cpu().numpy()
    one = (dev(A).bfloat16().float() @ dev(W).bfloat16().float().T).cpu().numpy()
    assert rel_err(C, ref) < TOL < 1e-3 < rel_err(one, ref)


def test_split_rows_strided_operands_and_unaligned_pointers(U):
    rng = np.random.default_rng(11)
    M, K, N = 333, 64, 128
    Ab = rng.standard_normal((M, K + 8)).astype(np.float32)
    Wb = rng.standard_normal((N, K + 4)).astype(np.float32)
    Cb = torch.zeros((M, N + 12), device="cuda")
    A, W = dev(Ab), dev(Wb)
    ref = Ab[:, :K].astype(np.float64) @ Wb[:, :K].astype(np.float64).T
    from u2gnn_b200 import engine as E
    U.LIB.call("u2gnn_gemm_split_rows", A.data_ptr(), M, K, K + 8, W.data_ptr(), 0, K + 4, N, 0, 0, 0, 0, 0, 0, 0, 0, 1.0, 0.0,
               Cb.data_ptr(), N + 12, 0, E._stream())
    assert rel_err(Cb[:, :N].cpu().numpy(), ref) < TOL
    assert float(Cb[:, N:].abs().max()) == 0.0                       # nothing written outside the N columns
    # odd element offsets: the scalar path
    A1 = dev(np.concatenate([[0.0], Ab[:, :K].reshape(-1)]).astype(np.float32))
    C1 = torch.zeros((M, N), device="cuda")
    U.LIB.call("u2gnn_gemm_split_rows", A1.data_ptr() + 4, M, K, K, W.data_ptr(), 0, K + 4, N, 0, 0, 0, 0, 0, 0, 0, 0, 1.0, 0.0,
               C1.data_ptr(), N, 0, E._stream())
    assert rel_err(C1.cpu().numpy(), ref) < TOL


@pytest.mark.parametrize("N", [2048, 1000, 50])
def test_split_rows_relu_dropout_epilogue_uses_the_engine_stream(U, N):
    """linear1 + ReLU + dropout (transformer.py:977-982): keep mask = the oracle's restatement of the counter-based stream on the
    linear index m * N + n."""
    rng = np.random.default_rng(N)
    M, K = 300, 64
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / 8).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    seed, stream, p = 0x1234ABCD5678, 23, 0.5
    keep, scale = O.dropout_keep_mask(seed, stream, M * N, p)
    ref = np.maximum(A.astype(np.float64) @ W.astype(np.float64).T + b, 0.0) * keep.reshape(M, N) * scale
    C = call_rows(U, dev(A), dev(W), 0, N, bias=dev(b), epi=1 | 2 | 4, seed=seed, stream=stream, thr=128).cpu().numpy()
    assert rel_err(C, ref) < TOL
    assert np.array_equal(C == 0.0, ref == 0.0) or np.mean((C == 0.0) != (ref == 0.0)) < 1e-4   # sign flips only where |pre-activation| ~ 0


def test_split_rows_aux_mask_epilogue(U):
    """dH -> dPre: gradient through ReLU and dropout by the saved hidden (u2gnn_sgemm epilogue bit 8)."""
    rng = np.random.default_rng(3)
    M, K, N = 260, 64, 1024
    df = rng.standard_normal((M, K)).astype(np.float32)
    W2 = (rng.standard_normal((K, N)) / 8).astype(np.float32)          # linear2.weight [d, ff] read as [K][N]
    hd = np.maximum(rng.standard_normal((M, N)), 0.0).astype(np.float32)
    ref = (df.astype(np.float64) @ W2.astype(np.float64)) * (hd > 0) * 2.0
    C = call_rows(U, dev(df), dev(W2), 1, N, epi=8, aux=dev(hd), aux_scale=2.0).cpu().numpy()
    assert rel_err(C, ref) < TOL


@pytest.mark.parametrize("M,N1,N2", [(1000, 192, 64), (77, 21, 7), (5000, 2048, 64), (4000, 64, 64), (300, 195, 65), (129, 130, 1),
                                      (64, 128, 64), (1, 5, 3)])
def test_split_wgrad_matches_fp64(U, M, N1, N2):
    from u2gnn_b200 import engine as E
    rng = np.random.default_rng(M + N1 + N2)
    A = rng.standard_normal((M, N1)).astype(np.float32)
    B = rng.standard_normal((M, N2)).astype(np.float32)
    dW0 = rng.standard_normal((N1, N2)).astype(np.float32)
    db0 = rng.standard_normal(N1).astype(np.float32)
    ref = dW0 + A.astype(np.float64).T @ B.astype(np.float64)
    refb = db0 + A.astype(np.float64).sum(0)
    dW, db, Ad, Bd = dev(dW0), dev(db0), dev(A), dev(B)         # (named: a temporary's memory could be reused by the next upload)
    U.LIB.call("u2gnn_gemm_split_wgrad", Ad.data_ptr(), M, N1, N1, Bd.data_ptr(), N2, N2, dW.data_ptr(), N2, 1, db.data_ptr(),
               E._stream())
    scale = np.sqrt(M)
    assert np.abs(dW.cpu().numpy() - ref).max() < TOL * max(np.abs(ref).max(), scale)
    assert np.abs(db.cpu().numpy() - refb).max() < TOL * max(np.abs(refb).max(), scale)
    # transposed destination (the orientation engine.wgrad_fp32 uses when the input side is the wide one), no db
    dWt = dev(np.ascontiguousarray(dW0.T))
    U.LIB.call("u2gnn_gemm_split_wgrad", Ad.data_ptr(), M, N1, N1, Bd.data_ptr(), N2, N2, dWt.data_ptr(), 1, N1, 0, E._stream())
    assert np.abs(dWt.cpu().numpy().T - ref).max() < TOL * max(np.abs(ref).max(), scale)


def test_engine_wgrad_fp32_both_orientations(U):
    from u2gnn_b200 import engine as E
    rng = np.random.default_rng(9)
    M, d, ff = 700, 64, 512
    df = rng.standard_normal((M, d)).astype(np.float32)
    hd = rng.standard_normal((M, ff)).astype(np.float32)
    dW2 = torch.zeros((d, ff), device="cuda")
    db2 = torch.zeros(d, device="cuda")
    E.wgrad_fp32(dev(df), M, d, dev(hd), ff, dW2, db2)               # n_in > n_out: transposed destination + colsum
    assert rel_err(dW2.cpu().numpy(), df.astype(np.float64).T @ hd) < TOL
    assert rel_err(db2.cpu().numpy(), df.astype(np.float64).sum(0)) < TOL
    dW1 = torch.zeros((ff, d), device="cuda")
    db1 = torch.zeros(ff, device="cuda")
    E.wgrad_fp32(dev(hd), M, ff, dev(df), d, dW1, db1)
    assert rel_err(dW1.cpu().numpy(), hd.astype(np.float64).T @ df) < TOL
    assert rel_err(db1.cpu().numpy(), hd.astype(np.float64).sum(0)) < TOL


@pytest.mark.parametrize("case", ["sup_neighbors_d64", "sup_nodes_small"])
def test_fixtures_still_pass_on_the_cuda_core_sgemm(U, case, monkeypatch):
    """engine.FP32_TC = False keeps the CUDA-core SGEMM reachable (it still carries the [S, S] score products of attn_axis="nodes")."""
    from u2gnn_b200 import engine as E
    from test_gpu_parity import build_sup
    monkeypatch.setattr(E, "FP32_TC", False)
    c = load_golden(case)
    m, gp, *_ = build_sup(U, c)
    m.eval()
    with torch.no_grad():
        s = m(dev(c["input_x"]), gp, dev(c["X"]))
    assert rel_err(s.cpu().numpy(), c["eval_scores"]) < 1e-4


# ------------------------------------------------------------------ plain bf16 mode of the K-looping rows kernel (wide bf16 FFN, configs[2])
def call_kloop(U, A, W, w_kn, N, K=None, bias=None, epi=0, seed=0, stream=0, thr=0, aux=None, aux_scale=1.0, out_bf16=False, packed=None):
    from u2gnn_b200 import engine as E
    M = A.shape[0]
    K = K or A.shape[1]
    C = torch.zeros((M, N), dtype=torch.bfloat16 if out_bf16 else torch.float32, device="cuda")
    U.LIB.call("u2gnn_gemm_tc_rows_kloop", A.data_ptr(), M, K, A.stride(0), W.data_ptr(), w_kn, W.stride(0), N, E._ptr(bias), epi, seed, stream,
               thr, 0, E._ptr(aux), int(aux is not None), 0 if aux is None else aux.stride(0), aux_scale, 0.0, C.data_ptr(), int(out_bf16),
               C.stride(0), E._ptr(packed), E._stream())
    torch.cuda.synchronize()
    return C


@pytest.mark.parametrize("M,K,N,w_kn,packed", [(1000, 72, 1024, 0, True), (300, 1024, 72, 1, True), (129, 64, 320, 0, False), (4097, 1024, 128, 1, False),
                                                (257, 104, 64, 0, True), (640, 2048, 80, 1, True)])
def test_kloop_plain_bf16_rows_match_fp64_on_rounded_operands(U, M, K, N, w_kn, packed):
    """u2gnn_gemm_tc_rows_kloop: A bf16 rows (a 128-column padded buffer, K = the columns actually used), one bf16 product per k-step,
    fp32 accumulation: against fp64 arithmetic on the SAME bf16-rounded operands (the weights are rounded inside the kernel)."""
    from u2gnn_b200 import engine as E
    rng = np.random.default_rng(M + K + N)
    lda = max(128, (K + 7) // 8 * 8)
    Ab = torch.zeros((M, lda), dtype=torch.bfloat16, device="cuda")
    Ab[:, :K] = dev(rng.standard_normal((M, K)).astype(np.float32)).bfloat16()
    W = dev((rng.standard_normal((K, N) if w_kn else (N, K)) / np.sqrt(K)).astype(np.float32))
    b = dev(rng.standard_normal(N).astype(np.float32))
    A64 = Ab[:, :K].float().cpu().numpy().astype(np.float64)
    W64 = W.bfloat16().float().cpu().numpy().astype(np.float64)
    ref = A64 @ (W64 if w_kn else W64.T) + b.cpu().numpy()
    pk = None
    if packed:
        nb = U.LIB.call("u2gnn_gemm_split_packed_bytes", N, K)
        pk = torch.empty(nb, dtype=torch.uint8, device="cuda")
        U.LIB.call("u2gnn_gemm_split_pack", W.data_ptr(), w_kn, W.stride(0), N, K, pk.data_ptr(), nb, E._stream())
    C = call_kloop(U, Ab, W, w_kn, N, K=K, bias=b, epi=1, packed=pk)
    assert rel_err(C.cpu().numpy(), ref) < 1e-5                       # fp32 accumulation of exact bf16 products
    Cb = call_kloop(U, Ab, W, w_kn, N, K=K, bias=b, epi=1, out_bf16=True, packed=pk)
    assert torch.equal(Cb, C.bfloat16())                              # the bf16 result is the fp32 one rounded once


def test_kloop_plain_relu_dropout_and_mask_epilogues(U):
    """linear1 + ReLU + dropout -> bf16 hidden, then dH masked by that hidden (engine.ffn_wide_fwd / _bwd): keep bits from the oracle's
    restatement of the dropout stream; the mask epilogue tests the bf16 hidden's bits (live and kept <=> non-zero)."""
    rng = np.random.default_rng(21)
    M, K, N = 700, 72, 1024
    Ab = torch.zeros((M, 128), dtype=torch.bfloat16, device="cuda")
    Ab[:, :K] = dev(rng.standard_normal((M, K)).astype(np.float32)).bfloat16()
    W = dev((rng.standard_normal((N, K)) / 8).astype(np.float32))
    b = dev(rng.standard_normal(N).astype(np.float32))
    seed, stream, p = 0xABCDEF123, 9, 0.5
    keep, scale = O.dropout_keep_mask(seed, stream, M * N, p)
    pre = Ab[:, :K].float().cpu().numpy().astype(np.float64) @ W.bfloat16().float().cpu().numpy().astype(np.float64).T + b.cpu().numpy()
    ref = np.maximum(pre, 0.0) * keep.reshape(M, N) * scale
    h = call_kloop(U, Ab, W, 0, N, K=K, bias=b, epi=1 | 2 | 4, seed=seed, stream=stream, thr=128, out_bf16=True)
    hf = h.float().cpu().numpy()
    assert rel_err(hf, ref) < 5e-3                                    # one bf16 rounding of the result
    assert np.mean((hf == 0.0) != (ref == 0.0)) < 1e-3
    df = torch.zeros((M, 128), dtype=torch.bfloat16, device="cuda")
    df[:, :K] = dev(rng.standard_normal((M, K)).astype(np.float32)).bfloat16()
    W2 = dev((rng.standard_normal((N, 128)) / 8).astype(np.float32))           # [N = hidden][K = feature] with a 128-column stride
    dh = call_kloop(U, df, W2, 0, N, K=K, epi=8, aux=h, aux_scale=2.0, out_bf16=True)
    refd = (df[:, :K].float().cpu().numpy().astype(np.float64) @ W2[:, :K].bfloat16().float().cpu().numpy().astype(np.float64).T) * (hf > 0) * 2.0
    assert rel_err(dh.float().cpu().numpy(), refd) < 5e-3
    assert not np.any(dh.float().cpu().numpy()[hf == 0.0])             # dropped / dead hidden units carry no gradient
