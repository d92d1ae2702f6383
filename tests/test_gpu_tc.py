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
    from u2gnn_b200._lib import probe_lib
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
    probe_lib().call("u2gnn_tc_selftest", mode, A.data_ptr(), B.data_ptr(), C.data_ptr(), K, N, scratch.data_ptr(), E._stream())
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
               t["beta"].data_ptr(), z.data_ptr(), stats.data_ptr(), xn.data_ptr(), 0, E._stream())
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
    # device arithmetic: S accumulated in fp32, rounded to bf16, bias added in bf16 (one rounding), ReLU, keep mask
    hb = np.maximum(_bf(_bf(_bf(y1).astype(np.float64) @ _bf(W1).T.astype(np.float64)) + _bf(b1)), 0) * (m2 > 0)
    z_emu = y64 + (hb.astype(np.float64) @ _bf(W2 * scale).T.astype(np.float64) + b2) * m3
    mean = z_ref.mean(1)
    rstd = 1.0 / np.sqrt(z_ref.var(1) + 1e-5)
    xn_ref = (z_ref - mean[:, None]) * rstd[:, None] * gamma + beta
    return z.cpu().numpy(), stats.cpu().numpy(), xn.cpu().numpy(), z_ref, z_emu, mean, rstd, xn_ref


@pytest.mark.parametrize("M,d,ff,p", [(256, 64, 256, 0.0), (1000, 64, 2048, 0.0), (37, 64, 128, 0.0), (257, 64, 1024, 0.5),
                                      (300, 7, 256, 0.5), (129, 12, 128, 0.0), (300, 64, 2048, 0.25),
                                      (148 * 512 + 5, 64, 256, 0.5), (148 * 256 * 4 + 300, 64, 128, 0.5),
                                      (148 * 256 * 3 + 129, 12, 128, 0.5)])
def test_ffn_tc_forward(U, M, d, ff, p):
    z, stats, xn, z_ref, z_emu, mean, rstd, xn_ref = _ffn_case(U, M, d, ff, p)
    scale = np.abs(z_ref).max()
    assert np.isfinite(z).all()
    assert np.abs(z - z_emu).max() / scale < 2e-3          # same bf16 operands, fp32 accumulation
    assert np.abs(z - z_ref).max() / scale < 2e-2          # tolerance of the bf16-FFN mode (north_star)
    assert np.abs(stats[:, 0] - mean).max() < 2e-2 * max(1.0, np.abs(mean).max())
    assert np.abs(stats[:, 1] / rstd - 1).max() < 2e-2
    assert np.abs(xn - xn_ref).max() / np.abs(xn_ref).max() < 2e-2


# ------------------------------------------------------------------ fused FFN backward
@pytest.mark.parametrize("M,d,ff,p", [(256, 64, 128, 0.0), (1000, 64, 2048, 0.5), (37, 64, 256, 0.0), (300, 7, 256, 0.5),
                                      (5000, 64, 1024, 0.25), (148 * 256 * 2 + 77, 64, 256, 0.5)])
def test_ffn_tc_backward(U, M, d, ff, p):
    from u2gnn_b200 import engine as E
    from oracle import u2gnn_oracle as O
    rng = np.random.default_rng(5 + M + d + ff)
    y1 = rng.standard_normal((M, d)).astype(np.float32)
    df = rng.standard_normal((M, d)).astype(np.float32)
    dz = rng.standard_normal((M, d)).astype(np.float32)
    W1 = (rng.standard_normal((ff, d)) / np.sqrt(d)).astype(np.float32)
    b1 = (0.1 * rng.standard_normal(ff)).astype(np.float32)
    W2 = (rng.standard_normal((d, ff)) / np.sqrt(ff)).astype(np.float32)
    b2 = np.zeros(d, dtype=np.float32)
    thr = E.dropout_threshold(p)
    scale = 256.0 / (256.0 - thr) if thr else 1.0
    dev = lambda a: torch.from_numpy(a).cuda()
    t = {k: dev(v) for k, v in dict(y1=y1, df=df, dz=dz, W1=W1, b1=b1, W2=W2, b2=b2).items()}
    nbytes = U.LIB.call("u2gnn_ffn_tc_packed_bytes", d, ff)
    packed = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    U.LIB.call("u2gnn_ffn_tc_prepare", t["W1"].data_ptr(), t["b1"].data_ptr(), t["W2"].data_ptr(), t["b2"].data_ptr(), d, ff,
               scale, packed.data_ptr(), nbytes, E._stream())
    dy1 = torch.zeros((M, d), device="cuda")
    dW1 = torch.zeros((ff, d), device="cuda"); db1 = torch.zeros(ff, device="cuda"); dW2 = torch.zeros((d, ff), device="cuda")
    SEED, S2 = 0x1234ABCD99, 18
    ws = torch.empty(U.LIB.call("u2gnn_ffn_tc_bwd_workspace_bytes", M), dtype=torch.uint8, device="cuda")
    U.LIB.call("u2gnn_ffn_tc_bwd", t["y1"].data_ptr(), t["df"].data_ptr(), 0, 0, 0, t["dz"].data_ptr(), M, d, ff, packed.data_ptr(), scale,
               SEED, S2, thr, dy1.data_ptr(), dW1.data_ptr(), db1.data_ptr(), dW2.data_ptr(), ws.data_ptr(), ws.numel(), E._stream())
    torch.cuda.synchronize()
    if thr:
        k2, _ = O.dropout_keep_mask(SEED, S2, M * ff, p)
        m2 = k2.reshape(M, ff).astype(np.float64) * scale
    else:
        m2 = np.ones((M, ff))
    y64, df64 = y1.astype(np.float64), df.astype(np.float64)
    hpre = y64 @ W1.T.astype(np.float64) + b1
    hd = np.maximum(hpre, 0) * m2
    dhd = df64 @ W2.astype(np.float64)
    dhpre = dhd * m2 * (hpre > 0)
    ref = dict(dy1=dz + dhpre @ W1.astype(np.float64), dW1=dhpre.T @ y64, db1=dhpre.sum(0), dW2=df64.T @ hd)
    # emulation of the tensor-core arithmetic: bf16 operands, wide accumulation, bf16 hidden / dPre.  The ReLU
    # mask is decided by the bf16-operand pre-activation, exactly as on the device.
    yb, dfb, W1b, W2sb = (_bf(a).astype(np.float64) for a in (y1, df, W1, W2 * scale))
    hpre_b = _bf(_bf(yb @ W1b.T) + _bf(b1)).astype(np.float64)      # bf16(bf16(S) + bf16(b1)), as in csrc/ffn_epi.cuh
    keep = (m2 > 0)
    Hb = _bf(np.maximum(hpre_b, 0) * keep).astype(np.float64)
    Pb = _bf((dfb @ W2sb) * keep * (hpre_b > 0)).astype(np.float64)
    emu = dict(dy1=dz + Pb @ W1b, dW1=Pb.T @ yb, db1=Pb.sum(0), dW2=(dfb.T @ Hb) * scale)
    got = dict(dy1=dy1, dW1=dW1, db1=db1, dW2=dW2)
    for name in ref:
        g = got[name].cpu().numpy().astype(np.float64)
        assert np.isfinite(g).all(), name
        e_emu = np.abs(g - emu[name]).max() / np.abs(emu[name]).max()
        assert e_emu < 3e-3, (name, "vs bf16 emulation", e_emu)
        # against exact fp32 semantics, norm-wise: the ReLU derivative is discontinuous, so the ~0.2% of hidden
        # units whose pre-activation changes sign under bf16 operand rounding each move a whole term of the
        # gradient; 5e-2 bounds that effect (the 2e-2 of BASELINE.json is stated for embeddings and loss, which
        # tests/test_gpu_tc.py::test_bf16_train_step_vs_fp32_path checks end to end).
        e_ref = np.linalg.norm(g - ref[name]) / np.linalg.norm(ref[name])
        assert e_ref < 5e-2, (name, "vs oracle", e_ref)


# ------------------------------------------------------------------ bf16-FFN mode end to end
def _golden_model(U, case, precision):
    from conftest import load_golden, split_case
    c = load_golden(case)
    params, grads, _ = split_case(c)
    k, d, ff, T, L, C = [int(v) for v in c["meta"]]
    m = U.TransformerU2GNN(d, ff, C, T, 0.5, L, attn_axis="neighbors", precision=precision).cuda()
    m.load_state_dict({n: torch.from_numpy(v) for n, v in params.items()})
    return c, m, grads, (k, d, ff, T, L, C)


def test_bf16_eval_scores_and_loss_within_2e2_of_reference(U):
    """north_star: embeddings and loss within 2e-2 relative in bf16-FFN mode (vs the reference fixtures)."""
    c, m, grads, (k, d, ff, T, L, C) = _golden_model(U, "sup_neighbors_d64", "bf16")
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    m.eval()
    with torch.no_grad():
        s = m(dev(c["input_x"]), dev(c["rowptr"]), dev(c["X"]))
    err = np.abs(s.cpu().numpy() - c["eval_scores"]).max() / np.abs(c["eval_scores"]).max()
    assert err < 2e-2, err
    m.train(); m.encoder_dropout = 0.0
    for dr in m.dropouts:
        dr.p = 0.0
    s = m(dev(c["input_x"]), dev(c["rowptr"]), dev(c["X"]))
    soft = U.label_smoothing(dev(c["labels"]), C)
    loss = torch.mean(torch.sum(-soft * torch.nn.functional.log_softmax(s, dim=1), 1))
    assert abs(loss.item() - float(c["loss"])) < 2e-2 * abs(float(c["loss"]))
    loss.backward()
    for n, p in m.named_parameters():
        ref = grads[n]
        if np.linalg.norm(ref) < 1e-3 * max(np.linalg.norm(v) for v in grads.values()):
            continue                                        # numerically-zero gradients (e.g. key bias)
        e = np.linalg.norm(p.grad.cpu().numpy() - ref) / np.linalg.norm(ref)
        assert e < 5e-2, (n, e)


def test_bf16_train_step_vs_fp32_path(U):
    """Same batch, same dropout stream: the tcgen05 path tracks the fp32 CUDA path through a full fused train step
    (forward + loss + backward + clip + Adam) on a cfg5-shaped batch."""
    from u2gnn_b200.synthetic import make_batch
    from u2gnn_b200.trainer import SupTrainer
    b = make_batch(3000, 16, 64, 2, seed=5)
    out = {}
    for prec in ("fp32", "bf16"):
        torch.manual_seed(7)
        m = U.TransformerU2GNN(64, 512, 2, 2, 0.5, 1, attn_axis="neighbors").cuda()
        tr = SupTrainer(m, lr=5e-4, precision=prec, seed=99)
        loss, scores = tr.forward_backward(b["input_x"], b["rowptr"], b["X"], b["labels"], train=True)
        out[prec] = (loss.item(), scores.clone(), tr.arena.g.clone())
    l32, s32, g32 = out["fp32"]
    l16, s16, g16 = out["bf16"]
    assert abs(l16 - l32) < 2e-2 * abs(l32)
    assert (s16 - s32).abs().max().item() < 2e-2 * s32.abs().max().item()
    assert ((g16 - g32).norm() / g32.norm()).item() < 5e-2


# ------------------------------------------------------------------ bf16 projection GEMMs
@pytest.mark.parametrize("M,K,N,w_kn,bias,beta", [(300, 64, 192, 0, True, 0.0), (1000, 64, 64, 1, False, 0.0), (517, 192, 64, 1, False, 1.0),
                                                  (129, 7, 21, 0, True, 0.0), (4096 * 5 + 3, 64, 192, 0, True, 0.0)])
def test_gemm_tc_rows(U, M, K, N, w_kn, bias, beta):
    from u2gnn_b200 import engine as E
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn((K, N) if w_kn else (N, K), device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    C = torch.randn(M, N, device="cuda", generator=g)
    ref = bf16_round(A) @ (bf16_round(W) if w_kn else bf16_round(W).t())
    if bias:
        ref = ref + b
    if beta:
        ref = ref + C
    U.LIB.call("u2gnn_gemm_tc_rows", A.data_ptr(), M, K, K, W.data_ptr(), w_kn, N, 0 if b is None else b.data_ptr(), beta,
               C.data_ptr(), N, E._stream())
    torch.cuda.synchronize()
    assert ((C - ref).abs().max() / ref.abs().max()).item() < 1e-4


@pytest.mark.parametrize("M,N1,N2", [(300, 64, 64), (1000, 192, 64), (77, 21, 7), (148 * 128 * 3 + 50, 192, 64)])
def test_gemm_tc_wgrad(U, M, N1, N2):
    from u2gnn_b200 import engine as E
    g = torch.Generator(device="cuda").manual_seed(M + N1 + N2)
    A = torch.randn(M, N1, device="cuda", generator=g)
    B = torch.randn(M, N2, device="cuda", generator=g)
    dW = torch.randn(N1, N2, device="cuda", generator=g)
    db = torch.randn(N1, device="cuda", generator=g)
    ref_w = dW.double() + bf16_round(A).double().t() @ bf16_round(B).double()
    ref_b = db.double() + bf16_round(A).double().sum(0)
    U.LIB.call("u2gnn_gemm_tc_wgrad", A.data_ptr(), M, N1, N1, B.data_ptr(), N2, N2, dW.data_ptr(), db.data_ptr(), E._stream())
    torch.cuda.synchronize()
    assert ((dW.double() - ref_w).abs().max() / ref_w.abs().max()).item() < 1e-4
    assert ((db.double() - ref_b).abs().max() / ref_b.abs().max()).item() < 1e-4


# ------------------------------------------------------------------ tensor-core attention core vs the fp32 kernels
@pytest.mark.parametrize("B,S,p", [(7, 17, 0.0), (50, 17, 0.5), (23, 5, 0.5), (9, 32, 0.0), (148 * 7 * 2 + 3, 17, 0.5)])
def test_seqattn_tc_matches_fp32_kernels(U, B, S, p):
    from u2gnn_b200 import engine as E
    d = 64
    g = torch.Generator(device="cuda").manual_seed(B + S)
    qkv = torch.randn(B * S, 3 * d, device="cuda", generator=g)
    dctx = torch.randn(B * S, d, device="cuda", generator=g)
    thr = E.dropout_threshold(p)
    ctx32 = torch.empty(B * S, d, device="cuda"); ctx16 = torch.empty_like(ctx32)
    dq32 = torch.empty(B * S, 3 * d, device="cuda"); dq16 = torch.empty_like(dq32)
    U.LIB.call("u2gnn_seqattn_fwd", qkv.data_ptr(), B, S, S, d, 77, 5, thr, ctx32.data_ptr(), E._stream())
    U.LIB.call("u2gnn_seqattn_tc_fwd", qkv.data_ptr(), B, S, d, 77, 5, thr, ctx16.data_ptr(), E._stream())
    U.LIB.call("u2gnn_seqattn_bwd", qkv.data_ptr(), dctx.data_ptr(), B, S, S, d, 77, 5, thr, dq32.data_ptr(), E._stream())
    U.LIB.call("u2gnn_seqattn_tc_bwd", qkv.data_ptr(), dctx.data_ptr(), B, S, d, 77, 5, thr, dq16.data_ptr(), E._stream())
    torch.cuda.synchronize()
    assert torch.isfinite(ctx16).all() and torch.isfinite(dq16).all()
    assert ((ctx16 - ctx32).norm() / ctx32.norm()).item() < 2e-2
    for k, name in enumerate(("dq", "dk", "dv")):
        a, b = dq16[:, k * d:(k + 1) * d], dq32[:, k * d:(k + 1) * d]
        assert ((a - b).norm() / b.norm()).item() < 3e-2, name


# ------------------------------------------------------------------ bf16 activation I/O variants (bit-identical to fp32 I/O)
@pytest.mark.parametrize("M,K,N", [(1000, 64, 192), (517, 192, 64), (300, 64, 64)])
def test_gemm_tc_rows_bf16_io_matches_fp32_io(U, M, K, N):
    """Storing the projection output as bf16 (and reading a bf16 input) must equal rounding the fp32-I/O result: every
    consumer of these tensors rounds them to bf16 anyway."""
    from u2gnn_b200 import engine as E
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / 8
    b = torch.randn(N, device="cuda", generator=g)
    ref = E.linear_tc(A, M, K, W, 0, N, bias=b)                                # fp32 in, fp32 out
    out_b = E.linear_tc(A, M, K, W, 0, N, bias=b, out_bf16=True)               # fp32 in, bf16 out
    assert out_b.dtype == torch.bfloat16
    assert torch.equal(out_b, ref.to(torch.bfloat16))
    A_b = A.to(torch.bfloat16)
    ref2 = E.linear_tc(A_b.float(), M, K, W, 0, N, bias=b)                     # the same rounded input through the fp32 path
    out2 = E.linear_tc(A_b, M, K, W, 0, N, bias=b)                             # bf16 in, fp32 out
    assert torch.equal(out2, ref2)


@pytest.mark.parametrize("M,N1,N2", [(1000, 192, 64), (700, 64, 64)])
def test_gemm_tc_wgrad_bf16_operands_match(U, M, N1, N2):
    from u2gnn_b200 import engine as E
    g = torch.Generator(device="cuda").manual_seed(M + N1)
    A = torch.randn(M, N1, device="cuda", generator=g).to(torch.bfloat16)
    B = torch.randn(M, N2, device="cuda", generator=g)
    dW0 = torch.zeros(N1, N2, device="cuda"); db0 = torch.zeros(N1, device="cuda")
    dW1 = torch.zeros(N1, N2, device="cuda"); db1 = torch.zeros(N1, device="cuda")
    E.wgrad_tc(A.float(), M, N1, B, N2, dW0, db0)
    E.wgrad_tc(A, M, N1, B, N2, dW1, db1)
    torch.cuda.synchronize()
    # same bf16 operands, fp32 accumulation; the flush uses atomics (order may differ between runs)
    assert (dW1 - dW0).abs().max().item() <= 1e-4 * dW0.abs().max().item()
    assert (db1 - db0).abs().max().item() <= 1e-4 * db0.abs().max().item()


@pytest.mark.parametrize("B,S,p", [(300, 17, 0.5), (77, 9, 0.0)])
def test_seqattn_tc_bf16_io_matches_fp32_io(U, B, S, p):
    from u2gnn_b200 import engine as E
    d = 64
    thr = E.dropout_threshold(p)
    g = torch.Generator(device="cuda").manual_seed(B + S)
    qkv = torch.randn(B * S, 3 * d, device="cuda", generator=g).to(torch.bfloat16)
    dctx = torch.randn(B * S, d, device="cuda", generator=g).to(torch.bfloat16)
    ctx0 = torch.empty(B * S, d, device="cuda"); ctx1 = torch.empty(B * S, d, device="cuda", dtype=torch.bfloat16)
    q32, g32 = qkv.float(), dctx.float()
    U.LIB.call("u2gnn_seqattn_tc_fwd", q32.data_ptr(), B, S, d, 5, 3, thr, ctx0.data_ptr(), E._stream())
    U.LIB.call("u2gnn_seqattn_tc_fwd_ex", qkv.data_ptr(), B, S, d, 5, 3, thr, ctx1.data_ptr(), 1, E._stream())
    dq0 = torch.empty(B * S, 3 * d, device="cuda"); dq1 = torch.empty(B * S, 3 * d, device="cuda", dtype=torch.bfloat16)
    U.LIB.call("u2gnn_seqattn_tc_bwd", q32.data_ptr(), g32.data_ptr(), B, S, d, 5, 3, thr, dq0.data_ptr(), E._stream())
    U.LIB.call("u2gnn_seqattn_tc_bwd_ex", qkv.data_ptr(), dctx.data_ptr(), B, S, d, 5, 3, thr, dq1.data_ptr(), 1, E._stream())
    torch.cuda.synchronize()
    assert torch.equal(ctx1, ctx0.to(torch.bfloat16))
    assert torch.equal(dq1, dq0.to(torch.bfloat16))




# ------------------------------------------------------------------ fused attention-block epilogues
@pytest.mark.parametrize("M,S,p,a_bf16", [(1000, 1, 0.5, 1), (4096 * 3 + 5, 1, 0.5, 1), (300, 17, 0.5, 1), (517, 1, 0.0, 0), (128, 1, 0.5, 0)])
def test_out_proj_ln_fused_equals_gemm_then_layernorm(U, M, S, p, a_bf16):
    """u2gnn_gemm_tc_rows_ln (out_proj + dropout + residual + LayerNorm1 in the GEMM epilogue) against the two kernels it
    replaces: same MMAs, same fp32 arithmetic in the same order -> z, y and stats bit-identical.  S > 1 reads the residual
    with a row stride (position 0 of each sequence, the last-timestep case)."""
    from u2gnn_b200 import engine as E
    d = 64
    thr = E.dropout_threshold(p)
    g = torch.Generator(device="cuda").manual_seed(M + S)
    ctx = torch.randn(M, d, device="cuda", generator=g)
    if a_bf16:
        ctx = ctx.to(torch.bfloat16)
    prm = {"self_attn.out_proj.weight": torch.randn(d, d, device="cuda", generator=g) / 8,
           "self_attn.out_proj.bias": torch.randn(d, device="cuda", generator=g),
           "norm1.weight": 1 + 0.1 * torch.randn(d, device="cuda", generator=g),
           "norm1.bias": 0.1 * torch.randn(d, device="cuda", generator=g)}
    x = torch.randn(M * S, d, device="cuda", generator=g)
    drop = (0x1234ABCD5, 21, thr)
    a = E.linear_tc(ctx, M, d, prm["self_attn.out_proj.weight"], 0, d, bias=prm["self_attn.out_proj.bias"])
    xq = x.view(M, S, d)[:, 0, :].contiguous()
    z0, y0, st0 = E.add_dropout_ln_fwd(xq, a, M, d, drop, prm["norm1.weight"], prm["norm1.bias"])
    z1, y1, st1, _ = E.out_proj_ln_tc(ctx, M, d, prm, x, S * d, drop)
    torch.cuda.synchronize()
    assert torch.equal(z1, z0)
    assert torch.equal(st1, st0)
    assert torch.equal(y1, y0)


@pytest.mark.parametrize("M,d,p", [(1000, 64, 0.5), (4096 * 2 + 7, 64, 0.5), (333, 32, 0.25), (200, 64, 0.0)])
def test_ln_bwd_bf16_da_and_folded_bias_gradient(U, M, d, p):
    """u2gnn_add_dropout_ln_bwd_ex: dz identical to the base entry point, da == bf16(da of the base entry point) bit for bit,
    dasum == column sums of the fp32 masked gradient (what u2gnn_colsum computed in a separate pass)."""
    from u2gnn_b200 import engine as E
    thr = E.dropout_threshold(p)
    g = torch.Generator(device="cuda").manual_seed(M + d)
    dy = torch.randn(M, d, device="cuda", generator=g)
    z = torch.randn(M, d, device="cuda", generator=g) * 2 + 0.3
    mean = z.mean(1)
    rstd = (z.var(1, unbiased=False) + 1e-5).rsqrt()
    stats = torch.stack([mean, rstd], 1).contiguous()
    gamma = 1 + 0.1 * torch.randn(d, device="cuda", generator=g)
    drop = (77, 5, thr)
    dg0 = torch.zeros(d, device="cuda"); db0 = torch.zeros(d, device="cuda")
    dg1 = torch.zeros(d, device="cuda"); db1 = torch.zeros(d, device="cuda"); ds1 = torch.zeros(d, device="cuda")
    dz0, da0 = E.add_dropout_ln_bwd(dy, z, stats, M, d, gamma, drop, dg0, db0)
    dz1, da1 = E.add_dropout_ln_bwd(dy, z, stats, M, d, gamma, drop, dg1, db1, da_bf16=True, dasum=ds1)
    torch.cuda.synchronize()
    assert torch.equal(dz1, dz0)
    if thr:
        assert da1.dtype == torch.bfloat16 and torch.equal(da1, da0.to(torch.bfloat16))
    else:
        assert da1 is dz1
    ref = da0.double().sum(0)
    assert (ds1.double() - ref).abs().max().item() <= 1e-5 * max(1.0, da0.abs().sum(0).max().item())
    assert (dg1 - dg0).abs().max().item() <= 1e-4 * max(1.0, dg0.abs().max().item())
    assert (db1 - db0).abs().max().item() <= 1e-4 * max(1.0, db0.abs().max().item())


def test_fused_attention_block_epilogues_equal_unfused_step(U):
    """Whole train-step gradients with the fused epilogues (out_proj+LN1, bf16 da, folded linear2 bias gradient) against the
    same step with the separate kernels: forward bit-identical, gradients equal up to the order of fp32 atomics."""
    from u2gnn_b200 import engine as E
    from u2gnn_b200.synthetic import make_batch
    from u2gnn_b200.trainer import SupTrainer
    b = make_batch(3000, 16, 64, 2, seed=9)
    out = {}
    defaults = (E.FUSE_OUT_PROJ_LN, E.FUSE_LN_BWD, E.FUSE_PROJ_BWD, E.FUSE_INPROJ_ATTN)
    try:
        for fused in (False, True):
            E.FUSE_OUT_PROJ_LN = E.FUSE_LN_BWD = E.FUSE_PROJ_BWD = E.FUSE_INPROJ_ATTN = fused
            torch.manual_seed(3)
            m = U.TransformerU2GNN(64, 512, 2, 3, 0.5, 1, attn_axis="neighbors").cuda()
            tr = SupTrainer(m, lr=5e-4, precision="bf16", seed=42)
            loss, scores = tr.forward_backward(b["input_x"], b["rowptr"], b["X"], b["labels"], train=True)
            out[fused] = (loss.item(), scores.clone(), tr.arena.g.clone())
    finally:
        E.FUSE_OUT_PROJ_LN, E.FUSE_LN_BWD, E.FUSE_PROJ_BWD, E.FUSE_INPROJ_ATTN = defaults
    assert out[True][0] == out[False][0]
    assert torch.equal(out[True][1], out[False][1])
    g1, g0 = out[True][2], out[False][2]
    assert ((g1 - g0).norm() / g0.norm()).item() < 1e-5


@pytest.mark.parametrize("B,S,p", [(300, 17, 0.5), (77, 9, 0.0), (4099, 17, 0.5)])
def test_seqattn_last_bf16_io_matches_fp32_io_on_rounded_inputs(U, B, S, p):
    """Last-timestep attention (query position 0 only) with bf16 qkv / dqkv: forward equals the fp32-I/O kernel fed the
    same rounded qkv bit for bit; dqkv equals the rounded fp32-I/O result."""
    from u2gnn_b200 import engine as E
    d = 64
    thr = E.dropout_threshold(p)
    g = torch.Generator(device="cuda").manual_seed(B * S)
    qkv_b = torch.randn(B * S, 3 * d, device="cuda", generator=g).to(torch.bfloat16)
    qkv_f = qkv_b.float()
    dctx = torch.randn(B, d, device="cuda", generator=g)
    SEED, ST = 99, 16
    ctx0 = torch.empty(B, d, device="cuda"); ctx1 = torch.empty(B, d, device="cuda")
    U.LIB.call("u2gnn_seqattn_last_fwd_ex", qkv_f.data_ptr(), 0, B, S, d, SEED, ST, thr, ctx0.data_ptr(), E._stream())
    U.LIB.call("u2gnn_seqattn_last_fwd_ex", qkv_b.data_ptr(), 1, B, S, d, SEED, ST, thr, ctx1.data_ptr(), E._stream())
    ref = torch.empty(B, d, device="cuda")
    U.LIB.call("u2gnn_seqattn_fwd", qkv_f.data_ptr(), B, S, 1, d, SEED, ST, thr, ref.data_ptr(), E._stream())
    dq0 = torch.empty(B * S, 3 * d, device="cuda")
    dq1 = torch.empty(B * S, 3 * d, device="cuda", dtype=torch.bfloat16)
    U.LIB.call("u2gnn_seqattn_last_bwd_ex", qkv_f.data_ptr(), dctx.data_ptr(), 0, B, S, d, SEED, ST, thr, dq0.data_ptr(), E._stream())
    U.LIB.call("u2gnn_seqattn_last_bwd_ex", qkv_b.data_ptr(), dctx.data_ptr(), 1, B, S, d, SEED, ST, thr, dq1.data_ptr(), E._stream())
    torch.cuda.synchronize()
    assert torch.equal(ctx0, ref)
    assert torch.equal(ctx1, ctx0)
    assert torch.equal(dq1, dq0.to(torch.bfloat16))


@pytest.mark.parametrize("M,N1,b_bf16,c_bf16,beta", [(1000, 192, 0, 0, 1.0), (148 * 128 * 3 + 77, 192, 0, 0, 1.0), (5000, 64, 1, 1, 0.0),
                                                     (333, 64, 0, 0, 0.0), (128 * 148 * 5 + 1, 64, 1, 1, 0.0), (100, 128, 0, 0, 0.0)])
def test_gemm_tc_dgrad_wgrad_equals_separate_kernels(U, M, N1, b_bf16, c_bf16, beta):
    """One pass over the output gradient (input gradient + weight / bias gradient) against the rows GEMM and the
    weight-gradient GEMM on the same operands: the input gradient is bit-identical (same MMAs, same epilogue arithmetic),
    the weight / bias gradients agree up to the order of the fp32 atomic flush."""
    from u2gnn_b200 import engine as E
    g = torch.Generator(device="cuda").manual_seed(M + N1)
    A = torch.randn(M, N1, device="cuda", generator=g).to(torch.bfloat16)
    Bm = torch.randn(M, 64, device="cuda", generator=g)
    if b_bf16:
        Bm = Bm.to(torch.bfloat16)
    W = torch.randn(N1, 64, device="cuda", generator=g) / 8
    C0 = torch.randn(M, 64, device="cuda", generator=g)
    dW0 = torch.zeros(N1, 64, device="cuda"); db0 = torch.zeros(N1, device="cuda")
    dW1 = torch.zeros(N1, 64, device="cuda"); db1 = torch.zeros(N1, device="cuda")
    E.wgrad_tc(A, M, N1, Bm, 64, dW0, db0)
    ref = E.linear_tc(A, M, N1, W, 1, 64, beta=beta, out=C0.clone() if beta else None, out_bf16=bool(c_bf16))
    out = E.proj_bwd_tc(A, M, N1, Bm, W, dW1, db1, out=C0.clone() if beta else None, out_bf16=bool(c_bf16), beta=beta)
    torch.cuda.synchronize()
    assert out.dtype == ref.dtype and torch.equal(out, ref)
    assert (dW1 - dW0).abs().max().item() <= 1e-4 * dW0.abs().max().item()
    assert (db1 - db0).abs().max().item() <= 1e-4 * db0.abs().max().item()


@pytest.mark.parametrize("B,S,p", [(300, 17, 0.5), (77, 9, 0.0), (148 * 7 * 3 + 5, 17, 0.5), (50, 32, 0.5)])
def test_inproj_attention_fused_equals_projection_then_attention(U, B, S, p):
    """u2gnn_inproj_seqattn_tc_fwd (qkv computed per tile inside the attention-forward kernel) against the projection GEMM
    followed by the attention kernel: qkv and ctx bit-identical."""
    from u2gnn_b200 import engine as E
    d = 64
    thr = E.dropout_threshold(p)
    g = torch.Generator(device="cuda").manual_seed(B * S + 1)
    x = torch.randn(B * S, d, device="cuda", generator=g)
    W = torch.randn(3 * d, d, device="cuda", generator=g) / 8
    b = torch.randn(3 * d, device="cuda", generator=g)
    SEED, ST = 4242, 16
    qkv0 = E.linear_tc(x, B * S, d, W, 0, 3 * d, bias=b, out_bf16=True)
    ctx0 = torch.empty(B * S, d, device="cuda", dtype=torch.bfloat16)
    U.LIB.call("u2gnn_seqattn_tc_fwd_ex", qkv0.data_ptr(), B, S, d, SEED, ST, thr, ctx0.data_ptr(), 1, E._stream())
    qkv1 = torch.full((B * S, 3 * d), float("nan"), device="cuda", dtype=torch.bfloat16)
    ctx1 = torch.full((B * S, d), float("nan"), device="cuda", dtype=torch.bfloat16)
    U.LIB.call("u2gnn_inproj_seqattn_tc_fwd", x.data_ptr(), 0, 0, B, S, d, W.data_ptr(), b.data_ptr(), SEED, ST, thr, qkv1.data_ptr(),
               ctx1.data_ptr(), 0, E._stream())
    torch.cuda.synchronize()
    assert torch.equal(qkv1, qkv0)
    assert torch.equal(ctx1, ctx0)




# ------------------------------------------------------------------ round 2: operand tile images written by the producers
def _image_of(rows_bf16):
    """Expected bf16 swizzled [128 x 64] tile images of an [M, 64] bf16 tensor (rows past M zero), as bytes."""
    M = rows_bf16.shape[0]
    nt = 2 * ((M + 255) // 256)
    pad = torch.zeros((nt * 128, 64), dtype=torch.bfloat16, device=rows_bf16.device)
    pad[:M] = rows_bf16
    chunks = pad.view(nt * 128, 8, 8)                                   # [row, 16-byte chunk, 8 bf16]
    r = torch.arange(nt * 128, device=pad.device)
    perm = (torch.arange(8, device=pad.device)[None, :] ^ (r[:, None] & 7))   # stored position p holds chunk p ^ (row & 7)
    return torch.gather(chunks, 1, perm[:, :, None].expand(-1, -1, 8)).contiguous().view(torch.uint8).reshape(-1)


@pytest.mark.parametrize("M,p", [(1000, 0.5), (128 * 7, 0.0), (148 * 128 * 3 + 77, 0.5), (5, 0.5)])
def test_producers_write_ffn_backward_tile_images(U, M, p):
    """u2gnn_add_dropout_ln_bwd_ex(da_bf16 = 2) and u2gnn_gemm_tc_rows_ln(y_img) write exactly the images the conversion pass of
    u2gnn_ffn_tc_bwd would build from their row-major outputs; the FFN backward fed with them gives bit-identical dy1."""
    from u2gnn_b200 import engine as E
    d, ff = 64, 256
    thr = E.dropout_threshold(p)
    g = torch.Generator(device="cuda").manual_seed(M)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
    dy, z = rnd(M, d), rnd(M, d) * 1.3 + 0.1
    st = torch.stack([z.mean(1), (z.var(1, unbiased=False) + 1e-5).rsqrt()], 1).contiguous()
    gamma = 1 + 0.1 * rnd(d)
    drop = (0xBEEF, 21, thr)
    zg = lambda: torch.zeros(d, device="cuda")
    dz0, da0 = E.add_dropout_ln_bwd(dy, z, st, M, d, gamma, drop, zg(), zg(), da_bf16=True, dasum=zg())
    dz1, img = E.add_dropout_ln_bwd(dy, z, st, M, d, gamma, drop, zg(), zg(), dasum=zg(), da_img=True)
    torch.cuda.synchronize()
    assert torch.equal(dz0, dz1)
    ref_rows = da0 if da0.dtype == torch.bfloat16 else dz0.to(torch.bfloat16)      # p = 0: no separate da
    n_real = ((M + 127) // 128) * 16384
    assert torch.equal(img[:n_real], _image_of(ref_rows)[:n_real])
    # out_proj + LN1 with the y image
    ctx = rnd(M, d).to(torch.bfloat16)
    prm = {"self_attn.out_proj.weight": rnd(d, d) / 8, "self_attn.out_proj.bias": 0.1 * rnd(d), "norm1.weight": gamma, "norm1.bias": 0.1 * rnd(d)}
    res = rnd(M, d)
    z1, y1, s1, yimg = E.out_proj_ln_tc(ctx, M, d, prm, res, d, (0xBEEF, 22, thr), want_img=True)
    torch.cuda.synchronize()
    assert torch.equal(yimg[:n_real], _image_of(y1.to(torch.bfloat16))[:n_real])
    # FFN backward: images from the producers vs the conversion pass
    prm2 = {"linear1.weight": rnd(ff, d) / 8, "linear1.bias": 0.1 * rnd(ff), "linear2.weight": rnd(d, ff) / 16, "linear2.bias": 0.1 * rnd(d)}
    packed = E.ffn_tc_pack(prm2, d, ff, thr)
    scale = 256.0 / (256.0 - thr) if thr else 1.0
    wsb = U.LIB.call("u2gnn_ffn_tc_bwd_workspace_bytes", M)
    df_rows = ref_rows.float().contiguous()
    outs = []
    for use_img in (False, True):
        dy1 = torch.empty(M, d, device="cuda")
        dW1, db1, dW2 = torch.zeros(ff, d, device="cuda"), torch.zeros(ff, device="cuda"), torch.zeros(d, ff, device="cuda")
        ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
        U.LIB.call("u2gnn_ffn_tc_bwd", y1.data_ptr(), df_rows.data_ptr(), yimg.data_ptr() if use_img else 0, img.data_ptr() if use_img else 0,
                   0, dz1.data_ptr(), M, d, ff, packed.data_ptr(), scale, 0xBEEF, 23, thr, dy1.data_ptr(), dW1.data_ptr(), db1.data_ptr(),
                   dW2.data_ptr(), ws.data_ptr(), wsb, E._stream())
        outs.append((dy1, dW1, db1, dW2))
    torch.cuda.synchronize()
    assert torch.equal(outs[0][0], outs[1][0])
    for a, b in zip(outs[0][1:], outs[1][1:]):
        assert (a - b).abs().max().item() <= 2e-4 * max(1.0, a.abs().max().item())


@pytest.mark.parametrize("M,ff,p", [(1000, 256, 0.5), (148 * 256 + 77, 512, 0.5), (300, 2048, 0.0), (129, 128, 0.25)])
def test_ffn_backward_with_forward_mask_equals_recomputed_mask(U, M, ff, p):
    """u2gnn_ffn_tc_fwd(mask_out) leaves one bit per hidden activation (ReLU live AND kept); u2gnn_ffn_tc_bwd(fwd_mask) must give the
    SAME results as the self-contained backward that recomputes the dropout stream and the sign of the hidden: dy1 bit-identical
    (same mask bits, same MMAs), weight gradients equal up to the order of the fp32 atomics."""
    from u2gnn_b200 import engine as E
    d = 64
    thr = E.dropout_threshold(p)
    scale = 256.0 / (256.0 - thr) if thr else 1.0
    g = torch.Generator(device="cuda").manual_seed(M + ff)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
    prm = {"linear1.weight": rnd(ff, d) / 8, "linear1.bias": 0.1 * rnd(ff), "linear2.weight": rnd(d, ff) / (ff ** 0.5), "linear2.bias": 0.1 * rnd(d)}
    packed = E.ffn_tc_pack(prm, d, ff, thr)
    y1, df, dz = rnd(M, d), rnd(M, d), rnd(M, d)
    gamma, beta = 1 + 0.1 * rnd(d), 0.1 * rnd(d)
    SEED, S_H, S_O = 0x77AA, 40, 41
    mask = torch.zeros(U.LIB.call("u2gnn_ffn_tc_mask_bytes", M, ff), dtype=torch.uint8, device="cuda")
    z = torch.empty(M, d, device="cuda"); st = torch.empty(M, 2, device="cuda"); xn = torch.empty(M, d, device="cuda")
    z0 = torch.empty(M, d, device="cuda"); st0 = torch.empty(M, 2, device="cuda"); xn0 = torch.empty(M, d, device="cuda")
    U.LIB.call("u2gnn_ffn_tc_fwd", y1.data_ptr(), M, d, ff, packed.data_ptr(), SEED, S_H, S_O, thr, gamma.data_ptr(), beta.data_ptr(),
               z.data_ptr(), st.data_ptr(), xn.data_ptr(), mask.data_ptr(), E._stream())
    U.LIB.call("u2gnn_ffn_tc_fwd", y1.data_ptr(), M, d, ff, packed.data_ptr(), SEED, S_H, S_O, thr, gamma.data_ptr(), beta.data_ptr(),
               z0.data_ptr(), st0.data_ptr(), xn0.data_ptr(), 0, E._stream())
    wsb = U.LIB.call("u2gnn_ffn_tc_bwd_workspace_bytes", M)
    outs = []
    for use_mask in (False, True):
        dy1 = torch.empty(M, d, device="cuda")
        dW1, db1, dW2 = torch.zeros(ff, d, device="cuda"), torch.zeros(ff, device="cuda"), torch.zeros(d, ff, device="cuda")
        ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
        U.LIB.call("u2gnn_ffn_tc_bwd", y1.data_ptr(), df.data_ptr(), 0, 0, mask.data_ptr() if use_mask else 0, dz.data_ptr(), M, d, ff,
                   packed.data_ptr(), scale, SEED, S_H, thr, dy1.data_ptr(), dW1.data_ptr(), db1.data_ptr(), dW2.data_ptr(), ws.data_ptr(),
                   wsb, E._stream())
        outs.append((dy1, dW1, db1, dW2, ws))
    torch.cuda.synchronize()
    assert torch.equal(z, z0) and torch.equal(xn, xn0)                   # emitting the mask does not change the forward
    # the self-contained backward writes its own mask words into the workspace: same bits for every real row
    tp = 2 * ((M + 255) // 256)
    own = outs[0][4][tp * 32768:tp * 32768 + mask.numel()].view(torch.int32).view(tp, ff // 128, 4, 128)
    fwd = mask.view(torch.int32).view(tp, ff // 128, 4, 128)
    rows = torch.arange(tp * 128, device="cuda").view(tp, 1, 1, 128).expand_as(fwd) < M
    assert torch.equal(own[rows], fwd[rows])
    assert torch.equal(outs[0][0], outs[1][0])
    for a, b in zip(outs[0][1:4], outs[1][1:4]):
        assert (a - b).abs().max().item() <= 2e-4 * max(1.0, a.abs().max().item())


@pytest.mark.parametrize("M,ff,p", [(37, 128, 0.5), (300, 256, 0.5), (129, 2048, 0.0)])
def test_ffn_kernels_write_nothing_outside_their_outputs(U, M, ff, p):
    """Bounds check of our own (compute-sanitizer is closed on this GPU pool, profiles/r02_sanitizer_unavailable.txt): every output
    of the fused FFN forward / backward sits between guard bands filled with a sentinel; the bulk stores, tensor-memory drains
    and atomics of the tcgen05 kernels must leave the bands untouched on shapes with partial tiles."""
    from u2gnn_b200 import engine as E
    d = 64
    thr = E.dropout_threshold(p)
    scale = 256.0 / (256.0 - thr) if thr else 1.0
    g = torch.Generator(device="cuda").manual_seed(M * 3 + ff)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
    prm = {"linear1.weight": rnd(ff, d) / 8, "linear1.bias": 0.1 * rnd(ff), "linear2.weight": rnd(d, ff) / (ff ** 0.5), "linear2.bias": 0.1 * rnd(d)}
    packed = E.ffn_tc_pack(prm, d, ff, thr)
    G = 4096                                                # guard elements on each side
    SENT = 12345.0

    class Guarded:
        def __init__(self, numel, dtype=torch.float32):
            self.buf = torch.full((numel + 2 * G,), SENT if dtype == torch.float32 else 0x5A, dtype=dtype, device="cuda")
            self.view = self.buf[G:G + numel]
            self.ref = self.buf.clone()

        def intact(self):
            return torch.equal(self.buf[:G], self.ref[:G]) and torch.equal(self.buf[-G:], self.ref[-G:])

    y1, df, dz, gamma, beta = rnd(M, d), rnd(M, d), rnd(M, d), 1 + 0.1 * rnd(d), 0.1 * rnd(d)
    out = {n: Guarded(k) for n, k in dict(z=M * d, st=M * 2, xn=M * d, dy1=M * d, dW1=ff * d, db1=ff, dW2=d * ff).items()}
    out["mask"] = Guarded(U.LIB.call("u2gnn_ffn_tc_mask_bytes", M, ff), torch.uint8)
    out["ws"] = Guarded(U.LIB.call("u2gnn_ffn_tc_bwd_workspace_bytes", M), torch.uint8)
    for n in ("dW1", "db1", "dW2"):
        out[n].view.zero_()
    ptr = lambda n: out[n].view.data_ptr()
    assert ptr("ws") % 128 == 0 and ptr("mask") % 128 == 0
    U.LIB.call("u2gnn_ffn_tc_fwd", y1.data_ptr(), M, d, ff, packed.data_ptr(), 7, 50, 51, thr, gamma.data_ptr(), beta.data_ptr(),
               ptr("z"), ptr("st"), ptr("xn"), ptr("mask"), E._stream())
    U.LIB.call("u2gnn_ffn_tc_bwd", y1.data_ptr(), df.data_ptr(), 0, 0, ptr("mask"), dz.data_ptr(), M, d, ff, packed.data_ptr(), scale, 7, 50, thr,
               ptr("dy1"), ptr("dW1"), ptr("db1"), ptr("dW2"), ptr("ws"), out["ws"].view.numel(), E._stream())
    torch.cuda.synchronize()
    for n, o in out.items():
        assert o.intact(), n
    for n in ("z", "xn", "dy1", "dW1", "dW2"):
        assert bool(torch.isfinite(out[n].view).all()) and not bool((out[n].view == SENT).any()), n


# ------------------------------------------------------------------ gather fused into its consumers
@pytest.mark.parametrize("nodes,S,T", [(300, 17, 3), (129, 9, 2), (1000, 17, 4)])
def test_fused_gather_layer_is_bit_identical_to_the_materialised_gather(U, nodes, S, T, monkeypatch):
    """u2gnn_layer_fwd / _bwd with the first timestep's kernels reading src[input_x] by index (in_proj + attention forward,
    out_proj + LayerNorm1 residual, in_proj backward) against the same layer on the gathered [N, S, d] tensor: the index-driven
    loads fetch the very same fp32 rows, so every output and every gradient must be bit-identical (weight gradients are
    accumulated with atomics in both: compared at 1e-5 of their largest entry)."""
    from u2gnn_b200 import engine as E
    d, ff = 64, 256
    g = torch.Generator(device="cuda").manual_seed(nodes)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
    src = rnd(nodes, d)
    input_x = torch.randint(0, nodes, (nodes, S), device="cuda", generator=g)
    input_x[:, 0] = torch.arange(nodes, device="cuda")
    params = [{"self_attn.in_proj_weight": rnd(3 * d, d) / 8, "self_attn.in_proj_bias": 0.1 * rnd(3 * d),
               "self_attn.out_proj.weight": rnd(d, d) / 8, "self_attn.out_proj.bias": 0.1 * rnd(d), "linear1.weight": rnd(ff, d) / 8,
               "linear1.bias": 0.1 * rnd(ff), "linear2.weight": rnd(d, ff) / 16, "linear2.bias": 0.1 * rnd(d),
               "norm1.weight": 1 + 0.1 * rnd(d), "norm1.bias": 0.1 * rnd(d), "norm2.weight": 1 + 0.1 * rnd(d), "norm2.bias": 0.1 * rnd(d)}
              for _ in range(T)]
    dout = rnd(nodes, d)
    drop = E.DropoutCfg(enabled=True, seed=77, p_enc=0.5, p_out=0.5)
    res = {}
    for fused in (True, False):
        monkeypatch.setattr(E, "FUSE_GATHER", fused)
        grads = [{n: torch.zeros_like(v) for n, v in p.items()} for p in params]
        out, saved = E.u2gnn_layer_fwd(src, input_x, params, 0, T, "neighbors", drop, "bf16")
        assert (saved.layers[0].x_idx is not None) == fused
        dsrc = E.u2gnn_layer_bwd(dout.clone(), saved, input_x, params, grads, 0, T, "neighbors", drop, need_dsrc=True,
                                 transpose=E.IndexTranspose(input_x, nodes))
        torch.cuda.synchronize()
        res[fused] = (out, dsrc, grads)
    assert torch.equal(res[True][0], res[False][0])
    assert torch.equal(res[True][1], res[False][1])
    for ga, gb in zip(res[True][2], res[False][2]):
        for n in ga:
            assert float((ga[n] - gb[n]).abs().max()) <= 1e-5 * max(float(gb[n].abs().max()), 1e-30), n


def test_fused_gather_out_of_range_index_gives_zero_row_and_sets_the_error_word(U):
    """An index outside the table - where F.embedding raises (pytorch_U2GNN_Sup.py:32) - is a zero input row and bit 1 of the
    device error word, exactly as in u2gnn_gather_rows."""
    from u2gnn_b200 import engine as E
    B, S, d = 40, 17, 64
    g = torch.Generator(device="cuda").manual_seed(4)
    table = torch.randn(100, d, device="cuda", generator=g)
    idx = torch.randint(0, 100, (B, S), device="cuda", generator=g)
    idx[3, 5] = 100
    idx[7, 0] = -1
    W = torch.randn(3 * d, d, device="cuda", generator=g) / 8
    b = torch.randn(3 * d, device="cuda", generator=g)
    qkv = torch.empty((B * S, 3 * d), dtype=torch.bfloat16, device="cuda")
    ctx = torch.empty((B * S, d), dtype=torch.bfloat16, device="cuda")
    err = E.err_word(table.device)
    err.zero_()
    U.LIB.call("u2gnn_inproj_seqattn_tc_fwd", table.data_ptr(), idx.data_ptr(), 100, B, S, d, W.data_ptr(), b.data_ptr(), 1, 2, 0,
               qkv.data_ptr(), ctx.data_ptr(), err.data_ptr(), E._stream())
    torch.cuda.synchronize()
    assert int(err.item()) & 2
    err.zero_()
    # a zero input row projects to the bias alone
    for r in (3 * S + 5, 7 * S):
        assert torch.equal(qkv[r], b.bfloat16())
    xg = E.gather_rows(table, idx.clamp(0, 99))
    ref = (xg.bfloat16().float() @ W.bfloat16().float().T + b)
    ok = torch.ones(B * S, dtype=torch.bool, device="cuda")
    ok[3 * S + 5] = False
    ok[7 * S] = False
    assert float((qkv.float()[ok] - ref[ok]).abs().max()) < 0.05
