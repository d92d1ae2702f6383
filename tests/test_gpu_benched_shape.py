"""Parity AT the benchmarked configuration: the bf16 tcgen05 path compared DIRECTLY with the fp64 numpy
oracle (oracle/u2gnn_oracle.py, pinned to reference fixtures by tests/test_oracle_golden.py) - not with sibling CUDA kernels.

  * whole train step, cfg5 shape (d 64, S 17, T 4, ff 2048, dropout ON, 4 096 nodes): scores / loss within 2e-2 (north_star
    bf16-FFN tolerance), every parameter gradient norm-wise within 5e-2 (ReLU derivative discontinuity, tests/test_gpu_tc.py)
  * one encoder layer, both timestep kinds (every row live / position 0 only): every saved intermediate of the forward
    (qkv, ctx, z1, y1, z2, y2) and every gradient (dx + the 12 parameter gradients) against oracle.encoder_layer_fwd/bwd, so each
    tcgen05 kernel (in_proj+attention, out_proj+LN1, FFN fwd, FFN wgrad/dgrad, LN backward, one-pass projection backward,
    attention backward, last-timestep attention) has an oracle comparison of its own output.
Reference arithmetic: pytorch_U2GNN_Sup.py:30-46, torch/nn/modules/transformer.py:944-982."""
import numpy as np
import pytest
import torch

from oracle import u2gnn_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def U():
    import u2gnn_b200
    u2gnn_b200.require_device()
    return u2gnn_b200


def _np(t):
    return t.detach().float().cpu().numpy().astype(np.float64)


def _nrm(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_cfg5_shape_train_step_bf16_vs_fp64_oracle(U):
    from u2gnn_b200.synthetic import make_batch
    from u2gnn_b200.trainer import SupTrainer
    N, k, d, ff, T, L, C = 4096, 16, 64, 2048, 4, 1, 2
    b = make_batch(N, k, d, C, seed=77)
    torch.manual_seed(5)
    m = U.TransformerU2GNN(d, ff, C, T, 0.5, L, attn_axis="neighbors").cuda()
    with torch.no_grad():                                      # biases / LayerNorm away from their 0 / 1 init
        for n_, p_ in m.named_parameters():
            if "norm" in n_ or n_.endswith("bias"):
                p_.add_(0.1 * torch.randn_like(p_))
    P = {n: _np(v) for n, v in m.state_dict().items()}
    tr = SupTrainer(m, lr=5e-4, precision="bf16", seed=31)
    loss, scores = tr.forward_backward(b["input_x"], b["rowptr"], b["X"], b["labels"], train=True)
    torch.cuda.synchronize()
    seed = (31 * 0x9E3779B97F4A7C15 + 1) & 0xFFFFFFFFFFFFFFFF
    ix, rp, X = b["input_x"].cpu().numpy(), b["rowptr"].cpu().numpy(), _np(b["X"])
    so, cache = O.sup_forward(P, ix, rp, X, L, T, "neighbors", O.DropoutSpec(True, seed, 0.5, 0.5))
    lo, ds = O.soft_cross_entropy(so, O.label_smoothing(b["labels"].cpu().numpy(), C, dtype=np.float64))
    g = O.sup_backward(ds, cache, P, ix, rp, X)
    s = _np(scores)
    assert np.isfinite(s).all()
    assert np.abs(s - so).max() <= 2e-2 * np.abs(so).max(), np.abs(s - so).max() / np.abs(so).max()
    assert abs(loss.item() - lo) <= 2e-2 * abs(lo)
    gmax = max(np.linalg.norm(v) for v in g.values())
    worst = {}
    for n, p in m.named_parameters():
        got = _np(tr.arena.gviews[n])
        if np.linalg.norm(g[n]) < 1e-3 * gmax:
            continue                                           # numerically-zero gradients (key bias)
        worst[n] = _nrm(got, g[n])
    assert worst and max(worst.values()) < 5e-2, sorted(worst.items(), key=lambda kv: -kv[1])[:4]


@pytest.mark.parametrize("last", [False, True])
@pytest.mark.parametrize("ff", [2048, 1024])
def test_encoder_layer_bf16_kernels_vs_fp64_oracle(U, last, ff):
    from u2gnn_b200 import engine as E
    B, S, d = 301, 17, 64                                      # 5 117 rows: partial tiles in every tcgen05 kernel
    Sq = 1 if last else S
    rng = np.random.default_rng(40 + ff + int(last))
    x = rng.standard_normal((B, S, d))
    p64 = {"self_attn.in_proj_weight": rng.standard_normal((3 * d, d)) / np.sqrt(d), "self_attn.in_proj_bias": 0.1 * rng.standard_normal(3 * d),
           "self_attn.out_proj.weight": rng.standard_normal((d, d)) / np.sqrt(d), "self_attn.out_proj.bias": 0.1 * rng.standard_normal(d),
           "linear1.weight": rng.standard_normal((ff, d)) / np.sqrt(d), "linear1.bias": 0.1 * rng.standard_normal(ff),
           "linear2.weight": rng.standard_normal((d, ff)) / np.sqrt(ff), "linear2.bias": 0.1 * rng.standard_normal(d),
           "norm1.weight": 1 + 0.1 * rng.standard_normal(d), "norm1.bias": 0.1 * rng.standard_normal(d),
           "norm2.weight": 1 + 0.1 * rng.standard_normal(d), "norm2.bias": 0.1 * rng.standard_normal(d)}
    p32 = {n: v.astype(np.float32) for n, v in p64.items()}
    p64 = {n: v.astype(np.float64) for n, v in p32.items()}
    x32 = x.astype(np.float32)
    dy = rng.standard_normal((B, Sq, d)).astype(np.float32)
    SEED, ids, thr = 0x1234567, [16, 17, 18, 19], 128
    spec = O.DropoutSpec(True, SEED, 0.5, 0.5)
    shapes = {0: (B, Sq, S), 1: (B, Sq, d), 2: (B, Sq, ff), 3: (B, Sq, d)}
    masks = {site: spec.mask(ids[site], shp, 0.5, np.float64) for site, shp in shapes.items()}
    yo, c = O.encoder_layer_fwd(x32.astype(np.float64), p64, masks, last_only=last)
    dxo, go = O.encoder_layer_bwd(dy.astype(np.float64), c, p64)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    pt = {n: dev(v) for n, v in p32.items()}
    gt = {n: torch.zeros_like(v) for n, v in pt.items()}
    y2, sv = E.encoder_layer_fwd(dev(x32.reshape(B * S, d)), B, S, Sq, pt, d, ff, ids, SEED, thr, False, "bf16")
    dx = E.encoder_layer_bwd(dev(dy.reshape(B * Sq, d)), sv, pt, gt, d, ff, ids, SEED, thr, False, need_dx=True)
    torch.cuda.synchronize()
    rel = lambda got, ref: float(np.abs(got - ref).max() / np.abs(ref).max())
    # ---- forward intermediates, kernel by kernel (bf16 operand rounding: 2e-2 of the tensor's range)
    qkv_o = np.concatenate([np.broadcast_to(c["q"], c["k"].shape) if not last else c["k"] * 0, c["k"], c["v"]], -1).reshape(B * S, 3 * d)
    qkv = _np(sv.qkv)
    if last:
        assert rel(qkv[:, d:], qkv_o[:, d:]) < 2e-2                                   # in_proj (K, V of every row)
        assert rel(qkv.reshape(B, S, 3 * d)[:, 0, :d], c["q"][:, 0]) < 2e-2           # Q of position 0
    else:
        assert rel(qkv, qkv_o) < 2e-2                                                 # in_proj inside the attention kernel
    assert rel(_np(sv.ctx).reshape(B, Sq, d), c["ctx"]) < 2e-2                        # attention core (softmax, prob dropout, PV)
    z1_o = c["xq"] + (c["ctx"] @ p64["self_attn.out_proj.weight"].T + p64["self_attn.out_proj.bias"]) * masks[1]
    assert rel(_np(sv.z1).reshape(B, Sq, d), z1_o) < 2e-2                             # out_proj + dropout + residual
    assert rel(_np(sv.y1).reshape(B, Sq, d), c["y1"]) < 2e-2                          # LayerNorm1
    z2_o = c["y1"] + (c["hd"] @ p64["linear2.weight"].T + p64["linear2.bias"]) * masks[3]
    assert rel(_np(sv.z2).reshape(B, Sq, d), z2_o) < 2e-2                             # fused FFN + dropout + residual
    assert rel(_np(y2).reshape(B, Sq, d), yo) < 2e-2                                  # LayerNorm2
    # ---- backward: gradients (norm-wise, ReLU / bf16 rounding as in tests/test_gpu_tc.py)
    assert _nrm(_np(dx).reshape(B, S, d), dxo) < 5e-2
    gmax = max(np.linalg.norm(v) for v in go.values())
    for n, ref in go.items():
        if np.linalg.norm(ref) < 1e-3 * gmax:
            continue
        assert _nrm(_np(gt[n]), ref) < 5e-2, n


def test_bf16_training_tracks_fp32_over_120_steps(U):
    """the bf16 tcgen05 path is not only close per step - its TRAINING trajectory follows the fp32 path's.
    Same initial weights, same batches, same dropout streams, 120 fused steps (forward, loss, backward, clip, Adam).  Training is
    a chaotic map (p = 0.5 dropout, ReLU), so the two runs cannot stay bit-close; what must hold is that the loss CURVES agree:
    10-step means within 25 % pointwise, the area under the curve within 5 %, and both fall by more than 30 %.  (Measured on
    B200 at lr 2e-3: fp32 7.93 5.00 2.93 2.24 2.10 1.73 1.23 0.86 0.69 0.76 0.64 0.60, bf16 7.94 5.05 2.96 2.75 1.86 1.45 1.26
    0.97 0.62 0.69 0.53 0.47.)"""
    from u2gnn_b200.synthetic import make_batch
    from u2gnn_b200.trainer import SupTrainer
    batches = [make_batch(2000, 16, 64, 2, seed=100 + i) for i in range(4)]
    for b in batches:                                                    # learnable labels: sign of the graph's mean first feature
        sizes = b["rowptr"][1:] - b["rowptr"][:-1]
        gid = torch.repeat_interleave(torch.arange(b["G"], device="cuda"), sizes)
        b["labels"] = (torch.zeros(b["G"], device="cuda").index_add_(0, gid, b["X"][:, 0]) > 0).long()
    curves = {}
    for prec in ("fp32", "bf16"):
        torch.manual_seed(11)
        m = U.TransformerU2GNN(64, 256, 2, 2, 0.5, 1, attn_axis="neighbors").cuda()
        tr = SupTrainer(m, lr=1e-3, precision=prec, seed=5)
        losses = []
        for s in range(120):
            b = batches[s % 4]
            losses.append(tr.step(b["input_x"], b["rowptr"], b["X"], b["labels"]).clone())      # step() returns the trainer's loss buffer
        curves[prec] = torch.stack([l.reshape(()) for l in losses]).cpu().numpy().reshape(12, 10).mean(1)
    f, h = curves["fp32"], curves["bf16"]
    assert np.isfinite(h).all()
    assert f[-1] < 0.7 * f[0] and h[-1] < 0.7 * h[0], (f, h)             # both learn
    assert abs(h.mean() - f.mean()) <= 5e-2 * f.mean(), (f, h)
    assert (np.abs(h - f) / f).max() <= 0.25, (f, h)


def test_d65_bf16_ffn_matches_reference_fixture(U):
    """configs[2] (IMDBBINARY shape, d = 65) in its named precision: for 64 < d <= 128 `precision="bf16"` runs the FFN as tcgen05 GEMMs
    with the hidden materialised in bf16 (engine.ffn_wide_fwd / _bwd), the attention block in fp32.  Against the reference-generated
    fixture `sup_neighbors_d65`: eval scores and p = 0 loss within 2e-2, every gradient norm-wise within 5e-2."""
    from conftest import load_golden, split_case
    c = load_golden("sup_neighbors_d65")
    params, grads, _ = split_case(c)
    k, d, ff, T, L, C = [int(v) for v in c["meta"]]
    assert d == 65
    m = U.TransformerU2GNN(d, ff, C, T, 0.5, L, attn_axis="neighbors", precision="bf16").cuda()
    m.load_state_dict({n: torch.from_numpy(v) for n, v in params.items()})
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    m.eval()
    with torch.no_grad():
        s = m(dev(c["input_x"]), dev(c["rowptr"]), dev(c["X"]))
    assert np.abs(s.cpu().numpy() - c["eval_scores"]).max() <= 2e-2 * np.abs(c["eval_scores"]).max()
    m.train(); m.encoder_dropout = 0.0
    for dr in m.dropouts:
        dr.p = 0.0
    s = m(dev(c["input_x"]), dev(c["rowptr"]), dev(c["X"]))
    soft = U.label_smoothing(dev(c["labels"]), C)
    loss = torch.mean(torch.sum(-soft * torch.nn.functional.log_softmax(s, dim=1), 1))
    assert abs(loss.item() - float(c["loss"])) <= 2e-2 * abs(float(c["loss"]))
    loss.backward()
    gmax = max(np.linalg.norm(v) for v in grads.values())
    for n, p in m.named_parameters():
        if np.linalg.norm(grads[n]) < 1e-3 * gmax:
            continue
        assert _nrm(_np(p.grad), grads[n].astype(np.float64)) < 5e-2, n


@pytest.mark.parametrize("d,ff,p", [(65, 1024, 0.5), (100, 320, 0.5), (128, 256, 0.0)])
def test_wide_ffn_forward_backward_vs_fp64_oracle(U, d, ff, p):
    """engine.ffn_wide_fwd / _bwd (64 < d <= 128) with dropout ON against the fp64 arithmetic of transformer.py:977-982 and its
    autograd, using the engine's dropout stream (oracle.dropout_keep_mask)."""
    from u2gnn_b200 import engine as E
    M = 700
    thr = E.dropout_threshold(p)
    scale = 256.0 / (256.0 - thr) if thr else 1.0
    rng = np.random.default_rng(d + ff)
    y1 = rng.standard_normal((M, d)).astype(np.float32)
    prm = {"linear1.weight": (rng.standard_normal((ff, d)) / np.sqrt(d)).astype(np.float32), "linear1.bias": (0.1 * rng.standard_normal(ff)).astype(np.float32),
           "linear2.weight": (rng.standard_normal((d, ff)) / np.sqrt(ff)).astype(np.float32), "linear2.bias": (0.1 * rng.standard_normal(d)).astype(np.float32)}
    df = rng.standard_normal((M, d)).astype(np.float32)
    dz = rng.standard_normal((M, d)).astype(np.float32)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    pt = {n: dev(v) for n, v in prm.items()}
    gt = {n: torch.zeros_like(v) for n, v in pt.items()}
    SEED, ST = 0xC0FFEE, 77
    f, h = E.ffn_wide_fwd(dev(y1), M, d, ff, pt, SEED, ST, thr)
    dzt = dev(dz)
    dy1 = E.ffn_wide_bwd(dev(df), dzt, h, M, d, ff, pt, gt, thr)
    torch.cuda.synchronize()
    keep = np.ones((M, ff))
    if thr:
        keep = O.dropout_keep_mask(SEED, ST, M * ff, p)[0].reshape(M, ff).astype(np.float64)
    P = {n: v.astype(np.float64) for n, v in prm.items()}
    pre = y1.astype(np.float64) @ P["linear1.weight"].T + P["linear1.bias"]
    hid = np.maximum(pre, 0) * keep * scale
    f_o = hid @ P["linear2.weight"].T + P["linear2.bias"]
    dh = (df.astype(np.float64) @ P["linear2.weight"]) * keep * scale * (pre > 0)
    ref = {"f": f_o, "dy1": dz + dh @ P["linear1.weight"], "linear1.weight": dh.T @ y1, "linear1.bias": dh.sum(0), "linear2.weight": df.T.astype(np.float64) @ hid}
    got = {"f": _np(f), "dy1": _np(dy1), "linear1.weight": _np(gt["linear1.weight"]), "linear1.bias": _np(gt["linear1.bias"]), "linear2.weight": _np(gt["linear2.weight"])}
    assert np.abs(got["f"] - ref["f"]).max() <= 2e-2 * np.abs(ref["f"]).max()
    for n in ref:
        assert _nrm(got[n], ref[n]) < 5e-2, n
