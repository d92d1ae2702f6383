"""CPU oracle for the U2GNN train-step hot path (TEST INFRASTRUCTURE ONLY).

This file is a numpy restatement of the arithmetic the reference runs for the
path BASELINE.json names.  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it.  Nothing under ``graph-transformer_b200/``
imports it, and the product path fails loudly without its CUDA library.

Parity status: the reference has no tests and no golden vectors,
so this oracle is pinned against outputs of the reference itself generated in
the build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference/U2GNN_pytorch`` verbatim) and committed under
``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every function here
against those fixtures.

What each function follows (paths relative to /root/reference):

* ``gather_rows``            U2GNN_pytorch/pytorch_U2GNN_Sup.py:32,39   (F.embedding)
* ``encoder_layer_fwd/bwd``  pytorch_U2GNN_Sup.py:20-21,35 -> torch.nn.TransformerEncoderLayer
                             (torch/nn/modules/transformer.py:944-982, post-norm, ReLU, 1 head;
                             MHA math torch/nn/functional.py multi_head_attention_forward)
* ``segment_sum``            pytorch_U2GNN_Sup.py:41  (torch.spmm(graph_pool, .) with the 0/1
                             pooling operator built at train_pytorch_U2GNN_Sup.py:73-89)
* ``sup_forward/backward``   pytorch_U2GNN_Sup.py:30-46
* ``label_smoothing``        pytorch_U2GNN_Sup.py:48-59
* ``soft_cross_entropy``     train_pytorch_U2GNN_Sup.py:140-142
* ``sampled_softmax_*``      U2GNN_pytorch/sampled_softmax.py:36-56
* ``unsup_forward/backward`` the assembled unsupervised model of SURVEY.md §8(c)
                             (ctor lines pytorch_U2GNN_UnSup.py:37-44; dataflow
                             U2GNN_tf/model_U2GNN_Unsup_multi.py:32-58)
* ``clip_grad_norm``/``adam_step``  train_pytorch_U2GNN_Sup.py:145,160-161
                             (torch.nn.utils.clip_grad_norm_, torch.optim.Adam defaults)
* ``dropout_keep_mask``      NOT reference arithmetic: the engine's own counter-based dropout
                             stream (graph-transformer_b200/csrc/rng.cuh) restated so that
                             train-mode runs can be compared with dropout switched on.  The
                             reference's masks come from torch's global generator and cannot be
                             matched (SURVEY.md "Hard parts").

Two attention layouts:
  attn_axis="nodes"      reference as written: sequence = the N nodes of the batch, only
                         column 0 of input_x is live.
  attn_axis="neighbors"  intended: sequence = [node, k sampled neighbours]; equals the reference
                         modules fed ``input_Tr.transpose(0, 1)``.
Weights are T independent sets per U2GNN layer.
"""
from __future__ import annotations

import math
import numpy as np

LN_EPS = 1e-5

# --------------------------------------------------------------------------------------
# engine dropout stream (restated from csrc/rng.cuh; not reference arithmetic)
# --------------------------------------------------------------------------------------
_M32 = np.uint64(0xFFFFFFFF)


def _mix32(x):
    x = np.asarray(x, dtype=np.uint64) & _M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & _M32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & _M32
    x ^= x >> np.uint64(16)
    return x


def rng_keys(seed: int, stream: int):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    lo, hi = seed & 0xFFFFFFFF, seed >> 32
    k0 = int(_mix32(lo ^ ((stream * 0x9E3779B1) & 0xFFFFFFFF)))
    k1 = int(_mix32((hi + stream + 0x7F4A7C15) & 0xFFFFFFFF))
    return k0, k1


def rng_word(k0: int, k1: int, g, plane: int):
    g = np.asarray(g, dtype=np.uint64)
    a = _mix32((g & _M32) ^ np.uint64(k0))
    b = (a + ((g >> np.uint64(32)) * np.uint64(0x9E3779B1)) + np.uint64(k1)
         + np.uint64((plane * 0x632BE5AB) & 0xFFFFFFFF)) & _M32
    return _mix32(b)


def dropout_threshold(p: float) -> int:
    """p is quantised to thr/256 (p=0.5 -> 128 exactly)."""
    thr = int(round(float(p) * 256.0))
    if thr < 0 or thr > 255:
        raise ValueError("dropout p out of range")
    return thr


def dropout_keep_mask(seed: int, stream: int, numel: int, p: float):
    """Return (keep[bool numel], scale).  Element e is kept iff the 8-bit number whose bit j is
    bit (e&31) of word(e>>5, plane j) is >= thr."""
    thr = dropout_threshold(p)
    if thr == 0:
        return np.ones(numel, dtype=bool), 1.0
    k0, k1 = rng_keys(seed, stream)
    ng = (numel + 31) // 32
    g = np.arange(ng, dtype=np.uint64)
    r = np.zeros(ng * 32, dtype=np.uint32)
    bitpos = np.arange(32, dtype=np.uint64)
    for plane in range(8):
        w = rng_word(k0, k1, g, plane)
        bits = ((w[:, None] >> bitpos[None, :]) & np.uint64(1)).astype(np.uint32).reshape(-1)
        r |= bits << np.uint32(plane)
    keep = r[:numel] >= thr
    return keep, 256.0 / (256.0 - thr)


def stream_id(layer: int, timestep: int, site: int, num_timesteps: int) -> int:
    """site: 0 attention probs, 1 post-attention, 2 post-ReLU, 3 post-FFN."""
    return ((layer * num_timesteps + timestep) * 4 + site) + 16


STREAM_POOLED = 0x40000000  # + layer  (dropout on pooled graph embeddings, p = args.dropout)
STREAM_CONCAT = 0x50000000  # dropout on concatenated node vectors (unsupervised)


class DropoutSpec:
    """Dropout configuration for one step.  enabled=False == model.eval() / p=0."""

    def __init__(self, enabled=False, seed=0, p_enc=0.5, p_out=0.5):
        self.enabled, self.seed, self.p_enc, self.p_out = enabled, seed, p_enc, p_out

    def mask(self, stream, shape, p, dtype):
        if not self.enabled or p == 0.0:
            return None
        keep, scale = dropout_keep_mask(self.seed, stream, int(np.prod(shape)), p)
        return (keep.reshape(shape).astype(dtype) * dtype(scale))


# --------------------------------------------------------------------------------------
# primitive ops
# --------------------------------------------------------------------------------------
def sample_neighbors_device_stream(g_rowptr, g_col, graph_start, batch_off, k, seed, stream):
    """Restatement of the engine's device batch builder (csrc/batch_builder.cu), which replaces the host loop of
    get_batch_data (train_pytorch_U2GNN_Sup.py:100-114): input_x[b] = [b, k neighbours of b drawn with replacement],
    isolated nodes repeat themselves; draw (b, j) = (rng_word(keys, b*k + j - 1, plane 0) * deg) >> 32.
    -> (input_x int64 [N, k+1] batch-local ids, node_global int64 [N])."""
    g_rowptr = np.asarray(g_rowptr, dtype=np.int64); g_col = np.asarray(g_col, dtype=np.int64)
    graph_start = np.asarray(graph_start, dtype=np.int64); batch_off = np.asarray(batch_off, dtype=np.int64)
    N = int(batch_off[-1])
    gi = np.searchsorted(batch_off, np.arange(N), side="right") - 1
    shift = batch_off[gi] - graph_start[gi]
    v = np.arange(N, dtype=np.int64) - shift
    deg = g_rowptr[v + 1] - g_rowptr[v]
    out = np.repeat(np.arange(N, dtype=np.int64)[:, None], k + 1, axis=1)
    if k > 0 and N > 0:
        k0, k1 = rng_keys(seed, stream)
        ctr = (np.arange(N, dtype=np.uint64)[:, None] * np.uint64(k) + np.arange(k, dtype=np.uint64)[None, :])
        r = rng_word(k0, k1, ctr.reshape(-1), 0).astype(np.uint64).reshape(N, k)
        pick = ((r * deg[:, None].astype(np.uint64)) >> np.uint64(32)).astype(np.int64)
        has = deg > 0
        idx = g_rowptr[v][:, None] + pick
        nb = g_col[np.where(has[:, None], idx, 0)] + shift[:, None]
        out[:, 1:] = np.where(has[:, None], nb, out[:, 1:])
    return out, v


def gather_rows(table, idx):
    """F.embedding(idx, table) (pytorch_U2GNN_Sup.py:32,39)."""
    return table[idx]


def scatter_add_rows(grad_out, idx, n_rows):
    """Backward of gather_rows: sums rows with duplicate indices."""
    g = np.zeros((n_rows, grad_out.shape[-1]), dtype=grad_out.dtype)
    np.add.at(g, idx.reshape(-1), grad_out.reshape(-1, grad_out.shape[-1]))
    return g


def rowptr_from_graph_pool(indices, num_graphs):
    """CSR row pointer of the COO pooling operator (train_pytorch_U2GNN_Sup.py:73-89):
    indices[0] = graph id (non-decreasing), indices[1] = node id (0..N-1 in order)."""
    rows = np.asarray(indices[0], dtype=np.int64)
    cols = np.asarray(indices[1], dtype=np.int64)
    if rows.size and (np.any(np.diff(rows) < 0) or np.any(cols != np.arange(cols.size))):
        raise ValueError("graph_pool is not the reference's contiguous block pooling operator")
    counts = np.bincount(rows, minlength=num_graphs).astype(np.int64)
    return np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)


def segment_sum(x, rowptr):
    """torch.spmm(graph_pool, x) (pytorch_U2GNN_Sup.py:41); sums in ascending node order."""
    G = len(rowptr) - 1
    out = np.zeros((G, x.shape[1]), dtype=x.dtype)
    for g in range(G):
        seg = x[rowptr[g]:rowptr[g + 1]]
        acc = np.zeros(x.shape[1], dtype=x.dtype)
        for r in seg:
            acc = acc + r
        out[g] = acc
    return out


def segment_sum_bwd(grad_out, rowptr, n):
    g = np.zeros((n, grad_out.shape[1]), dtype=grad_out.dtype)
    for i in range(len(rowptr) - 1):
        g[rowptr[i]:rowptr[i + 1]] = grad_out[i]
    return g


def layer_norm_fwd(z, w, b):
    mean = z.mean(-1, keepdims=True)
    var = ((z - mean) ** 2).mean(-1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + z.dtype.type(LN_EPS))
    xhat = (z - mean) * rstd
    return xhat * w + b, (xhat, rstd)


def layer_norm_bwd(dy, cache, w):
    xhat, rstd = cache
    dw = (dy * xhat).reshape(-1, xhat.shape[-1]).sum(0)
    db = dy.reshape(-1, xhat.shape[-1]).sum(0)
    dxhat = dy * w
    dz = rstd * (dxhat - dxhat.mean(-1, keepdims=True) - xhat * (dxhat * xhat).mean(-1, keepdims=True))
    return dz, dw, db


def _softmax(s):
    m = s.max(-1, keepdims=True)
    e = np.exp(s - m)
    return e / e.sum(-1, keepdims=True)


def encoder_layer_fwd(x, p, masks=None, last_only=False):
    """One post-norm TransformerEncoderLayer (nhead=1, ReLU) on x[B, S, d] with attention over
    axis 1.  p: dict with the reference state_dict names.  masks: None or dict site->array
    (already scaled by 1/(1-p)).  last_only: only sequence position 0 is produced for the
    out-projection / FFN (dead-row elimination, pytorch_U2GNN_Sup.py:36-37); K and V still use
    every position.  Returns (y[B, S or 1, d], cache)."""
    masks = masks or {}
    d = x.shape[-1]
    Wq, Wk, Wv = p["self_attn.in_proj_weight"][:d], p["self_attn.in_proj_weight"][d:2 * d], p["self_attn.in_proj_weight"][2 * d:]
    bq, bk, bv = p["self_attn.in_proj_bias"][:d], p["self_attn.in_proj_bias"][d:2 * d], p["self_attn.in_proj_bias"][2 * d:]
    xq = x[:, :1] if last_only else x
    q = xq @ Wq.T + bq
    k = x @ Wk.T + bk
    v = x @ Wv.T + bv
    scale = x.dtype.type(math.sqrt(1.0 / d))
    s = (q * scale) @ k.transpose(0, 2, 1)
    pr = _softmax(s)
    m0 = masks.get(0)
    prd = pr * m0 if m0 is not None else pr
    ctx = prd @ v
    a = ctx @ p["self_attn.out_proj.weight"].T + p["self_attn.out_proj.bias"]
    m1 = masks.get(1)
    ad = a * m1 if m1 is not None else a
    z1 = xq + ad
    y1, ln1 = layer_norm_fwd(z1, p["norm1.weight"], p["norm1.bias"])
    hpre = y1 @ p["linear1.weight"].T + p["linear1.bias"]
    h = np.maximum(hpre, 0)
    m2 = masks.get(2)
    hd = h * m2 if m2 is not None else h
    f = hd @ p["linear2.weight"].T + p["linear2.bias"]
    m3 = masks.get(3)
    fd = f * m3 if m3 is not None else f
    z2 = y1 + fd
    y2, ln2 = layer_norm_fwd(z2, p["norm2.weight"], p["norm2.bias"])
    cache = dict(x=x, xq=xq, q=q, k=k, v=v, pr=pr, prd=prd, ctx=ctx, y1=y1, ln1=ln1, hpre=hpre,
                 hd=hd, ln2=ln2, masks=masks, scale=scale, last_only=last_only)
    return y2, cache


def encoder_layer_bwd(dy2, c, p):
    """Backward of encoder_layer_fwd.  Returns (dx[B,S,d], grads dict)."""
    g = {}
    masks = c["masks"]
    d = c["x"].shape[-1]
    W = p["self_attn.in_proj_weight"]
    Wq, Wk, Wv = W[:d], W[d:2 * d], W[2 * d:]
    dz2, g["norm2.weight"], g["norm2.bias"] = layer_norm_bwd(dy2, c["ln2"], p["norm2.weight"])
    dfd = dz2
    df = dfd * masks[3] if masks.get(3) is not None else dfd
    f2 = lambda t: t.reshape(-1, t.shape[-1])
    g["linear2.weight"] = f2(df).T @ f2(c["hd"])
    g["linear2.bias"] = f2(df).sum(0)
    dhd = df @ p["linear2.weight"]
    dh = dhd * masks[2] if masks.get(2) is not None else dhd
    dhpre = dh * (c["hpre"] > 0)
    g["linear1.weight"] = f2(dhpre).T @ f2(c["y1"])
    g["linear1.bias"] = f2(dhpre).sum(0)
    dy1 = dz2 + dhpre @ p["linear1.weight"]
    dz1, g["norm1.weight"], g["norm1.bias"] = layer_norm_bwd(dy1, c["ln1"], p["norm1.weight"])
    dad = dz1
    da = dad * masks[1] if masks.get(1) is not None else dad
    g["self_attn.out_proj.weight"] = f2(da).T @ f2(c["ctx"])
    g["self_attn.out_proj.bias"] = f2(da).sum(0)
    dctx = da @ p["self_attn.out_proj.weight"]
    dprd = dctx @ c["v"].transpose(0, 2, 1)
    dv = c["prd"].transpose(0, 2, 1) @ dctx
    dpr = dprd * masks[0] if masks.get(0) is not None else dprd
    ds = c["pr"] * (dpr - (dpr * c["pr"]).sum(-1, keepdims=True))
    dq = (ds @ c["k"]) * c["scale"]
    dk = ds.transpose(0, 2, 1) @ (c["q"] * c["scale"])
    dW = np.concatenate([f2(dq).T @ f2(c["xq"]), f2(dk).T @ f2(c["x"]), f2(dv).T @ f2(c["x"])], 0)
    g["self_attn.in_proj_weight"] = dW
    g["self_attn.in_proj_bias"] = np.concatenate([f2(dq).sum(0), f2(dk).sum(0), f2(dv).sum(0)])
    dx = dk @ Wk + dv @ Wv
    dxq = dz1 + dq @ Wq
    if c["last_only"]:
        dx[:, :1] += dxq
    else:
        dx = dx + dxq
    return dx, g


# --------------------------------------------------------------------------------------
# U2GNN encoder stack (one U2GNN layer = T encoder layers)
# --------------------------------------------------------------------------------------
def _layer_params(params, l, t):
    pre = f"u2gnn_layers.{l}.layers.{t}."
    return {k[len(pre):]: v for k, v in params.items() if k.startswith(pre)}


def u2gnn_layer_fwd(src, input_x, params, l, T, attn_axis, drop: DropoutSpec):
    """One U2GNN layer: gather -> T encoder layers -> sequence position 0.
    src[N, d] (X_concat or the previous layer's output).  Returns (out[N, d], cache)."""
    dt = src.dtype.type
    if attn_axis == "neighbors":
        x = gather_rows(src, input_x)                      # [N, S, d]; batch N, sequence S
    elif attn_axis == "nodes":
        x = gather_rows(src, input_x[:, 0])[None]          # [1, N, d]; the only live column (F1)
    else:
        raise ValueError(attn_axis)
    caches = []
    for t in range(T):
        p = _layer_params(params, l, t)
        last = attn_axis == "neighbors" and t == T - 1
        B, S, d = x.shape
        So = 1 if last else S
        ff = p["linear1.weight"].shape[0]
        masks = {}
        if drop.enabled:
            shapes = {0: (B, So, S), 1: (B, So, d), 2: (B, So, ff), 3: (B, So, d)}
            for site, shp in shapes.items():
                masks[site] = drop.mask(stream_id(l, t, site, T), shp, drop.p_enc, dt)
        x, c = encoder_layer_fwd(x, p, masks, last_only=last)
        caches.append(c)
    out = x[:, 0] if attn_axis == "neighbors" else x[0]
    return out, dict(caches=caches, attn_axis=attn_axis, n_src=src.shape[0])


def u2gnn_layer_bwd(dout, cache, input_x, params, l, T, need_dsrc=True):
    grads = {}
    caches = cache["caches"]
    if cache["attn_axis"] == "neighbors":
        dx = dout[:, None, :]
    else:
        dx = dout[None]
    for t in reversed(range(T)):
        p = _layer_params(params, l, t)
        dx, g = encoder_layer_bwd(dx, caches[t], p)
        for k, v in g.items():
            grads[f"u2gnn_layers.{l}.layers.{t}.{k}"] = v
    if not need_dsrc:
        return None, grads
    if cache["attn_axis"] == "neighbors":
        dsrc = scatter_add_rows(dx, input_x, cache["n_src"])
    else:
        dsrc = scatter_add_rows(dx[0], input_x[:, 0], cache["n_src"])
    return dsrc, grads


# --------------------------------------------------------------------------------------
# supervised model + loss  (pytorch_U2GNN_Sup.py:30-46, train_pytorch_U2GNN_Sup.py:140-142)
# --------------------------------------------------------------------------------------
def label_smoothing(labels, classes, smoothing=0.1, dtype=np.float32):
    t = np.full((len(labels), classes), smoothing / (classes - 1), dtype=dtype)
    t[np.arange(len(labels)), labels] = 1.0 - smoothing
    return t


def soft_cross_entropy(scores, soft):
    m = scores.max(1, keepdims=True)
    lse = m + np.log(np.exp(scores - m).sum(1, keepdims=True))
    logp = scores - lse
    loss = (-(soft * logp).sum(1)).mean()
    dscores = (np.exp(logp) * soft.sum(1, keepdims=True) - soft) / scores.shape[0]
    return loss, dscores


def sup_forward(params, input_x, rowptr, X, L, T, attn_axis="neighbors", drop=None):
    drop = drop or DropoutSpec()
    dt = X.dtype.type
    G = len(rowptr) - 1
    src = X
    scores = 0
    cache = dict(layers=[], L=L, T=T)
    for l in range(L):
        out, c = u2gnn_layer_fwd(src, input_x, params, l, T, attn_axis, drop)
        ge = segment_sum(out, rowptr)
        m = drop.mask(STREAM_POOLED + l, ge.shape, drop.p_out, dt)
        ged = ge * m if m is not None else ge
        scores = scores + ged @ params[f"predictions.{l}.weight"].T + params[f"predictions.{l}.bias"]
        cache["layers"].append(dict(enc=c, out=out, ged=ged, m=m))
        src = out
    cache["G"] = G
    return scores, cache


def sup_backward(dscores, cache, params, input_x, rowptr, X):
    L, T = cache["L"], cache["T"]
    grads = {}
    dsrc_next = None
    for l in reversed(range(L)):
        c = cache["layers"][l]
        grads[f"predictions.{l}.weight"] = dscores.T @ c["ged"]
        grads[f"predictions.{l}.bias"] = dscores.sum(0)
        dged = dscores @ params[f"predictions.{l}.weight"]
        dge = dged * c["m"] if c["m"] is not None else dged
        dout = segment_sum_bwd(dge, rowptr, c["out"].shape[0])
        if dsrc_next is not None:
            dout = dout + dsrc_next
        dsrc_next, g = u2gnn_layer_bwd(dout, c["enc"], input_x, params, l, T, need_dsrc=(l > 0))
        grads.update(g)
    return grads


# --------------------------------------------------------------------------------------
# sampled softmax (sampled_softmax.py:36-56) and the assembled unsupervised model
# --------------------------------------------------------------------------------------
def sampled_softmax_fwd(x, labels, W, sample_ids):
    """loss_i = -log( exp(x_i.W[y_i]) / sum_s exp(x_i.W[s]) ); no max subtraction, no log-Q
    correction, true class not added to the denominator."""
    tw = W[labels]
    sw = W[np.asarray(sample_ids, dtype=np.int64)]
    tl = (x * tw).sum(1)
    sl = x @ sw.T
    true_e = np.exp(tl)
    samp_e = np.exp(sl)
    denom = samp_e.sum(1)
    loss = -np.log(true_e / denom)
    return loss, dict(tw=tw, sw=sw, samp_e=samp_e, denom=denom)


def sampled_softmax_bwd(dloss, c, x, labels, W_shape, sample_ids):
    """Gradients of sum_i dloss_i * loss_i wrt x and the dense class table."""
    psamp = c["samp_e"] / c["denom"][:, None]          # [N, ns]
    dsl = psamp * dloss[:, None]
    dtl = -dloss
    dx = dtl[:, None] * c["tw"] + dsl @ c["sw"]
    dW = np.zeros(W_shape, dtype=x.dtype)
    np.add.at(dW, labels, dtl[:, None] * x)
    np.add.at(dW, np.asarray(sample_ids, dtype=np.int64), dsl.T @ x)
    return dx, dW


def sampled_softmax_tf_fwd(x, labels, W, b, sample_ids, true_q, samp_q, remove_accidental_hits=True):
    """The TF model's loss (U2GNN_tf/model_U2GNN_Unsup_multi.py:54-58: tf.nn.sampled_softmax_loss with its defaults
    subtract_log_q=True, remove_accidental_hits=True) - NOT what the reference's PyTorch SampledSoftmax computes (above); restated
    here as the oracle of the SURVEY.md 8(f) row-4 extra that has no CUDA path yet.
        t_i  = x_i.W[y_i] + b[y_i] - log Q(y_i)              true_q[i]  = expected count of the label (Log_Uniform_Sampler.cpp:23-32)
        l_is = x_i.W[s]   + b[s]   - log Q(s)                samp_q[s]  = expected count of sampled id s
        l_is = -inf where s == y_i (accidental hit);   loss_i = logsumexp([t_i, l_i1 .. l_ins]) - t_i   (label at position 0)."""
    ids = np.asarray(sample_ids, dtype=np.int64)
    tw, sw = W[labels], W[ids]
    t = (x * tw).sum(1) + b[labels] - np.log(true_q)
    l = x @ sw.T + b[ids][None, :] - np.log(samp_q)[None, :]
    hit = (ids[None, :] == np.asarray(labels)[:, None]) if remove_accidental_hits else np.zeros(l.shape, dtype=bool)
    l = np.where(hit, -np.inf, l)
    m = np.maximum(t, l.max(1))
    et, el = np.exp(t - m), np.exp(l - m[:, None])
    denom = et + el.sum(1)
    loss = np.log(denom) + m - t
    return loss, dict(tw=tw, sw=sw, p_true=et / denom, p_samp=el / denom[:, None], ids=ids)


def sampled_softmax_tf_bwd(dloss, c, x, labels, W_shape):
    """Gradients of sum_i dloss_i * loss_i of sampled_softmax_tf_fwd wrt x, the class table and the bias."""
    dt = (c["p_true"] - 1.0) * dloss
    dl = c["p_samp"] * dloss[:, None]
    dx = dt[:, None] * c["tw"] + dl @ c["sw"]
    dW = np.zeros(W_shape, dtype=x.dtype)
    db = np.zeros(W_shape[0], dtype=x.dtype)
    np.add.at(dW, labels, dt[:, None] * x)
    np.add.at(dW, c["ids"], dl.T @ x)
    np.add.at(db, labels, dt)
    np.add.at(db, c["ids"], dl.sum(0))
    return dx, dW, db


def unsup_forward(params, X, input_x, input_y, sample_ids, L, T, attn_axis="neighbors", drop=None):
    """Assembled unsupervised model: per-layer position-0 vectors, concat over
    layers, dropout, SampledSoftmax.  Returns (per-node loss[N], cache)."""
    drop = drop or DropoutSpec()
    dt = X.dtype.type
    src = X
    outs, encs = [], []
    for l in range(L):
        out, c = u2gnn_layer_fwd(src, input_x, params, l, T, attn_axis, drop)
        outs.append(out)
        encs.append(c)
        src = out
    cat = np.concatenate(outs, 1)
    m = drop.mask(STREAM_CONCAT, cat.shape, drop.p_out, dt)
    catd = cat * m if m is not None else cat
    loss, sc = sampled_softmax_fwd(catd, input_y, params["ss.weight"], sample_ids)
    return loss, dict(encs=encs, m=m, catd=catd, sc=sc, L=L, T=T, d=X.shape[1])


def unsup_backward(dloss, cache, params, X, input_x, input_y, sample_ids):
    L, T, d = cache["L"], cache["T"], cache["d"]
    dcatd, dW = sampled_softmax_bwd(dloss, cache["sc"], cache["catd"], input_y,
                                    params["ss.weight"].shape, sample_ids)
    dcat = dcatd * cache["m"] if cache["m"] is not None else dcatd
    grads = {"ss.weight": dW}
    dsrc_next = None
    for l in reversed(range(L)):
        dout = dcat[:, l * d:(l + 1) * d]
        if dsrc_next is not None:
            dout = dout + dsrc_next
        dsrc_next, g = u2gnn_layer_bwd(dout, cache["encs"][l], input_x, params, l, T, need_dsrc=(l > 0))
        grads.update(g)
    return grads


# --------------------------------------------------------------------------------------
# optimiser step (train_pytorch_U2GNN_Sup.py:145,160-161)
# --------------------------------------------------------------------------------------
def clip_grad_norm(grads, max_norm=0.5):
    """torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1."""
    total = math.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads.values()))
    coef = min(1.0, max_norm / (total + 1e-6))
    return {k: g * g.dtype.type(coef) for k, g in grads.items()}, total


def adam_step(params, grads, state, lr, step, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults (no weight decay, no amsgrad); step is 1-based."""
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    for k in params:
        g = grads[k]
        m = state.setdefault(("m", k), np.zeros_like(g))
        v = state.setdefault(("v", k), np.zeros_like(g))
        m[...] = b1 * m + (1 - b1) * g
        v[...] = b2 * v + (1 - b2) * g * g
        denom = np.sqrt(v) / math.sqrt(bc2) + eps
        params[k] = params[k] - (lr / bc1) * (m / denom)
    return params
