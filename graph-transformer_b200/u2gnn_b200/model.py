"""Drop-in module surface of the reference on top of the CUDA engine.

  TransformerU2GNN          pytorch_U2GNN_Sup.py:7-46      forward(input_x, graph_pool, X_concat)
  TransformerU2GNNUnSup     pytorch_U2GNN_UnSup.py:12-93   forward(X_concat, input_x, input_y) (ctor/forward
                            signature and `.ss.weight` kept; body = the assembled model of SURVEY.md §8(c))
  SampledSoftmax            sampled_softmax.py:11-56       forward(inputs, labels) / sampled(inputs, labels, sample_values)
  LogUniformSampler         log_uniform/log_uniform.pyx:16-40
  label_smoothing           pytorch_U2GNN_Sup.py:48-59

Parameters are created by instantiating the same torch.nn modules in the same order as the
reference, so `torch.manual_seed(s)` gives bit-identical initial weights and identical state_dict
keys (T independent weight sets per U2GNN layer).  The modules are used as parameter
containers only: every forward/backward arithmetic step runs in libu2gnn_b200.so.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
from torch.nn import TransformerEncoder, TransformerEncoderLayer

from . import engine as E
from ._lib import LIB, require_device


def _tie(encoder: TransformerEncoder, tie):
    """tie=True: ONE weight set shared by all T timesteps - the Universal Transformer of the published U2GNN
    (U2GNN_tf/universal_transformer_modified_utils.py:251-252, 552-586).  The reference's PyTorch
    file clones T independent layers (torch/nn/modules/transformer.py _get_clones), which stays the default.  The same module
    object is listed T times, so state_dict() still carries the reference names layers.{t}.* (all aliasing one tensor)."""
    if tie:
        encoder.layers = nn.ModuleList([encoder.layers[0]] * len(encoder.layers))
    return encoder


def _layer_param_dicts(encoder: TransformerEncoder):
    out = []
    for layer in encoder.layers:
        sd = dict(layer.named_parameters())
        out.append({n: sd[n] for n in E.PARAM_NAMES})
    return out


class _StackFn(torch.autograd.Function):
    """One U2GNN layer (gather -> T encoder layers -> position 0) as a single autograd node."""

    @staticmethod
    def forward(ctx, src, input_x, cfg, *flat):
        l, T, axis, drop, transpose, precision = cfg
        params = [dict(zip(E.PARAM_NAMES, [t.detach() for t in flat[i * 12:(i + 1) * 12]])) for i in range(T)]
        out, saved = E.u2gnn_layer_fwd(src.detach().contiguous(), input_x, params, l, T, axis, drop, precision)
        ctx.saved_stack, ctx.params, ctx.cfg, ctx.input_x = saved, params, cfg, input_x
        ctx.need_dsrc = src.requires_grad
        return out

    @staticmethod
    def backward(ctx, dout):
        l, T, axis, drop, transpose, precision = ctx.cfg
        grads = [{n: torch.zeros_like(t) for n, t in p.items()} for p in ctx.params]
        dsrc = E.u2gnn_layer_bwd(dout.contiguous(), ctx.saved_stack, ctx.input_x, ctx.params, grads, l, T, axis, drop,
                                 need_dsrc=ctx.need_dsrc, transpose=transpose)
        ctx.saved_stack = None
        flat = [g[n] for g in grads for n in E.PARAM_NAMES]
        return (dsrc, None, None, *flat)


class _PoolFn(torch.autograd.Function):
    """torch.spmm(graph_pool, x) with the CSR form of the pooling operator (pytorch_U2GNN_Sup.py:41)."""

    @staticmethod
    def forward(ctx, x, rowptr):
        ctx.rowptr, ctx.n = rowptr, x.shape[0]
        return E.segment_sum(x.detach().contiguous(), rowptr)

    @staticmethod
    def backward(ctx, g):
        return E.segment_sum_bwd(g.contiguous(), ctx.rowptr, ctx.n), None


class _HeadFn(torch.autograd.Function):
    """dropout(ge) @ W^T + b (pytorch_U2GNN_Sup.py:42-44)."""

    @staticmethod
    def forward(ctx, ge, W, b, seed, stream, thr):
        ge, Wd, bd = ge.detach().contiguous(), W.detach(), b.detach()
        G, d = ge.shape
        C = W.shape[0]
        scores = torch.empty((G, C), dtype=torch.float32, device=ge.device)
        LIB.call("u2gnn_head_fwd", ge.data_ptr(), G, d, Wd.data_ptr(), bd.data_ptr(), C, seed, stream, thr,
                 scores.data_ptr(), 0, E._stream())
        ctx.args = (ge, Wd, seed, stream, thr)
        return scores

    @staticmethod
    def backward(ctx, ds):
        ge, W, seed, stream, thr = ctx.args
        G, d = ge.shape
        C = W.shape[0]
        ds = ds.contiguous()
        dW = torch.zeros_like(W)
        db = torch.zeros(C, dtype=torch.float32, device=W.device)
        dge = torch.empty_like(ge)
        LIB.call("u2gnn_head_bwd", ds.data_ptr(), ge.data_ptr(), G, d, W.data_ptr(), C, seed, stream, thr, dW.data_ptr(),
                 db.data_ptr(), dge.data_ptr(), E._stream())
        return dge, dW, db, None, None, None


class _DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, seed, stream, thr):
        x = x.detach().contiguous()
        y = torch.empty_like(x)
        LIB.call("u2gnn_dropout_apply", x.data_ptr(), x.numel(), seed, stream, thr, y.data_ptr(), E._stream())
        ctx.args = (seed, stream, thr)
        return y

    @staticmethod
    def backward(ctx, g):
        seed, stream, thr = ctx.args
        g = g.contiguous()
        y = torch.empty_like(g)
        LIB.call("u2gnn_dropout_apply", g.data_ptr(), g.numel(), seed, stream, thr, y.data_ptr(), E._stream())
        return y, None, None, None


class _U2GNNBase(nn.Module):
    def _init_engine(self, attn_axis, deterministic, precision="fp32"):
        if attn_axis not in ("nodes", "neighbors"):
            raise ValueError("attn_axis must be 'nodes' or 'neighbors'")
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.precision = precision            # 'bf16': FFN on tcgen05 tensor cores (2e-2 tolerance mode)
        self.attn_axis = attn_axis
        self.deterministic = deterministic
        self.encoder_dropout = 0.5            # hard-coded in the reference (pytorch_U2GNN_Sup.py:20)
        self._rng_seed = int(torch.initial_seed()) & 0x7FFFFFFFFFFFFFFF
        self._rng_step = 0

    def set_dropout_seed(self, seed, step=0):
        self._rng_seed, self._rng_step = int(seed), int(step)

    def _dropout_cfg(self, p_out):
        if not self.training:
            return E.DropoutCfg(enabled=False)
        self._rng_step += 1
        seed = (self._rng_seed * 0x9E3779B97F4A7C15 + self._rng_step) & 0xFFFFFFFFFFFFFFFF
        return E.DropoutCfg(enabled=True, seed=seed, p_enc=self.encoder_dropout, p_out=p_out)

    def _run_stack(self, l, src, input_x, drop, transpose):
        flat = [p[n] for p in _layer_param_dicts(self.u2gnn_layers[l]) for n in E.PARAM_NAMES]
        cfg = (l, self.num_self_att_layers, self.attn_axis, drop, transpose, self.precision)
        return _StackFn.apply(src, input_x, cfg, *flat)

    def _transpose_for(self, input_x, n_src):
        if self.deterministic and self.num_U2GNN_layers > 1 and self.attn_axis == "neighbors" and torch.is_grad_enabled():
            return E.IndexTranspose(input_x, n_src)
        return None


class TransformerU2GNN(_U2GNNBase):
    """Supervised U2GNN.  Same constructor and forward as pytorch_U2GNN_Sup.TransformerU2GNN; the
    extra keyword `attn_axis` selects the reference-as-written ("nodes", default) or the intended
    ("neighbors") attention layout."""

    def __init__(self, feature_dim_size, ff_hidden_size, num_classes, num_self_att_layers, dropout,
                 num_U2GNN_layers, attn_axis="nodes", deterministic=True, precision="fp32", tie_timesteps=False):
        super().__init__()
        self.feature_dim_size = feature_dim_size
        self.ff_hidden_size = ff_hidden_size
        self.num_classes = num_classes
        self.num_self_att_layers = num_self_att_layers
        self.num_U2GNN_layers = num_U2GNN_layers
        self.u2gnn_layers = nn.ModuleList()
        for _ in range(num_U2GNN_layers):
            enc = TransformerEncoderLayer(d_model=feature_dim_size, nhead=1, dim_feedforward=ff_hidden_size, dropout=0.5)
            self.u2gnn_layers.append(_tie(TransformerEncoder(enc, num_self_att_layers), tie_timesteps))
        self.predictions = nn.ModuleList()
        self.dropouts = nn.ModuleList()
        for _ in range(num_U2GNN_layers):
            self.predictions.append(nn.Linear(feature_dim_size, num_classes))
            self.dropouts.append(nn.Dropout(dropout))
        self._init_engine(attn_axis, deterministic, precision)

    def forward(self, input_x, graph_pool, X_concat):
        require_device()
        rowptr = graph_pool if graph_pool.dtype == torch.int64 and not graph_pool.is_sparse else E.rowptr_from_graph_pool(graph_pool)
        input_x = input_x.contiguous()
        drop = self._dropout_cfg(self.dropouts[0].p)
        transpose = self._transpose_for(input_x, X_concat.shape[0])
        src = X_concat.contiguous()
        scores = None
        for l in range(self.num_U2GNN_layers):
            out = self._run_stack(l, src, input_x, drop, transpose)
            ge = _PoolFn.apply(out, rowptr)
            s = _HeadFn.apply(ge, self.predictions[l].weight, self.predictions[l].bias, drop.seed, E.STREAM_POOLED + l,
                              dropout_thr(drop, self.dropouts[l].p))
            scores = s if scores is None else scores + s
            src = out
        return scores


def dropout_thr(drop, p):
    return E.dropout_threshold(p) if drop.enabled else 0


def label_smoothing(true_labels: torch.Tensor, classes: int, smoothing=0.1):
    """pytorch_U2GNN_Sup.py:48-59 (host-side helper kept for drop-in scripts; the fused trainer
    computes the smoothed targets inside the loss kernel)."""
    assert 0 <= smoothing < 1
    with torch.no_grad():
        dist = torch.full((true_labels.size(0), classes), smoothing / (classes - 1), device=true_labels.device)
        dist.scatter_(1, true_labels.data.unsqueeze(1), 1.0 - smoothing)
    return dist


# --------------------------------------------------------------------------------------
# sampler + sampled softmax
# --------------------------------------------------------------------------------------
class LogUniformSampler:
    """Device-side replacement of the Cython-wrapped C++ sampler (log_uniform.pyx:16-40).  The engine
    state is the reference's minstd_rand0 seeded with 1111 (Log_Uniform_Sampler.cpp:10); draws are
    recomputed in parallel on the GPU and match the reference's id sets and try counts."""

    def __init__(self, N, device=None):
        require_device()
        self.N = int(N)
        self.device = torch.device("cuda") if device is None else torch.device(device)
        self.state = torch.tensor([1111], dtype=torch.int32, device=self.device)   # minstd_rand0 state
        self.tries = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._ws = None

    def sample_device(self, size):
        """-> ids[size] int64 on the device (first-occurrence order); self.tries holds num_tries."""
        size = int(size)
        if size > self.N:
            raise ValueError("size > N: the reference sampler would never terminate")
        wb = LIB.call("u2gnn_logu_sample_workspace_bytes", size)
        if self._ws is None or self._ws.numel() < wb:
            self._ws = torch.empty(wb, dtype=torch.uint8, device=self.device)
        ids = torch.empty(size, dtype=torch.int64, device=self.device)
        LIB.call("u2gnn_logu_sample", self.N, size, self.state.data_ptr(), ids.data_ptr(), self.tries.data_ptr(),
                 self._ws.data_ptr(), self._ws.numel(), E._stream())
        return ids

    def expected_count_device(self, ids):
        ids = ids.contiguous()
        out = torch.empty(ids.numel(), dtype=torch.float32, device=self.device)
        LIB.call("u2gnn_logu_expected_count", self.N, self.tries.data_ptr(), ids.data_ptr(), ids.numel(), out.data_ptr(), E._stream())
        return out

    # ---- reference-compatible host API
    def sample(self, size, labels):
        ids = self.sample_device(size)
        lab = torch.as_tensor(np.asarray(labels), dtype=torch.int64, device=self.device)
        true_freq = self.expected_count_device(lab)
        sample_freq = self.expected_count_device(ids)
        return ids.tolist(), true_freq.tolist(), sample_freq.tolist()

    def sample_unique(self, size, labels):
        """log_uniform.pyx:25-27 -> Log_Uniform_Sampler.cpp:73-88: `size` distinct ids none of which is a label."""
        size = int(size)
        lab = torch.as_tensor(np.asarray(list(labels)), dtype=torch.int64, device=self.device).contiguous()
        if size + lab.numel() > self.N:
            raise ValueError("size + len(labels) > N: the reference sampler may never terminate")
        wb = LIB.call("u2gnn_logu_sample_unique_workspace_bytes", self.N, size)
        ws = torch.empty(wb, dtype=torch.uint8, device=self.device)
        ids = torch.empty(size, dtype=torch.int64, device=self.device)
        LIB.call("u2gnn_logu_sample_unique", self.N, size, lab.data_ptr(), lab.numel(), self.state.data_ptr(), ids.data_ptr(),
                 ws.data_ptr(), wb, E._stream())
        return ids.tolist()

    def probability(self, idx):
        return float(np.float32((math.log(idx + 2) - math.log(idx + 1)) / math.log(self.N + 1)))

    def accidental_match(self, labels, samples):
        pos = {int(v): i for i, v in enumerate(samples)}
        return [(i, pos[int(v)]) for i, v in enumerate(labels) if int(v) in pos]


class _SampledSoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, labels, ids):
        x, Wd = x.detach().contiguous(), W.detach()
        N, D = x.shape
        loss = torch.empty(N, dtype=torch.float32, device=x.device)
        denom = torch.empty(N, dtype=torch.float32, device=x.device)
        LIB.call("u2gnn_sampled_softmax_fwd", x.data_ptr(), labels.data_ptr(), N, D, Wd.data_ptr(), W.shape[0],
                 ids.data_ptr(), ids.numel(), 0, loss.data_ptr(), denom.data_ptr(), E.err_word(x.device).data_ptr(), E._stream())
        if E.DEBUG_CHECKS:
            E.check_device_errors(x.device)
        ctx.args = (x, Wd, labels, ids, denom)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        x, W, labels, ids, denom = ctx.args
        N, D = x.shape
        dloss = dloss.contiguous()
        dx = torch.empty_like(x)
        dW = torch.zeros_like(W)                      # dense, like the reference's W.grad
        LIB.call("u2gnn_sampled_softmax_bwd", dloss.data_ptr(), x.data_ptr(), labels.data_ptr(), N, D, W.data_ptr(),
                 W.shape[0], ids.data_ptr(), ids.numel(), 0, denom.data_ptr(), dx.data_ptr(), dW.data_ptr(), 0,
                 E.err_word(x.device).data_ptr(), E._stream())
        return dx, dW, None, None


class _SampledSoftmaxTFFn(torch.autograd.Function):
    """tf.nn.sampled_softmax_loss semantics (bias, log-Q correction, accidental hits removed, label inside the softmax)."""

    @staticmethod
    def forward(ctx, x, W, b, labels, ids, true_q, samp_q):
        x, Wd, bd = x.detach().contiguous(), W.detach(), b.detach()
        N, D = x.shape
        loss = torch.empty(N, dtype=torch.float32, device=x.device)
        denom = torch.empty(N, dtype=torch.float32, device=x.device)
        LIB.call("u2gnn_sampled_softmax_tf_fwd", x.data_ptr(), labels.data_ptr(), N, D, Wd.data_ptr(), bd.data_ptr(), W.shape[0],
                 ids.data_ptr(), ids.numel(), true_q.data_ptr(), samp_q.data_ptr(), loss.data_ptr(), denom.data_ptr(),
                 E.err_word(x.device).data_ptr(), E._stream())
        ctx.args = (x, Wd, bd, labels, ids, true_q, samp_q, denom)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        x, W, b, labels, ids, true_q, samp_q, denom = ctx.args
        N, D = x.shape
        dloss = dloss.contiguous()
        dx = torch.empty_like(x)
        dW = torch.zeros_like(W)
        db = torch.zeros_like(b)
        LIB.call("u2gnn_sampled_softmax_tf_bwd", dloss.data_ptr(), x.data_ptr(), labels.data_ptr(), N, D, W.data_ptr(), b.data_ptr(),
                 W.shape[0], ids.data_ptr(), ids.numel(), true_q.data_ptr(), samp_q.data_ptr(), denom.data_ptr(), dx.data_ptr(),
                 dW.data_ptr(), db.data_ptr(), E.err_word(x.device).data_ptr(), E._stream())
        return dx, dW, db, None, None, None, None


def sampled_softmax_tf(inputs, weight, bias, labels, ids, true_q, samp_q):
    """Per-node loss of the TF model (U2GNN_tf/model_U2GNN_Unsup_multi.py:54-58) on the device; `true_q` / `samp_q` are the
    expected counts from LogUniformSampler.expected_count_device."""
    f32 = lambda t: t.to(torch.float32).contiguous()
    return _SampledSoftmaxTFFn.apply(inputs, weight, bias, labels.contiguous(), ids.contiguous(), f32(true_q), f32(samp_q))


class SampledSoftmax(nn.Module):
    """sampled_softmax.py:11-56.  forward() draws the negatives on the device (no D2H sync);
    sampled() takes injected `sample_values = (ids, true_freq, sample_freq)` like the reference."""

    def __init__(self, ntokens, nsampled, nhid, device):
        super().__init__()
        self.ntokens, self.nsampled, self.device = ntokens, nsampled, device
        self.weight = nn.Parameter(torch.empty(ntokens, nhid))
        self.reset_parameters()
        self.sampler = None

    def reset_parameters(self):
        stdv = math.sqrt(6.0 / (self.weight.size(0) + self.weight.size(1)))
        self.weight.data.uniform_(-stdv, stdv)

    def forward(self, inputs, labels):
        if self.sampler is None:
            self.sampler = LogUniformSampler(self.ntokens, inputs.device)
        ids = self.sampler.sample_device(self.nsampled)
        return _SampledSoftmaxFn.apply(inputs, self.weight, labels.contiguous(), ids)

    def sampled(self, inputs, labels, sample_values):
        sample_ids = sample_values[0] if isinstance(sample_values, (tuple, list)) and len(sample_values) == 3 and \
            not np.isscalar(sample_values[0]) else sample_values
        ids = torch.as_tensor(np.asarray(sample_ids) if not torch.is_tensor(sample_ids) else sample_ids,
                              dtype=torch.int64).to(inputs.device).contiguous()
        return _SampledSoftmaxFn.apply(inputs, self.weight, labels.contiguous(), ids)


class TransformerU2GNNUnSup(_U2GNNBase):
    """Unsupervised U2GNN with the reference's constructor/forward signature
    (pytorch_U2GNN_UnSup.py:14-15, train_pytorch_U2GNN_UnSup.py:140-143,155) and `.ss.weight`.
    forward(X_concat, input_x, input_y) -> per-node sampled-softmax loss [N]."""

    def __init__(self, vocab_size, feature_dim_size, ff_hidden_size, sampled_num, num_self_att_layers,
                 num_U2GNN_layers, dropout, device, sampler_type="default", loss_type="default", adj_mat=None,
                 single_layer_only=True, attn_axis="nodes", deterministic=True, precision="fp32", tie_timesteps=False):
        super().__init__()
        if sampler_type != "default" or loss_type != "default":
            raise NotImplementedError("only the default sampler / sampled-softmax loss is part of the hot path")
        self.feature_dim_size = feature_dim_size
        self.ff_hidden_size = ff_hidden_size
        self.num_self_att_layers = num_self_att_layers
        self.num_U2GNN_layers = num_U2GNN_layers
        self.vocab_size = vocab_size
        self.sampled_num = sampled_num
        self.device = device
        self.u2gnn_layers = nn.ModuleList()
        for _ in range(num_U2GNN_layers):
            enc = TransformerEncoderLayer(d_model=feature_dim_size, nhead=1, dim_feedforward=ff_hidden_size, dropout=0.5)
            self.u2gnn_layers.append(_tie(TransformerEncoder(enc, num_self_att_layers), tie_timesteps))
        self.dropouts = nn.Dropout(dropout)
        self.ss = SampledSoftmax(vocab_size, sampled_num, feature_dim_size * num_U2GNN_layers, device)
        self._init_engine(attn_axis, deterministic, precision)

    def encode(self, X_concat, input_x, drop=None):
        drop = drop or E.DropoutCfg(enabled=False)
        input_x = input_x.contiguous()
        transpose = self._transpose_for(input_x, X_concat.shape[0])
        outs, src = [], X_concat.contiguous()
        for l in range(self.num_U2GNN_layers):
            out = self._run_stack(l, src, input_x, drop, transpose)
            outs.append(out)
            src = out
        return outs[0] if len(outs) == 1 else torch.cat(outs, 1)

    def forward(self, X_concat, input_x, input_y, sample_values=None):
        require_device()
        drop = self._dropout_cfg(self.dropouts.p)
        vec = self.encode(X_concat, input_x, drop)
        vec = _DropoutFn.apply(vec, drop.seed, E.STREAM_CONCAT, dropout_thr(drop, self.dropouts.p))
        if sample_values is not None:
            return self.ss.sampled(vec, input_y, sample_values)
        return self.ss(vec, input_y)
