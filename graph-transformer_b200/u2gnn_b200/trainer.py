"""Fused train step: forward + loss + backward + global-norm clip + Adam without autograd.

Replaces the body of `train()` in train_pytorch_U2GNN_Sup.py:149-164 / train_pytorch_U2GNN_UnSup.py:149-162
(zero_grad, model(...), loss, backward, clip_grad_norm_(0.5), Adam.step, loss.item()) by direct calls
into the C-ABI library over ONE flat parameter arena; data-parallel ranks all-reduce the flat
gradient (+ the squared norm) with NCCL before the fused clip+Adam kernel.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import engine as E
from ._lib import LIB
from .model import TransformerU2GNN, TransformerU2GNNUnSup, _layer_param_dicts
from .parallel import all_reduce_sum_


class FlatArena:
    """Moves every parameter of `model` into one contiguous fp32 buffer (parameters become views, so
    state_dict / optimizers keep working) with matching flat grad / Adam-moment buffers."""

    def __init__(self, model):
        params = [(n, p) for n, p in model.named_parameters()]
        dev = params[0][1].device
        offs, total = {}, 0
        for n, p in params:
            offs[n] = total
            total += (p.numel() + 3) // 4 * 4                 # 16-byte aligned slots
        self.total = total
        self.p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.views, self.gviews = {}, {}
        for n, p in params:
            o = offs[n]
            view = self.p[o:o + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            self.views[n] = p
            self.gviews[n] = self.g[o:o + p.numel()].view_as(p)
        self.model = model
        self.step_count = 0

    def zero_grad(self):
        self.g.zero_()

    def grad_dicts(self, params):
        """params[l][t] = {name: Parameter}  ->  the same nesting of gradient views, looked up by parameter identity so that
        timesteps sharing one weight set (tie_timesteps) accumulate into one gradient slot."""
        by_id = {id(p): self.gviews[n] for n, p in self.views.items()}
        return [[{n: by_id[id(p)] for n, p in d.items()} for d in layer] for layer in params]

    def grads_from_autograd(self):
        self.g.zero_()
        for n, p in self.views.items():
            if p.grad is not None:
                self.gviews[n].copy_(p.grad)

    def grads_to_autograd(self):
        for n, p in self.views.items():
            p.grad = self.gviews[n]

    def clip_adam_step(self, lr, max_norm=0.5, betas=(0.9, 0.999), eps=1e-8, all_reduce=False, want_norm=True):
        s = E._stream()
        self.sumsq.zero_()
        if all_reduce and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.g)                           # gradient all-reduce (sum), NCCL over NVLink
        LIB.call("u2gnn_grad_sqnorm", self.g.data_ptr(), self.total, self.sumsq.data_ptr(), s)
        self.step_count += 1
        E._acct_bytes("u2gnn_clip_adam", 28 * self.total)     # read p, g, m, v; write p, m, v (dense over every parameter, the class table included)
        LIB.call("u2gnn_clip_adam", self.p.data_ptr(), self.g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.total,
                 self.sumsq.data_ptr(), max_norm, lr, betas[0], betas[1], eps, self.step_count, s)
        return float(self.sumsq.sqrt().item()) if want_norm else None


def effective_smoothing(eps, classes, style="reference"):
    """Label-smoothing strength for the loss kernel, which implements the reference's targets (pytorch_U2GNN_Sup.py:48-59):
    1 - eps on the label, eps / (C - 1) elsewhere.  style="tf" gives the TF model's targets (tf.losses.softmax_cross_entropy
    label_smoothing, U2GNN_tf/model_U2GNN_Sup_multi.py:73-75): 1 - eps + eps / C on the label, eps / C elsewhere - the SAME
    family with eps' = eps (C - 1) / C, so no second kernel path is needed."""
    if style == "reference":
        return float(eps)
    if style == "tf":
        return float(eps) * (classes - 1) / classes
    raise ValueError("smoothing_style must be 'reference' or 'tf'")


def dominant_kernel(precision, d=64):
    """Entry point whose launches bench.py times for the roofline object: the FFN backward (the path's largest dense
    contraction, tcgen05) in bf16 mode; in fp32 mode the bf16-split tcgen05 rows GEMM (engine.FP32_TC; the CUDA-core SGEMM when
    that is switched off); for 64 < d <= 128 in bf16 mode the FFN runs as general tcgen05 rows GEMMs (engine.ffn_wide_*), which
    are then what is timed."""
    if precision == "fp32":
        return "u2gnn_gemm_split_rows" if E.FP32_TC else "u2gnn_sgemm"
    if d > 64:
        return "u2gnn_gemm_tc_rows_kloop" if E.WIDE_KLOOP else "u2gnn_gemm_tc_rows_ex"
    return "u2gnn_ffn_tc_bwd"


def roofline(model, precision, name, kernel_ms, launches, peaks, flops, ncu_summary=None):
    """bench.py roofline object for the dominant kernel: achieved = algorithmic flops / measured time."""
    peak = peaks.get("bf16_tflops_sustained")
    which = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    if peak is None:
        peak, which = 1590.0, "fallback (B200_PROFILING.md)"
    if name == "u2gnn_gemm_tc_rows_kloop":
        # wide bf16 FFN (64 < d <= 128): its four row products stream the materialised bf16 hidden - HBM-bound like the fp32 mode's
        hbm = peaks.get("hbm_gbs", 6550.0)
        gbs = E.BYTES.get(name, 0) / max(kernel_ms, 1e-9) / 1e6
        return {"bound": "hbm", "kernel": name, "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "traffic": None,
                "launches_timed": launches, "avg_launch_ms": kernel_ms / max(launches, 1),
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)",
                "tensor_tflops_algorithmic": flops.get(name, 0) / max(kernel_ms, 1e-9) / 1e9,
                "note": "bf16 FFN for 64 < d <= 128 on the K-looping tcgen05 rows kernel (one launch per product, fused ReLU / dropout / mask "
                        "epilogues, hidden materialised in bf16); achieved = the bytes every product must move once / CUDA-event time of its launches"}
    if name == "u2gnn_gemm_split_rows":
        # fp32 mode: the bf16-split GEMMs stream the materialised fp32 hidden - HBM is what bounds them
        hbm = peaks.get("hbm_gbs", 6550.0)
        gbs = E.BYTES.get(name, 0) / max(kernel_ms, 1e-9) / 1e6
        return {"bound": "hbm", "kernel": name, "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "traffic": None,
                "launches_timed": launches, "avg_launch_ms": kernel_ms / max(launches, 1),
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)",
                "tensor_tflops_algorithmic": flops.get(name, 0) / max(kernel_ms, 1e-9) / 1e9,
                "note": "fp32 mode on the tensor cores: three-product bf16 split (hi/lo) tcgen05 GEMMs with the [rows, ff] hidden materialised in fp32; "
                        "achieved = the bytes every product must move once (A, result, aux / old result, weights) / CUDA-event time of all its launches"}
    achieved = flops.get(name, 0) / max(kernel_ms, 1e-9) / 1e9
    traffic = None
    if name == "u2gnn_ffn_tc_bwd" and ncu_summary:
        # DRAM bytes per row from the committed `ncu` capture (image + wgrad + dgrad device kernels of this one entry
        # point), scaled to the average rows per timed launch; algorithmic flops per row = 8 d ff
        per_row = sum(v["traffic_bytes_per_launch"] / v["rows"] for v in ncu_summary.values() if isinstance(v, dict) and "rows" in v)
        rows = flops.get(name, 0) / (8.0 * model.feature_dim_size * model.ff_hidden_size) / max(launches, 1)
        traffic = per_row * rows
    return {"bound": "tensor", "kernel": name, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "traffic": traffic, "launches_timed": launches,
            "avg_launch_ms": kernel_ms / max(launches, 1), "peak_source": which,
            "note": ("fp32 mode on the tensor cores: three-product bf16 split (hi/lo) tcgen05 GEMMs, hidden materialised in fp32; achieved counts the algorithmic 2MNK once (the kernel executes 3x that in bf16 MMAs)" if name == "u2gnn_gemm_split_rows"
                     else "fp32 CUDA-core parity path; the tcgen05 path is --precision bf16" if precision == "fp32"
                     else "bf16 FFN for 64 < d <= 128: four of its six GEMMs (linear1, linear2, dH, dy1) through the general tcgen05 rows kernel, hidden materialised in bf16"
                     if name in ("u2gnn_gemm_tc_rows_ex", "u2gnn_gemm_tc_rows_kloop")
                     else "fused bf16 tcgen05 FFN backward (weight-gradient kernel + input-gradient kernel)")}


class SupTrainer:
    """Supervised fused step.  batch = (input_x[N,S] int64, rowptr[G+1] int64, X[N,d] f32, labels[G] int64)."""

    def __init__(self, model: TransformerU2GNN, lr=5e-4, smoothing=0.1, max_norm=0.5, seed=123, precision="fp32",
                 smoothing_style="reference"):
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        smoothing = effective_smoothing(smoothing, model.num_classes, smoothing_style)
        self.precision = precision
        model.precision = precision
        self.model, self.lr, self.smoothing, self.max_norm = model, lr, smoothing, max_norm
        self.arena = FlatArena(model)
        self.seed, self.steps = seed, 0
        self.L, self.T = model.num_U2GNN_layers, model.num_self_att_layers
        self.params = [_layer_param_dicts(model.u2gnn_layers[l]) for l in range(self.L)]
        self.grads = self.arena.grad_dicts(self.params)
        self.loss = torch.zeros(1, dtype=torch.float32, device=self.arena.p.device)

    def _drop(self, train):
        if not train:
            return E.DropoutCfg(enabled=False)
        self.steps += 1
        seed = (self.seed * 0x9E3779B97F4A7C15 + self.steps) & 0xFFFFFFFFFFFFFFFF
        return E.DropoutCfg(enabled=True, seed=seed, p_enc=self.model.encoder_dropout, p_out=self.model.dropouts[0].p)

    def forward_backward(self, input_x, rowptr, X, labels, train=True, G_total=None):
        m, s = self.model, E._stream()
        drop = self._drop(train)
        axis = m.attn_axis
        N, d = X.shape
        G = rowptr.numel() - 1
        C = m.num_classes
        thr_out = drop.thr_out()
        transpose = E.IndexTranspose(input_x, N) if (self.L > 1 and axis == "neighbors" and m.deterministic) else None
        scores = torch.empty((G, C), dtype=torch.float32, device=X.device)
        src, saved, outs, ges = X, [], [], []
        for l in range(self.L):
            pl = [{n: t.data for n, t in p.items()} for p in self.params[l]]
            out, sv = E.u2gnn_layer_fwd(src, input_x, pl, l, self.T, axis, drop, self.precision)
            ge = E.segment_sum(out, rowptr)
            W, b = m.predictions[l].weight.data, m.predictions[l].bias.data
            LIB.call("u2gnn_head_fwd", ge.data_ptr(), G, d, W.data_ptr(), b.data_ptr(), C, drop.seed, E.STREAM_POOLED + l,
                     thr_out, scores.data_ptr(), int(l > 0), s)
            saved.append((sv, pl)); outs.append(out); ges.append(ge)
            src = out
        dscores = torch.empty_like(scores)
        LIB.call("u2gnn_soft_ce_fwd_bwd", scores.data_ptr(), labels.data_ptr(), G, C, self.smoothing,
                 G if G_total is None else G_total, self.loss.data_ptr(), dscores.data_ptr(), s)
        self.arena.zero_grad()
        dsrc_next = None
        for l in reversed(range(self.L)):
            W = m.predictions[l].weight.data
            dge = torch.empty((G, d), dtype=torch.float32, device=X.device)
            LIB.call("u2gnn_head_bwd", dscores.data_ptr(), ges[l].data_ptr(), G, d, W.data_ptr(), C, drop.seed,
                     E.STREAM_POOLED + l, thr_out, self.arena.gviews["predictions.%d.weight" % l].data_ptr(),
                     self.arena.gviews["predictions.%d.bias" % l].data_ptr(), dge.data_ptr(), s)
            if dsrc_next is None:
                dout = E.segment_sum_bwd(dge, rowptr, N)
            else:
                dout = E.segment_sum_bwd(dge, rowptr, N, out=dsrc_next, accumulate=True)
            sv, pl = saved[l]
            dsrc_next = E.u2gnn_layer_bwd(dout, sv, input_x, pl, self.grads[l], l, self.T, axis, drop,
                                          need_dsrc=(l > 0), transpose=transpose)
            saved[l] = None
        return self.loss, scores

    def dominant_kernel(self):
        return dominant_kernel(self.precision, self.model.feature_dim_size)

    def roofline(self, name, kernel_ms, launches, peaks, flops, ncu_summary=None):
        return roofline(self.model, self.precision, name, kernel_ms, launches, peaks, flops, ncu_summary)

    def step(self, input_x, rowptr, X, labels, G_total=None):
        loss, _ = self.forward_backward(input_x, rowptr, X, labels, True, G_total)
        self.arena.clip_adam_step(self.lr, self.max_norm, all_reduce=True, want_norm=False)
        return loss


class UnSupTrainer:
    """Unsupervised fused step.  batch = (X[N,d], input_x[N,S], input_y[N]); negatives are drawn on the
    device by the log-uniform sampler (or injected with `sample_ids`).  loss = sum of per-node losses
    (train_pytorch_U2GNN_UnSup.py:156)."""

    def __init__(self, model: TransformerU2GNNUnSup, lr=5e-3, max_norm=0.5, seed=123, row_shard=None, global_vocab=None):
        """row_shard (parallel.RowShard): data-parallel mode with the class table row-sharded in contiguous blocks
        aligned with graph ownership — `model.ss.weight` then holds only this rank's rows [lo, hi) and `global_vocab`
        is the size of the whole table (the sampler's range)."""
        self.model, self.lr, self.max_norm = model, lr, max_norm
        self.row_shard, self.global_vocab = row_shard, global_vocab
        self.arena = FlatArena(model)
        self.seed, self.steps = seed, 0
        self.L, self.T = model.num_U2GNN_layers, model.num_self_att_layers
        self.params = [_layer_param_dicts(model.u2gnn_layers[l]) for l in range(self.L)]
        self.grads = self.arena.grad_dicts(self.params)
        self.sampler = None

    def step(self, X, input_x, input_y, sample_ids=None, apply=True):
        from .model import LogUniformSampler
        m, s = self.model, E._stream()
        self.steps += 1
        seed = (self.seed * 0x9E3779B97F4A7C15 + self.steps) & 0xFFFFFFFFFFFFFFFF
        drop = E.DropoutCfg(enabled=True, seed=seed, p_enc=m.encoder_dropout, p_out=m.dropouts.p)
        axis = m.attn_axis
        N, d = X.shape
        D = d * self.L
        V = m.ss.weight.shape[0]
        rs = self.row_shard
        if sample_ids is None:
            if self.sampler is None:          # every rank draws the SAME ids (replicated engine state, seed 1111)
                self.sampler = LogUniformSampler(self.global_vocab if rs is not None else V, X.device)
            sample_ids = self.sampler.sample_device(m.sampled_num)
        transpose = E.IndexTranspose(input_x, N) if (self.L > 1 and axis == "neighbors" and m.deterministic) else None
        src, saved = X, []
        cat = torch.empty((N, D), dtype=torch.float32, device=X.device) if self.L > 1 else None
        for l in range(self.L):
            pl = [{n: t.data for n, t in p.items()} for p in self.params[l]]
            out, sv = E.u2gnn_layer_fwd(src, input_x, pl, l, self.T, axis, drop, m.precision)
            saved.append((sv, pl))
            if cat is not None:
                E.copy_rows(out, d, cat, D, N, d, dst_off=l * d)
            else:
                cat = out
            src = out
        thr = drop.thr_out()
        vec = torch.empty_like(cat)
        LIB.call("u2gnn_dropout_apply", cat.data_ptr(), cat.numel(), drop.seed, E.STREAM_CONCAT, thr, vec.data_ptr(), s)
        W = m.ss.weight.data
        ns = sample_ids.numel()
        self.arena.zero_grad()
        gW = self.arena.gviews["ss.weight"]
        err = E.err_word(X.device).data_ptr()
        if rs is None:
            labels, samp, dsamp, samp_ptr, dsamp_ptr = input_y, None, None, 0, 0
        else:
            # row-sharded table: true-class rows are local (graph-aligned blocks, labels re-based to the block; a label outside
            # it sets the device error word instead of reading out of bounds).  The ns sampled rows are assembled on every rank
            # by ONE all-reduce of an [ns, D] buffer (owners fill their rows, the gather zero-fills ids outside the local
            # block) and handed to the loss kernels as a second table; their gradient comes back the same way.  Nothing of
            # size V is copied or zeroed per step besides the gradient arena itself.
            local_ids = (sample_ids - rs.lo).contiguous()
            samp = torch.empty((ns, D), dtype=torch.float32, device=X.device)
            LIB.call("u2gnn_gather_rows", W.data_ptr(), V, D, local_ids.data_ptr(), ns, 1, samp.data_ptr(), 0, s)
            all_reduce_sum_(samp)
            labels = (input_y - rs.lo).contiguous()
            dsamp = torch.zeros((ns, D), dtype=torch.float32, device=X.device)
            samp_ptr, dsamp_ptr = samp.data_ptr(), dsamp.data_ptr()
        node_loss = torch.empty(N, dtype=torch.float32, device=X.device)
        denom = torch.empty(N, dtype=torch.float32, device=X.device)
        LIB.call("u2gnn_sampled_softmax_fwd", vec.data_ptr(), labels.data_ptr(), N, D, W.data_ptr(), V,
                 sample_ids.data_ptr(), ns, samp_ptr, node_loss.data_ptr(), denom.data_ptr(), err, s)
        dloss = torch.ones(N, dtype=torch.float32, device=X.device)
        dvec = torch.empty_like(vec)
        LIB.call("u2gnn_sampled_softmax_bwd", dloss.data_ptr(), vec.data_ptr(), labels.data_ptr(), N, D, W.data_ptr(), V,
                 sample_ids.data_ptr(), ns, samp_ptr, denom.data_ptr(), dvec.data_ptr(), gW.data_ptr(), dsamp_ptr, err, s)
        if rs is not None:
            all_reduce_sum_(dsamp)                                                              # sampled-row gradients of all ranks
            LIB.call("u2gnn_scatter_add_rows", dsamp.data_ptr(), ns, D, local_ids.data_ptr(), 1, gW.data_ptr(), V, s)
        dcat = torch.empty_like(dvec)
        LIB.call("u2gnn_dropout_apply", dvec.data_ptr(), dvec.numel(), drop.seed, E.STREAM_CONCAT, thr, dcat.data_ptr(), s)
        dsrc_next = None
        for l in reversed(range(self.L)):
            if self.L > 1:
                dout = torch.empty((N, d), dtype=torch.float32, device=X.device)
                E.copy_rows(dcat, D, dout, d, N, d, src_off=l * d)
                if dsrc_next is not None:
                    LIB.call("u2gnn_axpy", 1.0, dsrc_next.data_ptr(), dout.data_ptr(), N * d, s)
            else:
                dout = dcat
            sv, pl = saved[l]
            dsrc_next = E.u2gnn_layer_bwd(dout, sv, input_x, pl, self.grads[l], l, self.T, axis, drop,
                                          need_dsrc=(l > 0), transpose=transpose)
            saved[l] = None
        if apply:
            if rs is None:
                self.arena.clip_adam_step(self.lr, self.max_norm, all_reduce=True, want_norm=False)
            else:
                self._sharded_clip_adam()
        return node_loss

    def dominant_kernel(self):
        return dominant_kernel(self.model.precision, self.model.feature_dim_size)

    def roofline(self, name, kernel_ms, launches, peaks, flops, ncu_summary=None):
        return roofline(self.model, self.model.precision, name, kernel_ms, launches, peaks, flops, ncu_summary)

    def _sharded_clip_adam(self):
        """Encoder gradients are all-reduced, table gradients stay local; the clip norm is global:
        |g|^2 = |g_encoder (reduced)|^2 + sum over ranks |g_table_local|^2."""
        a, s = self.arena, E._stream()
        n_table = self.model.ss.weight.numel()
        n_enc = a.total - ((n_table + 3) // 4 * 4)
        all_reduce_sum_(a.g[:n_enc])
        a.sumsq.zero_()
        LIB.call("u2gnn_grad_sqnorm", a.g.data_ptr() + 4 * n_enc, a.total - n_enc, a.sumsq.data_ptr(), s)
        all_reduce_sum_(a.sumsq)
        LIB.call("u2gnn_grad_sqnorm", a.g.data_ptr(), n_enc, a.sumsq.data_ptr(), s)
        a.step_count += 1
        E._acct_bytes("u2gnn_clip_adam", 28 * a.total)
        LIB.call("u2gnn_clip_adam", a.p.data_ptr(), a.g.data_ptr(), a.m.data_ptr(), a.v.data_ptr(), a.total,
                 a.sumsq.data_ptr(), self.max_norm, self.lr, 0.9, 0.999, 1e-8, a.step_count, s)
