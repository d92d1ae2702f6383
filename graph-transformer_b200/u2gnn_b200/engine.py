"""Host-side sequencing of the U2GNN train step over the C-ABI kernels.

Mirrors the dataflow of the reference forward (pytorch_U2GNN_Sup.py:30-46) and of
nn.TransformerEncoderLayer (torch/nn/modules/transformer.py:944-982); every arithmetic step is a
call into libu2gnn_b200.so.  PyTorch is used for device memory and streams only.

Two attention layouts:
  attn_axis="nodes"      reference as written (sequence = the N nodes of the batch, column 0 only)
  attn_axis="neighbors"  intended (sequence = [node, k sampled neighbours]); the last timestep of a
                         U2GNN layer only produces sequence position 0 (dead-row elimination)
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import torch

from ._lib import LIB, require_device

PARAM_NAMES = (
    "self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.out_proj.weight", "self_attn.out_proj.bias",
    "linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias",
    "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias",
)
STREAM_POOLED = 0x40000000
STREAM_CONCAT = 0x50000000
FUSE_OUT_PROJ_LN = os.environ.get("U2GNN_FUSE_EPILOGUES", "1") != "0"   # bf16 mode, d = 64: out_proj + dropout + residual + LayerNorm1 in one kernel (False: GEMM then LayerNorm kernel)
FUSE_LN_BWD = os.environ.get("U2GNN_FUSE_EPILOGUES", "1") != "0"        # bf16 mode: bf16 da out of the LayerNorm1 backward, linear2 bias gradient inside the LayerNorm2 backward
LAST_STEP_BF16 = os.environ.get("U2GNN_LAST_BF16", "1") != "0"            # bf16 mode, d = 64: bf16 qkv / dqkv at the dead-row-eliminated last timestep
FUSE_PROJ_BWD = os.environ.get("U2GNN_FUSE_PROJ_BWD", "1") != "0"         # bf16 mode, d = 64: projection input + weight gradients in one pass over the output gradient
FUSE_INPROJ_ATTN = os.environ.get("U2GNN_FUSE_INPROJ_ATTN", "1") != "0"   # bf16 mode, d = 64: in_proj inside the attention-forward kernel (qkv written once, never re-read in the forward)
FFN_FWD_MASK = os.environ.get("U2GNN_FFN_FWD_MASK", "1") != "0"          # bf16 mode: the FFN forward leaves the 1-bit live-and-kept mask of the hidden for the backward
FFN_BWD_IMAGES = os.environ.get("U2GNN_FFN_BWD_IMAGES", "1") != "0"      # bf16 mode, d = 64: y1 / dF leave their producers as bf16 tile images (no conversion pass in the FFN backward)
FUSE_GATHER = os.environ.get("U2GNN_FUSE_GATHER", "1") != "0"             # bf16 mode, d = 64: the first timestep's consumers gather their x rows by index (the [N, k+1, d] tensor is never written)
FP32_TC = os.environ.get("U2GNN_FP32_TC", "1") != "0"                     # fp32 mode: linear layers as three-product bf16-split tcgen05 GEMMs (csrc/gemm_split.cu) instead of the CUDA-core SGEMM
FLOPS = {}    # entry point -> algorithmic flops issued while LIB.timed is active (bench.py roofline)
BYTES = {}    # entry point -> algorithmic HBM bytes (the tensors the call must read + write once) while LIB.timed is active


def _acct_bytes(name, nbytes):
    if LIB.timed is not None:
        BYTES[name] = BYTES.get(name, 0) + int(nbytes)


def stream_id(layer, timestep, site, num_timesteps):
    """Dropout stream of encoder site (0 probs, 1 post-attention, 2 post-ReLU, 3 post-FFN).  Encoder streams live in
    [16, 0x40000000); the pooled-embedding (STREAM_POOLED + layer) and concatenated-vector (STREAM_CONCAT) streams sit
    above that range, so no (layer, timestep) count can make two sites share a stream."""
    sid = ((layer * num_timesteps + timestep) * 4 + site) + 16
    if sid >= STREAM_POOLED:
        raise ValueError("too many encoder layers x timesteps for the dropout stream layout")
    return sid


def dropout_threshold(p):
    thr = int(round(float(p) * 256.0))
    if not 0 <= thr <= 255:
        raise ValueError("dropout probability out of range: %r" % (p,))
    return thr


@dataclass
class DropoutCfg:
    """enabled=False is model.eval().  p is quantised to thr/256 (0.5 -> 128 exactly)."""
    enabled: bool = False
    seed: int = 0
    p_enc: float = 0.5
    p_out: float = 0.5

    def thr_enc(self):
        return dropout_threshold(self.p_enc) if self.enabled else 0

    def thr_out(self):
        return dropout_threshold(self.p_out) if self.enabled else 0


def _ptr(t):
    return 0 if t is None else t.data_ptr()


_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_RAW_DEVICE = getattr(torch._C, "_cuda_getDevice", None)


def _stream():
    """cudaStream_t of torch's current stream (the raw accessors avoid building a Stream object and the lazy-init check per C-ABI
    call: 0.4 us instead of 1.8 us, ~80 calls per step of the launch-bound small configurations)."""
    if _RAW_STREAM is not None and _RAW_DEVICE is not None:
        return _RAW_STREAM(_RAW_DEVICE())
    return torch.cuda.current_stream().cuda_stream


def _check(t, dtype, name):
    if t.device.type != "cuda":
        raise RuntimeError("%s must be a CUDA tensor (no CPU fallback in u2gnn_b200)" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


# --------------------------------------------------------------------------------------
# thin wrappers (tensor -> pointer)
# --------------------------------------------------------------------------------------
_ERR = {}          # device index -> int32 error word the kernels OR into (bit 0: label out of range, bit 1: gather index out of range)
ERR_MESSAGES = {1: "sampled softmax: a label lies outside [0, vocab) (the reference's index_select raises, sampled_softmax.py:45)",
                2: "row gather: an index lies outside the table (the reference's F.embedding raises, pytorch_U2GNN_Sup.py:32)"}
DEBUG_CHECKS = os.environ.get("U2GNN_DEBUG_CHECKS", "0") != "0"     # check the error word (one device sync) after every model forward


def err_word(device=None):
    """The device error word of `device` (allocated on first use).  Kernels set bits instead of reading out of bounds."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    w = _ERR.get(key)
    if w is None:
        w = _ERR[key] = torch.zeros(1, dtype=torch.int32, device=dev)
    return w


def check_device_errors(device=None):
    """Raises IndexError if any kernel since the last check met an out-of-range index (synchronises; call it where the host
    already waits for the device, e.g. next to loss.item())."""
    w = err_word(device)
    bits = int(w.item())
    if bits:
        w.zero_()
        raise IndexError("; ".join(m for b, m in ERR_MESSAGES.items() if bits & b))


def gather_rows(table, idx, idx_stride=1, n_idx=None):
    _check(table, torch.float32, "table"); _check(idx, torch.int64, "idx")
    n = idx.numel() // idx_stride if n_idx is None else n_idx
    out = torch.empty((n, table.shape[1]), dtype=torch.float32, device=table.device)
    # HBM-minimal traffic: every index and output row once, the table once (a row gathered k+1 times is re-read from L2)
    _acct_bytes("u2gnn_gather_rows", n * (8 + 4 * table.shape[1]) + min(n, table.shape[0]) * 4 * table.shape[1])
    LIB.call("u2gnn_gather_rows", _ptr(table), table.shape[0], table.shape[1], _ptr(idx), n, idx_stride, _ptr(out),
             _ptr(err_word(table.device)), _stream())
    return out


def scatter_add_rows(grad, idx, n_dst, idx_stride=1, out=None):
    d = grad.shape[-1]
    n = grad.numel() // d
    if out is None:
        out = torch.zeros((n_dst, d), dtype=torch.float32, device=grad.device)
    LIB.call("u2gnn_scatter_add_rows", _ptr(grad), n, d, _ptr(idx), idx_stride, _ptr(out), n_dst, _stream())
    return out


class IndexTranspose:
    """CSR transpose of an index list, built once per batch, for the deterministic scatter-add."""

    def __init__(self, idx, n_dst, idx_stride=1):
        n = idx.numel() // idx_stride
        dev = idx.device
        self.n_dst = n_dst
        self.rowptr = torch.empty(n_dst + 1, dtype=torch.int64, device=dev)
        self.pos = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        wb = LIB.call("u2gnn_index_transpose_workspace_bytes", n, n_dst)
        ws = torch.empty(wb, dtype=torch.uint8, device=dev)
        LIB.call("u2gnn_index_transpose_build", _ptr(idx), n, idx_stride, n_dst, _ptr(self.rowptr), _ptr(self.pos),
                 _ptr(ws), wb, _stream())

    def scatter_add(self, grad, out=None, accumulate=False):
        d = grad.shape[-1]
        if out is None:
            out = torch.empty((self.n_dst, d), dtype=torch.float32, device=grad.device)
            accumulate = False
        LIB.call("u2gnn_scatter_add_rows_det", _ptr(grad), d, _ptr(self.rowptr), _ptr(self.pos), _ptr(out), self.n_dst,
                 int(accumulate), _stream())
        return out


def rowptr_from_graph_pool(graph_pool):
    """CSR row pointer of the reference's COO pooling operator (train_pytorch_U2GNN_Sup.py:73-89)."""
    if graph_pool.layout != torch.sparse_coo:
        raise TypeError("graph_pool must be the reference's sparse COO pooling operator")
    idx = graph_pool._indices()
    G, N = graph_pool.shape
    if idx.shape[1] != N:
        raise ValueError("graph_pool must have exactly one entry per node")
    rows = idx[0].contiguous()
    rowptr = torch.empty(G + 1, dtype=torch.int64, device=rows.device)
    LIB.call("u2gnn_rowptr_from_coo", _ptr(rows), N, G, _ptr(rowptr), _stream())
    return rowptr


def segment_sum(x, rowptr):
    G = rowptr.numel() - 1
    out = torch.empty((G, x.shape[1]), dtype=torch.float32, device=x.device)
    _acct_bytes("u2gnn_segment_sum", 4 * x.shape[1] * (x.shape[0] + G) + 8 * (G + 1))
    LIB.call("u2gnn_segment_sum", _ptr(x), x.shape[0], x.shape[1], _ptr(rowptr), G, _ptr(out), _stream())
    return out


def segment_sum_bwd(gout, rowptr, n, out=None, accumulate=False):
    G, d = gout.shape
    if out is None:
        out = torch.empty((n, d), dtype=torch.float32, device=gout.device)
        accumulate = False
    LIB.call("u2gnn_segment_sum_bwd", _ptr(gout), G, d, _ptr(rowptr), _ptr(out), n, int(accumulate), _stream())
    return out


def sgemm(ta, tb, M, N, K, A, lda, B, ldb, C, ldc, alpha=1.0, beta=0.0, bias=None, relu=False, drop=None,
          aux=None, ldaux=0, aux_scale=1.0, splitk=1, a_off=0, b_off=0, c_off=0):
    """C = epi(alpha*op(A)op(B)+bias) (+beta*C).  *_off are element offsets into the tensors."""
    epi = 0
    seed = stream = thr = row0 = 0
    if bias is not None:
        epi |= 1
    if relu:
        epi |= 2
    if drop is not None and drop[2] > 0:
        epi |= 4
        seed, stream, thr = drop[0], drop[1], drop[2]
        row0 = drop[3] if len(drop) > 3 else 0
    if aux is not None:
        epi |= 8
    if splitk > 1 or splitk == -1:
        epi |= 16
        splitk = max(splitk, 1)
    if LIB.timed is not None:
        FLOPS["u2gnn_sgemm"] = FLOPS.get("u2gnn_sgemm", 0) + 2 * M * N * K
    LIB.call("u2gnn_sgemm", int(ta), int(tb), M, N, K, alpha, _ptr(A) + 4 * a_off, lda, _ptr(B) + 4 * b_off, ldb, beta,
             _ptr(C) + 4 * c_off, ldc, _ptr(bias), epi, seed, stream, thr, row0, _ptr(aux), ldaux, aux_scale, splitk, _stream())
    return C


def _splitk_for(rows):
    return int(max(1, min(296, rows // 2048)))


def wgrad(dout, M_rows, n_out, inp, n_in, dW, db=None):
    """dW[n_out, n_in] += dout[M, n_out]^T @ inp[M, n_in];  db[n_out] += colsum(dout)."""
    if M_rows == 0:
        return
    sk = _splitk_for(M_rows)
    sgemm(1, 0, n_out, n_in, M_rows, dout, n_out, inp, n_in, dW, n_in, splitk=-1 if sk == 1 else sk)
    if db is not None:
        LIB.call("u2gnn_colsum", _ptr(dout), M_rows, n_out, n_out, _ptr(db), 1, _stream())


def linear_fp32(A, M, K, W, w_kn, N, C, bias=None, relu=False, drop=None, aux=None, aux_scale=1.0, beta=0.0):
    """fp32 linear layer C[M,N] = epi(A[M,K] op(W) + bias) (+ beta C); W is [N,K] (w_kn=0, F.linear's layout) or [K,N] (w_kn=1, the
    input-gradient product).  FP32_TC: tensor cores with the three-product bf16 split (fp32 accuracy); else the CUDA-core SGEMM."""
    if not FP32_TC:
        return sgemm(0, 1 - w_kn, M, N, K, A, K, W, N if w_kn else K, C, N, bias=bias, relu=relu, drop=drop, aux=aux,
                     ldaux=N if aux is not None else 0, aux_scale=aux_scale, beta=beta)
    epi, seed, stream, thr = 0, 0, 0, 0
    if bias is not None:
        epi |= 1
    if relu:
        epi |= 2
    if drop is not None and drop[2] > 0:
        epi |= 4
        seed, stream, thr = drop[0], drop[1], drop[2]
    if aux is not None:
        epi |= 8
    if LIB.timed is not None:
        FLOPS["u2gnn_gemm_split_rows"] = FLOPS.get("u2gnn_gemm_split_rows", 0) + 2 * M * N * K
        # the bytes the product must move once: A, the result (twice with beta), the aux mask, the weights
        _acct_bytes("u2gnn_gemm_split_rows", 4 * (M * K + M * N * (1 + (1 if beta else 0) + (1 if aux is not None else 0)) + N * K))
    pk = _pack_w(W, w_kn, N if w_kn else K, N, K, M)
    LIB.call("u2gnn_gemm_split_rows", _ptr(A), M, K, K, _ptr(W), int(w_kn), N if w_kn else K, N, _ptr(bias), epi, seed, stream, thr, 0,
             _ptr(aux), N if aux is not None else 0, aux_scale, beta, _ptr(C), N, _ptr(pk), _stream())
    return C


def wgrad_fp32(dout, M_rows, n_out, inp, n_in, dW, db=None):
    """dW[n_out, n_in] += dout[M, n_out]^T @ inp[M, n_in];  db[n_out] += colsum(dout)  (fp32 mode)."""
    if not FP32_TC:
        return wgrad(dout, M_rows, n_out, inp, n_in, dW, db)
    if M_rows == 0:
        return
    if LIB.timed is not None:
        FLOPS["u2gnn_gemm_split_wgrad"] = FLOPS.get("u2gnn_gemm_split_wgrad", 0) + 2 * M_rows * n_out * n_in
    if n_in > n_out:
        # the wide side streams as the 128-column operand (fewer passes over the rows); the destination is then transposed
        LIB.call("u2gnn_gemm_split_wgrad", _ptr(inp), M_rows, n_in, n_in, _ptr(dout), n_out, n_out, _ptr(dW), 1, n_in, 0, _stream())
        if db is not None:
            LIB.call("u2gnn_colsum", _ptr(dout), M_rows, n_out, n_out, _ptr(db), 1, _stream())
    else:
        LIB.call("u2gnn_gemm_split_wgrad", _ptr(dout), M_rows, n_out, n_out, _ptr(inp), n_in, n_in, _ptr(dW), n_in, 1, _ptr(db), _stream())


def linear_tc(A, M, K, W, w_kn, N, bias=None, beta=0.0, out=None, out_bf16=False):
    """bf16 tensor-core projection: out[M,N] = A[M,K] W^T (+bias) (+beta*out).  A may be fp32 or bf16 (the dtype the
    producer stored); out_bf16 stores the result rounded to bf16 (every consumer rounds it anyway)."""
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=A.device)
    if LIB.timed is not None:
        FLOPS["u2gnn_gemm_tc_rows"] = FLOPS.get("u2gnn_gemm_tc_rows", 0) + 2 * M * N * K
    LIB.call("u2gnn_gemm_tc_rows_ex", _ptr(A), int(A.dtype == torch.bfloat16), M, K, K, _ptr(W), int(w_kn), N, _ptr(bias), beta,
             _ptr(out), int(out.dtype == torch.bfloat16), N, _stream())
    return out


def wgrad_tc(dout, M, n_out, inp, n_in, dW, db=None, inp_idx=None):
    """dW[n_out, n_in] += dout^T @ inp; db[n_out] += colsum(dout)  (tensor cores, bf16 operands).  inp_idx: row r of the input is
    inp[inp_idx[r]] (fused gather)."""
    if M == 0:
        return
    LIB.call("u2gnn_gemm_tc_wgrad_ex", _ptr(dout), int(dout.dtype == torch.bfloat16), M, n_out, n_out, _ptr(inp),
             int(inp.dtype == torch.bfloat16), n_in, n_in, _ptr(inp_idx), inp.shape[0] if inp_idx is not None else 0, _ptr(dW), _ptr(db),
             _stream())


def proj_bwd_tc(dout, M, n_out, inp, W, dW, db, out=None, out_bf16=False, beta=0.0, inp_idx=None):
    """Backward of a projection y = inp W^T + b in one pass over dout[M, n_out] (bf16): returns dinp[M, 64] = dout W (+ beta*out)
    and accumulates dW[n_out, 64] += dout^T inp, db += colsum(dout)."""
    if out is None:
        out = torch.empty((M, 64), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=dout.device)
    _acct_bytes("u2gnn_gemm_tc_dgrad_wgrad", M * (2 * n_out + (2 if inp.dtype == torch.bfloat16 else 4) * 64 +
                                                  (2 if (out_bf16 or (out is not None and out.dtype == torch.bfloat16)) else 4) * 64 * (2 if beta else 1)))
    if LIB.timed is not None:
        FLOPS["u2gnn_gemm_tc_dgrad_wgrad"] = FLOPS.get("u2gnn_gemm_tc_dgrad_wgrad", 0) + 4 * M * n_out * 64
    LIB.call("u2gnn_gemm_tc_dgrad_wgrad", _ptr(dout), M, n_out, n_out, _ptr(inp), int(inp.dtype == torch.bfloat16), 64, _ptr(inp_idx),
             inp.shape[0] if inp_idx is not None else 0, _ptr(W), _ptr(out), int(out.dtype == torch.bfloat16), 64, beta, _ptr(dW), _ptr(db),
             _stream())
    return out


def add_dropout_ln_fwd(res, a, M, d, drop, gamma, beta):
    dev = a.device
    z = torch.empty((M, d), dtype=torch.float32, device=dev)
    y = torch.empty((M, d), dtype=torch.float32, device=dev)
    stats = torch.empty((M, 2), dtype=torch.float32, device=dev)
    LIB.call("u2gnn_add_dropout_ln_fwd", _ptr(res), _ptr(a), M, d, drop[0], drop[1], drop[2], _ptr(gamma), _ptr(beta),
             _ptr(z), _ptr(y), _ptr(stats), _stream())
    return z, y, stats


def tile_images(M, device):
    """Buffer for bf16 swizzled [128 x 64] tile images of an [M, 64] operand (the FFN backward's operand format), padded to whole
    pairs of tiles; the rows past M (never written by the producers) are zeroed."""
    nbytes = LIB.call("u2gnn_ffn_tc_image_bytes", M)
    img = torch.empty(nbytes, dtype=torch.uint8, device=device)
    tail = (M // 128) * 16384
    if tail < nbytes:
        img[tail:].zero_()
    return img


def add_dropout_ln_bwd(dy, z, stats, M, d, gamma, drop, dgamma, dbeta, want_da=True, da_bf16=False, dasum=None, da_img=False):
    """da_bf16: store the dropout-masked gradient as bf16 (only when its consumers are tensor-core kernels).
    dasum[d] += colsum(masked gradient): the bias gradient of the layer in front of the dropout, folded into this pass.
    da_img (d == 64): the masked gradient is written as bf16 tile images for the FFN backward (returned as a uint8 buffer)."""
    dz = torch.empty((M, d), dtype=torch.float32, device=dy.device)
    if da_img:
        img = tile_images(M, dy.device)
        _acct_bytes("u2gnn_add_dropout_ln_bwd_ex", M * (4 * d * 3 + 8 + 2 * d))
        LIB.call("u2gnn_add_dropout_ln_bwd_ex", _ptr(dy), _ptr(z), _ptr(stats), M, d, _ptr(gamma), drop[0], drop[1], drop[2],
                 _ptr(dz), _ptr(img), 2, _ptr(dgamma), _ptr(dbeta), _ptr(dasum), _stream())
        return dz, img
    has_da = want_da and drop[2] > 0
    da = torch.empty((M, d), dtype=torch.bfloat16 if da_bf16 else torch.float32, device=dy.device) if has_da else None
    _acct_bytes("u2gnn_add_dropout_ln_bwd_ex", M * (4 * d * 3 + 8 + ((2 if da_bf16 else 4) * d if has_da else 0)))   # dy, z, stats -> dz (+ da)
    if (da_bf16 and has_da) or dasum is not None:
        LIB.call("u2gnn_add_dropout_ln_bwd_ex", _ptr(dy), _ptr(z), _ptr(stats), M, d, _ptr(gamma), drop[0], drop[1], drop[2],
                 _ptr(dz), _ptr(da), int(da_bf16 and has_da), _ptr(dgamma), _ptr(dbeta), _ptr(dasum), _stream())
    else:
        LIB.call("u2gnn_add_dropout_ln_bwd", _ptr(dy), _ptr(z), _ptr(stats), M, d, _ptr(gamma), drop[0], drop[1], drop[2],
                 _ptr(dz), _ptr(da), _ptr(dgamma), _ptr(dbeta), _stream())
    return dz, (da if da is not None else dz)


def out_proj_ln_tc(ctx, Mq, d, p, res, ldres, drop, want_img=False, res_idx=None):
    """out_proj + dropout + residual + LayerNorm1 as ONE kernel (bf16 mode, d = 64): the projection result never reaches HBM.
    want_img: y1 is also written as bf16 tile images, which the FFN backward bulk-copies (no conversion pass there)."""
    dev = ctx.device
    z = torch.empty((Mq, d), dtype=torch.float32, device=dev)
    y = torch.empty((Mq, d), dtype=torch.float32, device=dev)
    stats = torch.empty((Mq, 2), dtype=torch.float32, device=dev)
    img = tile_images(Mq, dev) if want_img else None
    _acct_bytes("u2gnn_gemm_tc_rows_ln", Mq * ((2 if ctx.dtype == torch.bfloat16 else 4) * d + 4 * d * 3 + 8 + (2 * d if want_img else 0)))   # ctx, residual -> z, y, stats (+ y image)
    if LIB.timed is not None:
        FLOPS["u2gnn_gemm_tc_rows_ln"] = FLOPS.get("u2gnn_gemm_tc_rows_ln", 0) + 2 * Mq * d * d
    LIB.call("u2gnn_gemm_tc_rows_ln", _ptr(ctx), int(ctx.dtype == torch.bfloat16), Mq, d, d, _ptr(p["self_attn.out_proj.weight"]), 0,
             _ptr(p["self_attn.out_proj.bias"]), _ptr(res), ldres, _ptr(res_idx), res.shape[0] if res_idx is not None else 0,
             drop[0], drop[1], drop[2], _ptr(p["norm1.weight"]),
             _ptr(p["norm1.bias"]), _ptr(z), _ptr(y), _ptr(stats), _ptr(img), _stream())
    return z, y, stats, img


def copy_rows(src, ld_src, dst, ld_dst, rows, d, accumulate=False, src_off=0, dst_off=0):
    LIB.call("u2gnn_copy_rows", _ptr(src) + 4 * src_off, ld_src, _ptr(dst) + 4 * dst_off, ld_dst, rows, d, int(accumulate), _stream())


# --------------------------------------------------------------------------------------
# one encoder layer (fp32 building-block path)
# --------------------------------------------------------------------------------------
@dataclass
class LayerSaved:
    x: torch.Tensor = None
    xq: torch.Tensor = None
    qkv: torch.Tensor = None
    probs: torch.Tensor = None
    pd: torch.Tensor = None
    ctx: torch.Tensor = None
    z1: torch.Tensor = None
    st1: torch.Tensor = None
    y1: torch.Tensor = None
    hd: torch.Tensor = None
    z2: torch.Tensor = None
    st2: torch.Tensor = None
    B: int = 0
    S: int = 0
    Sq: int = 0
    packed: torch.Tensor = None
    y1_img: torch.Tensor = None
    ffn_mask: torch.Tensor = None
    wide: bool = False
    attn_pad: tuple = None
    x_idx: torch.Tensor = None     # fused gather: x is the TABLE and row r of the layer input is x[x_idx[r]]


ATTN_PAD = os.environ.get("U2GNN_ATTN_PAD", "1") != "0"     # fp32 attention block for 64 < d <= 128: zero-padded q | k | v blocks (16-byte aligned rows) on the thread-per-row kernels


def attn_pad_size(d):
    """Padded feature size of the q | k | v blocks for 64 < d <= 128 (configs[2]: d = 65 -> 68), or 0.  Rows of 65 floats are not
    16-byte aligned, which leaves only the warp-per-node attention kernels (4-5x slower than the thread-per-row ones)."""
    if not ATTN_PAD or not 64 < d <= 128:
        return 0
    return next(D for D in (68, 80, 96, 112, 128) if d <= D)


def _attn_pad_weights(p, d, DP):
    """in_proj / out_proj weights in the padded layout: W_in [3 DP, d] (rows d..DP of each block zero, the q block times
    sqrt(DP / d): the kernels scale q by sqrt(1 / DP)), b_in [3 DP], W_out [d, DP] (columns d..DP zero).  Weights only: plumbing."""
    W, b, Wo = p["self_attn.in_proj_weight"], p["self_attn.in_proj_bias"], p["self_attn.out_proj.weight"]
    c = math.sqrt(DP / d)
    Wp = torch.zeros((3, DP, d), dtype=torch.float32, device=W.device)
    Wp[:, :d].copy_(W.view(3, d, d))
    Wp[0].mul_(c)
    bp = torch.zeros((3, DP), dtype=torch.float32, device=W.device)
    bp[:, :d].copy_(b.view(3, d))
    bp[0].mul_(c)
    Wop = torch.zeros((d, DP), dtype=torch.float32, device=W.device)
    Wop[:, :d].copy_(Wo)
    return Wp.view(3 * DP, d), bp.view(3 * DP), Wop, c


def gather_fusable(d, ff, S, Sq, precision, long_seq):
    """The first timestep can read its input rows by index (no materialised gather) when every consumer of x is one of the
    kernels that take an index array: in_proj + attention forward, out_proj + LayerNorm1 (residual), in_proj backward."""
    return (FUSE_GATHER and precision == "bf16" and not long_seq and d == 64 and Sq == S and S >= 2 and ffn_tc_supported(d, ff)
            and FUSE_INPROJ_ATTN and FUSE_OUT_PROJ_LN and FUSE_LN_BWD)


def encoder_layer_fwd(x, B, S, Sq, p, d, ff, drop_ids, seed, thr, long_seq, precision="fp32", for_backward=True, x_idx=None):
    """x[B*S, d] -> y[B*Sq, d].  p: dict name->tensor.  drop_ids: 4 stream ids.  long_seq selects the
    attn_axis="nodes" formulation (B == 1, scores materialised as [S, S]).  x_idx (only when gather_fusable): x is the
    gather TABLE and the layer's input row r is x[x_idx[r]]."""
    dev = x.device
    M, Mq = B * S, B * Sq
    f32 = dict(dtype=torch.float32, device=dev)
    sv = LayerSaved(x=x, B=B, S=S, Sq=Sq, x_idx=x_idx)
    assert x_idx is None or gather_fusable(d, ff, S, Sq, precision, long_seq)
    tc_proj = precision == "bf16" and not long_seq and d <= 64
    tc_attn = tc_proj and d == 64 and Sq == S and S >= 2        # tensor-core attention core: bf16 qkv / ctx between the kernels
    tc_last = tc_proj and d == 64 and Sq == 1 and S >= 2 and LAST_STEP_BF16     # last timestep: bf16 qkv / dqkv around the position-0 attention
    fused_in = tc_attn and FUSE_INPROJ_ATTN
    if fused_in:
        qkv = torch.empty((M, 3 * d), dtype=torch.bfloat16, device=dev)      # written once by the fused kernel, read by the backward
    elif tc_proj:
        qkv = linear_tc(x, M, d, p["self_attn.in_proj_weight"], 0, 3 * d, bias=p["self_attn.in_proj_bias"], out_bf16=tc_attn or tc_last)
    else:
        DP = attn_pad_size(d) if (not long_seq and Sq == S and S >= 2) else 0
        if DP:
            Wp, bp, Wop, cq = _attn_pad_weights(p, d, DP)
            sv.attn_pad = (DP, cq, Wp, Wop)
            qkv = torch.empty((M, 3 * DP), **f32)
            linear_fp32(x, M, d, Wp, 0, 3 * DP, qkv, bias=bp)
        else:
            qkv = torch.empty((M, 3 * d), **f32)
            linear_fp32(x, M, d, p["self_attn.in_proj_weight"], 0, 3 * d, qkv, bias=p["self_attn.in_proj_bias"])
    da_ = sv.attn_pad[0] if sv.attn_pad else d                   # feature size the attention core sees
    ctx = torch.empty((Mq, da_), dtype=torch.bfloat16 if tc_attn else torch.float32, device=dev)
    if long_seq:
        assert B == 1 and Sq == S
        scores = torch.empty((S, S), **f32)
        sgemm(0, 1, S, S, d, qkv, 3 * d, qkv, 3 * d, scores, S, alpha=math.sqrt(1.0 / d), b_off=d)
        pd = torch.empty((S, S), **f32) if thr > 0 else scores
        LIB.call("u2gnn_softmax_rows_fwd", _ptr(scores), S, S, _ptr(pd), seed, drop_ids[0], thr, _stream())
        sgemm(0, 0, S, d, S, pd, S, qkv, 3 * d, ctx, d, b_off=2 * d)
        sv.probs, sv.pd = scores, pd
    elif fused_in:
        # x (fp32) -> qkv, ctx (bf16); fused gather: the index per row and the table once (a row gathered k+1 times is re-read from L2)
        _acct_bytes("u2gnn_inproj_seqattn_tc_fwd", M * (2 * 3 * d + 2 * d) + (M * 4 * d if x_idx is None else M * 8 + min(M, x.shape[0]) * 4 * d))
        if LIB.timed is not None:
            FLOPS["u2gnn_inproj_seqattn_tc_fwd"] = FLOPS.get("u2gnn_inproj_seqattn_tc_fwd", 0) + 2 * M * 3 * d * d
        LIB.call("u2gnn_inproj_seqattn_tc_fwd", _ptr(x), _ptr(x_idx), x.shape[0] if x_idx is not None else 0, B, S, d,
                 _ptr(p["self_attn.in_proj_weight"]), _ptr(p["self_attn.in_proj_bias"]), seed, drop_ids[0], thr, _ptr(qkv), _ptr(ctx),
                 _ptr(err_word(dev)), _stream())
    elif tc_attn:
        LIB.call("u2gnn_seqattn_tc_fwd_ex", _ptr(qkv), B, S, d, seed, drop_ids[0], thr, _ptr(ctx), 1, _stream())
    elif tc_last:
        LIB.call("u2gnn_seqattn_last_fwd_ex", _ptr(qkv), 1, B, S, d, seed, drop_ids[0], thr, _ptr(ctx), _stream())
    else:
        LIB.call("u2gnn_seqattn_fwd", _ptr(qkv), B, S, Sq, da_, seed, drop_ids[0], thr, _ptr(ctx), _stream())
    if tc_proj and d == 64 and FUSE_OUT_PROJ_LN:
        # residual rows read in place (position 0 of each sequence when only that row is live)
        xq = None
        z1, y1, st1, sv.y1_img = out_proj_ln_tc(ctx, Mq, d, p, x, d if Sq == S else S * d, (seed, drop_ids[1], thr), want_img=FFN_BWD_IMAGES,
                                                res_idx=x_idx)
    else:
        if tc_proj:
            a = linear_tc(ctx, Mq, d, p["self_attn.out_proj.weight"], 0, d, bias=p["self_attn.out_proj.bias"])
        else:
            a = torch.empty((Mq, d), **f32)
            if sv.attn_pad:
                linear_fp32(ctx, Mq, da_, sv.attn_pad[3], 0, d, a, bias=p["self_attn.out_proj.bias"])
            else:
                linear_fp32(ctx, Mq, d, p["self_attn.out_proj.weight"], 0, d, a, bias=p["self_attn.out_proj.bias"])
        if Sq == S:
            xq = x
        else:
            xq = torch.empty((Mq, d), **f32)
            copy_rows(x, S * d, xq, d, B, d)
        z1, y1, st1 = add_dropout_ln_fwd(xq, a, Mq, d, (seed, drop_ids[1], thr), p["norm1.weight"], p["norm1.bias"])
    if precision == "bf16" and ffn_wide_supported(d, ff):
        # 64 < d <= 128 (configs[2]): bf16 FFN on the general tcgen05 GEMMs, hidden materialised in bf16; the attention block of
        # this feature size runs on the fp32 kernels ("bf16 FFN" exactly as the config names it)
        f, hd = ffn_wide_fwd(y1, Mq, d, ff, p, seed, drop_ids[2], thr)
        z2, y2, st2 = add_dropout_ln_fwd(y1, f, Mq, d, (seed, drop_ids[3], thr), p["norm2.weight"], p["norm2.bias"])
        sv.wide = True
    elif precision == "bf16":
        # fused tcgen05 FFN block: the [Mq, ff] hidden never reaches HBM and is recomputed in the backward
        packed = ffn_tc_pack(p, d, ff, thr)
        z2 = torch.empty((Mq, d), **f32)
        y2 = torch.empty((Mq, d), **f32)
        st2 = torch.empty((Mq, 2), **f32)
        if LIB.timed is not None:
            FLOPS["u2gnn_ffn_tc_fwd"] = FLOPS.get("u2gnn_ffn_tc_fwd", 0) + 4 * Mq * d * ff
        if for_backward and FFN_FWD_MASK:
            # one bit per hidden activation (ReLU live AND kept), 256 B per row at ff 2048: the backward kernels load it instead
            # of re-evaluating the dropout stream and the sign of the recomputed hidden
            sv.ffn_mask = torch.empty(LIB.call("u2gnn_ffn_tc_mask_bytes", Mq, ff), dtype=torch.uint8, device=dev)
        LIB.call("u2gnn_ffn_tc_fwd", _ptr(y1), Mq, d, ff, _ptr(packed), seed, drop_ids[2], drop_ids[3], thr,
                 _ptr(p["norm2.weight"]), _ptr(p["norm2.bias"]), _ptr(z2), _ptr(st2), _ptr(y2), _ptr(sv.ffn_mask), _stream())
        hd = None
        sv.packed = packed
    else:
        hd = torch.empty((Mq, ff), **f32)
        linear_fp32(y1, Mq, d, p["linear1.weight"], 0, ff, hd, bias=p["linear1.bias"], relu=True, drop=(seed, drop_ids[2], thr))
        f = torch.empty((Mq, d), **f32)
        linear_fp32(hd, Mq, ff, p["linear2.weight"], 0, d, f, bias=p["linear2.bias"])
        z2, y2, st2 = add_dropout_ln_fwd(y1, f, Mq, d, (seed, drop_ids[3], thr), p["norm2.weight"], p["norm2.bias"])
    sv.xq, sv.qkv, sv.ctx, sv.z1, sv.st1, sv.y1, sv.hd, sv.z2, sv.st2 = xq, qkv, ctx, z1, st1, y1, hd, z2, st2
    return y2, sv


def ffn_wide_supported(d, ff):
    """bf16 FFN for 64 < d <= 128: tcgen05 GEMMs of the general rows / weight-gradient kernels, hidden materialised in bf16."""
    return 64 < d <= 128 and ff % 64 == 0 and ff >= 64


WIDE_NS = int(os.environ.get("U2GNN_WIDE_NS", "256"))     # output-column slice of linear1 / dH per launch of the rows kernel
WIDE_KS = int(os.environ.get("U2GNN_WIDE_KS", "256"))     # K slice of linear2 / dy1 per launch
WIDE_DP = 128      # padded feature size of the wide path's operand copies (two 64-column K atoms, 256-byte bf16 rows)


def _pad_bf16(x, M, d):
    out = torch.empty((M, WIDE_DP), dtype=torch.bfloat16, device=x.device)
    LIB.call("u2gnn_pad_rows_bf16", _ptr(x), M, d, _ptr(out), WIDE_DP, _stream())
    return out


PACK_MIN_ROWS = int(os.environ.get("U2GNN_PACK_MIN_ROWS", "8192"))   # rows GEMMs of at least this many rows get their weights pre-packed (one extra tiny launch)


def _pack_w(W, w_kn, ldw, N, K, M):
    """Swizzled bf16 operand images of a weight for the K-looping rows GEMMs (u2gnn_gemm_split_pack), or None for small batches."""
    if M < PACK_MIN_ROWS:
        return None
    nbytes = LIB.call("u2gnn_gemm_split_packed_bytes", N, K)
    buf = torch.empty(nbytes, dtype=torch.uint8, device=W.device)
    LIB.call("u2gnn_gemm_split_pack", _ptr(W), int(w_kn), ldw, N, K, _ptr(buf), nbytes, _stream())
    return buf


WIDE_KLOOP = os.environ.get("U2GNN_WIDE_KLOOP", "1") != "0"   # wide bf16 FFN: K-looping rows GEMM with fused ReLU / dropout / mask epilogues (one launch per product)


def _rows_kloop(A, M, K, lda, W, w_kn, N, C, ldc, bias=None, relu=False, drop=None, aux=None, aux_scale=1.0):
    """C[M, N] = epi(A[M, K] op(W) + bias): A bf16 (lda), W fp32 with a WIDE_DP leading dimension, C / aux fp32 or bf16 by dtype."""
    epi, seed, stream, thr = 0, 0, 0, 0
    if bias is not None:
        epi |= 1
    if relu:
        epi |= 2
    if drop is not None and drop[2] > 0:
        epi |= 4
        seed, stream, thr = drop[0], drop[1], drop[2]
    if aux is not None:
        epi |= 8
    if LIB.timed is not None:
        FLOPS["u2gnn_gemm_tc_rows_kloop"] = FLOPS.get("u2gnn_gemm_tc_rows_kloop", 0) + 2 * M * N * K
        _acct_bytes("u2gnn_gemm_tc_rows_kloop", M * (2 * K + C.element_size() * N + (aux.element_size() * N if aux is not None else 0)) + 4 * N * K)
    pk = _pack_w(W, w_kn, WIDE_DP, N, K, M)
    LIB.call("u2gnn_gemm_tc_rows_kloop", _ptr(A), M, K, lda, _ptr(W), int(w_kn), WIDE_DP, N, _ptr(bias), epi, seed, stream, thr, 0, _ptr(aux),
             int(aux is not None and aux.dtype == torch.bfloat16), aux.shape[1] if aux is not None else 0, aux_scale, 0.0, _ptr(C),
             int(C.dtype == torch.bfloat16), ldc, _ptr(pk), _stream())
    return C


def _pad_cols(W, d):
    """[n, d] fp32 weight -> [n, WIDE_DP] zero-padded (weights only: plumbing, 4 * n * 128 bytes)."""
    out = torch.zeros((W.shape[0], WIDE_DP), dtype=torch.float32, device=W.device)
    out[:, :d].copy_(W)
    return out


def ffn_wide_fwd(y1, Mq, d, ff, p, seed, stream_hidden, thr):
    """f = linear2(dropout(relu(linear1(y1)))) for 64 < d <= 128 (transformer.py:977-982).  Returns (f [Mq, d] fp32, saved) with
    saved = (h [Mq, ff] bf16: the hidden after ReLU, dropout and its 1/(1-p) scale; y1p [Mq, 128] bf16: the padded operand copy).
    Every GEMM operand is a 16-byte-aligned bf16 / fp32 buffer with a 128-column feature stride; results are copied into the
    d-strided tensors of the module boundary by u2gnn_copy_rows."""
    dev = y1.device
    s = _stream()
    y1p = _pad_bf16(y1, Mq, d)
    W1p = _pad_cols(p["linear1.weight"], d)                       # [ff, 128]
    W2Tp = _pad_cols(p["linear2.weight"].t(), d)                  # [ff, 128]
    b2p = torch.zeros(WIDE_DP, dtype=torch.float32, device=dev)
    b2p[:d].copy_(p["linear2.bias"])
    b1 = p["linear1.bias"]
    h = torch.empty((Mq, ff), dtype=torch.bfloat16, device=dev)
    if WIDE_KLOOP:
        # one launch per product: linear1 + bias + ReLU + dropout -> h (bf16); linear2 over all of ff -> fp.  K of linear1 / N of
        # linear2 = d rounded up to 8 columns of the 128-column padded operands (the zero padding beyond is never loaded / written)
        d8 = (d + 7) // 8 * 8
        _rows_kloop(y1p, Mq, d8, WIDE_DP, W1p, 0, ff, h, ff, bias=b1, relu=True, drop=(seed, stream_hidden, thr))
        fp = torch.empty((Mq, WIDE_DP), dtype=torch.float32, device=dev)
        _rows_kloop(h, Mq, ff, ff, W2Tp, 1, d8, fp, WIDE_DP, bias=b2p)
        f = torch.empty((Mq, d), dtype=torch.float32, device=dev)
        LIB.call("u2gnn_copy_rows", _ptr(fp), WIDE_DP, _ptr(f), d, Mq, d, 0, s)
        return f, (h, y1p)
    for j in range(0, ff, WIDE_NS):                  # N slices of linear1 (the rows kernel holds N <= 256 accumulator columns)
        n = min(WIDE_NS, ff - j)
        LIB.call("u2gnn_gemm_tc_rows_ex", _ptr(y1p), 1, Mq, WIDE_DP, WIDE_DP, _ptr(W1p) + 4 * j * WIDE_DP, 0, n, _ptr(b1) + 4 * j, 0.0,
                 _ptr(h) + 2 * j, 1, ff, s)
    scale = 256.0 / (256.0 - thr) if thr else 1.0
    LIB.call("u2gnn_relu_dropout_bf16", _ptr(h), Mq, ff, seed, stream_hidden, thr, scale, s)
    fp = torch.empty((Mq, WIDE_DP), dtype=torch.float32, device=dev)
    for j in range(0, ff, WIDE_KS):                  # K slices of linear2, accumulated into fp
        k = min(WIDE_KS, ff - j)
        LIB.call("u2gnn_gemm_tc_rows_ex", _ptr(h) + 2 * j, 1, Mq, k, ff, _ptr(W2Tp) + 4 * j * WIDE_DP, 1, WIDE_DP, _ptr(b2p) if j == 0 else 0,
                 0.0 if j == 0 else 1.0, _ptr(fp), 0, WIDE_DP, s)
    f = torch.empty((Mq, d), dtype=torch.float32, device=dev)
    LIB.call("u2gnn_copy_rows", _ptr(fp), WIDE_DP, _ptr(f), d, Mq, d, 0, s)
    if LIB.timed is not None:                        # both GEMMs go through the rows entry point: 2 x (2 M d ff)
        FLOPS["u2gnn_gemm_tc_rows_ex"] = FLOPS.get("u2gnn_gemm_tc_rows_ex", 0) + 4 * Mq * d * ff
    return f, (h, y1p)


def ffn_wide_bwd(df, dz, saved, Mq, d, ff, p, g, thr):
    """Backward of ffn_wide_fwd: dy1 = dz + dPre W1 (accumulated INTO dz and returned), dW1 / db1 / dW2 accumulated into g
    (db2 = colsum(df) is the caller's).  df [Mq, d] fp32 = gradient at the linear2 output."""
    dev = df.device
    s = _stream()
    h, y1p = saved
    scale = 256.0 / (256.0 - thr) if thr else 1.0
    dfp = _pad_bf16(df, Mq, d)
    W1p = _pad_cols(p["linear1.weight"], d)                       # [ff, 128]: rows j.. are a contiguous [K, N] block for dy1
    W2Tp = _pad_cols(p["linear2.weight"].t(), d)                  # [ff, 128] = [N, K] for dH = df W2
    dh = torch.empty((Mq, ff), dtype=torch.bfloat16, device=dev)
    d8 = (d + 7) // 8 * 8
    if WIDE_KLOOP:
        # dPre = (df W2) masked by the saved hidden (ReLU live AND kept) and scaled, in the GEMM epilogue: one launch, no elementwise pass
        _rows_kloop(dfp, Mq, d8, WIDE_DP, W2Tp, 0, ff, dh, ff, aux=h, aux_scale=scale)
    else:
        for j in range(0, ff, WIDE_NS):
            n = min(WIDE_NS, ff - j)
            LIB.call("u2gnn_gemm_tc_rows_ex", _ptr(dfp), 1, Mq, WIDE_DP, WIDE_DP, _ptr(W2Tp) + 4 * j * WIDE_DP, 0, n, 0, 0.0, _ptr(dh) + 2 * j, 1, ff, s)
    # dW2[d, ff] += df^T h: the weight-gradient kernel takes at most 64 columns of its second operand, so 64-wide slices of h;
    # each slice's [128, 64] result (rows >= d are the zero padding) is added into its columns of dW2
    n_sl = ff // 64
    tmp = torch.zeros((n_sl, WIDE_DP * 64), dtype=torch.float32, device=dev)
    for i in range(n_sl):
        LIB.call("u2gnn_gemm_tc_wgrad_ex", _ptr(dfp), 1, Mq, WIDE_DP, WIDE_DP, _ptr(h) + 2 * 64 * i, 1, 64, ff, 0, 0, _ptr(tmp) + 4 * i * WIDE_DP * 64, 0, s)
        LIB.call("u2gnn_copy_rows", _ptr(tmp) + 4 * i * WIDE_DP * 64, 64, _ptr(g["linear2.weight"]) + 4 * 64 * i, ff, d, 64, 1, s)
    if not WIDE_KLOOP:
        LIB.call("u2gnn_relu_dropout_bwd_bf16", _ptr(dh), _ptr(h), Mq, ff, scale, s)        # dh is now dPre
    # dW1[ff, d] += dPre^T y1, db1 += colsum(dPre): 256-row slices of dW1 x two 64-column blocks of the padded y1
    blocks = [(0, 64), (64, d - 64)]
    n_a = (ff + 255) // 256
    tmp1 = torch.zeros((n_a, 2, 256 * 64), dtype=torch.float32, device=dev)
    for i in range(n_a):
        n1 = min(256, ff - 256 * i)
        for bi, (c0, nc) in enumerate(blocks):
            t = _ptr(tmp1) + 4 * (i * 2 + bi) * 256 * 64
            LIB.call("u2gnn_gemm_tc_wgrad_ex", _ptr(dh) + 2 * 256 * i, 1, Mq, n1, ff, _ptr(y1p) + 2 * c0, 1, 64, WIDE_DP, 0, 0, t,
                     (_ptr(g["linear1.bias"]) + 4 * 256 * i) if bi == 0 else 0, s)
            LIB.call("u2gnn_copy_rows", t, 64, _ptr(g["linear1.weight"]) + 4 * (256 * i * d + c0), d, n1, nc, 1, s)
    dyp = torch.empty((Mq, WIDE_DP), dtype=torch.float32, device=dev)
    if WIDE_KLOOP:
        _rows_kloop(dh, Mq, ff, ff, W1p, 1, d8, dyp, WIDE_DP)                               # dPre W1 over all of ff in one launch
    else:
        for j in range(0, ff, WIDE_KS):              # dPre W1 (K slices; W1p rows j.. are a contiguous [K, N] block)
            k = min(WIDE_KS, ff - j)
            LIB.call("u2gnn_gemm_tc_rows_ex", _ptr(dh) + 2 * j, 1, Mq, k, ff, _ptr(W1p) + 4 * j * WIDE_DP, 1, WIDE_DP, 0, 0.0 if j == 0 else 1.0,
                     _ptr(dyp), 0, WIDE_DP, s)
    LIB.call("u2gnn_copy_rows", _ptr(dyp), WIDE_DP, _ptr(dz), d, Mq, d, 1, s)               # dy1 = dz + dPre W1
    if LIB.timed is not None and not WIDE_KLOOP:                        # dH and dy1 go through the rows entry point (the two weight gradients through wgrad_ex)
        FLOPS["u2gnn_gemm_tc_rows_ex"] = FLOPS.get("u2gnn_gemm_tc_rows_ex", 0) + 4 * Mq * d * ff
    return dz


def ffn_tc_supported(d, ff):
    return d <= 64 and ff % 128 == 0 and 128 <= ff <= 2048


def ffn_tc_pack(p, d, ff, thr):
    """bf16 pre-swizzled weight images for the tcgen05 FFN kernels (rebuilt whenever the weights change)."""
    if not ffn_tc_supported(d, ff):
        raise RuntimeError("precision='bf16' needs feature_dim_size <= 64 and ff_hidden_size a multiple of 128 (<= 2048), or 64 < feature_dim_size <= 128 with ff_hidden_size a multiple of 64; "
                           "got d=%d ff=%d (use precision='fp32')" % (d, ff))
    nbytes = LIB.call("u2gnn_ffn_tc_packed_bytes", d, ff)
    packed = torch.empty(nbytes, dtype=torch.uint8, device=p["linear1.weight"].device)
    scale = 256.0 / (256.0 - thr) if thr else 1.0
    LIB.call("u2gnn_ffn_tc_prepare", _ptr(p["linear1.weight"]), _ptr(p["linear1.bias"]), _ptr(p["linear2.weight"]),
             _ptr(p["linear2.bias"]), d, ff, scale, _ptr(packed), nbytes, _stream())
    return packed


def encoder_layer_bwd(dy2, sv, p, g, d, ff, drop_ids, seed, thr, long_seq, need_dx=True):
    """Backward of encoder_layer_fwd.  g: dict name->gradient tensor (accumulated into).
    Returns dx[B*S, d] or None."""
    dev = dy2.device
    B, S, Sq = sv.B, sv.S, sv.Sq
    M, Mq = B * S, B * Sq
    f32 = dict(dtype=torch.float32, device=dev)
    drop_scale = 256.0 / (256.0 - thr) if thr else 1.0
    # LayerNorm2 + FFN
    fold = sv.packed is not None and d in (4, 8, 16, 32, 64, 128) and FUSE_LN_BWD     # linear2 bias gradient inside the LayerNorm2 backward
    df_img = None
    if fold and d == 64 and sv.y1_img is not None:
        # dF leaves the LayerNorm2 backward as bf16 tile images: the FFN backward bulk-copies them, the fp32 dF never exists
        dz2, df_img = add_dropout_ln_bwd(dy2, sv.z2, sv.st2, Mq, d, p["norm2.weight"], (seed, drop_ids[3], thr),
                                         g["norm2.weight"], g["norm2.bias"], dasum=g["linear2.bias"], da_img=True)
        df = None
    else:
        dz2, df = add_dropout_ln_bwd(dy2, sv.z2, sv.st2, Mq, d, p["norm2.weight"], (seed, drop_ids[3], thr),
                                     g["norm2.weight"], g["norm2.bias"], dasum=g["linear2.bias"] if fold else None)
    dy1 = dz2  # dy1 = dz2 + dhpre @ W1 (in place)
    if sv.wide:
        if df is dz2:                    # no output dropout: df aliases dz2, which ffn_wide_bwd accumulates into
            df = dz2.clone()
        LIB.call("u2gnn_colsum", _ptr(df), Mq, d, d, _ptr(g["linear2.bias"]), 1, _stream())
        ffn_wide_bwd(df, dz2, sv.hd, Mq, d, ff, p, g, thr)
    elif sv.packed is not None:
        # fused tcgen05 backward: hidden and its gradient recomputed on chip
        if df is dz2:                    # no output dropout: df aliases dz2, which the weight-gradient kernel still reads
            dy1 = torch.empty_like(dz2)
        if not fold:
            LIB.call("u2gnn_colsum", _ptr(df), Mq, d, d, _ptr(g["linear2.bias"]), 1, _stream())
        if LIB.timed is not None:
            FLOPS["u2gnn_ffn_tc_bwd"] = FLOPS.get("u2gnn_ffn_tc_bwd", 0) + 8 * Mq * d * ff
        wsb = LIB.call("u2gnn_ffn_tc_bwd_workspace_bytes", Mq)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        LIB.call("u2gnn_ffn_tc_bwd", _ptr(sv.y1), _ptr(df), _ptr(sv.y1_img if df_img is not None else None), _ptr(df_img), _ptr(sv.ffn_mask), _ptr(dz2),
                 Mq, d, ff, _ptr(sv.packed), drop_scale, seed, drop_ids[2], thr, _ptr(dy1), _ptr(g["linear1.weight"]),
                 _ptr(g["linear1.bias"]), _ptr(g["linear2.weight"]), _ptr(ws), wsb, _stream())
    else:
        wgrad_fp32(df, Mq, d, sv.hd, ff, g["linear2.weight"], g["linear2.bias"])
        dhpre = torch.empty((Mq, ff), **f32)
        linear_fp32(df, Mq, d, p["linear2.weight"], 1, ff, dhpre, aux=sv.hd, aux_scale=drop_scale)
        wgrad_fp32(dhpre, Mq, ff, sv.y1, d, g["linear1.weight"], g["linear1.bias"])
        linear_fp32(dhpre, Mq, ff, p["linear1.weight"], 1, d, dy1, beta=1.0)
        del dhpre
    return _encoder_attn_bwd(dy1, sv, p, g, d, ff, drop_ids, seed, thr, long_seq, need_dx)


def _encoder_attn_bwd(dy1, sv, p, g, d, ff, drop_ids, seed, thr, long_seq, need_dx):
    """LayerNorm1 + attention half of encoder_layer_bwd (dy1 = gradient at the LayerNorm1 output)."""
    dev = dy1.device
    B, S, Sq = sv.B, sv.S, sv.Sq
    M, Mq = B * S, B * Sq
    f32 = dict(dtype=torch.float32, device=dev)
    tc_proj = sv.packed is not None and not long_seq and d <= 64
    tc_attn = tc_proj and d == 64 and Sq == S and S >= 2
    fuse_bwd = tc_proj and d == 64 and FUSE_PROJ_BWD
    # da only feeds the two out_proj tensor-core GEMMs (weight gradient, input gradient): bf16, rounded once by its producer
    dz1, da = add_dropout_ln_bwd(dy1, sv.z1, sv.st1, Mq, d, p["norm1.weight"], (seed, drop_ids[1], thr),
                                 g["norm1.weight"], g["norm1.bias"], da_bf16=tc_proj and d == 64 and FUSE_LN_BWD)
    if fuse_bwd and da.dtype == torch.bfloat16:
        # out_proj backward: input gradient and weight gradient from one pass over da
        dctx = proj_bwd_tc(da, Mq, d, sv.ctx, p["self_attn.out_proj.weight"], g["self_attn.out_proj.weight"],
                           g["self_attn.out_proj.bias"], out_bf16=tc_attn)
    elif tc_proj:
        wgrad_tc(da, Mq, d, sv.ctx, d, g["self_attn.out_proj.weight"], g["self_attn.out_proj.bias"])
        dctx = linear_tc(da, Mq, d, p["self_attn.out_proj.weight"], 1, d, out_bf16=tc_attn)
    elif sv.attn_pad:
        return _encoder_attn_bwd_padded(da, dz1, sv, p, g, d, drop_ids, seed, thr, need_dx)
    else:
        wgrad_fp32(da, Mq, d, sv.ctx, d, g["self_attn.out_proj.weight"], g["self_attn.out_proj.bias"])
        dctx = torch.empty((Mq, d), **f32)
        linear_fp32(da, Mq, d, p["self_attn.out_proj.weight"], 1, d, dctx)
    tc_last = sv.qkv.dtype == torch.bfloat16 and Sq == 1
    dqkv = torch.empty((M, 3 * d), dtype=torch.bfloat16 if (tc_attn or tc_last) else torch.float32, device=dev)
    if long_seq:
        scale = math.sqrt(1.0 / d)
        dpd = torch.empty((S, S), **f32)
        sgemm(0, 1, S, S, d, dctx, d, sv.qkv, 3 * d, dpd, S, b_off=2 * d)                 # dP~ = dctx @ v^T
        sgemm(1, 0, S, d, S, sv.pd, S, dctx, d, dqkv, 3 * d, c_off=2 * d)                 # dv = P~^T @ dctx
        LIB.call("u2gnn_softmax_rows_bwd", _ptr(sv.probs), _ptr(dpd), S, S, seed, drop_ids[0], thr, _stream())
        sgemm(0, 0, S, d, S, dpd, S, sv.qkv, 3 * d, dqkv, 3 * d, alpha=scale, b_off=d)     # dq = ds @ k * scale
        sgemm(1, 0, S, d, S, dpd, S, sv.qkv, 3 * d, dqkv, 3 * d, alpha=scale, c_off=d)     # dk = ds^T @ q * scale
    elif tc_attn:
        _acct_bytes("u2gnn_seqattn_tc_bwd_ex", M * (2 * 3 * d * 2 + 2 * d))                # qkv, dctx -> dqkv (bf16)
        LIB.call("u2gnn_seqattn_tc_bwd_ex", _ptr(sv.qkv), _ptr(dctx), B, S, d, seed, drop_ids[0], thr, _ptr(dqkv), 1, _stream())
    elif tc_last:
        LIB.call("u2gnn_seqattn_last_bwd_ex", _ptr(sv.qkv), _ptr(dctx), 1, B, S, d, seed, drop_ids[0], thr, _ptr(dqkv), _stream())
    else:
        LIB.call("u2gnn_seqattn_bwd", _ptr(sv.qkv), _ptr(dctx), B, S, Sq, d, seed, drop_ids[0], thr, _ptr(dqkv), _stream())
    if fuse_bwd and need_dx and dqkv.dtype == torch.bfloat16:
        # in_proj backward in one pass over dqkv; the residual gradient dz1 is the old C of the epilogue when every row is live
        if Sq == S:
            return proj_bwd_tc(dqkv, M, 3 * d, sv.x, p["self_attn.in_proj_weight"], g["self_attn.in_proj_weight"],
                               g["self_attn.in_proj_bias"], out=dz1, beta=1.0, inp_idx=sv.x_idx)
        dx = proj_bwd_tc(dqkv, M, 3 * d, sv.x, p["self_attn.in_proj_weight"], g["self_attn.in_proj_weight"],
                         g["self_attn.in_proj_bias"])
        copy_rows(dz1, d, dx, S * d, B, d, accumulate=True)
        return dx
    if tc_proj:
        wgrad_tc(dqkv, M, 3 * d, sv.x, d, g["self_attn.in_proj_weight"], g["self_attn.in_proj_bias"], inp_idx=sv.x_idx)
    else:
        wgrad_fp32(dqkv, M, 3 * d, sv.x, d, g["self_attn.in_proj_weight"], g["self_attn.in_proj_bias"])
    if not need_dx:
        return None
    if tc_proj and Sq == S:
        # residual gradient folded into the projection epilogue: dz1 += dqkv W_in (saves the separate axpy pass)
        return linear_tc(dqkv, M, 3 * d, p["self_attn.in_proj_weight"], 1, d, beta=1.0, out=dz1)
    if FP32_TC and Sq == S:
        # residual gradient folded into the projection epilogue (beta = 1 over dz1)
        return linear_fp32(dqkv, M, 3 * d, p["self_attn.in_proj_weight"], 1, d, dz1, beta=1.0)
    if tc_proj:
        dx = linear_tc(dqkv, M, 3 * d, p["self_attn.in_proj_weight"], 1, d)
    else:
        dx = torch.empty((M, d), **f32)
        linear_fp32(dqkv, M, 3 * d, p["self_attn.in_proj_weight"], 1, d, dx)
    if Sq == S:
        LIB.call("u2gnn_axpy", 1.0, _ptr(dz1), _ptr(dx), M * d, _stream())
    else:
        copy_rows(dz1, d, dx, S * d, B, d, accumulate=True)
    return dx


def _encoder_attn_bwd_padded(da, dz1, sv, p, g, d, drop_ids, seed, thr, need_dx):
    """Attention half of the backward in the padded q | k | v layout of attn_pad_size (fp32, every query row live): the padded
    weights' gradients are computed in that layout and their live rows / columns added into the real gradients (the q block
    times the sqrt(DP / d) folded into its weights)."""
    DP, cq, Wp, Wop = sv.attn_pad
    B, S = sv.B, sv.S
    M = B * S
    dev = da.device
    f32 = dict(dtype=torch.float32, device=dev)
    dWop = torch.zeros((d, DP), **f32)
    wgrad_fp32(da, M, d, sv.ctx, DP, dWop, g["self_attn.out_proj.bias"])
    g["self_attn.out_proj.weight"].add_(dWop[:, :d])
    dctx = torch.empty((M, DP), **f32)
    linear_fp32(da, M, d, Wop, 1, DP, dctx)
    dqkv = torch.empty((M, 3 * DP), **f32)
    LIB.call("u2gnn_seqattn_bwd", _ptr(sv.qkv), _ptr(dctx), B, S, S, DP, seed, drop_ids[0], thr, _ptr(dqkv), _stream())
    dWp = torch.zeros((3, DP, d), **f32)
    dbp = torch.zeros((3, DP), **f32)
    wgrad_fp32(dqkv, M, 3 * DP, sv.x, d, dWp, dbp)
    dWp[0].mul_(cq)
    dbp[0].mul_(cq)
    g["self_attn.in_proj_weight"].view(3, d, d).add_(dWp[:, :d])
    g["self_attn.in_proj_bias"].view(3, d).add_(dbp[:, :d])
    if not need_dx:
        return None
    return linear_fp32(dqkv, M, 3 * DP, Wp, 1, d, dz1, beta=1.0)      # dx = dz1 + dqkv W_in (residual folded into the epilogue)


# --------------------------------------------------------------------------------------
# one U2GNN layer = gather -> T encoder layers -> position 0
# --------------------------------------------------------------------------------------
@dataclass
class StackSaved:
    layers: list = field(default_factory=list)
    n_src: int = 0
    N: int = 0
    S: int = 0


def u2gnn_layer_fwd(src, input_x, params, l, T, attn_axis, drop: DropoutCfg, precision="fp32", for_backward=True):
    """src[n_src, d], input_x[N, S] int64 -> out[N, d].  params: list (per timestep) of dicts."""
    require_device()
    _check(src, torch.float32, "src"); _check(input_x, torch.int64, "input_x")
    N, S = input_x.shape
    d = src.shape[1]
    ff = params[0]["linear1.weight"].shape[0]
    thr = drop.thr_enc()
    saved = StackSaved(n_src=src.shape[0], N=N, S=S)
    x_idx = None
    if attn_axis == "neighbors":
        if S > 32:
            raise ValueError("attn_axis='neighbors' supports num_neighbors <= 31")
        if T >= 2 and gather_fusable(d, ff, S, S, precision, False):
            x, x_idx = src, input_x                         # the first timestep's kernels read src[input_x[n, s]] themselves
        else:
            x = gather_rows(src, input_x)                   # [N*S, d]
        B, Sseq = N, S
    elif attn_axis == "nodes":
        x = gather_rows(src, input_x, idx_stride=S, n_idx=N)  # column 0 only
        B, Sseq = 1, N
    else:
        raise ValueError("attn_axis must be 'nodes' or 'neighbors'")
    for t in range(T):
        last = attn_axis == "neighbors" and t == T - 1
        Sq = 1 if last else Sseq
        ids = [stream_id(l, t, s, T) for s in range(4)]
        x, sv = encoder_layer_fwd(x, B, Sseq, Sq, params[t], d, ff, ids, drop.seed, thr, attn_axis == "nodes", precision, for_backward,
                                  x_idx=x_idx if t == 0 else None)
        saved.layers.append(sv)
    return x, saved  # [N, d] in both layouts


def u2gnn_layer_bwd(dout, saved: StackSaved, input_x, params, grads, l, T, attn_axis, drop: DropoutCfg,
                    need_dsrc=True, transpose: IndexTranspose | None = None):
    d = dout.shape[1]
    ff = params[0]["linear1.weight"].shape[0]
    thr = drop.thr_enc()
    dx = dout
    for t in reversed(range(T)):
        ids = [stream_id(l, t, s, T) for s in range(4)]
        need_dx = need_dsrc or t > 0
        dx = encoder_layer_bwd(dx, saved.layers[t], params[t], grads[t], d, ff, ids, drop.seed, thr,
                               attn_axis == "nodes", need_dx=need_dx)
        saved.layers[t] = None
    if not need_dsrc:
        return None
    if attn_axis == "neighbors":
        if transpose is not None:
            return transpose.scatter_add(dx)
        return scatter_add_rows(dx, input_x, saved.n_src)
    return scatter_add_rows(dx, input_x, saved.n_src, idx_stride=saved.S)
