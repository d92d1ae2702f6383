// Fused bf16 FFN block on tcgen05 + TMEM, backward entry point.  The [rows, ff] hidden activation and its
// gradient are recomputed on chip (never stored in HBM).  Two implementations (u2gnn_ffn_tc_bwd_mode):
//
//   mode 0 (default)  two kernels with opposite loop orders (the hidden is recomputed in both):
//                       dgrad  (ffn_tc_dgrad.cu: rows outer, ff chunks inner)          dy1 = dz + dPre W1
//                       wgrad  (ffn_tc_wgrad.cu: one ff chunk per CTA, row tiles inner)  dW1, db1, dW2
//   mode 1            MERGED: one kernel per (ff chunk, row slice) computes dW1, db1, dW2 AND the chunk's partial of
//                     dy1 = dz + dPre W1, which it adds into dy1 with vector reductions (ffn_tc_wgrad.cu, MERGED).  The
//                     bf16 tile images of y1 / dF it streams are produced by rows_to_images_kernel below.
//                     Parity-tested (tests/test_gpu_tc.py::test_ffn_tc_backward[mode 1]) but SLOWER on B200: 11.31 ms per
//                     4.19 M rows against 10.81 ms for mode 0.  The 16 chunk partials are 18 GB of L2 reductions per
//                     launch (3.8 ms at the measured 4.84 TB/s) and they do not overlap the epilogue warps that issue
//                     them: the kernel takes wgrad (6.3 ms) + reductions instead of hiding them.  Removing the second
//                     recomputation needs fewer partials per row (a cluster of CTAs reducing dY through distributed
//                     shared memory), not a faster reduction.
//
// with X = y1 (block input), dF = gradient at the linear2 output (after the output dropout),
// dz = gradient at the residual sum.  These are the autograd of linear1/ReLU/dropout/linear2 in
// nn.TransformerEncoderLayer._ff_block (torch/nn/modules/transformer.py:977-982).
#include "common.cuh"
#include "tc_common.cuh"
#include "ffn_epi.cuh"

struct FfnLnBwd {       // LayerNorm2 backward fused into the dgrad loader (ffn_tc_dgrad.cu)
    const float* dy2;
    const float* z2;
    const float* st2;
    const float* gamma2;
    uint32_t stream_out;
    float* dgamma2;
    float* dbeta2;
    float* db2;
};
int ffn_tc_dgrad_launch(const float* y1, const float* df, const float* dz, float* dy1, int64_t M, int d, int ff,
                        const void* packed, uint64_t seed, uint32_t stream_hidden, int thr, void* xb, void* fb, cudaStream_t st,
                        const FfnLnBwd* ln);
int ffn_tc_wgrad_launch(const void* xb, const void* fb, int64_t M, int d, int ff, const void* packed, float hidden_scale,
                        uint64_t seed, uint32_t stream_hidden, int thr, float* dW1, float* db1, float* dW2, void* mask,
                        cudaStream_t st);

namespace {

// fp32 rows [M, d] (d <= 64, zero padded) of TWO tensors -> bf16 swizzled [128 x 64] tile images (16 KB per tile), the
// operand format the backward kernel bulk-copies.  One CTA per tile, coalesced 128-bit loads, 8-byte stores.
__global__ void __launch_bounds__(256) rows_to_images_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t M, int d,
                                                             uint8_t* __restrict__ ia, uint8_t* __restrict__ ib) {
    const int64_t n_tiles = (M + 127) / 128;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * 128;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            const float* src = which ? b : a;
            uint8_t* img = (which ? ib : ia) + (size_t)tile * 16384;
            if (d == 64) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = u * 256 + threadIdx.x;
                    const int64_t rg = row0 + (e >> 4);
                    v[u] = (rg < M) ? __ldg(reinterpret_cast<const float4*>(src + rg * 64) + (e & 15)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = u * 256 + threadIdx.x;
                    uint2 w;
                    w.x = epi::cvt2(v[u].x, v[u].y);
                    w.y = epi::cvt2(v[u].z, v[u].w);
                    *reinterpret_cast<uint2*>(img + tc::sw128_offset(e >> 4, (e & 15) * 4)) = w;
                }
            } else {
                for (int e = threadIdx.x; e < 128 * 64; e += 256) {
                    const int r = e >> 6, k = e & 63;
                    const int64_t rg = row0 + r;
                    const float x = (rg < M && k < d) ? src[rg * d + k] : 0.0f;
                    *reinterpret_cast<__nv_bfloat16*>(img + tc::sw128_offset(r, k)) = __float2bfloat16(x);
                }
            }
        }
    }
}

int g_bwd_mode = 0;

}  // namespace

extern "C" int u2gnn_ffn_tc_bwd_mode(int mode) {
    g_bwd_mode = mode;
    return U2GNN_OK;
}

extern "C" size_t u2gnn_ffn_tc_bwd_workspace_bytes(int64_t M) {
    if (M < 0) return 0;
    return (size_t)(2 * ((M + 255) / 256)) * 2 * 16384;      // two images (y1, df) of 16 KB per 128-row tile, pairs of tiles
}

extern "C" int u2gnn_ffn_tc_bwd(const float* y1, const float* df, const float* dz, int64_t M, int d, int ff,
                                const void* packed, float hidden_scale, uint64_t seed, uint32_t stream_hidden, int thr,
                                float* dy1, float* dW1, float* db1, float* dW2, void* workspace, size_t workspace_bytes,
                                u2gnn_stream_t stream) {
    if (!y1 || !df || !dz || !packed || !dy1 || !dW1 || !db1 || !dW2 || !workspace || M < 0 || thr < 0 || thr > 255) return U2GNN_EINVAL;
    if (workspace_bytes < u2gnn_ffn_tc_bwd_workspace_bytes(M)) return U2GNN_EWORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) % 128) return U2GNN_EALIGN;
    if (d < 1 || d > 64 || ff < 128 || ff % 128 || ff > 2048) return U2GNN_EUNSUPPORTED;
    if (d == 64 && ((reinterpret_cast<uintptr_t>(y1) | reinterpret_cast<uintptr_t>(df) | reinterpret_cast<uintptr_t>(dz) |
                     reinterpret_cast<uintptr_t>(dy1)) % 16))
        return U2GNN_EALIGN;
    if (M == 0) return U2GNN_OK;
    // workspace: bf16 swizzled tile images of y1 and df (written by dgrad, bulk-copied by wgrad)
    const size_t half = u2gnn_ffn_tc_bwd_workspace_bytes(M) / 2;
    uint8_t* xb = static_cast<uint8_t*>(workspace);
    uint8_t* fb = xb + half;
    int rc;
    rc = ffn_tc_dgrad_launch(y1, df, dz, dy1, M, d, ff, packed, seed, stream_hidden, thr, xb, fb, as_stream(stream), nullptr);
    if (rc != U2GNN_OK) return rc;
    rc = ffn_tc_wgrad_launch(xb, fb, M, d, ff, packed, hidden_scale, seed, stream_hidden, thr, dW1, db1, dW2, nullptr, as_stream(stream));
    if (rc != U2GNN_OK) return rc;
    U2GNN_CHECK_LAUNCH();
}

// LayerNorm2 backward + output dropout + FFN backward (d == 64): dy2 is the gradient at the LayerNorm2 output, z2 / st2 the
// saved pre-norm rows and (mean, rstd).  The dgrad kernel's dF loader computes dz = LN backward(dy2), dF = dropout(dz) on the
// fly (the separate LayerNorm-backward pass read 520 and wrote 512 bytes per row), accumulates dgamma2 / dbeta2 and the
// linear2 bias gradient db2 = colsum(dF), and the drain adds dPre W1 to the parked dz.
extern "C" int u2gnn_ffn_tc_bwd_ln(const float* y1, const float* dy2, const float* z2, const float* st2, const float* gamma2,
                                   uint32_t stream_out, int64_t M, int d, int ff, const void* packed, float hidden_scale,
                                   uint64_t seed, uint32_t stream_hidden, int thr, float* dy1, float* dW1, float* db1, float* dW2,
                                   float* db2, float* dgamma2, float* dbeta2, void* workspace, size_t workspace_bytes,
                                   u2gnn_stream_t stream) {
    if (!y1 || !dy2 || !z2 || !st2 || !gamma2 || !packed || !dy1 || !dW1 || !db1 || !dW2 || !db2 || !dgamma2 || !dbeta2 ||
        !workspace || M < 0 || thr < 0 || thr > 255)
        return U2GNN_EINVAL;
    if (workspace_bytes < u2gnn_ffn_tc_bwd_workspace_bytes(M)) return U2GNN_EWORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) % 128) return U2GNN_EALIGN;
    if (d != 64 || ff < 128 || ff % 128 || ff > 2048) return U2GNN_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(y1) | reinterpret_cast<uintptr_t>(dy2) | reinterpret_cast<uintptr_t>(z2) |
         reinterpret_cast<uintptr_t>(dy1) | reinterpret_cast<uintptr_t>(gamma2)) % 16 || reinterpret_cast<uintptr_t>(st2) % 8)
        return U2GNN_EALIGN;
    if (dy1 == dy2 || dy1 == z2) return U2GNN_EINVAL;          // dy1 is written while later tiles of dy2 / z2 are still unread
    if (M == 0) return U2GNN_OK;
    const size_t half = u2gnn_ffn_tc_bwd_workspace_bytes(M) / 2;
    uint8_t* xb = static_cast<uint8_t*>(workspace);
    uint8_t* fb = xb + half;
    FfnLnBwd ln;
    ln.dy2 = dy2; ln.z2 = z2; ln.st2 = st2; ln.gamma2 = gamma2; ln.stream_out = stream_out;
    ln.dgamma2 = dgamma2; ln.dbeta2 = dbeta2; ln.db2 = db2;
    int rc = ffn_tc_dgrad_launch(y1, nullptr, dy1, dy1, M, d, ff, packed, seed, stream_hidden, thr, xb, fb, as_stream(stream), &ln);
    if (rc != U2GNN_OK) return rc;
    rc = ffn_tc_wgrad_launch(xb, fb, M, d, ff, packed, hidden_scale, seed, stream_hidden, thr, dW1, db1, dW2, nullptr, as_stream(stream));
    if (rc != U2GNN_OK) return rc;
    U2GNN_CHECK_LAUNCH();
}
