// Fused bf16 FFN block on tcgen05 + TMEM, backward entry point.  The [rows, ff] hidden activation and its
// gradient are recomputed on chip (never stored in HBM).  Three launches:
//   images  rows_to_images_kernel: fp32 rows of y1 / dF -> bf16 swizzled [128 x 64] tile images (the operand format both
//           tensor-core kernels bulk-copy)
//   wgrad   (ffn_tc_wgrad.cu: one ff chunk per CTA, row tiles inner, computed transposed)  dW1, db1, dW2  + 1 mask bit per
//           hidden activation (ReLU live AND dropout keep)
//   dgrad   (ffn_tc_dgrad.cu: rows outer, ff chunks inner)  dy1 = dz + dPre W1, dPre = (dF W2) & mask - no second
//           recomputation of the hidden: 4 + 2 = 6 executed GEMM units per (tile, chunk) for 4 algorithmic
//           (round 1: 4 + 3, with the hidden and its gradient staged through shared memory in wgrad).
// Measured dead end (round 1, kept as a note): ONE merged kernel adding its 16 chunk partials of dy1 into L2 with
// red.global.add.v4.f32 was slower (11.3 against 10.8 ms per 4.19 M rows): the reductions are LSU-issue bound (0.85-1.3
// cycles per lane and instruction).  Bulk reductions (cp.reduce.async.bulk, tools/probe_red.py: 5.25 TB/s, issued by the copy
// engine) would fix the issue cost but need a 34 KB staging buffer per tile in flight, which the merged kernel's shared
// memory does not have.
//
// with X = y1 (block input), dF = gradient at the linear2 output (after the output dropout),
// dz = gradient at the residual sum.  These are the autograd of linear1/ReLU/dropout/linear2 in
// nn.TransformerEncoderLayer._ff_block (torch/nn/modules/transformer.py:977-982).
#include "common.cuh"
#include "tc_common.cuh"
#include "ffn_epi.cuh"

int ffn_tc_dgrad_launch(const float* dz, float* dy1, int64_t M, int d, int ff, const void* packed, const void* fb, const void* mask,
                        cudaStream_t st);
int ffn_tc_wgrad_launch(const void* xb, const void* fb, int64_t M, int d, int ff, const void* packed, float hidden_scale,
                        uint64_t seed, uint32_t stream_hidden, int thr, float* dW1, float* db1, float* dW2, void* mask,
                        bool mask_from_forward, cudaStream_t st);

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
            if (!src) continue;
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

}  // namespace

namespace {
// workspace layout: [y1 images | dF images | mask words], all sized for whole PAIRS of 128-row tiles (the dgrad kernel walks pairs)
inline size_t tiles_padded(int64_t M) { return (size_t)(2 * ((M + 255) / 256)); }
}  // namespace

extern "C" size_t u2gnn_ffn_tc_bwd_workspace_bytes(int64_t M) {
    if (M < 0) return 0;
    return tiles_padded(M) * (2 * 16384 + 16 * 2048);      // two 16 KB images + 2 KB of mask words per ff chunk (ff <= 2048) per tile
}

extern "C" size_t u2gnn_ffn_tc_image_bytes(int64_t M) { return M < 0 ? 0 : tiles_padded(M) * 16384; }
extern "C" size_t u2gnn_ffn_tc_mask_bytes(int64_t M, int ff) { return (M < 0 || ff < 128) ? 0 : tiles_padded(M) * (size_t)(ff / 128) * 2048; }

extern "C" int u2gnn_ffn_tc_bwd(const float* y1, const float* df, const void* y1_img, const void* df_img, const void* fwd_mask,
                                const float* dz, int64_t M, int d, int ff, const void* packed, float hidden_scale, uint64_t seed,
                                uint32_t stream_hidden, int thr,
                                float* dy1, float* dW1, float* db1, float* dW2, void* workspace, size_t workspace_bytes,
                                u2gnn_stream_t stream) {
    if ((!y1 && !y1_img) || (!df && !df_img) || !dz || !packed || !dy1 || !dW1 || !db1 || !dW2 || !workspace || M < 0 || thr < 0 || thr > 255)
        return U2GNN_EINVAL;
    if ((y1_img && reinterpret_cast<uintptr_t>(y1_img) % 128) || (df_img && reinterpret_cast<uintptr_t>(df_img) % 128)) return U2GNN_EALIGN;
    if (workspace_bytes < u2gnn_ffn_tc_bwd_workspace_bytes(M)) return U2GNN_EWORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) % 128) return U2GNN_EALIGN;
    if (d < 1 || d > 64 || ff < 128 || ff % 128 || ff > 2048) return U2GNN_EUNSUPPORTED;
    if (d == 64 && ((reinterpret_cast<uintptr_t>(y1) | reinterpret_cast<uintptr_t>(df) | reinterpret_cast<uintptr_t>(dz) |
                     reinterpret_cast<uintptr_t>(dy1)) % 16))
        return U2GNN_EALIGN;
    if (M == 0) return U2GNN_OK;
    const size_t tp = tiles_padded(M);
    uint8_t* xb = static_cast<uint8_t*>(workspace);
    uint8_t* fb = xb + tp * 16384;
    uint8_t* mask = fb + tp * 16384;
    const int64_t n_tiles = (M + 127) / 128;
    // operands the producers did not already leave as tile images (u2gnn_gemm_tc_rows_ln: y1, u2gnn_add_dropout_ln_bwd_ex: dF)
    // are converted here; a null source pointer skips that tensor
    if (!y1_img || !df_img)
        rows_to_images_kernel<<<(int)(n_tiles < 4 * U2GNN_NUM_SMS ? n_tiles : 4 * U2GNN_NUM_SMS), 256, 0, as_stream(stream)>>>(
            y1_img ? nullptr : y1, df_img ? nullptr : df, M, d, xb, fb);
    if (y1_img) xb = const_cast<uint8_t*>(static_cast<const uint8_t*>(y1_img));
    if (df_img) fb = const_cast<uint8_t*>(static_cast<const uint8_t*>(df_img));
    if (fwd_mask) mask = const_cast<uint8_t*>(static_cast<const uint8_t*>(fwd_mask));      // written by u2gnn_ffn_tc_fwd(mask_out): nothing to recompute
    int rc = ffn_tc_wgrad_launch(xb, fb, M, d, ff, packed, hidden_scale, seed, stream_hidden, thr, dW1, db1, dW2, mask, fwd_mask != nullptr,
                                 as_stream(stream));
    if (rc != U2GNN_OK) return rc;
    rc = ffn_tc_dgrad_launch(dz, dy1, M, d, ff, packed, fb, mask, as_stream(stream));
    if (rc != U2GNN_OK) return rc;
    U2GNN_CHECK_LAUNCH();
}
