// Elementwise halves of the bf16 FFN for 64 < feature_dim_size <= 128 (BASELINE.json configs[2]: IMDBBINARY shape, d = 65).
// The fused FFN kernels (ffn_tc*.cu) fix the feature tile at one 64-column swizzle atom; for wider features the engine runs the
// FFN as tcgen05 GEMMs of the general rows / weight-gradient kernels (gemm_tc.cu) with the hidden activation MATERIALISED in bf16
// (engine.ffn_wide_fwd / ffn_wide_bwd).  These two kernels are what sits between the GEMMs:
//   forward   h <- relu(h) * keep * scale          in place on the [M, ff] bf16 hidden (linear1's bias is added by the GEMM epilogue)
//   backward  dh <- (h > 0) ? dh * scale : 0        h is the saved forward result, so h > 0 means ReLU live AND kept
// i.e. ReLU + dropout of nn.TransformerEncoderLayer._ff_block (torch/nn/modules/transformer.py:977-982) and their autograd.
// The keep bits are the engine's dropout stream (rng.cuh): element (row, col) -> group (row * ff + col) >> 5, bit col & 31.
#include <cuda_bf16.h>
#include "common.cuh"
#include "rng.cuh"

namespace {

__global__ void __launch_bounds__(256) relu_dropout_bf16_kernel(__nv_bfloat16* __restrict__ h, int64_t n8, RngKeys keys, int thr, int low,
                                                                float scale) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n8; t += (int64_t)gridDim.x * blockDim.x) {
        uint4 w = reinterpret_cast<const uint4*>(h)[t];
        const uint64_t e0 = (uint64_t)t * 8;
        uint32_t kw = 0xFFu;
        if (thr) kw = rng_keep_word_lo(keys, e0 >> 5, thr, low) >> (e0 & 31);
        __nv_bfloat162* v = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 f = __bfloat1622float2(v[j]);
            f.x = ((kw >> (2 * j)) & 1u) ? fmaxf(f.x, 0.0f) * scale : 0.0f;
            f.y = ((kw >> (2 * j + 1)) & 1u) ? fmaxf(f.y, 0.0f) * scale : 0.0f;
            v[j] = __floats2bfloat162_rn(f.x, f.y);
        }
        reinterpret_cast<uint4*>(h)[t] = w;
    }
}

__global__ void __launch_bounds__(256) relu_dropout_bwd_bf16_kernel(__nv_bfloat16* __restrict__ dh, const __nv_bfloat16* __restrict__ h,
                                                                    int64_t n8, float scale) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n8; t += (int64_t)gridDim.x * blockDim.x) {
        uint4 g = reinterpret_cast<const uint4*>(dh)[t];
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(h) + t);
        __nv_bfloat162* gv = reinterpret_cast<__nv_bfloat162*>(&g);
        const __nv_bfloat162* av = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(gv[j]), x = __bfloat1622float2(av[j]);
            gv[j] = __floats2bfloat162_rn(x.x > 0.0f ? f.x * scale : 0.0f, x.y > 0.0f ? f.y * scale : 0.0f);
        }
        reinterpret_cast<uint4*>(dh)[t] = g;
    }
}

// fp32 rows [M, d] -> bf16 rows [M, dp] zero-padded (dp a multiple of 8, >= d): the aligned operand copy the general GEMM kernels
// stream with 16-byte accesses (rows of 65 floats are not 16-byte aligned: their scalar staging path ran 10x below HBM speed)
__global__ void __launch_bounds__(256) pad_rows_bf16_kernel(const float* __restrict__ src, int64_t M, int d, __nv_bfloat16* __restrict__ dst, int dp) {
    const int64_t total = M * (int64_t)(dp / 2);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = t / (dp / 2);
        const int c = (int)(t - row * (dp / 2)) * 2;
        const float a = (c < d) ? __ldg(src + row * d + c) : 0.0f;
        const float b = (c + 1 < d) ? __ldg(src + row * d + c + 1) : 0.0f;
        reinterpret_cast<__nv_bfloat162*>(dst)[t] = __floats2bfloat162_rn(a, b);
    }
}

}  // namespace

extern "C" int u2gnn_pad_rows_bf16(const float* src, int64_t M, int d, void* dst, int dp, u2gnn_stream_t stream) {
    if (!src || !dst || M < 0 || d < 1 || dp < d || (dp & 7)) return U2GNN_EINVAL;
    if (reinterpret_cast<uintptr_t>(dst) % 16) return U2GNN_EALIGN;
    if (M == 0) return U2GNN_OK;
    pad_rows_bf16_kernel<<<grid_for(M * (int64_t)(dp / 2), 256, 8), 256, 0, as_stream(stream)>>>(src, M, d, static_cast<__nv_bfloat16*>(dst), dp);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_relu_dropout_bf16(void* h, int64_t M, int ff, uint64_t seed, uint32_t rng_stream, int thr, float scale,
                                       u2gnn_stream_t stream) {
    if (!h || M < 0 || ff < 32 || (ff & 31) || thr < 0 || thr > 255) return U2GNN_EINVAL;
    if (reinterpret_cast<uintptr_t>(h) % 16) return U2GNN_EALIGN;
    if (M == 0) return U2GNN_OK;
    const int64_t n8 = M * (int64_t)ff / 8;
    relu_dropout_bf16_kernel<<<grid_for(n8, 256, 8), 256, 0, as_stream(stream)>>>(static_cast<__nv_bfloat16*>(h), n8, rng_keys(seed, rng_stream),
                                                                                thr, rng_thr_low(thr), scale);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_relu_dropout_bwd_bf16(void* dh, const void* h, int64_t M, int ff, float scale, u2gnn_stream_t stream) {
    if (!dh || !h || M < 0 || ff < 8 || (ff & 7)) return U2GNN_EINVAL;
    if ((reinterpret_cast<uintptr_t>(dh) | reinterpret_cast<uintptr_t>(h)) % 16) return U2GNN_EALIGN;
    if (M == 0) return U2GNN_OK;
    const int64_t n8 = M * (int64_t)ff / 8;
    relu_dropout_bwd_bf16_kernel<<<grid_for(n8, 256, 8), 256, 0, as_stream(stream)>>>(static_cast<__nv_bfloat16*>(dh),
                                                                                     static_cast<const __nv_bfloat16*>(h), n8, scale);
    U2GNN_CHECK_LAUNCH();
}
