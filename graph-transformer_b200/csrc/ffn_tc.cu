// Fused bf16 FFN block on the 5th-generation tensor cores (tcgen05 + TMEM), forward.
//
//   z = y1 + dropout3( dropout2(relu(y1 W1^T + b1)) W2^T + b2 ),  stats = LayerNorm statistics of z,
//   xnext = LayerNorm(z) * gamma + beta
// i.e. linear1 -> ReLU -> dropout -> linear2 -> dropout -> residual -> norm2 of
// nn.TransformerEncoderLayer (torch/nn/modules/transformer.py:950-958,977-982).  The [rows, ff]
// hidden activation never leaves the SM: per 128-row tile and 128-wide ff chunk
//     S  = X W1c^T            tcgen05.mma SS  (X bf16 in smem, W1c bulk-copied pre-swizzled image)
//     H  = act(S)             epilogue warps: tcgen05.ld -> bias/ReLU/dropout -> bf16 -> tcgen05.st
//     Y += H W2c^T            tcgen05.mma TS  (H read from tensor memory)
// Persistent CTAs (one per SM) walk pairs of row tiles so that the tensor pipe works on one tile
// while the epilogue warps convert the other.  Weights stream from L2 through a 4-stage
// bulk-copy/mbarrier ring shared by both tiles.
//
// Warp roles (384 threads): warp 0 weight producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-7 epilogue of tile 0, warps 8-11 epilogue of tile 1 (warp % 4 selects the TMEM lane quarter).
// TMEM columns: Y0 [0,64) Y1 [64,128) S0 [128,256) S1 [256,384) H0 [384,448) H1 [448,512).
#include "common.cuh"
#include "rng.cuh"
#include "tc_common.cuh"

namespace {

constexpr int DP = 64;            // padded feature size (K of GEMM1, N of GEMM2)
constexpr int CH = 128;           // ff chunk
constexpr int TM = 128;           // rows per tile
constexpr int STAGES = 4;
constexpr uint32_t W1_BYTES = CH * DP * 2;       // 16 KB  [128 x 64] K-major image
constexpr uint32_t W2_BYTES = DP * CH * 2;       // 16 KB  two [64 x 64] K-major images
constexpr uint32_t FWD_BLOCK = W1_BYTES + W2_BYTES;
// packed weights: per chunk [W2c | W1c | W2Tc | W1Tc] (fwd: first two; dgrad: last three; wgrad: middle two), then b1, b2 (fp32)
constexpr uint32_t CHUNK_BYTES = 4 * 16384;
constexpr int kThreads = 384;

struct FwdParams {
    const float* y1;
    int64_t M;
    int d, ff;
    const uint8_t* packed;
    RngKeys keys2, keys3;
    int thr;
    float scale3;
    const float* gamma;
    const float* beta;
    float* z;
    float* stats;
    float* xnext;
};

__device__ __forceinline__ const float* packed_b1(const uint8_t* packed, int ff) {
    return reinterpret_cast<const float*>(packed + (size_t)(ff / CH) * CHUNK_BYTES);
}

// ---------------------------------------------------------------------------------------------
// weight packing: fp32 master weights -> pre-swizzled bf16 shared-memory images
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ffn_pack_kernel(const float* __restrict__ W1, const float* __restrict__ b1,
                                                       const float* __restrict__ W2, const float* __restrict__ b2, int d,
                                                       int ff, float hidden_scale, uint8_t* __restrict__ packed) {
    const int c = blockIdx.x;  // chunk
    uint8_t* blk = packed + (size_t)c * CHUNK_BYTES;
    for (int e = threadIdx.x; e < CH * DP; e += blockDim.x) {
        const int r = e / DP, k = e % DP;           // r: hidden unit in chunk, k: feature
        const int h = c * CH + r;
        const float w1 = (k < d) ? W1[(size_t)h * d + k] : 0.0f;
        const float w2 = (k < d) ? W2[(size_t)k * ff + h] * hidden_scale : 0.0f;
        // W1c  : B of GEMM1  [N = hidden r][K = feature k]
        *reinterpret_cast<__nv_bfloat16*>(blk + 16384 + tc::sw128_offset(r, k)) = __float2bfloat16(w1);
        // W2c  : B of GEMM2  [N = feature k][K = hidden r]  (two K atoms of 64)
        *reinterpret_cast<__nv_bfloat16*>(blk + (r >> 6) * 8192 + tc::sw128_offset(k, r & 63)) = __float2bfloat16(w2);
        // W2Tc : B of dH = dF W2c   [N = hidden r][K = feature k]
        *reinterpret_cast<__nv_bfloat16*>(blk + 32768 + tc::sw128_offset(r, k)) = __float2bfloat16(w2);
        // W1Tc : B of dy1 += dPre W1c  [N = feature k][K = hidden r]
        *reinterpret_cast<__nv_bfloat16*>(blk + 49152 + (r >> 6) * 8192 + tc::sw128_offset(k, r & 63)) = __float2bfloat16(w1);
    }
    float* bias = reinterpret_cast<float*>(packed + (size_t)(ff / CH) * CHUNK_BYTES);
    if (c == 0) {
        for (int e = threadIdx.x; e < ff; e += blockDim.x) bias[e] = b1[e];
        for (int e = threadIdx.x; e < DP; e += blockDim.x) bias[ff + e] = (e < d) ? b2[e] : 0.0f;
    }
}

// ---------------------------------------------------------------------------------------------
// forward kernel
// ---------------------------------------------------------------------------------------------
struct __align__(8) FwdBars {
    uint64_t w_full[STAGES], w_empty[STAGES];
    uint64_t x_full[2], x_free[2], s_full[2], h_full[2], h_free[2], y_full[2], y_free[2];
};

__global__ void __launch_bounds__(kThreads, 1) ffn_tc_fwd_kernel(const FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sX = smem;                                    // 2 x 16 KB
    uint8_t* sW = smem + 2 * 16384;                        // STAGES x 32 KB
    float* sB1 = reinterpret_cast<float*>(sW + STAGES * FWD_BLOCK);   // ff floats
    float* sB2 = sB1 + p.ff;                               // 64 floats
    float* sG = sB2 + DP;                                  // gamma, beta (2 x 64)
    __shared__ FwdBars bars;
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NC = p.ff / CH;
    const int64_t n_pairs = (p.M + 2 * TM - 1) / (2 * TM);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(&bars.w_full[s], 1);
            tc::mbar_init(&bars.w_empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&bars.x_full[i], 128);
            tc::mbar_init(&bars.x_free[i], 1);
            tc::mbar_init(&bars.s_full[i], 1);
            tc::mbar_init(&bars.h_full[i], 128);
            tc::mbar_init(&bars.h_free[i], 1);
            tc::mbar_init(&bars.y_full[i], 1);
            tc::mbar_init(&bars.y_free[i], 128);
        }
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc<512>(&tmem_slot);
    {   // biases / LayerNorm affine into shared memory
        const float* b1g = packed_b1(p.packed, p.ff);
        for (int e = threadIdx.x; e < p.ff + DP; e += kThreads) sB1[e] = b1g[e];
        for (int e = threadIdx.x; e < DP; e += kThreads) {
            sG[e] = (e < p.d) ? p.gamma[e] : 0.0f;
            sG[DP + e] = (e < p.d) ? p.beta[e] : 0.0f;
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        // ================= weight producer =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
                for (int c = 0; c < NC; ++c, ++it) {
                    const uint32_t s = it % STAGES, n = it / STAGES;
                    if (n > 0) tc::mbar_wait(&bars.w_empty[s], (n - 1) & 1);
                    tc::mbar_arrive_expect_tx(&bars.w_full[s], FWD_BLOCK);
                    tc::bulk_g2s(sW + s * FWD_BLOCK, p.packed + (size_t)c * CHUNK_BYTES, FWD_BLOCK, &bars.w_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc1 = tc::make_idesc(TM, CH, 0, 0);
            const uint32_t idesc2 = tc::make_idesc(TM, DP, 0, 0);
            uint32_t it = 0;        // global chunk counter (weights ring)
            uint32_t q = 0;         // pair counter
            uint32_t hcount[2] = {0, 0};
            for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++q) {
                auto g1 = [&](int i, uint32_t chunk_it) {
                    const uint32_t s = chunk_it % STAGES;
                    const uint32_t a0 = tc::smem_u32(sX + i * 16384), b0 = tc::smem_u32(sW + s * FWD_BLOCK + W2_BYTES);
#pragma unroll
                    for (int ks = 0; ks < DP / 16; ++ks)
                        tc::mma_ss(tmem + 128 + 128 * i, tc::make_desc_sw128(a0 + ks * 32, 16, 1024),
                                   tc::make_desc_sw128(b0 + ks * 32, 16, 1024), idesc1, ks > 0);
                    tc::mma_commit(&bars.s_full[i]);
                };
                // prologue: GEMM1 of chunk 0 for both tiles
                tc::mbar_wait(&bars.w_full[it % STAGES], (it / STAGES) & 1);
                for (int i = 0; i < 2; ++i) {
                    tc::mbar_wait(&bars.x_full[i], q & 1);
                    tc::tc_fence_after();
                    g1(i, it);
                }
                for (int c = 0; c < NC; ++c, ++it) {
                    const uint32_t s = it % STAGES;
                    if (c + 1 < NC) tc::mbar_wait(&bars.w_full[(it + 1) % STAGES], ((it + 1) / STAGES) & 1);
                    for (int i = 0; i < 2; ++i) {
                        tc::mbar_wait(&bars.h_full[i], hcount[i] & 1);   // H_i(c) in TMEM, S_i consumed
                        ++hcount[i];
                        if (c == 0 && q > 0) tc::mbar_wait(&bars.y_free[i], (q - 1) & 1);
                        tc::tc_fence_after();
                        const uint32_t b0 = tc::smem_u32(sW + s * FWD_BLOCK);
#pragma unroll
                        for (int ks = 0; ks < CH / 16; ++ks)
                            tc::mma_ts(tmem + 64 * i, tmem + 384 + 64 * i + ks * 8,
                                       tc::make_desc_sw128(b0 + (ks >> 2) * 8192 + (ks & 3) * 32, 16, 1024), idesc2,
                                       (c > 0 || ks > 0));
                        tc::mma_commit(&bars.h_free[i]);
                        if (c == NC - 1) tc::mma_commit(&bars.y_full[i]);
                        if (c + 1 < NC) {
                            g1(i, it + 1);
                            if (c + 2 == NC) tc::mma_commit(&bars.x_free[i]);   // last GEMM1 of this pair read X_i
                        }
                    }
                    if (NC == 1) { tc::mma_commit(&bars.x_free[0]); tc::mma_commit(&bars.x_free[1]); }
                    tc::mma_commit(&bars.w_empty[s]);                 // chunk c weights fully consumed
                }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue groups =================
        const int i = (warp - 4) >> 2;                  // tile within the pair
        const int wq = warp & 3;                        // TMEM lane quarter
        const int tg = (warp - 4 - 4 * i) * 32 + lane;  // thread index within the group (0..127) == row in tile
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
        const float scale2_unused = 1.0f;
        (void)scale2_unused;
        uint32_t q = 0, scount = 0, hfree_count = 0;
        for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++q) {
            const int64_t row0 = pair * (2 * TM) + (int64_t)i * TM;
            // ---- (a) X tile: fp32 rows -> bf16 K-major swizzled tile
            if (q > 0) tc::mbar_wait(&bars.x_free[i], (q - 1) & 1);
            {
                uint8_t* xt = sX + i * 16384;
                if (p.d == DP) {
                    float4 v[16];           // 16 independent 128-bit loads in flight per thread
#pragma unroll
                    for (int itx = 0; itx < 16; ++itx) {
                        const int e = itx * 128 + tg;
                        const int r = e >> 4, c4 = e & 15;
                        v[itx] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (row0 + r < p.M) v[itx] = __ldg(reinterpret_cast<const float4*>(p.y1 + (row0 + r) * DP) + c4);
                    }
#pragma unroll
                    for (int itx = 0; itx < 16; ++itx) {
                        const int e = itx * 128 + tg;
                        const int r = e >> 4, c4 = e & 15;
                        uint2 w;
                        w.x = tc::pack_bf16(v[itx].x, v[itx].y);
                        w.y = tc::pack_bf16(v[itx].z, v[itx].w);
                        *reinterpret_cast<uint2*>(xt + tc::sw128_offset(r, c4 * 4)) = w;
                    }
                } else {
                    for (int e = tg; e < TM * DP; e += 128) {
                        const int r = e / DP, k = e % DP;
                        const float v = (row0 + r < p.M && k < p.d) ? p.y1[(row0 + r) * p.d + k] : 0.0f;
                        *reinterpret_cast<__nv_bfloat16*>(xt + tc::sw128_offset(r, k)) = __float2bfloat16(v);
                    }
                }
                tc::fence_proxy_async();
                tc::mbar_arrive(&bars.x_full[i]);
            }
            const int64_t row = row0 + tg;
            // ---- (b) per chunk: S -> H
            for (int c = 0; c < NC; ++c) {
                tc::mbar_wait(&bars.s_full[i], scount & 1);
                ++scount;
                tc::tc_fence_after();
                uint32_t hp[64];
#pragma unroll
                for (int pc = 0; pc < 4; ++pc) {
                    uint32_t v[32];
                    tc::tmem_ld32(tmem + lane_base + 128 + 128 * i + 32 * pc, v);
                    uint32_t keep = 0xFFFFFFFFu;
                    if (p.thr) keep = rng_keep_word(p.keys2, (uint64_t)row * (uint64_t)(p.ff >> 5) + (uint64_t)(4 * c + pc), p.thr);
                    const float* bb = sB1 + c * CH + 32 * pc;
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        float a = fmaxf(__uint_as_float(v[j]) + bb[j], 0.0f);
                        float b = fmaxf(__uint_as_float(v[j + 1]) + bb[j + 1], 0.0f);
                        a = ((keep >> j) & 1u) ? a : 0.0f;
                        b = ((keep >> (j + 1)) & 1u) ? b : 0.0f;
                        hp[pc * 16 + (j >> 1)] = tc::pack_bf16(a, b);
                    }
                }
                if (hfree_count > 0) tc::mbar_wait(&bars.h_free[i], (hfree_count - 1) & 1);   // GEMM2 of the previous chunk done with H_i
                ++hfree_count;
                {
                    uint32_t lo[32], hi[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        lo[j] = hp[j];
                        hi[j] = hp[32 + j];
                    }
                    tc::tmem_st32(tmem + lane_base + 384 + 64 * i, lo);
                    tc::tmem_st32(tmem + lane_base + 384 + 64 * i + 32, hi);
                }
                tc::tmem_st_wait();
                tc::tc_fence_before();
                tc::mbar_arrive(&bars.h_full[i]);
            }
            // ---- (c) Y -> z, LayerNorm statistics, xnext
            tc::mbar_wait(&bars.y_full[i], q & 1);
            tc::tc_fence_after();
            {
                uint32_t y0[32], y1r[32];
                tc::tmem_ld32(tmem + lane_base + 64 * i, y0);
                tc::tmem_ld32(tmem + lane_base + 64 * i + 32, y1r);
                tc::tmem_ld_wait();
                tc::tc_fence_before();
                tc::mbar_arrive(&bars.y_free[i]);
                if (row < p.M) {
                    float zv[DP];
                    uint32_t k0 = 0xFFFFFFFFu, k1 = 0xFFFFFFFFu;
                    const bool fast = (p.d == DP);
                    if (p.thr && fast) {
                        k0 = rng_keep_word(p.keys3, (uint64_t)row * 2ull, p.thr);
                        k1 = rng_keep_word(p.keys3, (uint64_t)row * 2ull + 1ull, p.thr);
                    }
                    float sum = 0.0f;
#pragma unroll
                    for (int j = 0; j < DP; ++j) {
                        float f = __uint_as_float(j < 32 ? y0[j] : y1r[j - 32]) + sB2[j];
                        float mult;
                        if (fast) mult = (((j < 32 ? k0 : k1) >> (j & 31)) & 1u) ? p.scale3 : 0.0f;
                        else mult = (j < p.d) ? rng_dropout_mult(p.keys3, (uint64_t)row * (uint64_t)p.d + (uint64_t)j, p.thr, p.scale3) : 0.0f;
                        if (!p.thr) mult = 1.0f;
                        zv[j] = (j < p.d) ? f * mult : 0.0f;
                    }
                    if (fast) {   // residual: 128-bit loads of the fp32 input row
                        const float4* rr = reinterpret_cast<const float4*>(p.y1 + row * DP);
#pragma unroll
                        for (int j = 0; j < DP; j += 4) {
                            const float4 r4 = __ldg(rr + (j >> 2));
                            zv[j] += r4.x; zv[j + 1] += r4.y; zv[j + 2] += r4.z; zv[j + 3] += r4.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < DP; ++j)
                            if (j < p.d) zv[j] += __ldg(p.y1 + row * p.d + j);
                    }
#pragma unroll
                    for (int j = 0; j < DP; ++j) sum += zv[j];
                    const float inv_d = 1.0f / (float)p.d;
                    const float mean = sum * inv_d;
                    float sq = 0.0f;
#pragma unroll
                    for (int j = 0; j < DP; ++j) {
                        const float tt = (j < p.d) ? zv[j] - mean : 0.0f;
                        sq = fmaf(tt, tt, sq);
                    }
                    const float rstd = rsqrtf(sq * inv_d + 1e-5f);
                    if (p.stats) {
                        p.stats[2 * row] = mean;
                        p.stats[2 * row + 1] = rstd;
                    }
                    if (fast) {
                        float4* zo = reinterpret_cast<float4*>(p.z + row * DP);
                        float4* xo = p.xnext ? reinterpret_cast<float4*>(p.xnext + row * DP) : nullptr;
#pragma unroll
                        for (int j = 0; j < DP; j += 4) {
                            zo[j >> 2] = make_float4(zv[j], zv[j + 1], zv[j + 2], zv[j + 3]);
                            if (xo)
                                xo[j >> 2] = make_float4((zv[j] - mean) * rstd * sG[j] + sG[DP + j],
                                                         (zv[j + 1] - mean) * rstd * sG[j + 1] + sG[DP + j + 1],
                                                         (zv[j + 2] - mean) * rstd * sG[j + 2] + sG[DP + j + 2],
                                                         (zv[j + 3] - mean) * rstd * sG[j + 3] + sG[DP + j + 3]);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < DP; ++j)
                            if (j < p.d) {
                                p.z[row * p.d + j] = zv[j];
                                if (p.xnext) p.xnext[row * p.d + j] = (zv[j] - mean) * rstd * sG[j] + sG[DP + j];
                            }
                    }
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc<512>(tmem);
}

size_t packed_bytes(int ff) { return (size_t)(ff / CH) * CHUNK_BYTES + (size_t)(ff + DP) * sizeof(float); }

}  // namespace

extern "C" size_t u2gnn_ffn_tc_packed_bytes(int d, int ff) {
    if (d < 1 || d > DP || ff < CH || ff % CH) return 0;
    return packed_bytes(ff);
}

extern "C" int u2gnn_ffn_tc_prepare(const float* W1, const float* b1, const float* W2, const float* b2, int d, int ff,
                                    float hidden_scale, void* packed, size_t packed_size, u2gnn_stream_t stream) {
    if (!W1 || !b1 || !W2 || !b2 || !packed) return U2GNN_EINVAL;
    if (d < 1 || d > DP || ff < CH || ff % CH) return U2GNN_EUNSUPPORTED;
    if (packed_size < packed_bytes(ff)) return U2GNN_EWORKSPACE;
    if (reinterpret_cast<uintptr_t>(packed) % 128) return U2GNN_EALIGN;
    ffn_pack_kernel<<<ff / CH, 256, 0, as_stream(stream)>>>(W1, b1, W2, b2, d, ff, hidden_scale, static_cast<uint8_t*>(packed));
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_ffn_tc_fwd(const float* y1, int64_t M, int d, int ff, const void* packed, uint64_t seed,
                                uint32_t stream_hidden, uint32_t stream_out, int thr, const float* gamma,
                                const float* beta, float* z, float* stats, float* xnext, u2gnn_stream_t stream) {
    if (!y1 || !packed || !gamma || !beta || !z || M < 0 || thr < 0 || thr > 255) return U2GNN_EINVAL;
    if (d < 1 || d > DP || ff < CH || ff % CH || ff > 8192) return U2GNN_EUNSUPPORTED;
    if (d == DP && ((reinterpret_cast<uintptr_t>(y1) | reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(xnext)) % 16))
        return U2GNN_EALIGN;
    if (M == 0) return U2GNN_OK;
    FwdParams p;
    p.y1 = y1; p.M = M; p.d = d; p.ff = ff;
    p.packed = static_cast<const uint8_t*>(packed);
    p.keys2 = rng_keys(seed, stream_hidden);
    p.keys3 = rng_keys(seed, stream_out);
    p.thr = thr;
    p.scale3 = thr ? rng_keep_scale(thr) : 1.0f;
    p.gamma = gamma; p.beta = beta; p.z = z; p.stats = stats; p.xnext = xnext;
    const size_t smem = 1024 + 2 * 16384 + (size_t)STAGES * FWD_BLOCK + (size_t)(ff + 3 * DP) * sizeof(float);
    if (smem > 227 * 1024) return U2GNN_EUNSUPPORTED;
    cudaFuncSetAttribute(ffn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t n_pairs = (M + 2 * TM - 1) / (2 * TM);
    const int grid = (int)(n_pairs < U2GNN_NUM_SMS ? n_pairs : U2GNN_NUM_SMS);
    ffn_tc_fwd_kernel<<<grid, kThreads, smem, as_stream(stream)>>>(p);
    U2GNN_CHECK_LAUNCH();
}
