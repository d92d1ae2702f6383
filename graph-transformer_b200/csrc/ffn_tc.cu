// Fused bf16 FFN block on the 5th-generation tensor cores (tcgen05 + TMEM), forward.
//
//   z = y1 + dropout3( dropout2(relu(y1 W1^T + b1)) W2^T + b2 ),  stats = LayerNorm statistics of z,
//   xnext = LayerNorm(z) * gamma + beta
// i.e. linear1 -> ReLU -> dropout -> linear2 -> dropout -> residual -> norm2 of
// nn.TransformerEncoderLayer (torch/nn/modules/transformer.py:950-958,977-982).  The [rows, ff]
// hidden activation never leaves the SM: per 128-row tile and 128-wide ff chunk
//     S  = X W1c^T            tcgen05.mma TS  (X bf16 in tensor memory, W1c bulk-copied pre-swizzled image)
//     H  = act(S)             epilogue warps: tcgen05.ld -> bias/ReLU/dropout (packed bf16x2) -> tcgen05.st
//     Y += H W2c^T            tcgen05.mma TS  (H read from tensor memory, aliased over the S columns it came from)
// Persistent CTAs (one per SM) walk pairs of row tiles so that the tensor pipe works on one tile
// while the epilogue warps convert the other.  Weights stream from L2 through a 4-stage
// bulk-copy/mbarrier ring shared by both tiles.
//
// Measured facts behind the structure (tools/probe_mma.py, B200): an MMA issued from divergent single-thread code
// with per-instruction descriptor arithmetic costs ~96-118 cycles regardless of N; issued from warp-uniform code
// with immediate descriptor increments it costs 49 (N=64, TS) / 70 (N=128, TS) / 77 (N=128, SS) cycles.  Hence
// the whole MMA warp runs uniformly (elect.sync only around the instruction) and both GEMMs take A from TMEM.
//
// Warp roles (640 threads): warp 0 weight producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-11 epilogue
// of tile 0, warps 12-19 epilogue of tile 1 (two 64-column halves x four TMEM lane quarters per tile).
// TMEM columns: Y0 [0,64) Y1 [64,128) S0 [128,256) S1 [256,384) X0 [384,416) X1 [416,448);
//               H_i (packed bf16) overwrites S_i columns [0,32) and [64,96).
#include "common.cuh"
#include "rng.cuh"
#include "tc_common.cuh"
#include "ffn_epi.cuh"

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
constexpr int kThreads = 640;     // 4 control warps + 16 epilogue warps
constexpr uint32_t COL_Y = 0, COL_S = 128, COL_X = 384;

struct FwdParams {
    const float* y1;
    int64_t M;
    int d, ff;
    const uint8_t* packed;
    RngKeys keys2, keys3;
    int thr, low;
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
    uint64_t x_full[2], x_free[2], s_full[2], h_full[2], y_full[2], y_free[2];
};

// 4 k-steps of S_i = X_i W1c^T (A = packed X in TMEM, 8 columns per k-step; B = W1c image), warp-uniform
__device__ __forceinline__ void issue_gemm1(uint32_t tmem_s, uint32_t tmem_x, uint64_t b_desc, uint32_t idesc) {
    if (tc::elect_one()) {
        tc::mma_ts(tmem_s, tmem_x, b_desc, idesc, 0);
        tc::mma_ts_acc(tmem_s, tmem_x + 8, b_desc + 2, idesc);
        tc::mma_ts_acc(tmem_s, tmem_x + 16, b_desc + 4, idesc);
        tc::mma_ts_acc(tmem_s, tmem_x + 24, b_desc + 6, idesc);
    }
    __syncwarp();
}
// 8 k-steps of Y_i (+)= H_i W2c^T (A = packed H at S_i columns [0,32) and [64,96); B = two [64 x 64] atoms of W2c)
__device__ __forceinline__ void issue_gemm2(uint32_t tmem_y, uint32_t tmem_h, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    if (tc::elect_one()) {
        tc::mma_ts(tmem_y, tmem_h, b_desc, idesc, acc);
        tc::mma_ts_acc(tmem_y, tmem_h + 8, b_desc + 2, idesc);
        tc::mma_ts_acc(tmem_y, tmem_h + 16, b_desc + 4, idesc);
        tc::mma_ts_acc(tmem_y, tmem_h + 24, b_desc + 6, idesc);
        tc::mma_ts_acc(tmem_y, tmem_h + 64, b_desc + 512, idesc);           // second K atom: +8192 B
        tc::mma_ts_acc(tmem_y, tmem_h + 72, b_desc + 514, idesc);
        tc::mma_ts_acc(tmem_y, tmem_h + 80, b_desc + 516, idesc);
        tc::mma_ts_acc(tmem_y, tmem_h + 88, b_desc + 518, idesc);
    }
    __syncwarp();
}
__device__ __forceinline__ void commit_to(uint64_t* bar) {
    if (tc::elect_one()) tc::mma_commit(bar);
    __syncwarp();
}

__global__ void __launch_bounds__(kThreads, 1) ffn_tc_fwd_kernel(const FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sW = smem;                                    // STAGES x 32 KB
    uint32_t* sB1h = reinterpret_cast<uint32_t*>(sW + STAGES * FWD_BLOCK);   // b1 as packed bf16 pairs (ff/2 words)
    float* sB2 = reinterpret_cast<float*>(sB1h + p.ff / 2);                  // 64 floats
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
            tc::mbar_init(&bars.x_full[i], 4);      // one arrival per warp of the loading warpgroup
            tc::mbar_init(&bars.x_free[i], 1);
            tc::mbar_init(&bars.s_full[i], 1);
            tc::mbar_init(&bars.h_full[i], 8);      // one arrival per epilogue warp of the tile
            tc::mbar_init(&bars.y_full[i], 1);
            tc::mbar_init(&bars.y_free[i], 4);
        }
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc<512>(&tmem_slot);
    {   // biases / LayerNorm affine into shared memory
        const float* b1g = packed_b1(p.packed, p.ff);
        for (int e = threadIdx.x; e < p.ff / 2; e += kThreads) sB1h[e] = epi::cvt2(b1g[2 * e], b1g[2 * e + 1]);
        for (int e = threadIdx.x; e < DP; e += kThreads) {
            sB2[e] = b1g[p.ff + e];
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
        // ================= MMA issuer: the whole warp runs this code uniformly =================
        const uint32_t idesc1 = tc::make_idesc(TM, CH, 0, 0);
        const uint32_t idesc2 = tc::make_idesc(TM, DP, 0, 0);
        const uint64_t w_desc0 = tc::make_desc_sw128(tc::smem_u32(sW), 16, 1024);     // stage 0, W2c image
        uint32_t it = 0, q = 0;
        uint32_t hcount[2] = {0, 0};
        for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++q) {
            // prologue: GEMM1 of chunk 0 for both tiles
            tc::mbar_wait(&bars.w_full[it % STAGES], (it / STAGES) & 1);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                tc::mbar_wait(&bars.x_full[i], q & 1);
                tc::tc_fence_after();
                issue_gemm1(tmem + COL_S + 128 * i, tmem + COL_X + 32 * i,
                            w_desc0 + (uint64_t)((it % STAGES) * (FWD_BLOCK >> 4) + (W2_BYTES >> 4)), idesc1);
                commit_to(&bars.s_full[i]);
            }
            for (int c = 0; c < NC; ++c, ++it) {
                const uint32_t s = it % STAGES;
                const uint64_t w2_desc = w_desc0 + (uint64_t)(s * (FWD_BLOCK >> 4));
                if (c + 1 < NC) tc::mbar_wait(&bars.w_full[(it + 1) % STAGES], ((it + 1) / STAGES) & 1);
                const uint64_t w1_next = w_desc0 + (uint64_t)(((it + 1) % STAGES) * (FWD_BLOCK >> 4) + (W2_BYTES >> 4));
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    tc::mbar_wait(&bars.h_full[i], hcount[i] & 1);   // H_i(c) in TMEM (over S_i)
                    ++hcount[i];
                    if (c == 0 && q > 0) tc::mbar_wait(&bars.y_free[i], (q - 1) & 1);
                    tc::tc_fence_after();
                    issue_gemm2(tmem + COL_Y + 64 * i, tmem + COL_S + 128 * i, w2_desc, idesc2, c > 0);
                    if (c == NC - 1) commit_to(&bars.y_full[i]);
                    if (c + 1 < NC) {
                        // the tensor pipe executes MMAs in issue order: this GEMM1 overwrites S_i (and the H_i aliased
                        // over it) only after the GEMM2 above has consumed H_i
                        issue_gemm1(tmem + COL_S + 128 * i, tmem + COL_X + 32 * i, w1_next, idesc1);
                        commit_to(&bars.s_full[i]);
                    } else {
                        commit_to(&bars.x_free[i]);                  // last GEMM1 of this pair has read X_i
                    }
                }
                commit_to(&bars.w_empty[s]);                         // chunk c weights fully consumed
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue groups: 8 warps per tile = two 64-column halves x four TMEM lane quarters =====
        const int i = (warp - 4) >> 3;                  // tile within the pair
        const int wg = ((warp - 4) >> 2) & 1;           // column half of the 128-wide chunk
        const int wq = warp & 3;                        // TMEM lane quarter
        const int tr = wq * 32 + lane;                  // row in tile
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
        uint32_t q = 0, scount = 0;
        for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++q) {
            const int64_t row0 = pair * (2 * TM) + (int64_t)i * TM;
            const int64_t row = row0 + tr;
            // ---- (a) X tile: the first warpgroup converts one fp32 row per thread to packed bf16 in tensor memory
            if (wg == 0) {
                if (q > 0) tc::mbar_wait(&bars.x_free[i], (q - 1) & 1);
                uint32_t xp[32];
                if (p.d == DP && row < p.M) {
                    const float4* src = reinterpret_cast<const float4*>(p.y1 + row * DP);
                    float4 v[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) v[u] = __ldg(src + u);
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        xp[2 * u] = epi::cvt2(v[u].x, v[u].y);
                        xp[2 * u + 1] = epi::cvt2(v[u].z, v[u].w);
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < 32; ++u) {
                        const float a = (row < p.M && 2 * u < p.d) ? p.y1[row * p.d + 2 * u] : 0.0f;
                        const float b = (row < p.M && 2 * u + 1 < p.d) ? p.y1[row * p.d + 2 * u + 1] : 0.0f;
                        xp[u] = epi::cvt2(a, b);
                    }
                }
                tc::tmem_st32(tmem + lane_base + COL_X + 32 * i, xp);
                tc::tmem_st_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars.x_full[i]);
            }
            // ---- (b) per chunk: this thread turns 64 columns of its S row into packed bf16 H (written over them)
            for (int c = 0; c < NC; ++c) {
                tc::mbar_wait(&bars.s_full[i], scount & 1);
                ++scount;
                tc::tc_fence_after();
                const uint32_t s_addr = tmem + lane_base + COL_S + 128 * i + 64 * wg;
                uint32_t v0[32], v1[32];
                tc::tmem_ld32(s_addr, v0);                      // both 32-column pieces in flight, one wait
                tc::tmem_ld32(s_addr + 32, v1);
                uint32_t k0 = 0xFFFFFFFFu, k1 = 0xFFFFFFFFu;
                if (p.thr) {
                    const uint64_t g = (uint64_t)row * (uint64_t)(p.ff >> 5) + (uint64_t)(4 * c + 2 * wg);
                    k0 = rng_keep_word_lo(p.keys2, g, p.thr, p.low);
                    k1 = rng_keep_word_lo(p.keys2, g + 1, p.thr, p.low);
                }
                const uint4* bb = reinterpret_cast<const uint4*>(sB1h + ((c * CH + 64 * wg) >> 1));
                tc::tmem_ld_wait();
                {
                    uint32_t km[16];
                    if (p.thr) epi::keep_masks16(k0, km);
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const uint4 b4 = bb[q4];
                        const uint32_t bw[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int j = 4 * q4 + u;
                            uint32_t h2 = epi::relu_bias2(epi::cvt2(__uint_as_float(v0[2 * j]), __uint_as_float(v0[2 * j + 1])), bw[u]);
                            if (p.thr) h2 &= km[j];
                            v0[j] = h2;                            // in place: entries 2j, 2j+1 are already consumed
                        }
                    }
                    if (p.thr) epi::keep_masks16(k1, km);
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const uint4 b4 = bb[4 + q4];
                        const uint32_t bw[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int j = 4 * q4 + u;
                            uint32_t h2 = epi::relu_bias2(epi::cvt2(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1])), bw[u]);
                            if (p.thr) h2 &= km[j];
                            v1[j] = h2;
                        }
                    }
                }
                tc::tmem_st32_2x16(s_addr, v0, v1);    // H columns [64*wg, 64*wg + 32) of the S_i region: only this thread's own S data lived there
                tc::tmem_st_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars.h_full[i]);
            }
            if (wg != 0) continue;                      // the per-pair output epilogue is done by one warpgroup per tile
            // ---- (c) Y -> z, LayerNorm statistics, xnext
            tc::mbar_wait(&bars.y_full[i], q & 1);
            tc::tc_fence_after();
            {
                uint32_t y0[32], y1r[32];
                tc::tmem_ld32(tmem + lane_base + COL_Y + 64 * i, y0);
                tc::tmem_ld32(tmem + lane_base + COL_Y + 64 * i + 32, y1r);
                tc::tmem_ld_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars.y_free[i]);
                if (row < p.M) {
                    float zv[DP];
                    const bool fast = (p.d == DP);
                    if (fast) {
                        uint32_t k0 = 0xFFFFFFFFu, k1 = 0xFFFFFFFFu;
                        if (p.thr) {
                            k0 = rng_keep_word_lo(p.keys3, (uint64_t)row * 2ull, p.thr, p.low);
                            k1 = rng_keep_word_lo(p.keys3, (uint64_t)row * 2ull + 1ull, p.thr, p.low);
                        }
                        const float sc = p.thr ? p.scale3 : 1.0f;
                        const float4* rr = reinterpret_cast<const float4*>(p.y1 + row * DP);
#pragma unroll
                        for (int j = 0; j < DP; j += 4) {
                            const float4 r4 = __ldg(rr + (j >> 2));
                            const float res[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int jj = j + u;
                                const float f = __uint_as_float(jj < 32 ? y0[jj] : y1r[jj - 32]) + sB2[jj];
                                const float mult = (((jj < 32 ? k0 : k1) >> (jj & 31)) & 1u) ? sc : 0.0f;
                                zv[jj] = res[u] + f * mult;
                            }
                        }
                    } else {
#pragma unroll 1
                        for (int j = 0; j < p.d; ++j) {
                            const float f = __uint_as_float(j < 32 ? y0[j] : y1r[j - 32]) + sB2[j];
                            const float mult = rng_dropout_mult(p.keys3, (uint64_t)row * (uint64_t)p.d + (uint64_t)j, p.thr, p.scale3);
                            zv[j] = __ldg(p.y1 + row * p.d + j) + f * mult;
                        }
                        for (int j = p.d; j < DP; ++j) zv[j] = 0.0f;
                    }
                    float sum = 0.0f;
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
#pragma unroll 1
                        for (int j = 0; j < p.d; ++j) {
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

extern "C" int u2gnn_ffn_tc_debug(int flags) {
    (void)flags;                       // experiment switches were removed after the pipeline study (profiles/README.md)
    return U2GNN_OK;
}

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
    p.low = rng_thr_low(thr);
    p.scale3 = thr ? rng_keep_scale(thr) : 1.0f;
    p.gamma = gamma; p.beta = beta; p.z = z; p.stats = stats; p.xnext = xnext;
    const size_t smem = 1024 + (size_t)STAGES * FWD_BLOCK + (size_t)(ff / 2 + 3 * DP) * sizeof(float);
    if (smem > 227 * 1024) return U2GNN_EUNSUPPORTED;
    cudaFuncSetAttribute(ffn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t n_pairs = (M + 2 * TM - 1) / (2 * TM);
    const int grid = (int)(n_pairs < U2GNN_NUM_SMS ? n_pairs : U2GNN_NUM_SMS);
    ffn_tc_fwd_kernel<<<grid, kThreads, smem, as_stream(stream)>>>(p);
    U2GNN_CHECK_LAUNCH();
}
