// Fused bf16 FFN block on tcgen05 + TMEM, backward.  The [rows, ff] hidden activation and its
// gradient are recomputed on chip (never stored in HBM), so the backward is two kernels with opposite
// loop orders (DESIGN.md "FFN backward"):
//
//   dgrad  (rows outer, ff chunks inner)         dy1 = dz + dPre W1
//       S = X W1c^T -> mask = [S+b1 > 0] & keep ; D = dF W2c (scaled) ; dPre = D * mask ; dY += dPre W1c
//   wgrad  (one ff chunk per CTA, row tiles inner)   dW1, db1, dW2
//       H = relu(S+b1) * keep ; dPre as above ; dW2c^T += H^T dF ; dW1c += dPre^T X ; db1c += dPre^T 1
//
// with X = y1 (block input), dF = gradient at the linear2 output (after the output dropout),
// dz = gradient at the residual sum.  These are the autograd of linear1/ReLU/dropout/linear2 in
// nn.TransformerEncoderLayer._ff_block (torch/nn/modules/transformer.py:977-982).
// Operand facts (tc_selftest.cu): a [128 x 64] bf16 swizzled tile serves as K-major A (S, D GEMMs) and
// as MN-major B (weight-gradient GEMMs) without being rewritten.
#include "common.cuh"
#include "rng.cuh"
#include "tc_common.cuh"
#include "ffn_epi.cuh"

namespace {

constexpr int DP = 64, CH = 128, TM = 128;
constexpr uint32_t CHUNK_BYTES = 4 * 16384;   // [W2c | W1c | W2Tc | W1Tc]
constexpr int kThreads = 384;

__device__ __forceinline__ const float* packed_b1(const uint8_t* packed, int ff) {
    return reinterpret_cast<const float*>(packed + (size_t)(ff / CH) * CHUNK_BYTES);
}

// fp32 rows [row0, row0+128) of src[M, d] -> bf16 [128 x 64] swizzled tile (zero padded); 128 threads
__device__ __forceinline__ void load_tile_bf16(uint8_t* tile, const float* __restrict__ src, int64_t row0, int64_t M, int d,
                                               int tg) {
    if (d == DP) {
        // all 16 loads of a thread are issued before the first use (latency paid once per tile)
        float4 v[16];
#pragma unroll
        for (int itx = 0; itx < 16; ++itx) {
            const int e = itx * 128 + tg;
            const int r = e >> 4, c4 = e & 15;
            v[itx] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + r < M) v[itx] = __ldg(reinterpret_cast<const float4*>(src + (row0 + r) * DP) + c4);
        }
#pragma unroll
        for (int itx = 0; itx < 16; ++itx) {
            const int e = itx * 128 + tg;
            const int r = e >> 4, c4 = e & 15;
            uint2 w;
            w.x = tc::pack_bf16(v[itx].x, v[itx].y);
            w.y = tc::pack_bf16(v[itx].z, v[itx].w);
            *reinterpret_cast<uint2*>(tile + tc::sw128_offset(r, c4 * 4)) = w;
        }
    } else {
        for (int e = tg; e < TM * DP; e += 128) {
            const int r = e / DP, k = e % DP;
            const float v = (row0 + r < M && k < d) ? src[(row0 + r) * d + k] : 0.0f;
            *reinterpret_cast<__nv_bfloat16*>(tile + tc::sw128_offset(r, k)) = __float2bfloat16(v);
        }
    }
}

// ============================================================================================
// dgrad
// ============================================================================================
constexpr int DG_STAGES = 3;
constexpr uint32_t DG_BLOCK = 3 * 16384;      // [W1c | W2Tc | W1Tc]

struct DgradParams {
    const float* y1;
    const float* df;
    const float* dz;
    float* dy1;
    int64_t M;
    int d, ff;
    const uint8_t* packed;
    RngKeys keys2;
    int thr;
};

struct __align__(8) DgradBars {
    uint64_t w_full[DG_STAGES], w_empty[DG_STAGES];
    uint64_t x_full[2], x_free[2], s_full[2], a_done[2], d_full[2], p_full[2], p_free[2], y_full[2], y_free[2];
};

// TMEM: dY0 [0,64) dY1 [64,128) R0 [128,256) R1 [256,384) P0 [384,448) P1 [448,512)
__global__ void __launch_bounds__(kThreads, 1) ffn_tc_dgrad_kernel(const DgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sX = smem;                          // 2 x 16 KB
    uint8_t* sF = smem + 2 * 16384;              // 2 x 16 KB
    uint8_t* sW = smem + 4 * 16384;              // DG_STAGES x 48 KB
    uint32_t* sB1h = reinterpret_cast<uint32_t*>(sW + DG_STAGES * DG_BLOCK);   // b1 as packed bf16 pairs
    __shared__ DgradBars bars;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NC = p.ff / CH;
    const int64_t n_pairs = (p.M + 2 * TM - 1) / (2 * TM);

    if (threadIdx.x == 0) {
        for (int s = 0; s < DG_STAGES; ++s) {
            tc::mbar_init(&bars.w_full[s], 1);
            tc::mbar_init(&bars.w_empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&bars.x_full[i], 128);
            tc::mbar_init(&bars.x_free[i], 1);
            tc::mbar_init(&bars.s_full[i], 1);
            tc::mbar_init(&bars.a_done[i], 128);
            tc::mbar_init(&bars.d_full[i], 1);
            tc::mbar_init(&bars.p_full[i], 128);
            tc::mbar_init(&bars.p_free[i], 1);
            tc::mbar_init(&bars.y_full[i], 1);
            tc::mbar_init(&bars.y_free[i], 128);
        }
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc<512>(&tmem_slot);
    {
        const float* b1g = packed_b1(p.packed, p.ff);
        for (int e = threadIdx.x; e < p.ff / 2; e += kThreads) sB1h[e] = epi::cvt2(b1g[2 * e], b1g[2 * e + 1]);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
                for (int c = 0; c < NC; ++c, ++it) {
                    const uint32_t s = it % DG_STAGES, n = it / DG_STAGES;
                    if (n > 0) tc::mbar_wait(&bars.w_empty[s], (n - 1) & 1);
                    tc::mbar_arrive_expect_tx(&bars.w_full[s], DG_BLOCK);
                    tc::bulk_g2s(sW + s * DG_BLOCK, p.packed + (size_t)c * CHUNK_BYTES + 16384, DG_BLOCK, &bars.w_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_n128 = tc::make_idesc(TM, CH, 0, 0);
            const uint32_t idesc_n64 = tc::make_idesc(TM, DP, 0, 0);
            uint32_t it = 0, q = 0;
            uint32_t acount[2] = {0, 0}, pcount[2] = {0, 0};
            for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++q) {
                // S_i = X_i W1c^T   (and, with a = sF / b = +16384, D_i = dF_i W2Tc^T)
                auto gemm_n128 = [&](int i, uint32_t chunk_it, const uint8_t* a_tile, uint32_t b_off, uint64_t* bar) {
                    const uint32_t s = chunk_it % DG_STAGES;
                    const uint32_t a0 = tc::smem_u32(a_tile + i * 16384), b0 = tc::smem_u32(sW + s * DG_BLOCK + b_off);
#pragma unroll
                    for (int ks = 0; ks < DP / 16; ++ks)
                        tc::mma_ss(tmem + 128 + 128 * i, tc::make_desc_sw128(a0 + ks * 32, 16, 1024),
                                   tc::make_desc_sw128(b0 + ks * 32, 16, 1024), idesc_n128, ks > 0);
                    tc::mma_commit(bar);
                };
                tc::mbar_wait(&bars.w_full[it % DG_STAGES], (it / DG_STAGES) & 1);
                for (int i = 0; i < 2; ++i) {
                    tc::mbar_wait(&bars.x_full[i], q & 1);
                    tc::tc_fence_after();
                    gemm_n128(i, it, sX, 0, &bars.s_full[i]);
                }
                for (int c = 0; c < NC; ++c, ++it) {
                    const uint32_t s = it % DG_STAGES;
                    for (int i = 0; i < 2; ++i) {
                        tc::mbar_wait(&bars.a_done[i], acount[i] & 1);     // epilogue has turned S_i into mask bits
                        ++acount[i];
                        tc::tc_fence_after();
                        gemm_n128(i, it, sF, 16384, &bars.d_full[i]);
                        if (c == NC - 1) tc::mma_commit(&bars.x_free[i]);
                    }
                    if (c + 1 < NC) tc::mbar_wait(&bars.w_full[(it + 1) % DG_STAGES], ((it + 1) / DG_STAGES) & 1);
                    for (int i = 0; i < 2; ++i) {
                        tc::mbar_wait(&bars.p_full[i], pcount[i] & 1);     // dPre_i in TMEM, D_i consumed
                        ++pcount[i];
                        if (c == 0 && q > 0) tc::mbar_wait(&bars.y_free[i], (q - 1) & 1);
                        tc::tc_fence_after();
                        const uint32_t b0 = tc::smem_u32(sW + s * DG_BLOCK + 32768);
#pragma unroll
                        for (int ks = 0; ks < CH / 16; ++ks)
                            tc::mma_ts(tmem + 64 * i, tmem + 384 + 64 * i + ks * 8,
                                       tc::make_desc_sw128(b0 + (ks >> 2) * 8192 + (ks & 3) * 32, 16, 1024), idesc_n64,
                                       (c > 0 || ks > 0));
                        tc::mma_commit(&bars.p_free[i]);
                        if (c == NC - 1) tc::mma_commit(&bars.y_full[i]);
                        if (c + 1 < NC) gemm_n128(i, it + 1, sX, 0, &bars.s_full[i]);
                    }
                    tc::mma_commit(&bars.w_empty[s]);
                }
            }
        }
    } else if (warp >= 4) {
        const int i = (warp - 4) >> 2;
        const int wq = warp & 3;
        const int tg = (warp - 4 - 4 * i) * 32 + lane;
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
        uint32_t q = 0, scount = 0, dcount = 0, pfree_count = 0;
        for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++q) {
            const int64_t row0 = pair * (2 * TM) + (int64_t)i * TM;
            if (q > 0) tc::mbar_wait(&bars.x_free[i], (q - 1) & 1);
            load_tile_bf16(sX + i * 16384, p.y1, row0, p.M, p.d, tg);
            load_tile_bf16(sF + i * 16384, p.df, row0, p.M, p.d, tg);
            tc::fence_proxy_async();
            tc::mbar_arrive(&bars.x_full[i]);
            const int64_t row = row0 + tg;
            for (int c = 0; c < NC; ++c) {
                // ---- S -> activation mask bits (ReLU derivative & hidden-dropout keep)
                tc::mbar_wait(&bars.s_full[i], scount & 1);
                ++scount;
                tc::tc_fence_after();
                uint32_t msk[64];          // per pair: 0xFFFF where relu'(S + b1) = 1 and the hidden dropout kept the unit
                {
                    uint32_t v[2][32];
                    const uint32_t r_addr = tmem + lane_base + 128 + 128 * i;
                    tc::tmem_ld32(r_addr, v[0]);
#pragma unroll
                    for (int pc = 0; pc < 4; ++pc) {
                        tc::tmem_ld_wait();
                        if (pc < 3) tc::tmem_ld32(r_addr + 32 * (pc + 1), v[(pc + 1) & 1]);
                        uint32_t km[16];
                        if (p.thr)
                            epi::keep_masks16(rng_keep_word(p.keys2, (uint64_t)row * (uint64_t)(p.ff >> 5) + (uint64_t)(4 * c + pc), p.thr), km);
                        const uint4* bb = reinterpret_cast<const uint4*>(sB1h + ((c * CH + 32 * pc) >> 1));
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            const uint4 b4 = bb[q4];
                            const uint32_t bw[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int j = 4 * q4 + u;
                                uint32_t m = epi::gt0_mask2(epi::relu_bias2(
                                    epi::cvt2(__uint_as_float(v[pc & 1][2 * j]), __uint_as_float(v[pc & 1][2 * j + 1])), bw[u]));
                                if (p.thr) m &= km[j];
                                msk[pc * 16 + j] = m;
                            }
                        }
                    }
                }
                tc::tc_fence_before();
                tc::mbar_arrive(&bars.a_done[i]);
                // ---- D -> dPre (bf16, tensor memory)
                tc::mbar_wait(&bars.d_full[i], dcount & 1);
                ++dcount;
                tc::tc_fence_after();
                uint32_t hp[64];
                {
                    uint32_t v[2][32];
                    const uint32_t r_addr = tmem + lane_base + 128 + 128 * i;
                    tc::tmem_ld32(r_addr, v[0]);
#pragma unroll
                    for (int pc = 0; pc < 4; ++pc) {
                        tc::tmem_ld_wait();
                        if (pc < 3) tc::tmem_ld32(r_addr + 32 * (pc + 1), v[(pc + 1) & 1]);
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            hp[pc * 16 + j] = epi::cvt2(__uint_as_float(v[pc & 1][2 * j]), __uint_as_float(v[pc & 1][2 * j + 1])) & msk[pc * 16 + j];
                    }
                }
                if (pfree_count > 0) tc::mbar_wait(&bars.p_free[i], (pfree_count - 1) & 1);
                ++pfree_count;
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
                tc::mbar_arrive(&bars.p_full[i]);
            }
            // ---- dY + dz -> dy1
            tc::mbar_wait(&bars.y_full[i], q & 1);
            tc::tc_fence_after();
            uint32_t y0[32], y1r[32];
            tc::tmem_ld32(tmem + lane_base + 64 * i, y0);
            tc::tmem_ld32(tmem + lane_base + 64 * i + 32, y1r);
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            tc::mbar_arrive(&bars.y_free[i]);
            if (row < p.M) {
                if (p.d == DP) {
                    const float4* zi = reinterpret_cast<const float4*>(p.dz + row * DP);
                    float4* o = reinterpret_cast<float4*>(p.dy1 + row * DP);
#pragma unroll
                    for (int j = 0; j < DP; j += 4) {
                        const float4 r4 = zi[j >> 2];
                        const uint32_t* src = (j < 32) ? &y0[j] : &y1r[j - 32];
                        o[j >> 2] = make_float4(r4.x + __uint_as_float(src[0]), r4.y + __uint_as_float(src[1]),
                                                r4.z + __uint_as_float(src[2]), r4.w + __uint_as_float(src[3]));
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < DP; ++j)
                        if (j < p.d) p.dy1[row * p.d + j] = p.dz[row * p.d + j] + __uint_as_float(j < 32 ? y0[j] : y1r[j - 32]);
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc<512>(tmem);
}

// ============================================================================================
// wgrad
// ============================================================================================
struct WgradParams {
    const float* y1;
    const float* df;
    int64_t M;
    int d, ff;
    const uint8_t* packed;
    RngKeys keys2;
    int thr;
    float hidden_scale;
    float* dW1;   // [ff, d]
    float* db1;   // [ff]
    float* dW2;   // [d, ff]
};

constexpr int WG_STAGES = 3;                  // ring of (X, dF) row tiles filled by the loader warps

struct __align__(8) WgradBars {
    uint64_t w_full, ld_full[WG_STAGES], ld_free[WG_STAGES], s_full[2], a_done[2], d_full[2], hp_full, hp_free, flush_full;
};

// fp32 rows -> bf16 swizzled tile with NT loader threads (batches of 16 independent 128-bit loads)
template <int NT>
__device__ __forceinline__ void load_tile_bf16_nt(uint8_t* tile, const float* __restrict__ src, int64_t row0, int64_t M, int d,
                                                  int tid) {
    if (d == DP) {
        for (int base = 0; base < TM * 16; base += NT * 16) {
            float4 v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int e = base + u * NT + tid;
                const int r = e >> 4, c4 = e & 15;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row0 + r < M) v[u] = __ldg(reinterpret_cast<const float4*>(src + (row0 + r) * DP) + c4);
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int e = base + u * NT + tid;
                const int r = e >> 4, c4 = e & 15;
                uint2 w;
                w.x = tc::pack_bf16(v[u].x, v[u].y);
                w.y = tc::pack_bf16(v[u].z, v[u].w);
                *reinterpret_cast<uint2*>(tile + tc::sw128_offset(r, c4 * 4)) = w;
            }
        }
    } else {
        for (int e = tid; e < TM * DP; e += NT) {
            const int r = e / DP, k = e % DP;
            const float v = (row0 + r < M && k < d) ? src[(row0 + r) * d + k] : 0.0f;
            *reinterpret_cast<__nv_bfloat16*>(tile + tc::sw128_offset(r, k)) = __float2bfloat16(v);
        }
    }
}

// TMEM: R0 [0,128) R1 [128,256) dW1c [256,320) dW2c^T [320,384) db1c [384,400)
// Warp roles: 0 weights, 1 MMA issuer, 2-3 row-tile loaders (+ TMEM allocation), 4-7 / 8-11 epilogue of even / odd tiles.
__global__ void __launch_bounds__(kThreads, 1) ffn_tc_wgrad_kernel(const WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sXF = smem;                                   // WG_STAGES x (X 16 KB | dF 16 KB)
    uint8_t* sH = smem + WG_STAGES * 32768;                // 32 KB: two [128 rows x 64 hidden] tiles
    uint8_t* sP = sH + 32768;                              // 32 KB
    uint8_t* sW = sP + 32768;                              // 32 KB: [W1c | W2Tc]
    uint8_t* sOnes = sW + 32768;                           // 2 KB of bf16 1.0 (layout-invariant B operand)
    uint32_t* sB1h = reinterpret_cast<uint32_t*>(sOnes + 2048);   // chunk bias as 64 packed bf16 pairs
    __shared__ WgradBars bars;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NC = p.ff / CH;
    const int c = blockIdx.x % NC;
    const int slice = blockIdx.x / NC, n_slices = gridDim.x / NC;
    const int64_t n_tiles = (p.M + TM - 1) / TM;
    const int64_t my_tiles = (n_tiles > slice) ? (n_tiles - slice + n_slices - 1) / n_slices : 0;

    if (threadIdx.x == 0) {
        tc::mbar_init(&bars.w_full, 1);
        for (int s = 0; s < WG_STAGES; ++s) {
            tc::mbar_init(&bars.ld_full[s], 64);
            tc::mbar_init(&bars.ld_free[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&bars.s_full[i], 1);
            tc::mbar_init(&bars.a_done[i], 128);
            tc::mbar_init(&bars.d_full[i], 1);
        }
        tc::mbar_init(&bars.hp_full, 128);
        tc::mbar_init(&bars.hp_free, 1);
        tc::mbar_init(&bars.flush_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc<512>(&tmem_slot);
    {
        const float* b1g = packed_b1(p.packed, p.ff) + c * CH;
        for (int e = threadIdx.x; e < CH / 2; e += kThreads) sB1h[e] = epi::cvt2(b1g[2 * e], b1g[2 * e + 1]);
        for (int e = threadIdx.x; e < 2048 / 4; e += kThreads) reinterpret_cast<uint32_t*>(sOnes)[e] = 0x3F803F80u;
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (my_tiles > 0) {
        if (warp == 0) {
            if (lane == 0) {
                tc::mbar_arrive_expect_tx(&bars.w_full, 32768);
                tc::bulk_g2s(sW, p.packed + (size_t)c * CHUNK_BYTES + 16384, 32768, &bars.w_full);
            }
        } else if (warp == 2 || warp == 3) {
            // ================= row-tile loaders: global fp32 -> bf16 swizzled tiles, WG_STAGES ahead =================
            const int tid = (warp - 2) * 32 + lane;
            for (int64_t n = 0; n < my_tiles; ++n) {
                const uint32_t s = (uint32_t)(n % WG_STAGES), u = (uint32_t)(n / WG_STAGES);
                if (u > 0) tc::mbar_wait(&bars.ld_free[s], (u - 1) & 1);
                const int64_t row0 = ((int64_t)slice + n * n_slices) * TM;
                load_tile_bf16_nt<64>(sXF + s * 32768, p.y1, row0, p.M, p.d, tid);
                load_tile_bf16_nt<64>(sXF + s * 32768 + 16384, p.df, row0, p.M, p.d, tid);
                tc::fence_proxy_async();
                tc::mbar_arrive(&bars.ld_full[s]);
            }
        } else if (warp == 1) {
            if (lane == 0) {
                const uint32_t idesc_n128 = tc::make_idesc(TM, CH, 0, 0);
                const uint32_t idesc_w = tc::make_idesc(CH, DP, 1, 1);      // M = hidden, N = d, both MN-major
                const uint32_t idesc_b = tc::make_idesc(CH, 16, 1, 0);      // B = ones, K-major [16 x 16]
                auto gemm_n128 = [&](int i, const uint8_t* a_tile, uint32_t b_off, uint64_t* bar) {
                    const uint32_t a0 = tc::smem_u32(a_tile), b0 = tc::smem_u32(sW + b_off);
#pragma unroll
                    for (int ks = 0; ks < DP / 16; ++ks)
                        tc::mma_ss(tmem + 128 * i, tc::make_desc_sw128(a0 + ks * 32, 16, 1024),
                                   tc::make_desc_sw128(b0 + ks * 32, 16, 1024), idesc_n128, ks > 0);
                    tc::mma_commit(bar);
                };
                tc::mbar_wait(&bars.w_full, 0);
                tc::mbar_wait(&bars.ld_full[0], 0);
                tc::tc_fence_after();
                gemm_n128(0, sXF, 0, &bars.s_full[0]);
                for (int64_t n = 0; n < my_tiles; ++n) {
                    const int i = (int)(n & 1);
                    const uint32_t ph = (uint32_t)(n >> 1) & 1;      // per-parity barriers complete every second tile
                    const uint8_t* st = sXF + (n % WG_STAGES) * 32768;
                    tc::mbar_wait(&bars.a_done[i], ph);
                    tc::tc_fence_after();
                    gemm_n128(i, st + 16384, 16384, &bars.d_full[i]);
                    if (n + 1 < my_tiles) {
                        const uint32_t s2 = (uint32_t)((n + 1) % WG_STAGES);
                        tc::mbar_wait(&bars.ld_full[s2], (uint32_t)((n + 1) / WG_STAGES) & 1);
                        tc::tc_fence_after();
                        gemm_n128((int)((n + 1) & 1), sXF + s2 * 32768, 0, &bars.s_full[(n + 1) & 1]);
                    }
                    tc::mbar_wait(&bars.hp_full, (uint32_t)n & 1);
                    tc::tc_fence_after();
                    const uint32_t h0 = tc::smem_u32(sH), p0 = tc::smem_u32(sP);
                    const uint32_t x0 = tc::smem_u32(st), f0 = tc::smem_u32(st + 16384), o0 = tc::smem_u32(sOnes);
#pragma unroll
                    for (int ks = 0; ks < TM / 16; ++ks) {
                        const uint32_t acc = (n > 0 || ks > 0);
                        // dW2c^T[hidden, d] += H^T dF ; dW1c[hidden, d] += dPre^T X ; db1c += dPre^T 1
                        tc::mma_ss(tmem + 320, tc::make_desc_sw128(h0 + ks * 2048, 16384, 1024),
                                   tc::make_desc_sw128(f0 + ks * 2048, 16384, 1024), idesc_w, acc);
                        tc::mma_ss(tmem + 256, tc::make_desc_sw128(p0 + ks * 2048, 16384, 1024),
                                   tc::make_desc_sw128(x0 + ks * 2048, 16384, 1024), idesc_w, acc);
                        tc::mma_ss(tmem + 384, tc::make_desc_sw128(p0 + ks * 2048, 16384, 1024),
                                   tc::make_desc_sw128(o0, 16, 1024), idesc_b, acc);
                    }
                    tc::mma_commit(&bars.hp_free);
                    tc::mma_commit(&bars.ld_free[n % WG_STAGES]);
                }
                tc::mma_commit(&bars.flush_full);
            }
        } else if (warp >= 4) {
            const int i = (warp - 4) >> 2;
            const int wq = warp & 3;
            const int tg = (warp - 4 - 4 * i) * 32 + lane;
            const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
            uint32_t k = 0;                                           // tiles handled by this group
            for (int64_t n = i; n < my_tiles; n += 2, ++k) {
                const int64_t row0 = ((int64_t)slice + n * n_slices) * TM;
                const int64_t row = row0 + tg;
                // ---- S -> H = relu(S + b1) * keep   (bf16, registers)
                tc::mbar_wait(&bars.s_full[i], k & 1);
                tc::tc_fence_after();
                uint32_t hreg[64], preg[64];
                {
                    uint32_t v[2][32];
                    const uint32_t r_addr = tmem + lane_base + 128 * i;
                    tc::tmem_ld32(r_addr, v[0]);
#pragma unroll
                    for (int pc = 0; pc < 4; ++pc) {
                        tc::tmem_ld_wait();
                        if (pc < 3) tc::tmem_ld32(r_addr + 32 * (pc + 1), v[(pc + 1) & 1]);
                        uint32_t km[16];
                        if (p.thr)
                            epi::keep_masks16(rng_keep_word(p.keys2, (uint64_t)row * (uint64_t)(p.ff >> 5) + (uint64_t)(4 * c + pc), p.thr), km);
                        const uint4* bb = reinterpret_cast<const uint4*>(sB1h + 16 * pc);
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            const uint4 b4 = bb[q4];
                            const uint32_t bw[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int j = 4 * q4 + u;
                                uint32_t h2 = epi::relu_bias2(
                                    epi::cvt2(__uint_as_float(v[pc & 1][2 * j]), __uint_as_float(v[pc & 1][2 * j + 1])), bw[u]);
                                if (p.thr) h2 &= km[j];
                                hreg[pc * 16 + j] = h2;
                            }
                        }
                    }
                }
                tc::tc_fence_before();
                tc::mbar_arrive(&bars.a_done[i]);
                // ---- D -> dPre = D * [H > 0]
                tc::mbar_wait(&bars.d_full[i], k & 1);
                tc::tc_fence_after();
                {
                    uint32_t v[2][32];
                    const uint32_t r_addr = tmem + lane_base + 128 * i;
                    tc::tmem_ld32(r_addr, v[0]);
#pragma unroll
                    for (int pc = 0; pc < 4; ++pc) {
                        tc::tmem_ld_wait();
                        if (pc < 3) tc::tmem_ld32(r_addr + 32 * (pc + 1), v[(pc + 1) & 1]);
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            preg[pc * 16 + j] = epi::cvt2(__uint_as_float(v[pc & 1][2 * j]), __uint_as_float(v[pc & 1][2 * j + 1])) &
                                                epi::gt0_mask2(hreg[pc * 16 + j]);
                    }
                }
                // ---- H, dPre -> shared memory (MN-major A operands of the weight-gradient GEMMs)
                if (n > 0) tc::mbar_wait(&bars.hp_free, (uint32_t)(n - 1) & 1);
#pragma unroll
                for (int ch = 0; ch < 16; ++ch) {
                    const uint32_t off = (uint32_t)((ch >> 3) * 16384) + tc::sw128_chunk(tg, ch & 7);
                    *reinterpret_cast<uint4*>(sH + off) = make_uint4(hreg[4 * ch], hreg[4 * ch + 1], hreg[4 * ch + 2], hreg[4 * ch + 3]);
                    *reinterpret_cast<uint4*>(sP + off) = make_uint4(preg[4 * ch], preg[4 * ch + 1], preg[4 * ch + 2], preg[4 * ch + 3]);
                }
                tc::fence_proxy_async();
                tc::mbar_arrive(&bars.hp_full);
            }
            // ---- flush the chunk's weight gradients (group 0; thread <-> hidden unit)
            if (i == 0) {
                tc::mbar_wait(&bars.flush_full, 0);
                tc::tc_fence_after();
                const int h = c * CH + tg;
                uint32_t v[32];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    tc::tmem_ld32(tmem + lane_base + 256 + 32 * half, v);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (32 * half + j < p.d) atomicAdd(p.dW1 + (size_t)h * p.d + 32 * half + j, __uint_as_float(v[j]));
                    tc::tmem_ld32(tmem + lane_base + 320 + 32 * half, v);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (32 * half + j < p.d)
                            atomicAdd(p.dW2 + (size_t)(32 * half + j) * p.ff + h, __uint_as_float(v[j]) * p.hidden_scale);
                }
                uint32_t b16[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(b16[0]), "=r"(b16[1]), "=r"(b16[2]), "=r"(b16[3]), "=r"(b16[4]), "=r"(b16[5]), "=r"(b16[6]),
                               "=r"(b16[7]), "=r"(b16[8]), "=r"(b16[9]), "=r"(b16[10]), "=r"(b16[11]), "=r"(b16[12]), "=r"(b16[13]),
                               "=r"(b16[14]), "=r"(b16[15])
                             : "r"(tmem + lane_base + 384)
                             : "memory");
                tc::tmem_ld_wait();
                atomicAdd(p.db1 + h, __uint_as_float(b16[0]));
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc<512>(tmem);
}

}  // namespace

extern "C" int u2gnn_ffn_tc_bwd(const float* y1, const float* df, const float* dz, int64_t M, int d, int ff,
                                const void* packed, float hidden_scale, uint64_t seed, uint32_t stream_hidden, int thr,
                                float* dy1, float* dW1, float* db1, float* dW2, u2gnn_stream_t stream) {
    if (!y1 || !df || !dz || !packed || !dy1 || !dW1 || !db1 || !dW2 || M < 0 || thr < 0 || thr > 255) return U2GNN_EINVAL;
    if (d < 1 || d > DP || ff < CH || ff % CH || ff > 2048) return U2GNN_EUNSUPPORTED;
    if (d == DP && ((reinterpret_cast<uintptr_t>(y1) | reinterpret_cast<uintptr_t>(df) | reinterpret_cast<uintptr_t>(dz) |
                     reinterpret_cast<uintptr_t>(dy1)) % 16))
        return U2GNN_EALIGN;
    if (M == 0) return U2GNN_OK;
    const int NC = ff / CH;
    {
        DgradParams p;
        p.y1 = y1; p.df = df; p.dz = dz; p.dy1 = dy1; p.M = M; p.d = d; p.ff = ff;
        p.packed = static_cast<const uint8_t*>(packed);
        p.keys2 = rng_keys(seed, stream_hidden);
        p.thr = thr;
        const size_t smem = 1024 + 4 * 16384 + (size_t)DG_STAGES * DG_BLOCK + (size_t)(ff / 2) * sizeof(float);
        if (smem > 227 * 1024) return U2GNN_EUNSUPPORTED;
        cudaFuncSetAttribute(ffn_tc_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int64_t n_pairs = (M + 2 * TM - 1) / (2 * TM);
        const int grid = (int)(n_pairs < U2GNN_NUM_SMS ? n_pairs : U2GNN_NUM_SMS);
        ffn_tc_dgrad_kernel<<<grid, kThreads, smem, as_stream(stream)>>>(p);
    }
    {
        WgradParams p;
        p.y1 = y1; p.df = df; p.M = M; p.d = d; p.ff = ff;
        p.packed = static_cast<const uint8_t*>(packed);
        p.keys2 = rng_keys(seed, stream_hidden);
        p.thr = thr;
        p.hidden_scale = hidden_scale;
        p.dW1 = dW1; p.db1 = db1; p.dW2 = dW2;
        const size_t smem = 1024 + (size_t)WG_STAGES * 32768 + 3 * 32768 + 2048 + 512;
        cudaFuncSetAttribute(ffn_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int64_t n_tiles = (M + TM - 1) / TM;
        int n_slices = U2GNN_NUM_SMS / NC;
        if (n_slices < 1) n_slices = 1;
        if (n_slices > n_tiles) n_slices = (int)n_tiles;
        ffn_tc_wgrad_kernel<<<NC * n_slices, kThreads, smem, as_stream(stream)>>>(p);
    }
    U2GNN_CHECK_LAUNCH();
}
