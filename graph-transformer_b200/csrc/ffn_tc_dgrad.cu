// Fused bf16 FFN block, input gradient (dgrad):  dy1 = dz + dPre W1,  dPre = (dF W2) * relu'(y1 W1^T + b1) * keep.
// Same skeleton as the forward kernel (ffn_tc.cu): persistent CTAs, pairs of 128-row tiles, 4 control warps + 16
// epilogue warps, weights streamed through a bulk-copy ring, every MMA A operand in tensor memory and issued from
// warp-uniform code.  Per tile i and 128-wide ff chunk c:
//     R_i = X_i  W1c^T         (TS, N 128)   epilogue A: packed mask  m = [bf16(S)+b1 > 0] & keep   (registers)
//     R_i = dF_i W2Tc^T        (TS, N 128)   epilogue B: dPre = bf16(D) & m  -> written over R_i (packed)
//     dY_i += dPre_i W1Tc^T    (TS, N 64)
// TMEM columns: dY0 [0,64) dY1 [64,128) R0 [128,256) R1 [256,384) X0/X1 [384,448) dF0/dF1 [448,512).
// Measured dead ends (tools/trace_ffn_bwd.py, tools/bench_ffn_bwd.py): two one-tile CTAs per SM (same time), register
// prefetch of the next pair's tiles (slower: spills), staggering the CTAs' start by 3-10 us to spread the row-I/O bursts
// (same time), sixteen warps x 32 columns on both tiles software-pipelined A(0) A(1) B(0) B(1) (11.01 ms against 10.81 ms for
// dgrad + wgrad).  What is left between pairs (19.5 K of 83 K cycles) is load latency + the dY drain.
#include "common.cuh"
#include "rng.cuh"
#include "tc_common.cuh"
#include "ffn_epi.cuh"

namespace {

constexpr int DP = 64, CH = 128, TM = 128;
constexpr uint32_t CHUNK_BYTES = 4 * 16384;   // [W2c | W1c | W2Tc | W1Tc]
constexpr uint32_t BLOCK = 3 * 16384;         // [W1c | W2Tc | W1Tc]
// NT = row tiles per CTA.  NT = 2: one CTA per SM walks pairs of tiles (640 threads, 512 TMEM columns, 4 weight stages).
// NT = 1: TWO independent CTAs per SM, one tile each (320 threads, 256 TMEM columns, 2 weight stages), so that the row
// I/O phase of one CTA (35 % of the NT = 2 kernel's time, tools/trace_ffn_bwd.py) overlaps the chunk loop of the other.
// Measured: the SAME 11.46 ms per 4 M rows as NT = 2 (11.49 ms) - the chunk loop is bound by the S -> mask -> D -> dPre
// -> dY chain latency of each tile, not by a shared resource, so only MORE tiles in flight would help and tensor memory
// holds two.  Only NT = 2 is instantiated (half the L2 weight traffic; NT = 1 no longer fits two CTAs once the staging exists).
template <int NT> struct Cfg {
    static constexpr int kCtrl = (NT == 2) ? 4 : 2;                 // warp 0 weight producer, warp 1 MMA issuer + TMEM owner
    static constexpr int kThreads = 32 * (kCtrl + 8 * NT);
    static constexpr int STAGES = (NT == 2) ? 3 : 2;                // 48 KB weight stages
    static constexpr uint32_t COL_Y = 0, COL_R = 64 * NT, COL_X = 192 * NT, COL_F = 224 * NT;
    static constexpr int kTmemCols = 256 * NT;
};

struct Params {
    const float* y1;
    const float* df;
    const float* dz;
    float* dy1;
    int64_t M;
    int d, ff;
    const uint8_t* packed;
    RngKeys keys2;
    int thr, low;
    uint8_t* xb;      // [n_tiles][128 x 64] bf16 swizzled images of y1 / df, consumed by the wgrad kernel (may be null)
    uint8_t* fb;
    uint32_t* trace;   // debug clock stamps of CTA 0 (u2gnn_ffn_tc_set_trace), slots 32..
    // LayerNorm2 backward fused into the dF loader (LNF kernels, d == 64): dF = dropout3(dz), dz = LN backward of dy2 at
    // the saved pre-norm z2 / stats; dz is parked in dy1 (p.dz == p.dy1) until the drain adds dPre W1 to it
    const float* dy2;
    const float* z2;
    const float* st2;
    const float* gamma2;
    RngKeys keys3;
    float scale3;
    float* dgamma2;
    float* dbeta2;
    float* db2;        // linear2 bias gradient = colsum(dF)
};

__device__ __forceinline__ float group16_sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
constexpr int TRACE_CAP = 1024;

struct __align__(8) Bars {
    uint64_t w_full[4], w_empty[4];
    uint64_t x_full[2], x_free[2], s_full[2], a_done[2], d_full[2], p_full[2], y_full[2], y_free[2];
};

__device__ __forceinline__ void issue_n128(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc) {
    if (tc::elect_one()) {
        tc::mma_ts(tmem_d, tmem_a, b_desc, idesc, 0);
        tc::mma_ts_acc(tmem_d, tmem_a + 8, b_desc + 2, idesc);
        tc::mma_ts_acc(tmem_d, tmem_a + 16, b_desc + 4, idesc);
        tc::mma_ts_acc(tmem_d, tmem_a + 24, b_desc + 6, idesc);
    }
    __syncwarp();
}
__device__ __forceinline__ void issue_n64(uint32_t tmem_y, uint32_t tmem_p, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    if (tc::elect_one()) {
        tc::mma_ts(tmem_y, tmem_p, b_desc, idesc, acc);
        tc::mma_ts_acc(tmem_y, tmem_p + 8, b_desc + 2, idesc);
        tc::mma_ts_acc(tmem_y, tmem_p + 16, b_desc + 4, idesc);
        tc::mma_ts_acc(tmem_y, tmem_p + 24, b_desc + 6, idesc);
        tc::mma_ts_acc(tmem_y, tmem_p + 64, b_desc + 512, idesc);
        tc::mma_ts_acc(tmem_y, tmem_p + 72, b_desc + 514, idesc);
        tc::mma_ts_acc(tmem_y, tmem_p + 80, b_desc + 516, idesc);
        tc::mma_ts_acc(tmem_y, tmem_p + 88, b_desc + 518, idesc);
    }
    __syncwarp();
}
__device__ __forceinline__ void commit_to(uint64_t* bar) {
    if (tc::elect_one()) tc::mma_commit(bar);
    __syncwarp();
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <bool TRACE, int NT, bool LNF = false>
__global__ void __launch_bounds__(Cfg<NT>::kThreads, 3 - NT) ffn_tc_dgrad_kernel(const Params p) {
    using C_ = Cfg<NT>;
    constexpr int kThreads = C_::kThreads, STAGES = C_::STAGES, kCtrl = C_::kCtrl;
    constexpr uint32_t COL_Y = C_::COL_Y, COL_R = C_::COL_R, COL_X = C_::COL_X, COL_F = C_::COL_F;
    extern __shared__ uint8_t smem_raw[];
    uint32_t tr_n = 0;
    const long long tr_t0 = TRACE ? clock64() : 0;
    auto stamp = [&](int slot) {
        if (TRACE && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && tr_n < (uint32_t)TRACE_CAP)
            p.trace[(32 + slot) * TRACE_CAP + tr_n++] = (uint32_t)(clock64() - tr_t0);
    };
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // offset on the shared-window address: keeps LDS / STS (a uintptr_t round trip makes every access generic)
    uint8_t* sW = smem;                                                        // STAGES x 48 KB
    uint32_t* sB1h = reinterpret_cast<uint32_t*>(sW + STAGES * BLOCK);         // b1 as packed bf16 pairs
    uint8_t* sIO = reinterpret_cast<uint8_t*>(sB1h + p.ff / 2);                // 16 KB per (tile, warpgroup): bf16 tile image, later dY staging
    __shared__ Bars bars;
    __shared__ uint32_t tmem_slot;
    __shared__ float s_acc[LNF ? 192 : 1];                 // dgamma2 | dbeta2 | db2 partial sums of this CTA
    if (LNF) for (int e = threadIdx.x; e < 192; e += Cfg<NT>::kThreads) s_acc[e] = 0.0f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NC = p.ff / CH;
    const int64_t n_pairs = (p.M + NT * TM - 1) / (NT * TM);      // groups of NT tiles ("pairs" for NT = 2)

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(&bars.w_full[s], 1);
            tc::mbar_init(&bars.w_empty[s], 1);
        }
        for (int i = 0; i < NT; ++i) {
            tc::mbar_init(&bars.x_full[i], 8);      // X (warpgroup 0) and dF (warpgroup 1): one arrival per warp
            tc::mbar_init(&bars.x_free[i], 1);
            tc::mbar_init(&bars.s_full[i], 1);
            tc::mbar_init(&bars.a_done[i], 8);
            tc::mbar_init(&bars.d_full[i], 1);
            tc::mbar_init(&bars.p_full[i], 8);
            tc::mbar_init(&bars.y_full[i], 1);
            tc::mbar_init(&bars.y_free[i], 8);      // both warpgroups of the tile drain half of dY each
        }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc<C_::kTmemCols>(&tmem_slot);
    {
        const float* b1g = reinterpret_cast<const float*>(p.packed + (size_t)NC * CHUNK_BYTES);
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
                    const uint32_t s = it % STAGES, n = it / STAGES;
                    if (n > 0) tc::mbar_wait(&bars.w_empty[s], (n - 1) & 1);
                    tc::mbar_arrive_expect_tx(&bars.w_full[s], BLOCK);
                    tc::bulk_g2s(sW + s * BLOCK, p.packed + (size_t)c * CHUNK_BYTES + 16384, BLOCK, &bars.w_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (warp-uniform) =================
        const uint32_t idesc_n128 = tc::make_idesc(TM, CH, 0, 0);
        const uint32_t idesc_n64 = tc::make_idesc(TM, DP, 0, 0);
        const uint64_t w_desc0 = tc::make_desc_sw128(tc::smem_u32(sW), 16, 1024);     // stage 0: [W1c | W2Tc | W1Tc]
        uint32_t it = 0, q = 0;
        uint32_t acount[2] = {0, 0}, pcount[2] = {0, 0};
        for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++q) {
            tc::mbar_wait(&bars.w_full[it % STAGES], (it / STAGES) & 1);
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                tc::mbar_wait(&bars.x_full[i], q & 1);
                tc::tc_fence_after();
                issue_n128(tmem + COL_R + 128 * i, tmem + COL_X + 32 * i, w_desc0 + (uint64_t)((it % STAGES) * (BLOCK >> 4)), idesc_n128);
                commit_to(&bars.s_full[i]);
            }
            for (int c = 0; c < NC; ++c, ++it) {
                const uint32_t s = it % STAGES;
                const uint64_t wd = w_desc0 + (uint64_t)(s * (BLOCK >> 4));
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                    stamp(0);
                    tc::mbar_wait(&bars.a_done[i], acount[i] & 1);     // epilogue has turned S_i into mask registers
                    stamp(0);
                    ++acount[i];
                    tc::tc_fence_after();
                    issue_n128(tmem + COL_R + 128 * i, tmem + COL_F + 32 * i, wd + 1024, idesc_n128);      // D_i = dF_i W2Tc^T
                    commit_to(&bars.d_full[i]);
                }
                if (c + 1 < NC) tc::mbar_wait(&bars.w_full[(it + 1) % STAGES], ((it + 1) / STAGES) & 1);
                const uint64_t w1_next = w_desc0 + (uint64_t)(((it + 1) % STAGES) * (BLOCK >> 4));
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                    stamp(0);
                    tc::mbar_wait(&bars.p_full[i], pcount[i] & 1);     // dPre_i in TMEM (over R_i)
                    stamp(0);
                    ++pcount[i];
                    if (c == 0 && q > 0) tc::mbar_wait(&bars.y_free[i], (q - 1) & 1);
                    tc::tc_fence_after();
                    issue_n64(tmem + COL_Y + 64 * i, tmem + COL_R + 128 * i, wd + 2048, idesc_n64, c > 0);  // dY_i += dPre_i W1Tc^T
                    if (c == NC - 1) {
                        commit_to(&bars.y_full[i]);
                        commit_to(&bars.x_free[i]);                    // X_i and dF_i no longer needed
                    } else {
                        issue_n128(tmem + COL_R + 128 * i, tmem + COL_X + 32 * i, w1_next, idesc_n128);     // in order after the GEMM above
                        commit_to(&bars.s_full[i]);
                    }
                }
                commit_to(&bars.w_empty[s]);
            }
        }
    } else if (warp >= kCtrl) {
        const int i = (warp - kCtrl) >> 3;
        const int wg = ((warp - kCtrl) >> 2) & 1;
        const int wq = warp & 3;
        const int tr = wq * 32 + lane;
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
        uint32_t q = 0, scount = 0, dcount = 0;
        for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++q) {
            const int64_t row = pair * (NT * TM) + (int64_t)i * TM + tr;
            // ---- operands into tensor memory: warpgroup 0 stages X, warpgroup 1 stages dF.  The 128 threads of a
            // warpgroup read the [128 x 64] fp32 tile with COALESCED 128-bit loads, round it into the swizzled bf16 tile
            // image in shared memory (bulk-stored from there for the wgrad kernel), and each thread then moves its own row
            // of the image into tensor memory.  (Thread-per-row global loads touched 32 lines per instruction and made the
            // row I/O 35 % of this kernel.)
            const float* src = (wg == 0) ? p.y1 : p.df;
            uint8_t* img_s = sIO + (size_t)(i * 2 + wg) * 16384;
            const int bar_id = 1 + i * 2 + wg;
            stamp(warp - kCtrl + 1);
            if (q > 0) tc::mbar_wait(&bars.x_free[i], (q - 1) & 1);
            stamp(warp - kCtrl + 1);
            {
                const int64_t row0 = pair * (NT * TM) + (int64_t)i * TM;
                if (tr == 0) tc::bulk_wait_read<0>();     // an image store of the previous pair may still read the region
                if (LNF && wg == 1) {
                    // LayerNorm2 backward in the loader (same arithmetic, same order as ln_bwd_vec_kernel<16>, layernorm.cu):
                    // 16 lanes x float4 = one row; dz -> dy1 (fp32, re-read by the drain), dF = dropout3(dz) -> bf16 image.
                    // Batches of RB rows per thread keep the loads of a batch (g, x, stats) in flight together.
                    const int l = tr & 15;
                    const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma2) + l);
                    float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = ag, as = ag;
#pragma unroll
                    constexpr int RB = 4;                       // rows per thread per batch (8: 600 bytes of spills at 102 registers)
#pragma unroll
                    for (int hb = 0; hb < 16 / RB; ++hb) {
                        float4 gv[RB], xv[RB];
                        float2 sv[RB];
#pragma unroll
                        for (int u = 0; u < RB; ++u) {
                            const int e = (hb * RB + u) * 128 + tr;
                            const int64_t rg = row0 + (e >> 4);
                            gv[u] = xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                            sv[u] = make_float2(0.f, 0.f);
                            if (rg < p.M) {
                                gv[u] = __ldg(reinterpret_cast<const float4*>(p.dy2 + rg * DP) + l);
                                xv[u] = __ldg(reinterpret_cast<const float4*>(p.z2 + rg * DP) + l);
                                sv[u] = __ldg(reinterpret_cast<const float2*>(p.st2) + rg);
                            }
                        }
                        if (hb == 0) named_bar_sync(bar_id, 128);      // everybody is done reading the staging of the previous pair
#pragma unroll
                        for (int u = 0; u < RB; ++u) {
                            const int e = (hb * RB + u) * 128 + tr;
                            const int64_t rg = row0 + (e >> 4);
                            const float4 g = gv[u], x = xv[u];
                            const float mean = sv[u].x, rstd = sv[u].y;
                            const float4 xh = make_float4((x.x - mean) * rstd, (x.y - mean) * rstd, (x.z - mean) * rstd, (x.w - mean) * rstd);
                            const float4 dh = make_float4(g.x * g4.x, g.y * g4.y, g.z * g4.z, g.w * g4.w);
                            ag.x = fmaf(g.x, xh.x, ag.x); ag.y = fmaf(g.y, xh.y, ag.y); ag.z = fmaf(g.z, xh.z, ag.z); ag.w = fmaf(g.w, xh.w, ag.w);
                            ab.x += g.x; ab.y += g.y; ab.z += g.z; ab.w += g.w;
                            const float m1 = group16_sum((dh.x + dh.y) + (dh.z + dh.w)) * (1.0f / 64.0f);
                            const float m2 = group16_sum((dh.x * xh.x + dh.y * xh.y) + (dh.z * xh.z + dh.w * xh.w)) * (1.0f / 64.0f);
                            float4 o = make_float4(rstd * (dh.x - m1 - xh.x * m2), rstd * (dh.y - m1 - xh.y * m2),
                                                   rstd * (dh.z - m1 - xh.z * m2), rstd * (dh.w - m1 - xh.w * m2));
                            if (rg < p.M) reinterpret_cast<float4*>(p.dy1 + rg * DP)[l] = o;
                            if (p.thr) {
                                const uint64_t el = (uint64_t)(rg * DP + 4 * l);
                                const uint32_t kw = rng_keep_word_lo(p.keys3, el >> 5, p.thr, p.low) >> (el & 31);
                                o.x = (kw & 1u) ? o.x * p.scale3 : 0.0f;
                                o.y = (kw & 2u) ? o.y * p.scale3 : 0.0f;
                                o.z = (kw & 4u) ? o.z * p.scale3 : 0.0f;
                                o.w = (kw & 8u) ? o.w * p.scale3 : 0.0f;
                            }
                            as.x += o.x; as.y += o.y; as.z += o.z; as.w += o.w;
                            uint2 w;
                            w.x = epi::cvt2(o.x, o.y);
                            w.y = epi::cvt2(o.z, o.w);
                            *reinterpret_cast<uint2*>(img_s + tc::sw128_offset(e >> 4, l * 4)) = w;
                        }
                    }
                    // the two rows of the warp, then one shared-memory atomic per column per warp; flushed once per CTA
                    ag.x += __shfl_xor_sync(0xffffffffu, ag.x, 16); ag.y += __shfl_xor_sync(0xffffffffu, ag.y, 16);
                    ag.z += __shfl_xor_sync(0xffffffffu, ag.z, 16); ag.w += __shfl_xor_sync(0xffffffffu, ag.w, 16);
                    ab.x += __shfl_xor_sync(0xffffffffu, ab.x, 16); ab.y += __shfl_xor_sync(0xffffffffu, ab.y, 16);
                    ab.z += __shfl_xor_sync(0xffffffffu, ab.z, 16); ab.w += __shfl_xor_sync(0xffffffffu, ab.w, 16);
                    as.x += __shfl_xor_sync(0xffffffffu, as.x, 16); as.y += __shfl_xor_sync(0xffffffffu, as.y, 16);
                    as.z += __shfl_xor_sync(0xffffffffu, as.z, 16); as.w += __shfl_xor_sync(0xffffffffu, as.w, 16);
                    if (lane < 16) {
                        atomicAdd(&s_acc[4 * l], ag.x); atomicAdd(&s_acc[4 * l + 1], ag.y); atomicAdd(&s_acc[4 * l + 2], ag.z); atomicAdd(&s_acc[4 * l + 3], ag.w);
                        atomicAdd(&s_acc[64 + 4 * l], ab.x); atomicAdd(&s_acc[64 + 4 * l + 1], ab.y); atomicAdd(&s_acc[64 + 4 * l + 2], ab.z); atomicAdd(&s_acc[64 + 4 * l + 3], ab.w);
                        atomicAdd(&s_acc[128 + 4 * l], as.x); atomicAdd(&s_acc[128 + 4 * l + 1], as.y); atomicAdd(&s_acc[128 + 4 * l + 2], as.z); atomicAdd(&s_acc[128 + 4 * l + 3], as.w);
                    }
                } else if (p.d == DP) {
                    // (prefetching these loads into registers before the previous pair's output phase was measured: slower,
                    // the 64 extra live registers spill)
                    float4 v[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const int e = u * 128 + tr;
                        const int64_t rg = row0 + (e >> 4);
                        v[u] = (rg < p.M) ? __ldg(reinterpret_cast<const float4*>(src + rg * DP) + (e & 15)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    named_bar_sync(bar_id, 128);          // everybody is done reading the staging of the previous pair
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const int e = u * 128 + tr;
                        uint2 w;
                        w.x = epi::cvt2(v[u].x, v[u].y);
                        w.y = epi::cvt2(v[u].z, v[u].w);
                        *reinterpret_cast<uint2*>(img_s + tc::sw128_offset(e >> 4, (e & 15) * 4)) = w;
                    }
                } else {
                    named_bar_sync(bar_id, 128);
                    for (int e = tr; e < TM * DP; e += 128) {
                        const int r = e >> 6, k = e & 63;
                        const int64_t rg = row0 + r;
                        const float x = (rg < p.M && k < p.d) ? src[rg * p.d + k] : 0.0f;
                        *reinterpret_cast<__nv_bfloat16*>(img_s + tc::sw128_offset(r, k)) = __float2bfloat16(x);
                    }
                }
                tc::fence_proxy_async();                  // generic writes -> visible to the bulk store below
                named_bar_sync(bar_id, 128);
                uint8_t* img = (wg == 0) ? p.xb : p.fb;
                if (img && tr == 0) {
                    tc::bulk_s2g(img + (size_t)(pair * NT + i) * 16384, img_s, 16384);
                    tc::bulk_commit();
                }
                uint32_t xp[32];
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    const uint4 w = *reinterpret_cast<const uint4*>(img_s + tc::sw128_chunk(tr, ch));
                    xp[4 * ch] = w.x; xp[4 * ch + 1] = w.y; xp[4 * ch + 2] = w.z; xp[4 * ch + 3] = w.w;
                }
                tc::tmem_st32(tmem + lane_base + (wg == 0 ? COL_X : COL_F) + 32 * i, xp);
                tc::tmem_st_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars.x_full[i]);
                stamp(warp - kCtrl + 1);
            }
            const uint32_t r_addr = tmem + lane_base + COL_R + 128 * i + 64 * wg;
            for (int c = 0; c < NC; ++c) {
                // ---- epilogue A: S -> packed mask
                stamp(warp - kCtrl + 1);
                tc::mbar_wait(&bars.s_full[i], scount & 1);
                stamp(warp - kCtrl + 1);
                ++scount;
                tc::tc_fence_after();
                uint32_t msk[32];
#pragma unroll
                for (int pc = 0; pc < 2; ++pc) {
                    uint32_t v[32];
                    tc::tmem_ld32(r_addr + 32 * pc, v);
                    uint32_t km[16];
                    if (p.thr)
                        epi::keep_masks16(rng_keep_word_lo(p.keys2, (uint64_t)row * (uint64_t)(p.ff >> 5) + (uint64_t)(4 * c + 2 * wg + pc),
                                                           p.thr, p.low), km);
                    const uint4* bb = reinterpret_cast<const uint4*>(sB1h + ((c * CH + 64 * wg + 32 * pc) >> 1));
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const uint4 b4 = bb[q4];
                        const uint32_t bw[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int j = 4 * q4 + u;
                            uint32_t m = epi::gt0_mask2(epi::relu_bias2(epi::cvt2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), bw[u]));
                            if (p.thr) m &= km[j];
                            msk[pc * 16 + j] = m;
                        }
                    }
                }
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars.a_done[i]);
                stamp(warp - kCtrl + 1);
                // ---- epilogue B: D -> dPre (packed, over this thread's own R columns)
                tc::mbar_wait(&bars.d_full[i], dcount & 1);
                stamp(warp - kCtrl + 1);
                ++dcount;
                tc::tc_fence_after();
                uint32_t hp[32];
#pragma unroll
                for (int pc = 0; pc < 2; ++pc) {
                    uint32_t v[32];
                    tc::tmem_ld32(r_addr + 32 * pc, v);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        hp[pc * 16 + j] = epi::cvt2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])) & msk[pc * 16 + j];
                }
                tc::tmem_st32(r_addr, hp);
                tc::tmem_st_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars.p_full[i]);
                stamp(warp - kCtrl + 1);
            }
            stamp(warp - kCtrl + 1);
            // ---- dY + dz -> dy1: each warpgroup drains 32 of the 64 columns; registers -> staging (its own 16 KB region,
            // [128 rows x 128 B], 16-byte chunks XOR-swizzled with the row) -> coalesced dz loads and dy1 stores
            tc::mbar_wait(&bars.y_full[i], q & 1);
            tc::tc_fence_after();
            uint32_t yv[32];
            tc::tmem_ld32(tmem + lane_base + COL_Y + 64 * i + 32 * wg, yv);
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bars.y_free[i]);
            if (p.d == DP) {
                if (tr == 0) tc::bulk_wait_read<0>();     // the image store of this pair has left the region
                named_bar_sync(bar_id, 128);
#pragma unroll
                for (int ch = 0; ch < 8; ++ch)
                    *reinterpret_cast<uint4*>(img_s + tr * 128 + ((ch ^ (tr & 7)) << 4)) = make_uint4(yv[4 * ch], yv[4 * ch + 1], yv[4 * ch + 2], yv[4 * ch + 3]);
                named_bar_sync(bar_id, 128);
                const int64_t row0 = pair * (NT * TM) + (int64_t)i * TM;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = u * 128 + tr;
                    const int r = e >> 3, c4 = e & 7;
                    const int64_t rg = row0 + r;
                    if (rg < p.M) {
                        const float4 o = *reinterpret_cast<const float4*>(img_s + r * 128 + ((c4 ^ (r & 7)) << 4));
                        // plain load: with LNF the row was written by this CTA's loader (ordered by the mbarrier chain x_full -> y_full)
                        const float4 z4 = LNF ? *(reinterpret_cast<const float4*>(p.dz + rg * DP + 32 * wg) + c4)
                                              : __ldg(reinterpret_cast<const float4*>(p.dz + rg * DP + 32 * wg) + c4);
                        reinterpret_cast<float4*>(p.dy1 + rg * DP + 32 * wg)[c4] = make_float4(o.x + z4.x, o.y + z4.y, o.z + z4.z, o.w + z4.w);
                    }
                }
            } else if (row < p.M) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int col = 32 * wg + j;
                    if (col < p.d) p.dy1[row * p.d + col] = p.dz[row * p.d + col] + __uint_as_float(yv[j]);
                }
            }
            if (pair + gridDim.x >= n_pairs && tr == 0) tc::bulk_wait_all<0>();   // last pair: image stores complete before exit
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<C_::kTmemCols>(tmem);
    if (LNF && threadIdx.x < 192) {
        float* dst = threadIdx.x < 64 ? p.dgamma2 : (threadIdx.x < 128 ? p.dbeta2 : p.db2);
        atomicAdd(dst + (threadIdx.x & 63), s_acc[threadIdx.x]);
    }
}

}  // namespace

// internal launch used by u2gnn_ffn_tc_bwd (ffn_tc_bwd.cu)
struct FfnLnBwd {       // LayerNorm2 backward inputs / outputs of the fused entry point (ffn_tc_bwd.cu)
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
                        const FfnLnBwd* ln) {
    Params p;
    p.y1 = y1; p.df = df; p.dz = dz; p.dy1 = dy1; p.M = M; p.d = d; p.ff = ff;
    p.dy2 = p.z2 = p.st2 = p.gamma2 = nullptr;
    p.dgamma2 = p.dbeta2 = p.db2 = nullptr;
    p.keys3 = rng_keys(seed, 0);
    p.scale3 = thr ? rng_keep_scale(thr) : 1.0f;
    if (ln) {
        if (d != DP) return U2GNN_EUNSUPPORTED;
        p.dy2 = ln->dy2; p.z2 = ln->z2; p.st2 = ln->st2; p.gamma2 = ln->gamma2;
        p.keys3 = rng_keys(seed, ln->stream_out);
        p.dgamma2 = ln->dgamma2; p.dbeta2 = ln->dbeta2; p.db2 = ln->db2;
        p.df = nullptr;
        p.dz = dy1;                                     // dz is parked in dy1 by the loader
    }
    p.packed = static_cast<const uint8_t*>(packed);
    p.keys2 = rng_keys(seed, stream_hidden);
    p.thr = thr;
    p.low = rng_thr_low(thr);
    p.xb = static_cast<uint8_t*>(xb);
    p.fb = static_cast<uint8_t*>(fb);
    extern uint32_t* g_ffn_trace;
    p.trace = g_ffn_trace;
    auto launch = [&](auto kern, int nt, int threads, int stages) -> int {
        const size_t smem = 1024 + (size_t)stages * BLOCK + (size_t)(ff / 2) * sizeof(uint32_t) + (size_t)nt * 2 * 16384;
        if (smem > 227 * 1024) return U2GNN_EUNSUPPORTED;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int64_t n_groups = (M + nt * TM - 1) / (nt * TM);
        const int64_t cap = (int64_t)U2GNN_NUM_SMS * (3 - nt);
        kern<<<(int)(n_groups < cap ? n_groups : cap), threads, smem, st>>>(p);
        return U2GNN_OK;
    };
    if (ln) return launch(ffn_tc_dgrad_kernel<false, 2, true>, 2, Cfg<2>::kThreads, Cfg<2>::STAGES);
    if (p.trace) return launch(ffn_tc_dgrad_kernel<true, 2>, 2, Cfg<2>::kThreads, Cfg<2>::STAGES);
    return launch(ffn_tc_dgrad_kernel<false, 2>, 2, Cfg<2>::kThreads, Cfg<2>::STAGES);
}

