// Fused bf16 FFN block, input gradient (dgrad):  dy1 = dz + dPre W1,  dPre = (dF W2) * relu'(y1 W1^T + b1) * keep.
// Runs AFTER the weight-gradient kernel (ffn_tc_wgrad.cu), which has already recomputed the hidden activation and left one
// bit per hidden unit and row (ReLU live AND dropout keep): this kernel never forms S = y1 W1^T again.  Persistent CTAs, pairs
// of 128-row tiles, per tile i and 128-wide ff chunk c:
//     R_i = dF_i W2Tc^T        (TS, N 128)   epilogue: dPre = bf16(D) & mask  -> written over R_i (packed)
//     dY_i += dPre_i W1Tc^T    (TS, N 64)
// i.e. 2 GEMM units per (tile, chunk) and a dependency chain D -> dPre -> dY (round 1: S -> mask -> D -> dPre -> dY, 3 units,
// 2 470 cycles per (pair, chunk) against 935 cycles of MMA issue).  The mask words of a (pair, chunk) travel with the chunk's
// weight images through the bulk-copy ring (32 KB [W2Tc | W1Tc] + 2 x 2 KB of mask words per stage) and are expanded to bf16
// pair masks BEFORE the wait for D, so only load -> convert -> and -> store sits between the two GEMMs.
// Row operands: the bf16 tile images of dF (written by the LayerNorm2 backward / the image pass) are bulk-copied one PAIR
// AHEAD into a double-buffered staging area by their own producer warp; each thread moves its row from there into tensor memory.
// TMEM columns: dY0 [0,64) dY1 [64,128) R0 [128,256) R1 [256,384) dF0 [384,416) dF1 [416,448).
#include "common.cuh"
#include "tc_common.cuh"
#include "ffn_epi.cuh"

namespace {

constexpr int DP = 64, CH = 128, TM = 128, NT = 2;
constexpr uint32_t CHUNK_BYTES = 4 * 16384;   // [W2c | W1c | W2Tc | W1Tc]
constexpr uint32_t W_BYTES = 2 * 16384;       // [W2Tc | W1Tc]
constexpr uint32_t MASK_TILE = 4 * TM * 4;    // mask words of one (tile, chunk): [4 groups of 32 hidden units][128 rows]
constexpr uint32_t STAGE = W_BYTES + NT * MASK_TILE;
constexpr int STAGES = 3;
constexpr int kCtrl = 4;                      // warp 0 weight + mask producer, warp 1 MMA issuer + TMEM owner, warp 2 row-image producer
constexpr int kThreads = 32 * (kCtrl + 8 * NT);
constexpr uint32_t COL_Y = 0, COL_R = 128, COL_F = 384;
constexpr int kTmemCols = 512;

struct Params {
    const float* dz;
    float* dy1;
    int64_t M;
    int d, ff;
    const uint8_t* packed;
    const uint8_t* fb;       // [n_pairs * 2][128 x 64] bf16 swizzled images of dF
    const uint8_t* mask;     // [n_pairs * 2][ff / 128][4][128] mask words from the weight-gradient kernel
    uint32_t* trace;         // debug clock stamps of CTA 0 (probe build only), slots 32..
};
constexpr int TRACE_CAP = 1024;

struct __align__(8) Bars {
    uint64_t w_full[STAGES], w_empty[STAGES];
    uint64_t img_full[2], img_free[2];
    uint64_t x_full[2], x_free[2], d_full[2], p_full[2], y_full[2], y_free[2];
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

template <bool TRACE>
__global__ void __launch_bounds__(kThreads, 1) ffn_tc_dgrad_kernel(const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint32_t tr_n = 0;
    const long long tr_t0 = TRACE ? clock64() : 0;
    auto stamp = [&](int slot) {
        if (TRACE && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && tr_n < (uint32_t)TRACE_CAP)
            p.trace[(32 + slot) * TRACE_CAP + tr_n++] = (uint32_t)(clock64() - tr_t0);
    };
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // offset on the shared-window address: keeps LDS / STS
    uint8_t* sW = smem;                                    // STAGES x (32 KB weights | 4 KB mask words)
    uint8_t* sImg = sW + STAGES * STAGE;                   // 2 buffers x 2 tiles x 16 KB dF images; the current pair's buffer is the dY staging of warpgroup 0 at the end of the pair
    uint8_t* sOut = sImg + 2 * NT * 16384;                 // 2 tiles x 16 KB: dY staging of warpgroup 1
    __shared__ Bars bars;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NC = p.ff / CH;
    const int64_t n_pairs = (p.M + NT * TM - 1) / (NT * TM);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(&bars.w_full[s], 1);
            tc::mbar_init(&bars.w_empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&bars.img_full[i], 1);
            tc::mbar_init(&bars.img_free[i], 16);
            tc::mbar_init(&bars.x_full[i], 8);      // both warpgroups of the tile store half of the dF row each
            tc::mbar_init(&bars.x_free[i], 1);
            tc::mbar_init(&bars.d_full[i], 1);
            tc::mbar_init(&bars.p_full[i], 8);
            tc::mbar_init(&bars.y_full[i], 1);
            tc::mbar_init(&bars.y_free[i], 8);      // both warpgroups of the tile drain half of dY each
        }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc<kTmemCols>(&tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        // ================= weight + mask producer =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
                for (int c = 0; c < NC; ++c, ++it) {
                    const uint32_t s = it % STAGES, n = it / STAGES;
                    if (n > 0) tc::mbar_wait(&bars.w_empty[s], (n - 1) & 1);
                    uint8_t* dst = sW + s * STAGE;
                    tc::mbar_arrive_expect_tx(&bars.w_full[s], STAGE);
                    tc::bulk_g2s(dst, p.packed + (size_t)c * CHUNK_BYTES + 32768, W_BYTES, &bars.w_full[s]);
#pragma unroll
                    for (int i = 0; i < NT; ++i)
                        tc::bulk_g2s(dst + W_BYTES + i * MASK_TILE, p.mask + ((size_t)(pair * NT + i) * NC + c) * MASK_TILE, MASK_TILE, &bars.w_full[s]);
                }
            }
        }
    } else if (warp == 2) {
        // ================= row-image producer: the dF images of the next pair land while this pair is in its chunk loop =================
        if (lane == 0) {
            uint32_t q = 0;
            for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++q) {
                const uint32_t b = q & 1;
                if (q >= 2) tc::mbar_wait(&bars.img_free[b], ((q >> 1) - 1) & 1);
                tc::mbar_arrive_expect_tx(&bars.img_full[b], NT * 16384);
                tc::bulk_g2s(sImg + b * (NT * 16384), p.fb + (size_t)pair * (NT * 16384), NT * 16384, &bars.img_full[b]);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (warp-uniform) =================
        const uint32_t idesc_n128 = tc::make_idesc(TM, CH, 0, 0);
        const uint32_t idesc_n64 = tc::make_idesc(TM, DP, 0, 0);
        const uint64_t w_desc0 = tc::make_desc_sw128(tc::smem_u32(sW), 16, 1024);     // stage 0: [W2Tc | W1Tc]
        uint32_t it = 0, q = 0;
        uint32_t pcount[2] = {0, 0};
        for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++q) {
            tc::mbar_wait(&bars.w_full[it % STAGES], (it / STAGES) & 1);
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                tc::mbar_wait(&bars.x_full[i], q & 1);
                tc::tc_fence_after();
                issue_n128(tmem + COL_R + 128 * i, tmem + COL_F + 32 * i, w_desc0 + (uint64_t)((it % STAGES) * (STAGE >> 4)), idesc_n128);   // D_i(0)
                commit_to(&bars.d_full[i]);
            }
            for (int c = 0; c < NC; ++c, ++it) {
                const uint32_t s = it % STAGES;
                const uint64_t wd = w_desc0 + (uint64_t)(s * (STAGE >> 4));
                if (c + 1 < NC) tc::mbar_wait(&bars.w_full[(it + 1) % STAGES], ((it + 1) / STAGES) & 1);
                const uint64_t w2_next = w_desc0 + (uint64_t)(((it + 1) % STAGES) * (STAGE >> 4));
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                    stamp(0);
                    tc::mbar_wait(&bars.p_full[i], pcount[i] & 1);     // dPre_i in TMEM (over R_i)
                    stamp(0);
                    ++pcount[i];
                    if (c == 0 && q > 0) tc::mbar_wait(&bars.y_free[i], (q - 1) & 1);
                    tc::tc_fence_after();
                    issue_n64(tmem + COL_Y + 64 * i, tmem + COL_R + 128 * i, wd + 1024, idesc_n64, c > 0);  // dY_i += dPre_i W1Tc^T
                    if (c == NC - 1) {
                        commit_to(&bars.y_full[i]);
                        commit_to(&bars.x_free[i]);                    // dF_i no longer needed
                    } else {
                        issue_n128(tmem + COL_R + 128 * i, tmem + COL_F + 32 * i, w2_next, idesc_n128);     // D_i(c+1), in order after the GEMM above
                        commit_to(&bars.d_full[i]);
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
        const uint32_t r_addr = tmem + lane_base + COL_R + 128 * i + 64 * wg;
        const int bar_id = 1 + i * 2 + wg;
        uint32_t q = 0, it = 0, dcount = 0;
        for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++q) {
            const int64_t row = pair * (NT * TM) + (int64_t)i * TM + tr;
            const uint32_t b = q & 1;
            uint8_t* img_s = sImg + b * (NT * 16384) + i * 16384;
            // ---- dF row into tensor memory: each warpgroup moves half of the row (4 of its 8 16-byte pieces)
            stamp(warp - kCtrl + 1);
            if (q > 0) tc::mbar_wait(&bars.x_free[i], (q - 1) & 1);
            tc::mbar_wait(&bars.img_full[b], (q >> 1) & 1);
            stamp(warp - kCtrl + 1);
            {
                uint32_t xp[32];
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    const uint4 w = *reinterpret_cast<const uint4*>(img_s + tc::sw128_chunk(tr, 4 * wg + ch));
                    xp[4 * ch] = w.x; xp[4 * ch + 1] = w.y; xp[4 * ch + 2] = w.z; xp[4 * ch + 3] = w.w;
                }
                tc::tc_fence_after();
                tc::tmem_st16(tmem + lane_base + COL_F + 32 * i + 16 * wg, xp);
                tc::tmem_st_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars.x_full[i]);
            }
            stamp(warp - kCtrl + 1);
            for (int c = 0; c < NC; ++c, ++it) {
                // ---- mask words of this thread's row -> bf16 pair masks (before D arrives)
                const uint32_t s = it % STAGES;
                tc::mbar_wait(&bars.w_full[s], (it / STAGES) & 1);
                const uint32_t* mw = reinterpret_cast<const uint32_t*>(sW + s * STAGE + W_BYTES + i * MASK_TILE) + (2 * wg) * TM + tr;
                uint32_t km[32];
                {
                    uint32_t k0[16], k1[16];
                    epi::flag_masks16(mw[0], k0);
                    epi::flag_masks16(mw[TM], k1);
#pragma unroll
                    for (int j = 0; j < 16; ++j) { km[j] = k0[j]; km[16 + j] = k1[j]; }
                }
                // ---- D -> dPre (packed, over this thread's own R columns)
                stamp(warp - kCtrl + 1);
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
                        hp[pc * 16 + j] = epi::cvt2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])) & km[pc * 16 + j];
                }
                tc::tmem_st32(r_addr, hp);
                tc::tmem_st_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars.p_full[i]);
                stamp(warp - kCtrl + 1);
            }
            stamp(warp - kCtrl + 1);
            // ---- dY + dz -> dy1: each warpgroup drains 32 of the 64 columns; registers -> staging (16 KB, [128 rows x 128 B],
            // 16-byte chunks XOR-swizzled with the row) -> coalesced dz loads and dy1 stores
            tc::mbar_wait(&bars.y_full[i], q & 1);
            tc::tc_fence_after();
            uint32_t yv[32];
            tc::tmem_ld32(tmem + lane_base + COL_Y + 64 * i + 32 * wg, yv);
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bars.y_free[i]);
            if (p.d == DP) {
                uint8_t* st_s = (wg == 0) ? img_s : (sOut + i * 16384);     // the dF image of this pair is dead (in tensor memory since x_full)
#pragma unroll
                for (int ch = 0; ch < 8; ++ch)
                    *reinterpret_cast<uint4*>(st_s + tr * 128 + ((ch ^ (tr & 7)) << 4)) = make_uint4(yv[4 * ch], yv[4 * ch + 1], yv[4 * ch + 2], yv[4 * ch + 3]);
                named_bar_sync(bar_id, 128);
                const int64_t row0 = pair * (NT * TM) + (int64_t)i * TM;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = u * 128 + tr;
                    const int r = e >> 3, c4 = e & 7;
                    const int64_t rg = row0 + r;
                    if (rg < p.M) {
                        const float4 o = *reinterpret_cast<const float4*>(st_s + r * 128 + ((c4 ^ (r & 7)) << 4));
                        const float4 z4 = __ldg(reinterpret_cast<const float4*>(p.dz + rg * DP + 32 * wg) + c4);
                        reinterpret_cast<float4*>(p.dy1 + rg * DP + 32 * wg)[c4] = make_float4(o.x + z4.x, o.y + z4.y, o.z + z4.z, o.w + z4.w);
                    }
                }
                named_bar_sync(bar_id, 128);          // everybody has read the staging back before the next pair's drain (or image) overwrites it
            } else if (row < p.M) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int col = 32 * wg + j;
                    if (col < p.d) p.dy1[row * p.d + col] = p.dz[row * p.d + col] + __uint_as_float(yv[j]);
                }
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bars.img_free[b]);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<kTmemCols>(tmem);
}

}  // namespace

#ifdef U2GNN_PROBE_BUILD
extern uint32_t* g_ffn_trace;
#endif

// internal launch used by u2gnn_ffn_tc_bwd (ffn_tc_bwd.cu)
int ffn_tc_dgrad_launch(const float* dz, float* dy1, int64_t M, int d, int ff, const void* packed, const void* fb, const void* mask,
                        cudaStream_t st) {
    Params p;
    p.dz = dz; p.dy1 = dy1; p.M = M; p.d = d; p.ff = ff;
    p.packed = static_cast<const uint8_t*>(packed);
    p.fb = static_cast<const uint8_t*>(fb);
    p.mask = static_cast<const uint8_t*>(mask);
    p.trace = nullptr;
    auto kern = ffn_tc_dgrad_kernel<false>;
#ifdef U2GNN_PROBE_BUILD
    p.trace = g_ffn_trace;
    if (p.trace) kern = ffn_tc_dgrad_kernel<true>;
#endif
    const size_t smem = 1024 + (size_t)STAGES * STAGE + (size_t)2 * NT * 16384 + (size_t)NT * 16384;
    if (smem > 227 * 1024) return U2GNN_EUNSUPPORTED;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t n_pairs = (M + NT * TM - 1) / (NT * TM);
    kern<<<(int)(n_pairs < U2GNN_NUM_SMS ? n_pairs : U2GNN_NUM_SMS), kThreads, smem, st>>>(p);
    return U2GNN_OK;
}
