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
constexpr int STAGES = 5;                        // weight ring: 16 KB stages, alternating W1c / W2c images
constexpr uint32_t STAGE_BYTES = CH * DP * 2;    // 16 KB  W1c: [128 x 64] K-major image, W2c: two [64 x 64] K-major images
constexpr uint32_t XS_ROW_BYTES = 272;           // fp32 staging row: 256 B + 16 B pad (conflict-free thread-per-row access)
constexpr uint32_t XS_TILE_BYTES = TM * XS_ROW_BYTES;
// packed weights: per chunk [W2c | W1c | W2Tc | W1Tc] (fwd: first two; dgrad: last three; wgrad: middle two), then b1, b2 (fp32)
constexpr uint32_t CHUNK_BYTES = 4 * 16384;
constexpr int kEpiWarps = 16;
// warp roles (704 threads): 0-15 chunk epilogue, 16-19 row I/O, 20 weight producer, 21 MMA issuer.  80 registers per
// thread for everyone: moving registers between warpgroups with setmaxnreg was tried and made the MMA and I/O warps spill.
constexpr int kIoWarp0 = 16, kProdWarp = 20, kMmaWarp = 21, kThreads = 704;
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
    uint32_t* mask;      // EMIT: [n_pairs * 2][ff / 128][4][128] mask words (see ffn_tc_wgrad.cu)
    uint32_t* trace;     // debug: per-warp clock stamps of CTA 0 (u2gnn_ffn_tc_set_trace); null in production
};
constexpr int TRACE_CAP = 1024;   // stamps per warp slot (0 = MMA warp, 1..16 = chunk-epilogue warps, 17..20 = I/O warps)

__device__ __forceinline__ const float* packed_b1(const uint8_t* packed, int ff) {
    return reinterpret_cast<const float*>(packed + (size_t)(ff / CH) * CHUNK_BYTES);
}

// ---------------------------------------------------------------------------------------------
// weight packing: fp32 master weights -> pre-swizzled bf16 shared-memory images
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ffn_pack_kernel(const float* __restrict__ W1, const float* __restrict__ b1,
                                                       const float* __restrict__ W2, const float* __restrict__ b2, int d,
                                                       int ff, float hidden_scale, uint8_t* __restrict__ packed) {
    const int c = blockIdx.x;  // chunk; blockIdx.y: one of gridDim.y slices of its hidden units (16 CTAs alone took 28 us)
    uint8_t* blk = packed + (size_t)c * CHUNK_BYTES;
    const int per = CH * DP / gridDim.y;
    for (int e = blockIdx.y * per + threadIdx.x; e < (blockIdx.y + 1) * per; e += blockDim.x) {
        const int r = e / DP, k = e % DP;           // r: hidden unit in chunk, k: feature
        const int h = c * CH + r;
        const float w1 = (k < d) ? W1[(size_t)h * d + k] : 0.0f;
        const float w2 = (k < d) ? W2[(size_t)k * ff + h] * hidden_scale : 0.0f;
        // W1c  : B of GEMM1  [N = hidden r][K = feature k]
        *reinterpret_cast<__nv_bfloat16*>(blk + 16384 + tc::sw128_offset(r, k)) = __float2bfloat16(w1);
        // W2c  : B of GEMM2  [N = feature k][K = hidden r]  (two K atoms of 64)
        // (times 0.5, exact: the forward's chunk epilogue leaves 2 h in tensor memory - epi::keep_factors16)
        *reinterpret_cast<__nv_bfloat16*>(blk + (r >> 6) * 8192 + tc::sw128_offset(k, r & 63)) = __float2bfloat16(0.5f * w2);
        // W2Tc : B of dH = dF W2c   [N = hidden r][K = feature k]
        *reinterpret_cast<__nv_bfloat16*>(blk + 32768 + tc::sw128_offset(r, k)) = __float2bfloat16(w2);
        // W1Tc : B of dy1 += dPre W1c  [N = feature k][K = hidden r]
        *reinterpret_cast<__nv_bfloat16*>(blk + 49152 + (r >> 6) * 8192 + tc::sw128_offset(k, r & 63)) = __float2bfloat16(w1);
    }
    float* bias = reinterpret_cast<float*>(packed + (size_t)(ff / CH) * CHUNK_BYTES);
    if (c == 0 && blockIdx.y == 0) {
        for (int e = threadIdx.x; e < ff; e += blockDim.x) bias[e] = b1[e];
        for (int e = threadIdx.x; e < DP; e += blockDim.x) bias[ff + e] = (e < d) ? b2[e] : 0.0f;
    }
}

// ---------------------------------------------------------------------------------------------
// forward kernel
// ---------------------------------------------------------------------------------------------
struct __align__(8) FwdBars {
    uint64_t w_full[STAGES], w_empty[STAGES];
    uint64_t xs_full[2][2];                 // fp32 staging rows of tile i landed (buffer b)
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
// 8 k-steps of Y_i (+)= H_i W2c^T.  A = packed H: the epilogue warp that owns S columns [32j, 32j+32) writes its 32
// hidden units as 16 packed columns at S column 32j, so k-step k reads columns 32*(k/2) + 8*(k%2).
// B = two [64 x 64] K-major atoms of W2c.
__device__ __forceinline__ void issue_gemm2(uint32_t tmem_y, uint32_t tmem_h, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    if (tc::elect_one()) {
        tc::mma_ts(tmem_y, tmem_h, b_desc, idesc, acc);
        tc::mma_ts_acc(tmem_y, tmem_h + 8, b_desc + 2, idesc);
        tc::mma_ts_acc(tmem_y, tmem_h + 32, b_desc + 4, idesc);
        tc::mma_ts_acc(tmem_y, tmem_h + 40, b_desc + 6, idesc);
        tc::mma_ts_acc(tmem_y, tmem_h + 64, b_desc + 512, idesc);           // second K atom: +8192 B
        tc::mma_ts_acc(tmem_y, tmem_h + 72, b_desc + 514, idesc);
        tc::mma_ts_acc(tmem_y, tmem_h + 96, b_desc + 516, idesc);
        tc::mma_ts_acc(tmem_y, tmem_h + 104, b_desc + 518, idesc);
    }
    __syncwarp();
}
__device__ __forceinline__ void commit_to(uint64_t* bar) {
    if (tc::elect_one()) tc::mma_commit(bar);
    __syncwarp();
}

// Warp roles (704 threads):
//   warps  0-15  chunk epilogue: ALL sixteen warps convert the S of ONE tile (32 columns x 32 lanes each) while the
//                tensor pipe runs the other tile's GEMMs, then swap: per-tile latency is half of an 8-warp split
//   warps 16-19  row I/O, off the critical path: per-row 256-byte bulk copies into a padded fp32 staging buffer
//                (prefetched one pair ahead), fp32 -> bf16 conversion into tensor memory, and the output epilogue
//                (Y + bias, dropout, residual from the staging row, LayerNorm) whose z / xnext rows leave through
//                per-row bulk stores - no uncoalesced global access anywhere
//   warp 20      weight producer (16 KB bulk copies, W1c(0) W2c(0) W1c(1) W2c(1) ...), warp 21 MMA issuer + TMEM owner
template <bool TRACE, bool EMIT>      // EMIT: write the 1-bit ReLU-and-keep mask of every hidden activation for the backward kernels
__global__ void __launch_bounds__(kThreads, 1) ffn_tc_fwd_kernel(const FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint32_t tr_n = 0;
    const long long tr_t0 = TRACE ? clock64() : 0;
    auto stamp = [&](int slot) {
        if (TRACE && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && tr_n < (uint32_t)TRACE_CAP)
            p.trace[slot * TRACE_CAP + tr_n++] = (uint32_t)(clock64() - tr_t0);
    };
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // offset on the shared-window address: keeps LDS / STS (a uintptr_t round trip makes every access generic)
    uint8_t* sW = smem;                                    // STAGES x 16 KB
    uint8_t* sX = sW + STAGES * STAGE_BYTES;               // [2 buffers][2 tiles][128 rows x 272 B] fp32 staging
    uint32_t* sB1h = reinterpret_cast<uint32_t*>(sX + 4 * XS_TILE_BYTES);    // -b1 as packed bf16 pairs (ff/2 words)
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
            tc::mbar_init(&bars.xs_full[0][i], 4);  // one arrival (+ its rows' bytes) per I/O warp
            tc::mbar_init(&bars.xs_full[1][i], 4);
            tc::mbar_init(&bars.x_full[i], 4);
            tc::mbar_init(&bars.x_free[i], 1);
            tc::mbar_init(&bars.s_full[i], 1);
            tc::mbar_init(&bars.h_full[i], kEpiWarps);
            tc::mbar_init(&bars.y_full[i], 1);
            tc::mbar_init(&bars.y_free[i], 4);
        }
        tc::fence_barrier_init();
    }
    if (warp == kMmaWarp) tc::tmem_alloc<512>(&tmem_slot);
    {   // biases / LayerNorm affine into shared memory
        const float* b1g = packed_b1(p.packed, p.ff);
        for (int e = threadIdx.x; e < p.ff / 2; e += kThreads) sB1h[e] = epi::cvt2(-b1g[2 * e], -b1g[2 * e + 1]);     // NEGATED (chunk epilogue: tn = -(c + b1))
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

    if (warp >= kProdWarp) {
      if (warp == kProdWarp) {
        // ================= weight producer =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
                for (int c2 = 0; c2 < 2 * NC; ++c2, ++it) {
                    const uint32_t s = it % STAGES, n = it / STAGES;
                    if (n > 0) tc::mbar_wait(&bars.w_empty[s], (n - 1) & 1);
                    tc::mbar_arrive_expect_tx(&bars.w_full[s], STAGE_BYTES);
                    // even: W1c(c) (second image of the chunk block), odd: W2c(c) (first image)
                    tc::bulk_g2s(sW + s * STAGE_BYTES, p.packed + (size_t)(c2 >> 1) * CHUNK_BYTES + ((c2 & 1) ? 0 : 16384),
                                 STAGE_BYTES, &bars.w_full[s]);
                }
            }
        }
      } else if (warp == kMmaWarp) {
        // ================= MMA issuer: the whole warp runs this code uniformly =================
        const uint32_t idesc1 = tc::make_idesc(TM, CH, 0, 0);
        const uint32_t idesc2 = tc::make_idesc(TM, DP, 0, 0);
        const uint64_t w_desc0 = tc::make_desc_sw128(tc::smem_u32(sW), 16, 1024);     // stage 0
        uint32_t it = 0, q = 0;
        uint32_t hcount[2] = {0, 0};
        for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++q) {
            // prologue: GEMM1 of chunk 0 for both tiles
            tc::mbar_wait(&bars.w_full[it % STAGES], (it / STAGES) & 1);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                tc::mbar_wait(&bars.x_full[i], q & 1);
                tc::tc_fence_after();
                issue_gemm1(tmem + COL_S + 128 * i, tmem + COL_X + 32 * i, w_desc0 + (uint64_t)((it % STAGES) * (STAGE_BYTES >> 4)), idesc1);
                commit_to(&bars.s_full[i]);
                if (NC == 1) commit_to(&bars.x_free[i]);
            }
            commit_to(&bars.w_empty[it % STAGES]);
            ++it;
            for (int c = 0; c < NC; ++c) {
                const uint32_t s2 = it % STAGES;                    // W2c(c)
                const uint32_t s1 = (it + 1) % STAGES;              // W1c(c + 1)
                tc::mbar_wait(&bars.w_full[s2], (it / STAGES) & 1);
                if (c + 1 < NC) tc::mbar_wait(&bars.w_full[s1], ((it + 1) / STAGES) & 1);
                const uint64_t w2_desc = w_desc0 + (uint64_t)(s2 * (STAGE_BYTES >> 4));
                const uint64_t w1_next = w_desc0 + (uint64_t)(s1 * (STAGE_BYTES >> 4));
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    stamp(0);
                    tc::mbar_wait(&bars.h_full[i], hcount[i] & 1);   // H_i(c) in TMEM (over S_i)
                    stamp(0);
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
                        if (c + 2 == NC) commit_to(&bars.x_free[i]);     // last GEMM1 of this pair reads X_i
                    }
                }
                commit_to(&bars.w_empty[s2]);
                ++it;
                if (c + 1 < NC) {
                    commit_to(&bars.w_empty[s1]);
                    ++it;
                }
            }
        }
      }
    } else if (warp < kEpiWarps) {
        const int ew = warp;
        // ================= chunk epilogue: 4 column quarters x 4 TMEM lane quarters, both tiles alternately =========
        const int cq = ew >> 2;                         // 32-column quarter of the 128-wide chunk
        const int wq = warp & 3;                        // TMEM lane quarter
        const int tr = wq * 32 + lane;                  // row in tile
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
        const uint32_t s_addr0 = tmem + lane_base + COL_S + 32 * cq;
        const uint32_t bar_s = tc::smem_u32(&bars.s_full[0]), bar_h = tc::smem_u32(&bars.h_full[0]);
        const uint32_t b1_addr = tc::smem_u32(sB1h) + 64u * cq;          // 16 packed-bias words per (chunk, quarter)
        const int thr = p.thr, low = p.low;
        const RngKeys keys2 = p.keys2;
        const uint32_t g_per_row = (uint32_t)(p.ff >> 5);
        const bool leader = (lane == 0);
        uint32_t scount = 0;
        auto keep_word = [&](int64_t pair, int c, int i) -> uint32_t {
            const uint64_t row = (uint64_t)(pair * (2 * TM) + (int64_t)i * TM + tr);
            return rng_keep_word_lo(keys2, row * g_per_row + (uint64_t)(4 * c + cq), thr, low);
        };
        for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
            for (int c = 0; c < NC; ++c, ++scount) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    // (measured, tools/trace_ffn.py: evaluating the keep word one step AHEAD under the previous step's tensor-memory
                    // load moves its ~25 instructions from here into the load phase and leaves the chunk-pair period where it was,
                    // 3 197 vs 3 211 cycles - the sixteen warps are bound by instruction issue, not by this serial chain)
                    uint32_t k0 = 0xFFFFFFFFu;
                    if (thr) k0 = keep_word(pair, c, i);
                    stamp(ew + 1);
                    tc::mbar_wait_addr(bar_s + 8u * i, scount & 1);
                    stamp(ew + 1);
                    tc::tc_fence_after();
                    const uint32_t s_addr = s_addr0 + 128 * i;
                    uint32_t v[32];
                    tc::tmem_ld32(s_addr, v);
                    // independent work under the load latency: the 16 keep-factor pairs of this step and its packed bias
                    uint32_t kp[16];
                    epi::keep_factors16(k0, kp);       // 2.0 kept / 0.0 dropped (no dropout: k0 = all ones)
                    uint32_t bw[16];
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4)
                        tc::lds128(b1_addr + (uint32_t)c * 256u + 16u * q4, bw[4 * q4], bw[4 * q4 + 1], bw[4 * q4 + 2], bw[4 * q4 + 3]);
                    tc::tmem_ld_wait();
                    stamp(ew + 1);
                    uint32_t nz = 0;                   // flag word (epi::flag_pos): hidden e of the group is live and kept
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        // tn = -(bf16(S) + b1);  g = 2 tn keep (sign set <=> live and kept; +0 when dropped or t = 0);  h' = relu(-g) = 2 relu(t) keep
                        const uint32_t g0 = epi::fma2(epi::fma2(epi::cvt2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), 0xBF80BF80u, bw[j]), kp[j], 0u);
                        const uint32_t g1 = epi::fma2(epi::fma2(epi::cvt2(__uint_as_float(v[2 * j + 2]), __uint_as_float(v[2 * j + 3])), 0xBF80BF80u, bw[j + 1]), kp[j + 1], 0u);
                        v[j] = epi::fma_relu2(g0, 0xBF80BF80u, 0u);          // in place: entries 2j .. 2j+3 are already consumed; W2c carries the 0.5
                        v[j + 1] = epi::fma_relu2(g1, 0xBF80BF80u, 0u);
                        if (EMIT) nz |= epi::flag_gather(g0, g1) & (0x01010101u << (j >> 1));
                    }
                    stamp(ew + 1);
                    tc::tmem_st16(s_addr, v);          // packed H over the first 16 of this thread's own 32 S columns
                    if (EMIT) p.mask[(((size_t)(pair * 2 + i) * NC + c) * 4 + cq) * TM + tr] = nz;      // 128 bytes per warp, coalesced
                    tc::tmem_st_wait();
                    tc::tc_fence_before();
                    __syncwarp();
                    if (leader) tc::mbar_arrive_addr(bar_h + 8u * i);
                    stamp(ew + 1);
                }
            }
        }
    } else if (warp >= kIoWarp0 && warp < kIoWarp0 + 4) {
        // ================= row I/O =================
        const int wq = warp & 3;
        const int tr = wq * 32 + lane;
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
        const bool fast = (p.d == DP);
        auto stage_row = [&](int b, int i) -> uint8_t* { return sX + (size_t)((b * 2 + i) * TM + tr) * XS_ROW_BYTES; };
        // rows of tile i of the CTA's pair number n -> staging buffer n & 1
        auto issue_load = [&](int64_t pair, uint32_t n, int i) {
            const int64_t row0 = pair * (2 * TM) + (int64_t)i * TM + wq * 32;
            const int64_t row = row0 + lane;
            uint8_t* dst = stage_row(n & 1, i);
            uint64_t* bar = &bars.xs_full[n & 1][i];
            tc::fence_proxy_async();                   // this thread's earlier generic accesses to the row
            if (fast) {
                int64_t nv = p.M - row0;
                nv = nv < 0 ? 0 : (nv > 32 ? 32 : nv);
                if (lane == 0) {
                    if (nv > 0) tc::mbar_arrive_expect_tx(bar, (uint32_t)nv * 256u);
                    else tc::mbar_arrive(bar);
                }
                __syncwarp();
                if (row < p.M) {
                    tc::bulk_g2s(dst, p.y1 + row * DP, 256, bar);
                } else {
#pragma unroll
                    for (int u = 0; u < 16; ++u) reinterpret_cast<float4*>(dst)[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
                float* drow = reinterpret_cast<float*>(dst);
                for (int j = 0; j < DP; ++j) drow[j] = (row < p.M && j < p.d) ? __ldg(p.y1 + row * p.d + j) : 0.0f;
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(bar);
            }
        };
        auto convert = [&](uint32_t n, int i) {
            tc::mbar_wait(&bars.xs_full[n & 1][i], (n >> 1) & 1);
            const float4* src = reinterpret_cast<const float4*>(stage_row(n & 1, i));
            uint32_t xp[32];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const float4 v = src[u];
                xp[2 * u] = epi::cvt2(v.x, v.y);
                xp[2 * u + 1] = epi::cvt2(v.z, v.w);
            }
            tc::tmem_st32(tmem + lane_base + COL_X + 32 * i, xp);
            tc::tmem_st_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bars.x_full[i]);
        };
        // prologue: two pairs of staging loads in flight, first pair converted
        {
            const int64_t p0 = blockIdx.x, p1 = p0 + gridDim.x;
            if (p0 < n_pairs) {
                issue_load(p0, 0, 0);
                issue_load(p0, 0, 1);
                if (p1 < n_pairs) {
                    issue_load(p1, 1, 0);
                    issue_load(p1, 1, 1);
                }
                convert(0, 0);
                convert(0, 1);
            }
        }
        uint32_t q = 0;
        for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++q) {
            const int64_t next = pair + gridDim.x, next2 = next + gridDim.x;
            // ---- (a) next pair's X into tensor memory as soon as this pair's last GEMM1 has read the current one
            if (next < n_pairs) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    tc::mbar_wait(&bars.x_free[i], q & 1);
                    tc::tc_fence_after();
                    convert(q + 1, i);
                }
            }
            // ---- (b) Y -> z (in place over the residual row), LayerNorm statistics
            float mean_[2], rstd_[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int64_t row = pair * (2 * TM) + (int64_t)i * TM + tr;
                float* srow = reinterpret_cast<float*>(stage_row(q & 1, i));
                stamp(17 + wq);
                tc::mbar_wait(&bars.y_full[i], q & 1);
                stamp(17 + wq);
                tc::tc_fence_after();
                uint32_t y0[32], y1r[32];
                tc::tmem_ld32(tmem + lane_base + COL_Y + 64 * i, y0);
                tc::tmem_ld32(tmem + lane_base + COL_Y + 64 * i + 32, y1r);
                tc::tmem_ld_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars.y_free[i]);
                float zv[DP];
                if (fast) {
                    uint32_t k0 = 0xFFFFFFFFu, k1 = 0xFFFFFFFFu;
                    if (p.thr) {
                        k0 = rng_keep_word_lo(p.keys3, (uint64_t)row * 2ull, p.thr, p.low);
                        k1 = rng_keep_word_lo(p.keys3, (uint64_t)row * 2ull + 1ull, p.thr, p.low);
                    }
                    const float sc = p.thr ? p.scale3 : 1.0f;
                    const float4* rr = reinterpret_cast<const float4*>(srow);
#pragma unroll
                    for (int j = 0; j < DP; j += 4) {
                        const float4 r4 = rr[j >> 2];
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
                        zv[j] = srow[j] + f * mult;
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
                mean_[i] = mean;
                rstd_[i] = rstd;
                if (p.stats && row < p.M) {
                    p.stats[2 * row] = mean;
                    p.stats[2 * row + 1] = rstd;
                }
#pragma unroll
                for (int j = 0; j < DP; j += 4) reinterpret_cast<float4*>(srow)[j >> 2] = make_float4(zv[j], zv[j + 1], zv[j + 2], zv[j + 3]);
                if (fast) {
                    tc::fence_proxy_async();
                    if (row < p.M) tc::bulk_s2g(p.z + row * DP, srow, 256);
                    tc::bulk_commit();
                } else if (row < p.M) {
#pragma unroll 1
                    for (int j = 0; j < p.d; ++j) p.z[row * p.d + j] = zv[j];
                }
            }
            // ---- (c) xnext = LayerNorm(z), written over the z row once its bulk store has read it
            if (p.xnext) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int64_t row = pair * (2 * TM) + (int64_t)i * TM + tr;
                    float* srow = reinterpret_cast<float*>(stage_row(q & 1, i));
                    const float mean = mean_[i], rstd = rstd_[i];
                    if (fast) {
                        tc::bulk_wait_read<1>();
                        float4* r4 = reinterpret_cast<float4*>(srow);
#pragma unroll
                        for (int j = 0; j < DP; j += 4) {
                            const float4 zq = r4[j >> 2];
                            r4[j >> 2] = make_float4((zq.x - mean) * rstd * sG[j] + sG[DP + j],
                                                     (zq.y - mean) * rstd * sG[j + 1] + sG[DP + j + 1],
                                                     (zq.z - mean) * rstd * sG[j + 2] + sG[DP + j + 2],
                                                     (zq.w - mean) * rstd * sG[j + 3] + sG[DP + j + 3]);
                        }
                        tc::fence_proxy_async();
                        if (row < p.M) tc::bulk_s2g(p.xnext + row * DP, srow, 256);
                        tc::bulk_commit();
                    } else if (row < p.M) {
#pragma unroll 1
                        for (int j = 0; j < p.d; ++j) p.xnext[row * p.d + j] = (srow[j] - mean) * rstd * sG[j] + sG[DP + j];
                    }
                }
            }
            // ---- (d) the buffer is free once the stores have read it: prefetch the pair after next
            if (fast) tc::bulk_wait_read<0>();
            if (next2 < n_pairs) {
                issue_load(next2, q + 2, 0);
                issue_load(next2, q + 2, 1);
            }
        }
        if (fast) tc::bulk_wait_all<0>();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tc::tmem_dealloc<512>(tmem);
}

size_t packed_bytes(int ff) { return (size_t)(ff / CH) * CHUNK_BYTES + (size_t)(ff + DP) * sizeof(float); }

}  // namespace

#ifdef U2GNN_PROBE_BUILD
// probe library only (libu2gnn_b200_probe.so, include/u2gnn_b200_probe.h): device buffer of clock stamps written by CTA 0 of the
// next FFN launches (slot 0 = MMA warp, 1..16 = epilogue warps; the backward kernels use slots 0.. and 32..); nullptr = off.
// The product library has no such state: its kernels are the TRACE = false instantiations.
uint32_t* g_ffn_trace = nullptr;
extern "C" int u2gnn_ffn_tc_set_trace(void* buf) {
    g_ffn_trace = static_cast<uint32_t*>(buf);
    return U2GNN_OK;
}
#endif

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
    ffn_pack_kernel<<<dim3(ff / CH, 8), 256, 0, as_stream(stream)>>>(W1, b1, W2, b2, d, ff, hidden_scale, static_cast<uint8_t*>(packed));
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_ffn_tc_fwd(const float* y1, int64_t M, int d, int ff, const void* packed, uint64_t seed,
                                uint32_t stream_hidden, uint32_t stream_out, int thr, const float* gamma,
                                const float* beta, float* z, float* stats, float* xnext, void* mask_out, u2gnn_stream_t stream) {
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
    p.mask = static_cast<uint32_t*>(mask_out);
    const size_t smem = 1024 + (size_t)STAGES * STAGE_BYTES + 4 * (size_t)XS_TILE_BYTES + (size_t)(ff / 2 + 3 * DP) * sizeof(float);
    if (smem > 227 * 1024) return U2GNN_EUNSUPPORTED;
    p.trace = nullptr;
    const int64_t n_pairs = (M + 2 * TM - 1) / (2 * TM);
    const int grid = (int)(n_pairs < U2GNN_NUM_SMS ? n_pairs : U2GNN_NUM_SMS);
    auto launch = [&](auto kern, int threads) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, threads, smem, as_stream(stream)>>>(p);
    };
#ifdef U2GNN_PROBE_BUILD
    p.trace = g_ffn_trace;
    if (p.trace) {
        if (p.mask) launch(ffn_tc_fwd_kernel<true, true>, kThreads);
        else launch(ffn_tc_fwd_kernel<true, false>, kThreads);
        U2GNN_CHECK_LAUNCH();
    }
#endif
    if (p.mask) launch(ffn_tc_fwd_kernel<false, true>, kThreads);
    else launch(ffn_tc_fwd_kernel<false, false>, kThreads);
    U2GNN_CHECK_LAUNCH();
}
