// Fused bf16 FFN block, weight gradients (wgrad):  dW1, db1, dW2  with the hidden recomputed on chip, and the 1-bit
// ReLU-and-keep mask of every hidden activation written out for the input-gradient kernel (ffn_tc_dgrad.cu).
// One 128-wide ff chunk per CTA (its W1c / W2Tc images stay resident in shared memory); the CTAs of a chunk split
// the 128-row tiles among themselves.  Everything is computed TRANSPOSED - tensor-memory lane = hidden unit, column =
// row of the tile - so that the hidden activation and its gradient are already the A operands (M = hidden, K = rows)
// of the two weight-gradient GEMMs and never pass through shared memory.  Per tile n (buffer i = n & 1):
//     R_i    = W1c  X^T          (SS, N 128)   epilogue A: H^T = relu(bf16(S^T)+b1) & keep -> packed bf16 over R_i's own columns
//     dW2c^T[hidden, d]      += H^T  dF        (TS, B = the dF tile read MN-major, N 64)
//     R_i    = W2Tc dF^T         (SS, N 128)   epilogue B: dPre^T = bf16(D^T) & [H^T > 0]        -> packed bf16 in P
//     [dW1c | db1c][hidden,] += dPre^T [X | 1] (TS, N 80: the second MN group of B is a tile of ones)
// accumulated in tensor memory over all tiles of the CTA and flushed once with atomics.
// The first version of this kernel (round 1) kept lane = row: H and dPre then had to be stored to shared memory as MN-major
// A operands (64 KB of stores and 100 KB of operand reads per tile, the kernel's floor at 3 300 cycles per tile).
// The dropout keep words are defined per (row, 32 hidden units); a warp owns 32 hidden units x 32 rows, so lane j evaluates
// the word of row j and a 5-step butterfly (transpose32) hands every lane the 32 row bits of its own hidden unit.  The
// mask words leave the same way: bit (hidden unit) of word (row), one coalesced 128-byte store per warp and tile.
// TMEM columns: R0 [0,128) R1 [128,256) dW2c^T [256,320) dW1c [320,384) db1c [384,400) P [400,464).
#include "common.cuh"
#include "rng.cuh"
#include "tc_common.cuh"
#include "ffn_epi.cuh"

namespace {

constexpr int DP = 64, CH = 128, TM = 128;
constexpr uint32_t CHUNK_BYTES = 4 * 16384;   // [W2c | W1c | W2Tc | W1Tc]
constexpr int kThreads = 640;
constexpr int WG_STAGES = 4;
constexpr uint32_t COL_R = 0, COL_DW2 = 256, COL_DW1 = 320, COL_P = 400;

struct Params {
    const uint8_t* xb;   // bf16 swizzled [128 x 64] tile images of y1 / dF
    const uint8_t* fb;
    int64_t M;
    int d, ff;
    const uint8_t* packed;
    RngKeys keys2;
    int thr, low;
    float hidden_scale;
    float* dW1;   // [ff, d]
    float* db1;   // [ff]
    float* dW2;   // [d, ff]
    uint32_t* mask;    // [n_tiles][ff / 128][4][128] words: word (tile, chunk, q, row) holds the live-and-kept bits of hidden units
                       // 128 chunk + 32 q + [0, 32) of that row in flag-word order (ffn_epi.cuh, epi::flag_pos).  FWD_MASK: read; else written (may be null)
    uint32_t* trace;   // debug clock stamps of CTA 0 (probe build only)
};
constexpr int TRACE_CAP = 1024;

struct __align__(8) Bars {
    uint64_t w_full, ld_full[WG_STAGES], ld_free[WG_STAGES], s_full[2], a_done[2], d_full[2], b_done[2], p_full, p_free, flush_full;
};

__device__ __forceinline__ void commit_to(uint64_t* bar) {
    if (tc::elect_one()) tc::mma_commit(bar);
    __syncwarp();
}
// R = W_image(K-major A, 4 k-steps) * tile^T (K-major B), N = 128
__device__ __forceinline__ void issue_n128(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    if (tc::elect_one()) {
        tc::mma_ss(tmem_d, a_desc, b_desc, idesc, 0);
        tc::mma_ss_acc(tmem_d, a_desc + 2, b_desc + 2, idesc);
        tc::mma_ss_acc(tmem_d, a_desc + 4, b_desc + 4, idesc);
        tc::mma_ss_acc(tmem_d, a_desc + 6, b_desc + 6, idesc);
    }
    __syncwarp();
}
// acc += A^T-in-TMEM (lanes = hidden, packed row pairs along the columns) * tile (MN-major B: 8 k-steps of 16 rows, 2048 B each).
// SPLIT: the packed columns sit at the start of every 32-column group (written in place over the accumulator they came from).
template <bool SPLIT>
__device__ __forceinline__ void issue_wgrad(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    if (tc::elect_one()) {
        tc::mma_ts(tmem_d, tmem_a, b_desc, idesc, acc);
#pragma unroll
        for (int ks = 1; ks < 8; ++ks)
            tc::mma_ts_acc(tmem_d, tmem_a + (SPLIT ? 32 * (ks >> 1) + 8 * (ks & 1) : 8 * ks), b_desc + 128 * ks, idesc);
    }
    __syncwarp();
}

// 32 x 32 bit-matrix transpose across the lanes of a warp: lane i ends up with bit b = bit i of lane b's word
__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const uint32_t mask = (s == 16) ? 0x0000FFFFu : (s == 8) ? 0x00FF00FFu : (s == 4) ? 0x0F0F0F0Fu : (s == 2) ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, s);
        const bool lo = (lane & s) == 0;
        const uint32_t t = lo ? (y << s) : (y >> s);
        const uint32_t m2 = lo ? mask : ~mask;
        x = (x & m2) | (t & ~m2);
    }
    return x;
}
// FWD_MASK: the forward kernel already left the mask words (ffn_tc.cu, EMIT); this kernel then neither evaluates the dropout
// stream nor extracts / writes the mask - it loads the 32 words of its (32 rows x 32 hidden units) block and transposes them.
template <bool TRACE, bool FWD_MASK>
__global__ void __launch_bounds__(kThreads, 1) ffn_tc_wgrad_kernel(const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint32_t tr_n = 0;
    const long long tr_t0 = TRACE ? clock64() : 0;
    auto stamp = [&](int slot) {
        if (TRACE && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && tr_n < (uint32_t)TRACE_CAP)
            p.trace[slot * TRACE_CAP + tr_n++] = (uint32_t)(clock64() - tr_t0);
    };
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // offset on the shared-window address: keeps LDS / STS (a uintptr_t round trip makes every access generic)
    uint8_t* sXF = smem;                                   // WG_STAGES x (X 16 KB | dF 16 KB)
    uint8_t* sOnes = smem + WG_STAGES * 32768;             // 16 KB tile of bf16 1.0 (after the ring: LBO to it is positive)
    uint8_t* sW = sOnes + 16384;                           // 32 KB: [W1c | W2Tc]
    __shared__ Bars bars;
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
            tc::mbar_init(&bars.ld_full[s], 1);            // producer arrive + 32 KB of bulk-copy bytes
            tc::mbar_init(&bars.ld_free[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&bars.s_full[i], 1);
            tc::mbar_init(&bars.a_done[i], 16);
            tc::mbar_init(&bars.d_full[i], 1);
            tc::mbar_init(&bars.b_done[i], 16);
        }
        tc::mbar_init(&bars.p_full, 16);
        tc::mbar_init(&bars.p_free, 1);
        tc::mbar_init(&bars.flush_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc<512>(&tmem_slot);
    for (int e = threadIdx.x; e < 16384 / 4; e += kThreads) reinterpret_cast<uint32_t*>(sOnes)[e] = 0x3F803F80u;
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
        } else if (warp == 2) {
            // ================= row-tile producer: two 16 KB bulk copies per tile ==========
            if (lane == 0) {
                for (int64_t n = 0; n < my_tiles; ++n) {
                    const uint32_t s = (uint32_t)(n % WG_STAGES), u = (uint32_t)(n / WG_STAGES);
                    if (u > 0) tc::mbar_wait(&bars.ld_free[s], (u - 1) & 1);
                    const int64_t tile = (int64_t)slice + n * n_slices;
                    tc::mbar_arrive_expect_tx(&bars.ld_full[s], 32768);
                    tc::bulk_g2s(sXF + s * 32768, p.xb + (size_t)tile * 16384, 16384, &bars.ld_full[s]);
                    tc::bulk_g2s(sXF + s * 32768 + 16384, p.fb + (size_t)tile * 16384, 16384, &bars.ld_full[s]);
                }
            }
        } else if (warp == 1) {
            // ================= MMA issuer (warp-uniform) =================
            const uint32_t idesc_n128 = tc::make_idesc(CH, TM, 0, 0);          // M = hidden, N = rows of the tile
            const uint32_t idesc_w2 = tc::make_idesc(CH, DP, 0, 1);            // A from TMEM, B MN-major (rows = K)
            const uint32_t idesc_w1 = tc::make_idesc(CH, DP + 16, 0, 1);       // N = 80: [X | ones]
            const uint64_t xf0 = tc::make_desc_sw128(tc::smem_u32(sXF), 16, 1024);         // K-major view of stage 0 X tile
            const uint64_t w1d = tc::make_desc_sw128(tc::smem_u32(sW), 16, 1024);
            const uint64_t w2td = w1d + 1024;
            // Issue order (the tensor pipe executes in order):
            //   S(0) | W2(0) D(0) S(1) | W2(1) D(1) S(2) W1(0) | W2(2) D(2) S(3) W1(1) | ...
            // W2(n) reads H^T(n) out of R_i before D(n) overwrites R_i; S(n+1) goes to the other buffer as soon as D(n-1) has
            // been read out of it; W1(n-1) needs dPre^T(n-1) from the epilogue and comes last so that it never delays S / D.
            tc::mbar_wait(&bars.w_full, 0);
            tc::mbar_wait(&bars.ld_full[0], 0);
            tc::tc_fence_after();
            issue_n128(tmem + COL_R, w1d, xf0, idesc_n128);
            commit_to(&bars.s_full[0]);
            auto issue_w1 = [&](int64_t n) {
                const uint32_t s = (uint32_t)(n % WG_STAGES);
                stamp(0);
                tc::mbar_wait(&bars.p_full, (uint32_t)n & 1);
                stamp(0);
                tc::tc_fence_after();
                // MN-major B view of [X | ones]: two 64-wide groups, the second one LBO bytes further (the ones tile)
                const uint32_t x_addr = tc::smem_u32(sXF) + s * 32768;
                const uint64_t xd_mn = tc::make_desc_sw128(x_addr, tc::smem_u32(sOnes) - x_addr, 1024);
                issue_wgrad<false>(tmem + COL_DW1, tmem + COL_P, xd_mn, idesc_w1, n > 0);
                commit_to(&bars.p_free);
                commit_to(&bars.ld_free[s]);
                stamp(0);
            };
            for (int64_t n = 0; n < my_tiles; ++n) {
                const uint32_t i = (uint32_t)(n & 1), ph = (uint32_t)(n >> 1) & 1;
                const uint32_t s = (uint32_t)(n % WG_STAGES);
                const uint32_t x_addr = tc::smem_u32(sXF) + s * 32768;
                const uint64_t xd = xf0 + (uint64_t)(s * 2048);            // X tile of this stage (K-major view)
                stamp(0);
                tc::mbar_wait(&bars.a_done[i], ph);
                stamp(0);
                tc::tc_fence_after();
                issue_wgrad<true>(tmem + COL_DW2, tmem + COL_R + 128 * i, tc::make_desc_sw128(x_addr + 16384, 16384, 1024), idesc_w2, n > 0);   // W2(n)
                issue_n128(tmem + COL_R + 128 * i, w2td, xd + 1024, idesc_n128);            // D(n) = W2Tc dF^T
                commit_to(&bars.d_full[i]);
                if (n + 1 < my_tiles) {
                    const uint32_t s2 = (uint32_t)((n + 1) % WG_STAGES);
                    tc::mbar_wait(&bars.ld_full[s2], (uint32_t)((n + 1) / WG_STAGES) & 1);
                    if (n >= 1) tc::mbar_wait(&bars.b_done[i ^ 1], (uint32_t)((n - 1) >> 1) & 1);   // D(n-1) has left R_{i^1}
                    tc::tc_fence_after();
                    issue_n128(tmem + COL_R + 128 * (i ^ 1), w1d, xf0 + (uint64_t)(s2 * 2048), idesc_n128);   // S(n+1)
                    commit_to(&bars.s_full[i ^ 1]);
                }
                if (n >= 1) issue_w1(n - 1);
            }
            issue_w1(my_tiles - 1);
            commit_to(&bars.flush_full);
        } else if (warp >= 4) {
            // ================= epilogue: 16 warps = 4 TMEM lane quarters (32 hidden units) x 4 column groups (32 rows), EVERY tile ======
            // Software-pipelined per warp:  A(n) | B(n-1) | A(n+1) | B(n) | ...  so that D(n) and S(n+1) are computed by the
            // tensor pipe while the warps finish tile n-1.
            const int ew = warp - 4;
            const int cg = ew >> 2;                         // rows [32 cg, 32 cg + 32) of the tile = accumulator columns
            const int wq = warp & 3;                        // hidden units [32 wq, 32 wq + 32) of the chunk = TMEM lanes
            const int hl = wq * 32 + lane;
            const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
            const uint32_t r_addr0 = tmem + lane_base + COL_R + 32 * cg;
            const uint32_t p_addr = tmem + lane_base + COL_P + 16 * cg;
            const int thr = p.thr, low = p.low;
            const RngKeys keys2 = p.keys2;
            const uint32_t g_per_row = (uint32_t)(p.ff >> 5);
            uint32_t bb;                                    // b1 of this thread's hidden unit, both halves
            {
                const float b = reinterpret_cast<const float*>(p.packed + (size_t)NC * CHUNK_BYTES)[c * CH + hl];
                bb = epi::cvt2(b, b);
            }
            uint32_t nz_cur = 0, nz_prev = 0;               // split-pair words: bit j = row 2j of the group is live, bit 16 + j = row 2j + 1
            uint32_t mw_next = 0;                           // FWD_MASK: mask word of this lane's row in the NEXT tile
            if (FWD_MASK) mw_next = __ldg(p.mask + (((size_t)slice * NC + c) * 4 + wq) * TM + 32 * cg + lane);
            auto phase_a = [&](int64_t n) {                // S^T(n) -> H^T = relu(bf16(S^T) + b1) & keep, packed over R_i
                const uint32_t i = (uint32_t)(n & 1);
                const int64_t tile = (int64_t)slice + n * n_slices;
                const int64_t row0 = tile * TM + 32 * cg;
                uint32_t kw = 0xFFFFFFFFu;                  // bit r = this hidden unit is kept (FWD_MASK: kept AND live) in row row0 + r
                if (FWD_MASK) {
                    // lane j holds the word of row row0 + j (loaded one tile AHEAD: a load issued here would expose its full
                    // latency); after the transpose lane k holds the row bits of the hidden unit at flag-word position k, so
                    // hidden unit `lane` fetches them from lane flag_pos(lane)
                    const uint32_t w = mw_next;
                    if (n + 1 < my_tiles)
                        mw_next = __ldg(p.mask + (((size_t)(tile + n_slices) * NC + c) * 4 + wq) * TM + 32 * cg + lane);
                    kw = __shfl_sync(0xffffffffu, transpose32(w, lane), epi::flag_pos(lane));
                } else if (thr) {
                    kw = transpose32(rng_keep_word_lo(keys2, (uint64_t)(row0 + lane) * g_per_row + (uint64_t)(4 * c + wq), thr, low), lane);
                }
                stamp(warp - 3);
                tc::mbar_wait(&bars.s_full[i], (uint32_t)(n >> 1) & 1);
                stamp(warp - 3);
                tc::tc_fence_after();
                uint32_t v[32];
                tc::tmem_ld32(r_addr0 + 128 * i, v);
                // phase A sits on the kernel's critical chain (S(n+1) -> A -> W2 / D(n+1)) and was ALU-pipe bound (cvt, and, prmt
                // per pair + the word transpose): the mask is applied as a bf16 MULTIPLICATION like in the forward kernel
                // (epi::keep_factors16: one IMAD + one LOP3 per pair; tools/probe_epi_ops.cu) - H^T leaves as 2 h, exact, and the
                // flush of dW2 carries the 0.5
                uint32_t kp[16];
                epi::keep_factors16(kw, kp);               // 2.0 kept (and live) / 0.0; no mask: kw = all ones
                tc::tmem_ld_wait();
                uint32_t nz = 0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t t2 = epi::fma2(epi::cvt2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), 0x3F803F80u, bb);
                    const uint32_t h2 = epi::fma_relu2(t2, kp[j], 0u);        // 2 relu(t) keep
                    v[j] = h2;
                    if (!FWD_MASK) nz |= epi::gt0_mask2(h2) & (0x00010001u << j);
                }
                tc::tmem_st16(r_addr0 + 128 * i, v);       // packed H^T over the first 16 of this warp's own 32 columns
                tc::tmem_st_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars.a_done[i]);
                nz_cur = FWD_MASK ? kw : nz;                // FWD_MASK: natural row order (bit r = row r); else split-pair over the rows
                if (!FWD_MASK && p.mask) {
                    // hidden units are moved to their flag-word lane first, so that after the transpose lane k holds the word of
                    // row split_elem(k) (nz is split-pair over the rows) with the hidden bits in flag-word order (epi::flag_pos) - the
                    // format the forward kernel writes
                    const uint32_t w = transpose32(__shfl_sync(0xffffffffu, nz, epi::flag_elem(lane)), lane);
                    p.mask[(((size_t)tile * NC + c) * 4 + wq) * TM + 32 * cg + epi::split_elem(lane)] = w;
                }
                stamp(warp - 3);
            };
            auto phase_b = [&](int64_t n) {                // D^T(n) -> dPre^T = bf16(D^T) & [H^T > 0], packed into P
                const uint32_t i = (uint32_t)(n & 1);
                stamp(warp - 3);
                tc::mbar_wait(&bars.d_full[i], (uint32_t)(n >> 1) & 1);
                stamp(warp - 3);
                tc::tc_fence_after();
                uint32_t v[32];
                tc::tmem_ld32(r_addr0 + 128 * i, v);
                uint32_t pm[16];
                if (FWD_MASK) epi::keep_masks16(nz_prev, pm);
                else epi::split_masks16(nz_prev, pm);
                tc::tmem_ld_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars.b_done[i]);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = epi::cvt2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])) & pm[j];
                stamp(warp - 3);
                if (n > 0) tc::mbar_wait(&bars.p_free, (uint32_t)(n - 1) & 1);
                stamp(warp - 3);
                tc::tc_fence_after();
                tc::tmem_st16(p_addr, v);
                tc::tmem_st_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars.p_full);
                stamp(warp - 3);
            };
            for (int64_t n = 0; n < my_tiles; ++n) {
                phase_a(n);
                if (n > 0) phase_b(n - 1);
                nz_prev = nz_cur;
            }
            phase_b(my_tiles - 1);
            // ---- flush the chunk's weight gradients (column group 0: one warp per lane quarter; thread <-> hidden unit)
            if (cg == 0) {
                tc::mbar_wait(&bars.flush_full, 0);
                tc::tc_fence_after();
                const int h = c * CH + hl;
                uint32_t v[32];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    tc::tmem_ld32(tmem + lane_base + COL_DW1 + 32 * half, v);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (32 * half + j < p.d) atomicAdd(p.dW1 + (size_t)h * p.d + 32 * half + j, __uint_as_float(v[j]));
                    tc::tmem_ld32(tmem + lane_base + COL_DW2 + 32 * half, v);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (32 * half + j < p.d)
                            atomicAdd(p.dW2 + (size_t)(32 * half + j) * p.ff + h, __uint_as_float(v[j]) * (0.5f * p.hidden_scale));   // H^T was 2 h
                }
                uint32_t b16[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(b16[0]), "=r"(b16[1]), "=r"(b16[2]), "=r"(b16[3]), "=r"(b16[4]), "=r"(b16[5]), "=r"(b16[6]),
                               "=r"(b16[7]), "=r"(b16[8]), "=r"(b16[9]), "=r"(b16[10]), "=r"(b16[11]), "=r"(b16[12]), "=r"(b16[13]),
                               "=r"(b16[14]), "=r"(b16[15])
                             : "r"(tmem + lane_base + COL_DW1 + 64)
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

#ifdef U2GNN_PROBE_BUILD
extern uint32_t* g_ffn_trace;
#endif

size_t ffn_tc_mask_bytes(int64_t M, int ff) { return (size_t)((M + TM - 1) / TM) * (size_t)(ff / CH) * 4 * TM * sizeof(uint32_t); }

// internal launch used by u2gnn_ffn_tc_bwd (ffn_tc_bwd.cu)
int ffn_tc_wgrad_launch(const void* xb, const void* fb, int64_t M, int d, int ff, const void* packed, float hidden_scale,
                        uint64_t seed, uint32_t stream_hidden, int thr, float* dW1, float* db1, float* dW2, void* mask,
                        bool mask_from_forward, cudaStream_t st) {
    Params p;
    p.xb = static_cast<const uint8_t*>(xb); p.fb = static_cast<const uint8_t*>(fb); p.M = M; p.d = d; p.ff = ff;
    p.packed = static_cast<const uint8_t*>(packed);
    p.keys2 = rng_keys(seed, stream_hidden);
    p.thr = thr;
    p.low = rng_thr_low(thr);
    p.hidden_scale = hidden_scale;
    p.dW1 = dW1; p.db1 = db1; p.dW2 = dW2;
    p.mask = static_cast<uint32_t*>(mask);
    p.trace = nullptr;
    const size_t smem = 1024 + (size_t)WG_STAGES * 32768 + 16384 + 32768;
    if (mask_from_forward && !mask) return U2GNN_EINVAL;
    auto kern = mask_from_forward ? ffn_tc_wgrad_kernel<false, true> : ffn_tc_wgrad_kernel<false, false>;
#ifdef U2GNN_PROBE_BUILD
    p.trace = g_ffn_trace;
    if (p.trace) kern = mask_from_forward ? ffn_tc_wgrad_kernel<true, true> : ffn_tc_wgrad_kernel<true, false>;
#endif
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int NC = ff / CH;
    const int64_t n_tiles = (M + TM - 1) / TM;
    // 148 = 16 * 9 + 4: a whole number of slices per chunk leaves 4 SMs without a CTA.  Measured (round 2, tools/ab_lib.sh,
    // profiles/r02_ab_wgrad_tail_ctas.txt): four extra CTAs that take the last 1/37 of the tiles for four chunks each (flush + weight
    // reload between passes; per-CTA tile counts equal, results identical) make the kernel SLOWER - backward 8.37 -> 8.55-8.92 ms per
    // 4.46 M rows, step 53.6 -> 54.0 ms at the same 1 845 MHz: the step runs at the board's power cap, where 2.7 % more busy SMs buy
    // no throughput and the extra passes' fill / drain / flush are pure cost.  Not kept.
    int n_slices = U2GNN_NUM_SMS / NC;
    if (n_slices < 1) n_slices = 1;
    if (n_slices > n_tiles) n_slices = (int)n_tiles;
    kern<<<NC * n_slices, kThreads, smem, st>>>(p);
    return U2GNN_OK;
}
