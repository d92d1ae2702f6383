// Fused bf16 FFN block, weight gradients (wgrad):  dW1, db1, dW2  with the hidden recomputed on chip.
// One 128-wide ff chunk per CTA (its W1c / W2Tc images stay resident in shared memory); the CTAs of a chunk split
// the 128-row tiles among themselves.  Per tile (parity i selects the TMEM region and the epilogue group):
//     R_i = X  W1c^T            (SS, N 128)   epilogue A: H = relu(bf16(S)+b1) & keep              (registers)
//     R_i = dF W2Tc^T           (SS, N 128)   epilogue B: dPre = bf16(D) & [H > 0]                 (registers)
//     H, dPre -> shared memory as [128 rows x 128 hidden] swizzled tiles = MN-major A operands
//     dW2c^T[hidden, d]      += H^T    dF          (SS, MN-major A and B, N 64)
//     [dW1c | db1c][hidden,] += dPre^T [X | 1]     (SS, N 80: the second MN group of B is a tile of ones)
// accumulated in tensor memory over all tiles of the CTA and flushed once with atomics.
// X / dF row tiles are converted to bf16 swizzled tiles by two loader warps through a 3-stage ring; the same
// tile is the K-major A operand of the S / D GEMMs and the MN-major B operand of the gradient GEMMs.
// TMEM columns: R0 [0,128) R1 [128,256) dW2c^T [256,320) dW1c [320,384) db1c [384,400) dY partial [400,464).
//
// MERGED variant (u2gnn_ffn_tc_bwd_mode(1); measured slower than dgrad + wgrad, see ffn_tc_bwd.cu): the kernel ALSO computes
// the input gradient.  After the weight-gradient GEMMs of
// a tile it issues  dYp = dPre W1Tc^T  (SS: the dPre tile already in shared memory is the K-major A operand, N 64) and the
// epilogue warps add the partial of their chunk into dy1 (pre-set to dz) with red.global.add.v4.f32 - a quad of lanes
// transposes its 4 x 4 block of 16-byte pieces with shuffles first so that one instruction covers 64 contiguous bytes of
// a row (full 32-byte sectors; tools/probe_red.py: coalesced L2 reductions sustain 4.8 TB/s, half-sector ones 2.7).
// This replaces the separate dgrad kernel: the hidden is recomputed once instead of twice (7 executed GEMM units per
// tile-chunk instead of 9).
#include "common.cuh"
#include "rng.cuh"
#include "tc_common.cuh"
#include "ffn_epi.cuh"

namespace {

constexpr int DP = 64, CH = 128, TM = 128;
constexpr uint32_t CHUNK_BYTES = 4 * 16384;   // [W2c | W1c | W2Tc | W1Tc]
constexpr int kThreads = 640;
constexpr int WG_STAGES = 3;
constexpr uint32_t COL_R = 0, COL_DW2 = 256, COL_DW1 = 320, COL_DY = 400;

struct Params {
    const uint8_t* xb;   // bf16 swizzled tile images written by the dgrad kernel
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
    float* dy1;   // MERGED: [M, d] input gradient, pre-set to dz; the chunk partials are added into it
    uint32_t* trace;   // debug clock stamps of CTA 0 (u2gnn_ffn_tc_set_trace)
};
constexpr int TRACE_CAP = 1024;

struct __align__(8) Bars {
    uint64_t w_full, ld_full[WG_STAGES], ld_free[WG_STAGES], s_full[2], a_done[2], d_full[2], b_done[2], hp_full, hp_free, flush_full, dy_full, dy_free;
};

__device__ __forceinline__ void commit_to(uint64_t* bar) {
    if (tc::elect_one()) tc::mma_commit(bar);
    __syncwarp();
}
// R = A_tile(K-major, 4 k-steps) * B_image^T, N = 128
__device__ __forceinline__ void issue_n128(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    if (tc::elect_one()) {
        tc::mma_ss(tmem_d, a_desc, b_desc, idesc, 0);
        tc::mma_ss_acc(tmem_d, a_desc + 2, b_desc + 2, idesc);
        tc::mma_ss_acc(tmem_d, a_desc + 4, b_desc + 4, idesc);
        tc::mma_ss_acc(tmem_d, a_desc + 6, b_desc + 6, idesc);
    }
    __syncwarp();
}
// acc += A^T B over the 128 rows of the tile: both operands MN-major, 8 k-steps of 16 rows (2048 B each)
__device__ __forceinline__ void issue_wgrad(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    if (tc::elect_one()) {
        tc::mma_ss(tmem_d, a_desc, b_desc, idesc, acc);
#pragma unroll
        for (int ks = 1; ks < 8; ++ks) tc::mma_ss_acc(tmem_d, a_desc + 128 * ks, b_desc + 128 * ks, idesc);
    }
    __syncwarp();
}

template <bool TRACE, bool MERGED>
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
    uint8_t* sH = sOnes + 16384;                           // 32 KB: two [128 rows x 64 hidden] tiles
    uint8_t* sP = sH + 32768;                              // 32 KB
    uint8_t* sW = sP + 32768;                              // 32 KB: [W1c | W2Tc]  (MERGED: 48 KB, + W1Tc)
    constexpr uint32_t W_BYTES = MERGED ? 49152 : 32768;
    uint32_t* sB1h = reinterpret_cast<uint32_t*>(sW + W_BYTES);   // chunk bias as 64 packed bf16 pairs
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
        tc::mbar_init(&bars.hp_full, 16);
        tc::mbar_init(&bars.hp_free, 1);
        tc::mbar_init(&bars.flush_full, 1);
        tc::mbar_init(&bars.dy_full, 1);
        tc::mbar_init(&bars.dy_free, 16);
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc<512>(&tmem_slot);
    {
        const float* b1g = reinterpret_cast<const float*>(p.packed + (size_t)NC * CHUNK_BYTES) + c * CH;
        for (int e = threadIdx.x; e < CH / 2; e += kThreads) sB1h[e] = epi::cvt2(b1g[2 * e], b1g[2 * e + 1]);
        for (int e = threadIdx.x; e < 16384 / 4; e += kThreads) reinterpret_cast<uint32_t*>(sOnes)[e] = 0x3F803F80u;
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (my_tiles > 0) {
        if (warp == 0) {
            if (lane == 0) {
                tc::mbar_arrive_expect_tx(&bars.w_full, W_BYTES);
                tc::bulk_g2s(sW, p.packed + (size_t)c * CHUNK_BYTES + 16384, W_BYTES, &bars.w_full);
            }
        } else if (warp == 2) {
            // ================= row-tile producer: two 16 KB bulk copies per tile (images written by dgrad) ==========
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
            const uint32_t idesc_n128 = tc::make_idesc(TM, CH, 0, 0);
            const uint32_t idesc_w2 = tc::make_idesc(CH, DP, 1, 1);        // M = hidden, N = d, both MN-major
            const uint32_t idesc_w1 = tc::make_idesc(CH, DP + 16, 1, 1);   // N = 80: [X | ones]
            const uint32_t idesc_dy = tc::make_idesc(TM, DP, 0, 0);        // dYp: K-major A (dPre) and B (W1Tc)
            const uint64_t xf0 = tc::make_desc_sw128(tc::smem_u32(sXF), 16, 1024);         // K-major view of stage 0 X tile
            const uint64_t w1d = tc::make_desc_sw128(tc::smem_u32(sW), 16, 1024);
            const uint64_t w2td = w1d + 1024;
            const uint64_t hd = tc::make_desc_sw128(tc::smem_u32(sH), 16384, 1024);        // MN-major A: LBO = next 64-hidden tile
            const uint64_t pd = tc::make_desc_sw128(tc::smem_u32(sP), 16384, 1024);
            // Issue order (tensor pipe executes in order):  S(0) | D(0) S(1) | D(1) S(2) W(0) | D(2) S(3) W(1) | ...
            // i.e. the S / D GEMMs run one tile ahead of the weight-gradient GEMMs, so that D(n+1) does not queue behind
            // W(n) and the epilogue group of tile n+1 computes dPre while the other group is still storing H / dPre of
            // tile n.  S(n+2) overwrites the region D(n) lived in: it waits for b_done(n) (D(n) is in registers).
            tc::mbar_wait(&bars.w_full, 0);
            tc::mbar_wait(&bars.ld_full[0], 0);
            tc::tc_fence_after();
            issue_n128(tmem + COL_R, xf0, w1d, idesc_n128);
            commit_to(&bars.s_full[0]);
            auto issue_w = [&](int64_t n) {
                const uint32_t s = (uint32_t)(n % WG_STAGES);
                stamp(0);
                tc::mbar_wait(&bars.hp_full, (uint32_t)n & 1);
                stamp(0);
                tc::tc_fence_after();
                // MN-major B views of the same X / dF tiles: dF has one 64-wide group; [X | ones] has two, the second
                // one LBO bytes further (the ones tile)
                const uint32_t x_addr = tc::smem_u32(sXF) + s * 32768;
                const uint64_t fd_mn = tc::make_desc_sw128(x_addr + 16384, 16384, 1024);
                const uint64_t xd_mn = tc::make_desc_sw128(x_addr, tc::smem_u32(sOnes) - x_addr, 1024);
                issue_wgrad(tmem + COL_DW2, hd, fd_mn, idesc_w2, n > 0);
                issue_wgrad(tmem + COL_DW1, pd, xd_mn, idesc_w1, n > 0);
                if (MERGED) {
                    // dYp(n) = dPre(n) W1Tc^T: A = the dPre tile as K-major operand (two 64-hidden tiles), B = W1Tc image (two K atoms)
                    if (n > 0) tc::mbar_wait(&bars.dy_free, (uint32_t)(n - 1) & 1);
                    tc::tc_fence_after();
                    if (tc::elect_one()) {
                        const uint64_t pk = tc::make_desc_sw128(tc::smem_u32(sP), 16, 1024);
                        const uint64_t wt = tc::make_desc_sw128(tc::smem_u32(sW) + 32768, 16, 1024);
                        tc::mma_ss(tmem + COL_DY, pk, wt, idesc_dy, 0);
                        tc::mma_ss_acc(tmem + COL_DY, pk + 2, wt + 2, idesc_dy);
                        tc::mma_ss_acc(tmem + COL_DY, pk + 4, wt + 4, idesc_dy);
                        tc::mma_ss_acc(tmem + COL_DY, pk + 6, wt + 6, idesc_dy);
                        tc::mma_ss_acc(tmem + COL_DY, pk + 1024, wt + 512, idesc_dy);
                        tc::mma_ss_acc(tmem + COL_DY, pk + 1026, wt + 514, idesc_dy);
                        tc::mma_ss_acc(tmem + COL_DY, pk + 1028, wt + 516, idesc_dy);
                        tc::mma_ss_acc(tmem + COL_DY, pk + 1030, wt + 518, idesc_dy);
                    }
                    __syncwarp();
                    commit_to(&bars.dy_full);
                }
                commit_to(&bars.hp_free);
                commit_to(&bars.ld_free[s]);
                stamp(0);
            };
            for (int64_t n = 0; n < my_tiles; ++n) {
                const uint32_t i = (uint32_t)(n & 1), ph = (uint32_t)(n >> 1) & 1;
                const uint32_t s = (uint32_t)(n % WG_STAGES);
                const uint64_t xd = xf0 + (uint64_t)(s * 2048);            // X tile of this stage (K-major view)
                stamp(0);
                tc::mbar_wait(&bars.a_done[i], ph);
                stamp(0);
                tc::tc_fence_after();
                issue_n128(tmem + COL_R + 128 * i, xd + 1024, w2td, idesc_n128);            // D(n) = dF W2Tc^T
                commit_to(&bars.d_full[i]);
                if (n + 1 < my_tiles) {
                    const uint32_t s2 = (uint32_t)((n + 1) % WG_STAGES);
                    tc::mbar_wait(&bars.ld_full[s2], (uint32_t)((n + 1) / WG_STAGES) & 1);
                    if (n >= 1) tc::mbar_wait(&bars.b_done[i ^ 1], (uint32_t)((n - 1) >> 1) & 1);   // D(n-1) has left R_{i^1}
                    tc::tc_fence_after();
                    issue_n128(tmem + COL_R + 128 * (i ^ 1), xf0 + (uint64_t)(s2 * 2048), w1d, idesc_n128);   // S(n+1)
                    commit_to(&bars.s_full[i ^ 1]);
                }
                if (n >= 1) issue_w(n - 1);
            }
            issue_w(my_tiles - 1);
            commit_to(&bars.flush_full);
        } else if (warp >= 4) {
            // ================= epilogue: 16 warps = 4 hidden quarters (32 columns) x 4 TMEM lane quarters, EVERY tile ======
            // Software-pipelined per warp:  A(n) | B(n-1) store(n-1) | A(n+1) | B(n) store(n) | ...  so that D(n) and
            // S(n+1) are computed by the tensor pipe while the warps finish tile n-1.
            const int ew = warp - 4;
            const int cq = ew >> 2;                         // hidden columns [32 cq, 32 cq + 32) of the chunk
            const int wq = warp & 3;
            const int tr = wq * 32 + lane;
            const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
            const uint32_t r_addr0 = tmem + lane_base + COL_R + 32 * cq;
            const uint32_t b1_addr = tc::smem_u32(sB1h) + 64u * cq;
            const uint32_t hp_off = (uint32_t)((cq >> 1) * 16384) + (uint32_t)tr * 128u;
            const int ch0 = (cq & 1) * 4;
            const int thr = p.thr, low = p.low;
            const RngKeys keys2 = p.keys2;
            const uint32_t g_per_row = (uint32_t)(p.ff >> 5);
            uint32_t hcur[16], hprev[16];
            uint32_t bw[16];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) tc::lds128(b1_addr + 16u * q4, bw[4 * q4], bw[4 * q4 + 1], bw[4 * q4 + 2], bw[4 * q4 + 3]);
            auto phase_a = [&](int64_t n) {                // S(n) -> H = relu(bf16(S) + b1) & keep      (registers hcur)
                const uint32_t i = (uint32_t)(n & 1);
                const int64_t row = ((int64_t)slice + n * n_slices) * TM + tr;
                uint32_t k0 = 0xFFFFFFFFu;
                if (thr) k0 = rng_keep_word_lo(keys2, (uint64_t)row * g_per_row + (uint64_t)(4 * c + cq), thr, low);
                stamp(warp - 3);
                tc::mbar_wait(&bars.s_full[i], (uint32_t)(n >> 1) & 1);
                stamp(warp - 3);
                tc::tc_fence_after();
                uint32_t v[32];
                tc::tmem_ld32(r_addr0 + 128 * i, v);
                uint32_t km[16];
                if (thr) epi::keep_masks16(k0, km);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    uint32_t h2 = epi::relu_bias2(epi::cvt2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), bw[j]);
                    if (thr) h2 &= km[j];
                    hcur[j] = h2;
                }
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars.a_done[i]);
                stamp(warp - 3);
            };
            auto phase_b = [&](int64_t n) {                // D(n) -> dPre = bf16(D) & [H > 0];  H, dPre -> shared memory
                const uint32_t i = (uint32_t)(n & 1);
                stamp(warp - 3);
                tc::mbar_wait(&bars.d_full[i], (uint32_t)(n >> 1) & 1);
                stamp(warp - 3);
                tc::tc_fence_after();
                uint32_t v[32];
                tc::tmem_ld32(r_addr0 + 128 * i, v);
                tc::tmem_ld_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars.b_done[i]);
                uint32_t pr[16];
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    pr[j] = epi::cvt2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])) & epi::gt0_mask2(hprev[j]);
                stamp(warp - 3);
                if (n > 0) tc::mbar_wait(&bars.hp_free, (uint32_t)(n - 1) & 1);
                stamp(warp - 3);
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    const uint32_t off = hp_off + (uint32_t)(((ch0 + ch) ^ (tr & 7)) << 4);
                    *reinterpret_cast<uint4*>(sH + off) = make_uint4(hprev[4 * ch], hprev[4 * ch + 1], hprev[4 * ch + 2], hprev[4 * ch + 3]);
                    *reinterpret_cast<uint4*>(sP + off) = make_uint4(pr[4 * ch], pr[4 * ch + 1], pr[4 * ch + 2], pr[4 * ch + 3]);
                }
                tc::fence_proxy_async();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars.hp_full);
                stamp(warp - 3);
            };
            // MERGED: this warp's 32 rows x 16 columns of dYp(n) -> dy1.  Thread = row holds 16 consecutive floats (4 pieces of
            // 16 bytes); the four lanes of a quad exchange pieces (4 x 4 transpose) so that lane i of the quad owns piece i of
            // the quad's four rows: one red.global.add.v4.f32 then covers 64 contiguous bytes of one row per quad.
            auto drain = [&](int64_t n) {
                tc::mbar_wait(&bars.dy_full, (uint32_t)n & 1);
                tc::tc_fence_after();
                uint32_t y[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(y[0]), "=r"(y[1]), "=r"(y[2]), "=r"(y[3]), "=r"(y[4]), "=r"(y[5]), "=r"(y[6]), "=r"(y[7]),
                               "=r"(y[8]), "=r"(y[9]), "=r"(y[10]), "=r"(y[11]), "=r"(y[12]), "=r"(y[13]), "=r"(y[14]), "=r"(y[15])
                             : "r"(tmem + lane_base + COL_DY + 16 * cq)
                             : "memory");
                tc::tmem_ld_wait();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars.dy_free);
                // 4 x 4 transpose of 16-byte pieces inside each quad: after it, lane q of the quad holds piece q of rows 0..3
                const int ql = lane & 3;
#pragma unroll
                for (int step = 1; step <= 2; step <<= 1) {
                    // exchange with lane ^ step: the pieces whose index bit `step` differs from this lane's bit
#pragma unroll
                    for (int pc = 0; pc < 4; ++pc) {
                        if ((pc & step) == 0) {
                            const int hi = pc | step;                 // pieces (pc, hi) form a pair for this step
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const bool up = (ql & step) != 0;     // lanes with the bit set keep `hi`, send `pc`
                                const uint32_t send = up ? y[4 * pc + k] : y[4 * hi + k];
                                const uint32_t got = __shfl_xor_sync(0xffffffffu, send, step);
                                if (up) y[4 * pc + k] = got; else y[4 * hi + k] = got;
                            }
                        }
                    }
                }
                // now y[4 * j .. 4 * j + 3] = piece `ql` of the quad's row j
                const int64_t row_q = ((int64_t)slice + n * n_slices) * TM + (tr & ~3);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int64_t rg = row_q + j;
                    if (rg < p.M) {
                        if (p.d == DP) {
                            float* dst = p.dy1 + rg * DP + 16 * cq + 4 * ql;
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__uint_as_float(y[4 * j])),
                                         "f"(__uint_as_float(y[4 * j + 1])), "f"(__uint_as_float(y[4 * j + 2])), "f"(__uint_as_float(y[4 * j + 3]))
                                         : "memory");
                        } else {                                   // unpadded rows of d < 64 floats
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int col = 16 * cq + 4 * ql + k;
                                if (col < p.d) atomicAdd(p.dy1 + rg * p.d + col, __uint_as_float(y[4 * j + k]));
                            }
                        }
                    }
                }
            };
            for (int64_t n = 0; n < my_tiles; ++n) {
                phase_a(n);
                if (MERGED && n >= 2) drain(n - 2);
                if (n > 0) phase_b(n - 1);
#pragma unroll
                for (int j = 0; j < 16; ++j) hprev[j] = hcur[j];
            }
            if (MERGED && my_tiles >= 2) drain(my_tiles - 2);
            phase_b(my_tiles - 1);
            if (MERGED) drain(my_tiles - 1);
            const int i = ew >> 3, wg = (ew >> 2) & 1;      // flush: first warpgroup
            // ---- flush the chunk's weight gradients (first warpgroup; thread <-> hidden unit)
            if (i == 0 && wg == 0) {
                tc::mbar_wait(&bars.flush_full, 0);
                tc::tc_fence_after();
                const int h = c * CH + tr;
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
                            atomicAdd(p.dW2 + (size_t)(32 * half + j) * p.ff + h, __uint_as_float(v[j]) * p.hidden_scale);
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

// internal launch used by u2gnn_ffn_tc_bwd (ffn_tc_bwd.cu)
int ffn_tc_wgrad_launch(const void* xb, const void* fb, int64_t M, int d, int ff, const void* packed, float hidden_scale,
                        uint64_t seed, uint32_t stream_hidden, int thr, float* dW1, float* db1, float* dW2, float* dy1_merged,
                        cudaStream_t st) {
    Params p;
    p.xb = static_cast<const uint8_t*>(xb); p.fb = static_cast<const uint8_t*>(fb); p.M = M; p.d = d; p.ff = ff;
    p.packed = static_cast<const uint8_t*>(packed);
    p.keys2 = rng_keys(seed, stream_hidden);
    p.thr = thr;
    p.low = rng_thr_low(thr);
    p.hidden_scale = hidden_scale;
    p.dW1 = dW1; p.db1 = db1; p.dW2 = dW2; p.dy1 = dy1_merged;
    const size_t smem = 1024 + (size_t)WG_STAGES * 32768 + 16384 + 3 * 32768 + (dy1_merged ? 16384 : 0) + 512;
    extern uint32_t* g_ffn_trace;
    p.trace = g_ffn_trace;
    auto kern = dy1_merged ? (p.trace ? ffn_tc_wgrad_kernel<true, true> : ffn_tc_wgrad_kernel<false, true>)
                           : (p.trace ? ffn_tc_wgrad_kernel<true, false> : ffn_tc_wgrad_kernel<false, false>);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int NC = ff / CH;
    const int64_t n_tiles = (M + TM - 1) / TM;
    int n_slices = U2GNN_NUM_SMS / NC;
    if (n_slices < 1) n_slices = 1;
    if (n_slices > n_tiles) n_slices = (int)n_tiles;
    kern<<<NC * n_slices, kThreads, smem, st>>>(p);
    return U2GNN_OK;
}
