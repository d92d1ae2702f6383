// Short-sequence self-attention core on the tensor cores (bf16 mode, attn_axis="neighbors", d = 64, S <= 32).
//
// A CTA tile holds NB = 128 / S whole node sequences (S = 17 -> 7 nodes, 119 of 128 rows).  The score matrix of the
// tile is ONE 128 x 128 x 64 MMA (S_all = Q K^T); only its block diagonal (each node's S x S block) is meaningful
// and is what the per-row softmax threads read back from tensor memory.  The off-diagonal work is wasted flops on
// a pipe that is otherwise idle here; what it buys is that every product of the forward and backward is a dense
// 128-row MMA on operands that are staged once:
//   forward :  S = Q K^T,  P~ = dropout(softmax(S / sqrt(d))) (block diagonal),  ctx = P~ V
//   backward:  S = Q K^T,  dP = G V^T,  per row: p, ds = p (dp - sum p dp) / sqrt(d),
//              dV = P~^T G,  dK = dS^T Q,  dQ = dS K
// Q, K, V, G tiles are [128 x 64] bf16 swizzled tiles that serve as K-major A / B operands and, unchanged, as
// MN-major B operands; P~ and dS are written to shared memory as [128 x 128] tiles (zero outside the diagonal
// blocks) and used both K-major (dQ, ctx) and MN-major (dV, dK).  Same arithmetic as seqattn.cu up to bf16
// operand rounding; same dropout stream.  Phase-serial CTAs, two per SM.
#include "common.cuh"
#include "rng.cuh"
#include "tc_common.cuh"
#include "ffn_epi.cuh"

namespace {

constexpr int D = 64, TM = 128;
constexpr int kThreads = 256;

struct AttnRng {
    RngKeys keys;
    int thr;
    float scale;
};

// rows [row0, row0 + 128) of a strided matrix (64 columns at column offset `coff`; fp32, or bf16 when `bf16_src`: then the
// producer has already rounded and the tile is a plain copy at half the bytes) -> bf16 swizzled tile
// idx (fp32 sources only): row r of the tile is row idx[row0 + r] of a table with n_table rows - the gather of
// F.embedding(input_x, X_concat) (pytorch_U2GNN_Sup.py:32) fused into its consumer; an index outside the table gives a zero
// row and sets bit 1 of the device error word (as u2gnn_gather_rows does)
__device__ __forceinline__ void stage_tile(uint8_t* tile, const void* __restrict__ src_, int bf16_src, int64_t ld, int coff, int64_t row0,
                                           int valid_rows, int tid, const int64_t* __restrict__ idx = nullptr, int64_t n_table = 0,
                                           int* err = nullptr) {
    if (bf16_src) {                                             // (the kernel stages all bf16 tiles at once: stage_tiles_bf16)
        const __nv_bfloat16* src = static_cast<const __nv_bfloat16*>(src_);
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = u * kThreads + tid;
            const int r = e >> 3, c8 = e & 7;
            v[u] = make_uint4(0u, 0u, 0u, 0u);
            if (r < valid_rows) v[u] = __ldg(reinterpret_cast<const uint4*>(src + (row0 + r) * ld + coff) + c8);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = u * kThreads + tid;
            *reinterpret_cast<uint4*>(tile + tc::sw128_chunk(e >> 3, e & 7)) = v[u];
        }
        return;
    }
    const float* src = static_cast<const float*>(src_);
    for (int base = 0; base < TM * 16; base += kThreads * 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = base + u * kThreads + tid;
            const int r = e >> 4, c4 = e & 15;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < valid_rows) {
                int64_t row = row0 + r;
                if (idx) {
                    row = __ldg(idx + row);
                    if (row < 0 || row >= n_table) {
                        if (c4 == 0 && err) atomicOr(err, 2);
                        continue;
                    }
                }
                v[u] = __ldg(reinterpret_cast<const float4*>(src + row * ld + coff) + c4);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = base + u * kThreads + tid;
            const int r = e >> 4, c4 = e & 15;
            uint2 w;
            w.x = epi::cvt2(v[u].x, v[u].y);
            w.y = epi::cvt2(v[u].z, v[u].w);
            *reinterpret_cast<uint2*>(tile + tc::sw128_offset(r, c4 * 4)) = w;
        }
    }
}

// one row of a [128 x 128] bf16 operand (two [128 x 64] swizzled tiles `t0`, `t1`): zero it, then place `n` values at
// columns col0 .. col0 + n - 1.  Fully unrolled with predicates so that `vals` stays in registers.
template <int SMAX>
__device__ __forceinline__ void write_diag_row(uint8_t* t0, uint8_t* t1, int r, int col0, int n, const float (&vals)[SMAX]) {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
        *reinterpret_cast<uint4*>(t0 + tc::sw128_chunk(r, ch)) = z;
        *reinterpret_cast<uint4*>(t1 + tc::sw128_chunk(r, ch)) = z;
    }
#pragma unroll
    for (int j = 0; j < SMAX; ++j) {
        if (j < n) {
            const int c = col0 + j;
            uint8_t* t = (c < 64) ? t0 : t1;
            *reinterpret_cast<__nv_bfloat16*>(t + tc::sw128_offset(r, c & 63)) = __float2bfloat16(vals[j]);
        }
    }
}

// D[128 x N] (+)= A(K-major tile(s), K = 64 * kt) * B(K-major tile)^T : used for S = Q K^T and dP = G V^T (kt = 1, N = 128)
__device__ __forceinline__ void mma_kk(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, uint32_t idesc) {
    const uint64_t ad = tc::make_desc_sw128(a_addr, 16, 1024), bd = tc::make_desc_sw128(b_addr, 16, 1024);
    tc::mma_ss(tmem_d, ad, bd, idesc, 0);
    tc::mma_ss_acc(tmem_d, ad + 2, bd + 2, idesc);
    tc::mma_ss_acc(tmem_d, ad + 4, bd + 4, idesc);
    tc::mma_ss_acc(tmem_d, ad + 6, bd + 6, idesc);
}

struct Params {
    const void* qkv;
    const void* dctx;      // backward only
    void* out;             // forward: ctx[B*S, 64]; backward: dqkv[B*S, 192]
    int io_bf16;           // qkv, dctx and out are bf16 row-major instead of fp32
    int64_t B;
    int S;
    AttnRng rng;
    int low;
    // INPROJ forward: the tile's qkv rows are computed here (qkv = x W_in^T + b_in, nn.MultiheadAttention in_proj) instead of
    // being read back: x fp32 [B*S, 64], w_in [192, 64], b_in [192]; qkv_out bf16 [B*S, 192] is still written (the backward
    // reads it)
    const float* x;
    const float* w_in;
    const float* b_in;
    void* qkv_out;
    const int64_t* x_idx;  // optional: x row r = x[x_idx[r]] (gather fused into the kernel), x_rows = rows of that table
    int64_t x_rows;
    int* err;
};

// shared memory layout (bytes): Q 0, K 16K, V 32K; forward: P~ over Q / K; backward: P~ tile 0 over V, G 48K, P~1 64K, dS 80K / 96K
template <bool BWD, int SMAX, bool INPROJ = false>   // SMAX: compile-time bound of the sequence length (register arrays)
__global__ void __launch_bounds__(kThreads, BWD ? 2 : 3) attn_tc_kernel(const Params p) {
    static_assert(!(BWD && INPROJ), "the fused in_proj exists for the forward only");
    // no static shared memory and a 1024-byte aligned dynamic segment: the backward's seven 16 KB tiles (+ 16 bytes for
    // the barrier and the TMEM slot behind them) then fit TWICE into an SM, so two phase-serial CTAs overlap
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((tc::smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* sQ = smem;
    uint8_t* sK = smem + 16384;
    uint8_t* sV = smem + 32768;
    uint8_t* sG = smem + 49152;                                  // backward only
    // forward: P~ goes over Q / K (dead once S = Q K^T has completed), so a forward CTA needs three tiles and 128 tensor-memory
    // columns (ctx reuses the S columns) and THREE CTAs fit an SM: the tile is a serial chain (load -> S -> softmax -> ctx ->
    // store), only more tiles in flight hide it
    uint8_t* sP0 = BWD ? sV : sQ;                                // P~ columns 0..63   (backward: over V, dead after dP)
    uint8_t* sP1 = BWD ? smem + 65536 : sK;                      // P~ columns 64..127
    uint8_t* sS0 = smem + 81920;                                 // dS columns 0..63   (backward only)
    uint8_t* sS1 = smem + 98304;
    // INPROJ: the x tile is staged over V (dead until the V columns are drained), W_in as a [192 x 64] K-major image behind V
    uint8_t* sX = sV;
    uint8_t* sWin = smem + 49152;
    constexpr int kBarOff = BWD ? 114688 : (INPROJ ? 73728 : 49152);
    uint64_t& bar = *reinterpret_cast<uint64_t*>(smem + kBarOff);
    uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(smem + kBarOff + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::fence_barrier_init();
    }
    constexpr int kTmemCols = BWD ? 256 : 128;
    if (warp == 0) tc::tmem_alloc<kTmemCols>(&tmem_slot);
    if constexpr (INPROJ) {
        for (int e = tid; e < 3 * D * D; e += kThreads)        // W_in[n][k] -> K-major B image, row n = output column
            *reinterpret_cast<__nv_bfloat16*>(sWin + tc::sw128_offset(e >> 6, e & 63)) = __float2bfloat16(p.w_in[e]);
        tc::fence_proxy_async();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int S = p.S, NB = TM / S;
    const float qscale = 0.125f;                                 // sqrt(1/64)
    const uint32_t idesc_s = tc::make_idesc(128, 128, 0, 0);     // S, dP: K-major A and B
    const uint32_t idesc_mn = tc::make_idesc(128, 64, 1, 1);     // dV, dK: MN-major A and B
    const uint32_t idesc_kmn = tc::make_idesc(128, 64, 0, 1);    // ctx, dQ: K-major A, MN-major B
    const int wq = warp & 3;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const int r = wq * 32 + lane;                                // row handled in the per-row phases
    const int64_t n_tiles = (p.B + NB - 1) / NB;
    uint32_t phase = 0;
    // Backward with bf16 I/O: the four operand tiles of the NEXT tile travel with cp.async (no registers, 16-byte chunks straight
    // into the swizzled images, rows past the tile zero-filled) while this tile's results are drained - the tile's global
    // loads were 25 % of the stall samples of the phase-serial version (profiles/r01_ncu_attn_bwd_v27_stalls.txt).
    constexpr bool kAsync = BWD && !INPROJ;
    auto prefetch_tiles = [&](int64_t t) {
        const int64_t n0 = t * NB;
        const int nrows = (int)((p.B - n0 < NB) ? p.B - n0 : NB) * S;
        const int64_t r0 = n0 * S;
        const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(p.qkv);
        const __nv_bfloat16* gsrc = static_cast<const __nv_bfloat16*>(p.dctx);
        const uint32_t base = tc::smem_u32(smem);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = u * kThreads + tid;
            const int rr = e >> 3, c8 = e & 7;
            const bool ok = rr < nrows;
            const uint32_t off = tc::sw128_chunk(rr, c8);
            const __nv_bfloat16* rowp = q + (r0 + (ok ? rr : 0)) * (3 * D) + 8 * c8;
            const uint32_t nb = ok ? 16u : 0u;
#pragma unroll
            for (int t3 = 0; t3 < 3; ++t3)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(base + (uint32_t)t3 * 16384u + off), "l"(rowp + t3 * D), "r"(nb) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(base + 49152u + off), "l"(gsrc + (r0 + (ok ? rr : 0)) * D + 8 * c8), "r"(nb) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (kAsync && p.io_bf16 && (int64_t)blockIdx.x < n_tiles) prefetch_tiles(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t node0 = tile * NB;
        const int nodes = (int)((p.B - node0 < NB) ? p.B - node0 : NB);
        const int rows = nodes * S;
        const int64_t row0 = node0 * S;
        // ---- 1. stage operands
        if constexpr (INPROJ) {
            // qkv rows of the tile from x: two MMA passes through the SAME 128 tensor-memory columns (Q | K, then V) so that the
            // CTA stays at 128 columns and three CTAs per SM; thread = row drains them (+ bias, rounded to bf16 once - the value
            // the separate projection kernel stores) into the swizzled operand tiles
            const int half = warp >> 2;
            stage_tile(sX, p.x, 0, D, 0, row0, rows, tid, p.x_idx, p.x_rows, p.err);
            tc::fence_proxy_async();
            tc::tc_fence_before();
            __syncthreads();
            tc::tc_fence_after();
            if (warp == 0) {
                if (tc::elect_one()) {
                    mma_kk(tmem, tc::smem_u32(sX), tc::smem_u32(sWin), tc::make_idesc(128, 128, 0, 0));
                    tc::mma_commit(&bar);
                }
                __syncwarp();
            }
            tc::mbar_wait(&bar, phase);
            phase ^= 1;
            tc::tc_fence_after();
            {
                uint8_t* tile = half ? sK : sQ;               // warps 0-3 drain Q (columns 0..63), warps 4-7 drain K
#pragma unroll
                for (int pc = 0; pc < 2; ++pc) {
                    uint32_t v[32];
                    tc::tmem_ld32(tmem + lane_base + 64 * half + 32 * pc, v);
                    tc::tmem_ld_wait();
                    const float4* b4 = reinterpret_cast<const float4*>(p.b_in + 64 * half + 32 * pc);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float4 ba = __ldg(b4 + 2 * c), bb = __ldg(b4 + 2 * c + 1);
                        uint4 w = make_uint4(0u, 0u, 0u, 0u);
                        if (r < rows) {
                            w.x = epi::cvt2(__uint_as_float(v[8 * c]) + ba.x, __uint_as_float(v[8 * c + 1]) + ba.y);
                            w.y = epi::cvt2(__uint_as_float(v[8 * c + 2]) + ba.z, __uint_as_float(v[8 * c + 3]) + ba.w);
                            w.z = epi::cvt2(__uint_as_float(v[8 * c + 4]) + bb.x, __uint_as_float(v[8 * c + 5]) + bb.y);
                            w.w = epi::cvt2(__uint_as_float(v[8 * c + 6]) + bb.z, __uint_as_float(v[8 * c + 7]) + bb.w);
                        }
                        *reinterpret_cast<uint4*>(tile + tc::sw128_chunk(r, 4 * pc + c)) = w;
                    }
                }
            }
            tc::tc_fence_before();
            __syncthreads();
            tc::tc_fence_after();
            if (warp == 0) {
                if (tc::elect_one()) {
                    mma_kk(tmem, tc::smem_u32(sX), tc::smem_u32(sWin) + 16384, tc::make_idesc(128, 64, 0, 0));   // V = x W_v^T
                    tc::mma_commit(&bar);
                }
                __syncwarp();
            }
            tc::mbar_wait(&bar, phase);
            phase ^= 1;
            tc::tc_fence_after();
            {
                uint32_t v[32];
                tc::tmem_ld32(tmem + lane_base + 32 * half, v);   // warps 0-3: V columns 0..31, warps 4-7: 32..63 (x is dead now)
                tc::tmem_ld_wait();
                const float4* b4 = reinterpret_cast<const float4*>(p.b_in + 128 + 32 * half);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 ba = __ldg(b4 + 2 * c), bb = __ldg(b4 + 2 * c + 1);
                    uint4 w = make_uint4(0u, 0u, 0u, 0u);
                    if (r < rows) {
                        w.x = epi::cvt2(__uint_as_float(v[8 * c]) + ba.x, __uint_as_float(v[8 * c + 1]) + ba.y);
                        w.y = epi::cvt2(__uint_as_float(v[8 * c + 2]) + ba.z, __uint_as_float(v[8 * c + 3]) + ba.w);
                        w.z = epi::cvt2(__uint_as_float(v[8 * c + 4]) + bb.x, __uint_as_float(v[8 * c + 5]) + bb.y);
                        w.w = epi::cvt2(__uint_as_float(v[8 * c + 6]) + bb.z, __uint_as_float(v[8 * c + 7]) + bb.w);
                    }
                    *reinterpret_cast<uint4*>(sV + tc::sw128_chunk(r, 4 * half + c)) = w;
                }
            }
        } else if (kAsync && p.io_bf16) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");     // this tile's operands were requested during the previous tile's epilogue
        } else if (p.io_bf16) {
            // bf16 sources: tiles are plain copies; every 16-byte load of all three / four tiles is in flight before the first
            // shared-memory store (one exposed memory latency per tile-set instead of one per tile)
            const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(p.qkv);
            const __nv_bfloat16* gsrc = static_cast<const __nv_bfloat16*>(p.dctx);
            uint4 v[BWD ? 4 : 3][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = u * kThreads + tid;
                const int rr = e >> 3, c8 = e & 7;
                const bool ok = rr < rows;
                const uint4* rowp = reinterpret_cast<const uint4*>(q + (row0 + rr) * (3 * D));
#pragma unroll
                for (int t = 0; t < 3; ++t) v[t][u] = ok ? __ldg(rowp + 8 * t + c8) : make_uint4(0u, 0u, 0u, 0u);
                if (BWD) v[BWD ? 3 : 0][u] = ok ? __ldg(reinterpret_cast<const uint4*>(gsrc + (row0 + rr) * D) + c8) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = u * kThreads + tid;
                const uint32_t off = tc::sw128_chunk(e >> 3, e & 7);
                *reinterpret_cast<uint4*>(sQ + off) = v[0][u];
                *reinterpret_cast<uint4*>(sK + off) = v[1][u];
                *reinterpret_cast<uint4*>(sV + off) = v[2][u];
                if (BWD) *reinterpret_cast<uint4*>(sG + off) = v[BWD ? 3 : 0][u];
            }
        } else {
            stage_tile(sQ, p.qkv, 0, 3 * D, 0, row0, rows, tid);
            stage_tile(sK, p.qkv, 0, 3 * D, D, row0, rows, tid);
            stage_tile(sV, p.qkv, 0, 3 * D, 2 * D, row0, rows, tid);
            if (BWD) stage_tile(sG, p.dctx, 0, D, 0, row0, rows, tid);
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        // ---- 2. S = Q K^T (and dP = G V^T)
        if (warp == 0) {
            if (tc::elect_one()) {
                mma_kk(tmem, tc::smem_u32(sQ), tc::smem_u32(sK), idesc_s);
                if (BWD) mma_kk(tmem + 128, tc::smem_u32(sG), tc::smem_u32(sV), idesc_s);
                tc::mma_commit(&bar);
            }
            __syncwarp();
        }
        if constexpr (INPROJ) {
            // the three operand tiles -> qkv_out (bf16 row-major, coalesced 16-byte pieces) while the S MMA runs; the barrier
            // keeps the P~ rows (written over Q / K below) behind every thread's reads
            __nv_bfloat16* qo = static_cast<__nv_bfloat16*>(p.qkv_out);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = u * kThreads + tid;
                const int rr = e >> 3, c8 = e & 7;
                if (rr < rows) {
                    const uint32_t off = tc::sw128_chunk(rr, c8);
                    uint4* dst = reinterpret_cast<uint4*>(qo + (row0 + rr) * (3 * D)) + c8;
                    dst[0] = *reinterpret_cast<const uint4*>(sQ + off);
                    dst[8] = *reinterpret_cast<const uint4*>(sK + off);
                    dst[16] = *reinterpret_cast<const uint4*>(sV + off);
                }
            }
            __syncthreads();
        }
        tc::mbar_wait(&bar, phase);
        phase ^= 1;
        tc::tc_fence_after();
        // ---- 3. per-row softmax (threads 0..127: row r = node * S + i reads the S columns of its own node)
        // Forward: warps 0-3 (thread = row).  Backward: BOTH warp sets work on every row - warps 0-3 produce the P~ row,
        // warps 4-7 (same TMEM lane quarters) recompute the cheap softmax and produce the dS row - which halves the
        // longest phase of the backward tile and the registers each thread holds.
        if (warp < 4 || BWD) {
            const bool role_ds = BWD && warp >= 4;
            const bool live = r < rows;
            const int node = live ? r / S : 0, i = live ? r - node * S : 0;
            const int col0 = node * S;
            // tcgen05.ld takes ONE (warp-uniform) column address, but the 32 rows of a warp belong to up to three nodes
            // (S >= 11 for the tiles used here; four for shorter sequences).  The warp therefore loads the 32 columns
            // behind EACH candidate block start (all warp-uniform) and every lane keeps the registers of its own node:
            // selects on registers instead of a 96-column window indexed per lane in local memory.
            const int node_lo = (wq * 32) / S;                               // node of the warp's first row
            const int sel = node - node_lo;                                  // 0 .. 3
            const int n_cand = ((wq * 32 + 31) / S) - node_lo + 1;           // warp-uniform
            float sv[SMAX], dv[BWD ? SMAX : 1];                              // scores -> p / P~ ; dP -> dS
#pragma unroll
            for (int j = 0; j < SMAX; ++j) { sv[j] = 0.f; dv[BWD ? j : 0] = 0.f; }
            for (int cnd = 0; cnd < n_cand; ++cnd) {                         // warp-uniform trip count
                uint32_t t0[32];
                const int cb = (node_lo + cnd) * S;                          // <= 127: columns cb .. cb + 31 stay inside the allocation
                tc::tmem_ld32(tmem + lane_base + cb, t0);
                tc::tmem_ld_wait();
                const bool mine = live && (sel == cnd);
#pragma unroll
                for (int j = 0; j < SMAX; ++j) sv[j] = (mine && j < S) ? __uint_as_float(t0[j]) : sv[j];
                if (BWD && role_ds) {                                        // warp-uniform
                    // dP lives in columns [128, 256) = the end of the allocation: the 32-column window is clamped to it and
                    // the registers are shifted down by the (warp-uniform) difference with a barrel shifter on registers
                    const int cbl = cb < 96 ? cb : 96, dsh = cb - cbl;
                    tc::tmem_ld32(tmem + lane_base + 128 + cbl, t0);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int bit = 16; bit >= 1; bit >>= 1) {
                        if (dsh & bit) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) t0[j] = (j + bit < 32) ? t0[j + bit < 32 ? j + bit : 31] : 0u;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < SMAX; ++j) dv[BWD ? j : 0] = (mine && j < S) ? __uint_as_float(t0[j]) : dv[BWD ? j : 0];
                }
            }
            // dropout keep bits of this row: elements ebase .. ebase + S - 1 of the [B, S, S] probability tensor lie in at
            // most two 32-element groups of the engine's stream (one RNG word each instead of one per element)
            const uint64_t ebase = (uint64_t)((node0 + node) * S + i) * (uint64_t)S;
            uint32_t keep_bits = 0xFFFFFFFFu;                                // bit j = element ebase + j kept
            if (p.rng.thr && live) {
                const uint32_t w0 = rng_keep_word_lo(p.rng.keys, ebase >> 5, p.rng.thr, p.low);
                const uint32_t w1 = rng_keep_word_lo(p.rng.keys, (ebase >> 5) + 1, p.rng.thr, p.low);
                const uint32_t sh = (uint32_t)(ebase & 31);
                keep_bits = sh ? ((w0 >> sh) | (w1 << (32 - sh))) : w0;
            }
            const float dscale = p.rng.thr ? p.rng.scale : 1.0f;
            if (live) {
                float m = -INFINITY;
#pragma unroll
                for (int j = 0; j < SMAX; ++j)
                    if (j < S) m = fmaxf(m, sv[j] * qscale);
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < SMAX; ++j) {
                    sv[j] = (j < S) ? __expf(sv[j] * qscale - m) : 0.f;
                    sum += sv[j];
                }
                const float inv = 1.0f / sum;
                if (!role_ds) {
#pragma unroll
                    for (int j = 0; j < SMAX; ++j) sv[j] = sv[j] * inv * (((keep_bits >> j) & 1u) ? dscale : 0.0f);     // P~
                } else {
                    float tsum = 0.f;
#pragma unroll
                    for (int j = 0; j < SMAX; ++j) {
                        sv[j] *= inv;                                                                              // p
                        dv[BWD ? j : 0] *= (((keep_bits >> j) & 1u) ? dscale : 0.0f);                               // dp
                        tsum = fmaf(sv[j], dv[BWD ? j : 0], tsum);
                    }
#pragma unroll
                    for (int j = 0; j < SMAX; ++j) dv[BWD ? j : 0] = sv[j] * (dv[BWD ? j : 0] - tsum) * qscale;      // dS
                }
            }
            // the TMEM reads above must be finished before anyone overwrites S / dP (step 4 reuses the columns); the
            // shared-memory rows are written after the block-wide barrier below for the backward (P~ reuses V)
            tc::tc_fence_before();
            if (!role_ds) write_diag_row<SMAX>(sP0, sP1, r, col0, live ? S : 0, sv);
            if constexpr (BWD) {
                if (role_ds) write_diag_row<SMAX>(sS0, sS1, r, col0, live ? S : 0, dv);
            }
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        // ---- 4. second round of MMAs
        if (warp == 0) {
            if (tc::elect_one()) {
                if (!BWD) {
                    // ctx = P~ V : A = P~ K-major (K = 128 keys: 8 k-steps over two tiles), B = V MN-major (16 key rows per k-step)
                    const uint64_t bd = tc::make_desc_sw128(tc::smem_u32(sV), 16384, 1024);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint64_t ad = tc::make_desc_sw128(tc::smem_u32(ks < 4 ? sP0 : sP1) + (ks & 3) * 32, 16, 1024);
                        tc::mma_ss(tmem, ad, bd + 128 * ks, idesc_kmn, ks > 0);
                    }
                } else {
                    const uint32_t p0 = tc::smem_u32(sP0), s0 = tc::smem_u32(sS0);
                    const uint64_t pd_mn = tc::make_desc_sw128(p0, tc::smem_u32(sP1) - p0, 1024);     // MN-major A: M = 128 keys
                    const uint64_t ds_mn = tc::make_desc_sw128(s0, 16384, 1024);
                    const uint64_t g_mn = tc::make_desc_sw128(tc::smem_u32(sG), 16384, 1024);
                    const uint64_t q_mn = tc::make_desc_sw128(tc::smem_u32(sQ), 16384, 1024);
                    const uint64_t k_mn = tc::make_desc_sw128(tc::smem_u32(sK), 16384, 1024);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        tc::mma_ss(tmem + 128, pd_mn + 128 * ks, g_mn + 128 * ks, idesc_mn, ks > 0);       // dV = P~^T G
                        tc::mma_ss(tmem + 64, ds_mn + 128 * ks, q_mn + 128 * ks, idesc_mn, ks > 0);        // dK = dS^T Q
                        const uint64_t ad = tc::make_desc_sw128(tc::smem_u32(ks < 4 ? sS0 : sS1) + (ks & 3) * 32, 16, 1024);
                        tc::mma_ss(tmem, ad, k_mn + 128 * ks, idesc_kmn, ks > 0);                          // dQ = dS K
                    }
                }
                tc::mma_commit(&bar);
            }
            __syncwarp();
        }
        tc::mbar_wait(&bar, phase);
        phase ^= 1;
        tc::tc_fence_after();
        // ---- 5. write results (fp32): forward ctx[row, 64]; backward dqkv[row, 192] = [dQ | dK | dV].
        // Tensor memory -> registers (thread = row) -> fp32 staging in the now dead operand tiles (one 32 KB tile per 64
        // output columns, 16-byte chunks XOR-swizzled with the row so that both the row-per-thread writes and the
        // chunk-per-thread reads are conflict free) -> COALESCED 128-bit global stores.
        if (kAsync && p.io_bf16) {
            // every operand tile is dead: tiles 0-3 take the NEXT tile's operands (cp.async, in flight under the drain below),
            // tiles 4-6 stage this tile's dqkv rows as bf16 - 384 contiguous bytes per row, 16-byte chunks XOR-swizzled inside
            // each 128-byte piece - for plain coalesced 16-byte copies to global memory
            if (tile + gridDim.x < n_tiles) prefetch_tiles(tile + gridDim.x);
            const int half = warp >> 2;
            uint8_t* st = smem + 65536 + r * 384;
#pragma unroll
            for (int pc = 0; pc < 3; ++pc) {
                uint32_t v[32];
                tc::tmem_ld32(tmem + lane_base + 64 * pc + 32 * half, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 w;
                    w.x = epi::cvt2(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1]));
                    w.y = epi::cvt2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
                    w.z = epi::cvt2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
                    w.w = epi::cvt2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
                    *reinterpret_cast<uint4*>(st + 128 * pc + (((4 * half + j) ^ (r & 7)) << 4)) = w;
                }
            }
            tc::tc_fence_before();
            __syncthreads();
            __nv_bfloat16* outp = static_cast<__nv_bfloat16*>(p.out);
#pragma unroll
            for (int u = 0; u < 12; ++u) {
                const int e = u * kThreads + tid;
                const int rr = e / 24, c = e - rr * 24;
                if (rr < rows) {
                    const uint4 w = *reinterpret_cast<const uint4*>(smem + 65536 + rr * 384 + ((c & ~7) << 4) + (((c & 7) ^ (rr & 7)) << 4));
                    *reinterpret_cast<uint4*>(outp + (row0 + rr) * (3 * D) + 8 * c) = w;
                }
            }
        } else {
            constexpr int NP = BWD ? 3 : 1;                     // 64-column pieces
            const int half = warp >> 2;                         // 32-column half of a piece handled by this warp
#pragma unroll
            for (int pc = 0; pc < NP; ++pc) {
                uint32_t v[32];
                tc::tmem_ld32(tmem + lane_base + 64 * pc + 32 * half, v);
                tc::tmem_ld_wait();
                uint8_t* st = smem + pc * 32768 + r * 256;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<uint4*>(st + (((8 * half + j) ^ (r & 15)) << 4)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            tc::tc_fence_before();
            __syncthreads();
            if (p.io_bf16) {
#pragma unroll
                for (int u = 0; u < NP * 4; ++u) {
                    const int e = u * kThreads + tid;
                    const int pc = e >> 10, rr = (e >> 3) & 127, c8 = e & 7;
                    if (rr < rows) {
                        const uint8_t* st = smem + pc * 32768 + rr * 256;
                        const uint4 o0 = *reinterpret_cast<const uint4*>(st + (((2 * c8) ^ (rr & 15)) << 4));
                        const uint4 o1 = *reinterpret_cast<const uint4*>(st + (((2 * c8 + 1) ^ (rr & 15)) << 4));
                        uint4 w;
                        w.x = epi::cvt2(__uint_as_float(o0.x), __uint_as_float(o0.y));
                        w.y = epi::cvt2(__uint_as_float(o0.z), __uint_as_float(o0.w));
                        w.z = epi::cvt2(__uint_as_float(o1.x), __uint_as_float(o1.y));
                        w.w = epi::cvt2(__uint_as_float(o1.z), __uint_as_float(o1.w));
                        *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + (row0 + rr) * (BWD ? 3 * D : D) + 64 * pc + 8 * c8) = w;
                    }
                }
            } else
#pragma unroll
            for (int u = 0; u < NP * 8; ++u) {
                const int e = u * kThreads + tid;
                const int pc = e >> 11, rr = (e >> 4) & 127, c4 = e & 15;
                if (rr < rows) {
                    const uint4 o = *reinterpret_cast<const uint4*>(smem + pc * 32768 + rr * 256 + ((c4 ^ (rr & 15)) << 4));
                    const int64_t off = (row0 + rr) * (BWD ? 3 * D : D) + 64 * pc + 4 * c4;
                    if (p.io_bf16) {
                        uint2 w;
                        w.x = epi::cvt2(__uint_as_float(o.x), __uint_as_float(o.y));
                        w.y = epi::cvt2(__uint_as_float(o.z), __uint_as_float(o.w));
                        *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.out) + off) = w;
                    } else {
                        *reinterpret_cast<uint4*>(static_cast<float*>(p.out) + off) = o;
                    }
                }
            }
        }
        tc::tc_fence_before();
        __syncthreads();
    }
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<kTmemCols>(tmem);
}

AttnRng make_rng(uint64_t seed, uint32_t stream, int thr) {
    AttnRng r;
    r.keys = rng_keys(seed, stream);
    r.thr = thr;
    r.scale = thr ? rng_keep_scale(thr) : 1.0f;
    return r;
}

template <bool BWD, bool INPROJ = false>
int launch(const Params& p, cudaStream_t st) {
    const size_t smem = (BWD ? 114688 : (INPROJ ? 73728 : 49152)) + 16;   // tiles (+ W_in image) + barrier + TMEM slot
    auto k = (p.S <= 20) ? attn_tc_kernel<BWD, 20, INPROJ> : attn_tc_kernel<BWD, 32, INPROJ>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int NB = TM / p.S;
    const int64_t n_tiles = (p.B + NB - 1) / NB;
    const int64_t cap = (int64_t)U2GNN_NUM_SMS * (BWD ? 2 : 3);   // backward: two CTAs per SM (256 TMEM columns, seven tiles each); forward: three
    k<<<(int)(n_tiles < cap ? n_tiles : cap), kThreads, smem, st>>>(p);
    return U2GNN_OK;
}

}  // namespace

extern "C" int u2gnn_seqattn_tc_fwd_ex(const void* qkv, int64_t B, int S, int d, uint64_t seed, uint32_t rng_stream, int thr,
                                       void* ctx, int io_bf16, u2gnn_stream_t stream) {
    if (!qkv || !ctx || B < 0 || thr < 0 || thr > 255) return U2GNN_EINVAL;
    if (d != D || S < 2 || S > 32) return U2GNN_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(ctx)) % 16) return U2GNN_EALIGN;
    if (B == 0) return U2GNN_OK;
    Params p;
    p.qkv = qkv; p.dctx = nullptr; p.out = ctx; p.io_bf16 = io_bf16; p.B = B; p.S = S;
    p.rng = make_rng(seed, rng_stream, thr);
    p.low = rng_thr_low(thr);
    launch<false>(p, as_stream(stream));
    U2GNN_CHECK_LAUNCH();
}

// in_proj + attention core in one kernel (bf16 mode): qkv = x W_in^T + b_in is computed per tile on the tensor cores, written
// once (bf16, for the backward) and consumed from shared memory; ctx as u2gnn_seqattn_tc_fwd_ex(io_bf16 = 1).
extern "C" int u2gnn_inproj_seqattn_tc_fwd(const float* x, const int64_t* x_idx, int64_t x_rows, int64_t B, int S, int d,
                                           const float* w_in, const float* b_in, uint64_t seed, uint32_t rng_stream, int thr,
                                           void* qkv_out, void* ctx, int* err, u2gnn_stream_t stream) {
    if (!x || !w_in || !b_in || !qkv_out || !ctx || B < 0 || thr < 0 || thr > 255 || (x_idx && x_rows < 1)) return U2GNN_EINVAL;
    if (d != D || S < 2 || S > 32) return U2GNN_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(qkv_out) | reinterpret_cast<uintptr_t>(ctx) |
         reinterpret_cast<uintptr_t>(b_in)) % 16)
        return U2GNN_EALIGN;
    if (B == 0) return U2GNN_OK;
    Params p;
    p.qkv = nullptr; p.dctx = nullptr; p.out = ctx; p.io_bf16 = 1; p.B = B; p.S = S;
    p.rng = make_rng(seed, rng_stream, thr);
    p.low = rng_thr_low(thr);
    p.x = x; p.w_in = w_in; p.b_in = b_in; p.qkv_out = qkv_out;
    p.x_idx = x_idx; p.x_rows = x_rows; p.err = err;
    launch<false, true>(p, as_stream(stream));
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_seqattn_tc_fwd(const float* qkv, int64_t B, int S, int d, uint64_t seed, uint32_t rng_stream, int thr,
                                    float* ctx, u2gnn_stream_t stream) {
    return u2gnn_seqattn_tc_fwd_ex(qkv, B, S, d, seed, rng_stream, thr, ctx, 0, stream);
}

extern "C" int u2gnn_seqattn_tc_bwd_ex(const void* qkv, const void* dctx, int64_t B, int S, int d, uint64_t seed,
                                       uint32_t rng_stream, int thr, void* dqkv, int io_bf16, u2gnn_stream_t stream) {
    if (!qkv || !dctx || !dqkv || B < 0 || thr < 0 || thr > 255) return U2GNN_EINVAL;
    if (d != D || S < 2 || S > 32) return U2GNN_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(dctx) | reinterpret_cast<uintptr_t>(dqkv)) % 16) return U2GNN_EALIGN;
    if (B == 0) return U2GNN_OK;
    Params p;
    p.qkv = qkv; p.dctx = dctx; p.out = dqkv; p.io_bf16 = io_bf16; p.B = B; p.S = S;
    p.rng = make_rng(seed, rng_stream, thr);
    p.low = rng_thr_low(thr);
    launch<true>(p, as_stream(stream));
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_seqattn_tc_bwd(const float* qkv, const float* dctx, int64_t B, int S, int d, uint64_t seed,
                                    uint32_t rng_stream, int thr, float* dqkv, u2gnn_stream_t stream) {
    return u2gnn_seqattn_tc_bwd_ex(qkv, dctx, B, S, d, seed, rng_stream, thr, dqkv, 0, stream);
}
