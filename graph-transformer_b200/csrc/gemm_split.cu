// fp32 GEMMs of the exact-precision (1e-4) mode on the tensor cores: three-product bf16 split.
//
// Every fp32 operand value a is staged as two bf16 values, hi = bf16(a) and lo = bf16(a - hi) (|a - hi - lo| <= 2^-17 |a|),
// and every product is evaluated as  A_hi B_hi + A_lo B_hi + A_hi B_lo  with fp32 accumulation in tensor memory - three
// tcgen05.mma per k-step instead of one; the dropped A_lo B_lo term is <= 2^-16 of the product.  Operands stay fp32 in HBM
// (the split happens while a tile is staged into shared memory), so these kernels are drop-ins for u2gnn_sgemm on the
// projections and the FFN of precision="fp32":
//   u2gnn_gemm_split_rows    C[M, N]   = epi(A[M, K] op(W) + bias) (+ beta C)     W = [N][K] (w_kn 0) or [K][N] (w_kn 1)
//   u2gnn_gemm_split_wgrad   dW[N1,N2] += A[M, N1]^T B[M, N2],  db[N1] += colsum(A)
// i.e. F.linear and its autograd inside nn.TransformerEncoderLayer (torch/nn/modules/transformer.py:944-982, linear1 /
// linear2, and multi_head_attention_forward's in_proj / out_proj) at fp32 accuracy.  Any M, K, N and leading dimensions:
// 16-byte aligned shapes take 128-bit loads, others a scalar path.
//
// Structure: phase-serial CTAs of 256 threads, two per SM (one converts / stores while the other's MMAs run).  A step of
// the rows kernel is one 64-wide K chunk of a 128-row tile for one slice of <= 128 output columns; the fp32 words of the NEXT
// step are already in registers while the current step's MMAs run, so HBM latency is covered by the tensor pipe.
#include "common.cuh"
#include "tc_common.cuh"
#include "ffn_epi.cuh"
#include "rng.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int TM = 128;     // rows per tile of the rows kernel
constexpr int KC = 64;      // K chunk = one 128-byte swizzle atom of bf16
constexpr int NSMAX = 128;  // output columns per slice (tensor-memory columns of a CTA)

// two fp32 -> packed bf16 hi pair and packed bf16 lo pair (lo = bf16(a - float(hi)))
__device__ __forceinline__ void split2(float a0, float a1, uint32_t& hi, uint32_t& lo) {
    hi = epi::cvt2(a0, a1);
    lo = epi::cvt2(a0 - __uint_as_float(hi << 16), a1 - __uint_as_float(hi & 0xFFFF0000u));
}

// A [rows x cols] block (cols = 1 << cl2, a multiple of 64) of a row-major fp32 matrix, as 4 * NV words per thread.
// Element group g = tid + 256 u (u < NV) of n_groups: vec -> the float4 at (g / (cols/4), 4 (g % (cols/4)));
// scalar -> the four single elements 4 g' ... of the flattened block, g' = the same index: (e / cols, e % cols), e = 1024 u + tid + 256 j.
// Elements at rows >= rv or columns >= cv read as zero.
template <int NV>
__device__ __forceinline__ void blk_load(float (&v)[4 * NV], const float* __restrict__ src, int64_t ld, int rv, int cv, int cl2,
                                         bool vec, int n_groups, int tid) {
    if (vec) {
#pragma unroll
        for (int u = 0; u < NV; ++u) {
            const int g = u * kThreads + tid;
            const int r = g >> (cl2 - 2), c = (g & ((1 << (cl2 - 2)) - 1)) << 2;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g < n_groups && r < rv && c < cv) x = __ldg(reinterpret_cast<const float4*>(src + (int64_t)r * ld + c));
            v[4 * u] = x.x; v[4 * u + 1] = x.y; v[4 * u + 2] = x.z; v[4 * u + 3] = x.w;
        }
    } else {
#pragma unroll
        for (int u = 0; u < 4 * NV; ++u) {
            const int e = u * kThreads + tid;
            const int r = e >> cl2, c = e & ((1 << cl2) - 1);
            float x = 0.f;
            if (e < 4 * n_groups && r < rv && c < cv) x = __ldg(src + (int64_t)r * ld + c);
            v[u] = x;
        }
    }
}

// the same words -> swizzled bf16 hi / lo images: 64-column blocks of `blk_bytes` (= rows * 128) each
template <int NV>
__device__ __forceinline__ void blk_store(uint8_t* hi, uint8_t* lo, const float (&v)[4 * NV], int cl2, uint32_t blk_bytes, bool vec,
                                          int n_groups, int tid) {
    if (vec) {
#pragma unroll
        for (int u = 0; u < NV; ++u) {
            const int g = u * kThreads + tid;
            if (g >= n_groups) continue;
            const int r = g >> (cl2 - 2), c = (g & ((1 << (cl2 - 2)) - 1)) << 2;
            const uint32_t off = (uint32_t)(c >> 6) * blk_bytes + tc::sw128_offset(r, c & 63);
            uint2 h, l;
            split2(v[4 * u], v[4 * u + 1], h.x, l.x);
            split2(v[4 * u + 2], v[4 * u + 3], h.y, l.y);
            *reinterpret_cast<uint2*>(hi + off) = h;
            *reinterpret_cast<uint2*>(lo + off) = l;
        }
    } else {
#pragma unroll
        for (int u = 0; u < 4 * NV; ++u) {
            const int e = u * kThreads + tid;
            if (e >= 4 * n_groups) continue;
            const int r = e >> cl2, c = e & ((1 << cl2) - 1);
            const uint32_t off = (uint32_t)(c >> 6) * blk_bytes + tc::sw128_offset(r, c & 63);
            const __nv_bfloat16 h = __float2bfloat16(v[u]);
            *reinterpret_cast<__nv_bfloat16*>(hi + off) = h;
            *reinterpret_cast<__nv_bfloat16*>(lo + off) = __float2bfloat16(v[u] - __bfloat162float(h));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// rows GEMM
// ---------------------------------------------------------------------------------------------------------------
struct RowsP {
    const void* A;        // fp32 (split mode) or bf16 (plain mode) row-major
    int a_bf16, c_bf16, aux_bf16;   // plain mode (split = 0): bf16 operands / results stored by tensor-core producers and consumers
    int split;            // 1: three-product hi / lo split of fp32 operands; 0: one bf16 product per k-step
    int64_t M, lda;
    int K;
    const float* W;
    const uint8_t* Wp;    // optional: the weights ALREADY as swizzled bf16 images (gemm_split_pack_kernel): a step's weights are one bulk copy
    int w_kn;
    int64_t ldw;
    int N;
    const float* bias;
    int flags;            // 1 bias, 2 ReLU, 4 dropout, 8 aux mask (the u2gnn_sgemm epilogue bits)
    RngKeys keys;
    int thr, low;
    float drop_scale;
    int64_t rng_row0;
    const void* aux;
    int64_t ldaux;
    float aux_scale;
    float beta;
    void* C;
    int64_t ldc;
    int a_vec, w_vec, c_vec;
};

// bf16 rows -> one swizzled [128 x 64] tile: a plain copy of 16-byte chunks (4 per thread), zero past rv / cv (cv a multiple of 8).
// The chunk bits travel in the first 16 words of the fp32 prefetch registers (one register file for both modes).
__device__ __forceinline__ void tile_load_bf16(float (&v)[32], const __nv_bfloat16* __restrict__ src, int64_t ld, int rv, int cv, int tid) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int g = u * kThreads + tid;
        const int r = g >> 3, c8 = g & 7;
        uint4 x = make_uint4(0u, 0u, 0u, 0u);
        if (r < rv && 8 * c8 < cv) x = __ldg(reinterpret_cast<const uint4*>(src + (int64_t)r * ld) + c8);
        v[4 * u] = __uint_as_float(x.x); v[4 * u + 1] = __uint_as_float(x.y); v[4 * u + 2] = __uint_as_float(x.z); v[4 * u + 3] = __uint_as_float(x.w);
    }
}
__device__ __forceinline__ void tile_store_bf16(uint8_t* tile, const float (&v)[32], int tid) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int g = u * kThreads + tid;
        *reinterpret_cast<uint4*>(tile + tc::sw128_chunk(g >> 3, g & 7)) =
            make_uint4(__float_as_uint(v[4 * u]), __float_as_uint(v[4 * u + 1]), __float_as_uint(v[4 * u + 2]), __float_as_uint(v[4 * u + 3]));
    }
}

// shared memory: A hi | A lo (16 KB each) | two weight buffers of W hi | W lo (16 KB each: <= 128 output columns x 64).  The
// fp32 staging of the epilogue (32 KB) lies over the weight buffer the step's own MMAs have just finished with; the other
// buffer already holds the next step's weights (staged while those MMAs ran).
constexpr uint32_t kRowsSmem = 2 * 16384 + 2 * 32768;

template <int EPI, bool PLAIN>   // EPI 0: generic epilogue (runtime flags, any alignment); 1 / 2 / 3: vectorised specialisations (see the copy-out)
                                 // PLAIN: bf16 A rows, one product per k-step, optional bf16 result / aux (compiled out of the fp32 split kernels)
__global__ void __launch_bounds__(kThreads, 2) gemm_split_rows_kernel(const RowsP p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((tc::smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* sAh = smem;
    uint8_t* sAl = smem + 16384;
    uint8_t* sW = smem + 32768;                              // buffer b: hi at sW + b * 32768, lo 16 KB behind it
    __shared__ uint64_t wbar[2];                             // packed weights: bytes of buffer b have landed
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::mbar_init(&wbar[0], 1);
        tc::mbar_init(&wbar[1], 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc<NSMAX>(&tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int KB = (p.K + KC - 1) / KC;
    const int NSL = (p.N + NSMAX - 1) / NSMAX;
    const int64_t n_tiles = (p.M + TM - 1) / TM;
    const bool w_resident = (KB == 1 && NSL == 1);
    const int wq = warp & 3, half = warp >> 2;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;

    // geometry of a step
    auto a_src = [&](int64_t tile, int kb) { return static_cast<const float*>(p.A) + tile * TM * p.lda + (int64_t)kb * KC; };
    auto a_src16 = [&](int64_t tile, int kb) { return static_cast<const __nv_bfloat16*>(p.A) + tile * TM * p.lda + (int64_t)kb * KC; };
    auto a_rv = [&](int64_t tile) { const int64_t r = p.M - tile * TM; return (int)(r < TM ? r : TM); };
    auto k_cv = [&](int kb) { const int c = p.K - kb * KC; return c < KC ? c : KC; };
    auto n_w = [&](int ns) { const int c = p.N - ns * NSMAX; return c < NSMAX ? c : NSMAX; };
    // W block of (ns, kb): K-major image [nps rows x 64] (w_kn 0) or MN-major image [64 rows x npw cols] (w_kn 1)
    auto w_groups = [&](int ns) {
        const int nw = n_w(ns);
        return p.w_kn ? 16 * ((nw + 63) / 64 * 64) : 16 * ((nw + 15) / 16 * 16);
    };
    float a[32];
    auto load_a = [&](int64_t tile, int kb) {
        if constexpr (PLAIN) tile_load_bf16(a, a_src16(tile, kb), p.lda, a_rv(tile), k_cv(kb), tid);
        else blk_load<8>(a, a_src(tile, kb), p.lda, a_rv(tile), k_cv(kb), 6, p.a_vec, 2048, tid);
    };
    auto store_a = [&]() {
        if constexpr (PLAIN) tile_store_bf16(sAh, a, tid);
        else blk_store<8>(sAh, sAl, a, 6, 0u, p.a_vec, 2048, tid);
    };
    // weights of (ns, kb) -> buffer wb, through short-lived registers (an L2 read: the weights are a few hundred KB)
    const uint32_t w_bytes = PLAIN ? 16384u : 32768u;        // packed image of one (slice, chunk): hi (+ lo)
    auto stage_w = [&](int ns, int kb, int wb) {
        if (p.Wp) {                                          // one asynchronous bulk copy, no registers, no conversion
            if (tid == 0) {
                tc::mbar_arrive_expect_tx(&wbar[wb], w_bytes);
                tc::bulk_g2s(sW + wb * 32768, p.Wp + ((size_t)ns * KB + kb) * 32768, w_bytes, &wbar[wb]);
            }
            return;
        }
        float w[32];
        const int nw = n_w(ns);
        uint8_t* hi = sW + wb * 32768;
        if (!p.w_kn) {
            blk_load<8>(w, p.W + (int64_t)ns * NSMAX * p.ldw + (int64_t)kb * KC, p.ldw, nw, k_cv(kb), 6, p.w_vec, w_groups(ns), tid);
            blk_store<8>(hi, hi + 16384, w, 6, 0u, p.w_vec, w_groups(ns), tid);
        } else {
            const int cl2 = (nw > 64) ? 7 : 6;
            blk_load<8>(w, p.W + (int64_t)kb * KC * p.ldw + (int64_t)ns * NSMAX, p.ldw, k_cv(kb), nw, cl2, p.w_vec, w_groups(ns), tid);
            blk_store<8>(hi, hi + 16384, w, cl2, 8192u, p.w_vec, w_groups(ns), tid);
        }
    };

    int64_t tile = blockIdx.x;
    int ns = 0, kb = 0;
    bool valid = tile < n_tiles, pending = false;
    uint32_t phase = 0;
    int wb = 0;                                              // weight buffer of the current step
    uint32_t wphase[2] = {0u, 0u};
    bool w_waited = false;
    if (valid) {
        load_a(tile, 0);
        stage_w(0, 0, 0);
    }
    const int rr0 = tid >> 4, c4 = tid & 15;                 // copy-out: 16 lanes x float4 per row, rows rr0 + 16 u
    while (valid) {
        const bool a_new = !(KB == 1 && ns > 0);             // K <= 64: the A tile stays staged across the output slices
        if (pending) {                                       // the previous step's MMAs have read the A tile (and the other weight buffer)
            tc::mbar_wait(&bar, phase);
            phase ^= 1;
            pending = false;
        }
        if (a_new) store_a();
        // next step
        int64_t ntile = tile;
        int nns = ns, nkb = kb + 1;
        if (nkb == KB) {
            nkb = 0;
            if (++nns == NSL) {
                nns = 0;
                ntile += gridDim.x;
            }
        }
        const bool nvalid = ntile < n_tiles;
        const bool last_k = (kb == KB - 1);
        const bool next_a = nvalid && !(KB == 1 && nns > 0);
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        const int nw = n_w(ns);
        const int nps = (nw + 15) / 16 * 16;
        if (p.Wp && (!w_resident || !w_waited)) {            // (uniform) the step's packed weights have landed
            if (warp == 0) tc::mbar_wait(&wbar[wb], wphase[wb]);
            wphase[wb] ^= 1;
            w_waited = true;
        }
        if (warp == 0) {
            if (tc::elect_one()) {
                const uint32_t idesc = tc::make_idesc(TM, nps, 0, p.w_kn);
                const uint64_t ah = tc::make_desc_sw128(tc::smem_u32(sAh), 16, 1024), al = tc::make_desc_sw128(tc::smem_u32(sAl), 16, 1024);
                const uint32_t w0 = tc::smem_u32(sW) + (uint32_t)wb * 32768u;
                const uint64_t wh = p.w_kn ? tc::make_desc_sw128(w0, 8192, 1024) : tc::make_desc_sw128(w0, 16, 1024);
                const uint64_t wl = p.w_kn ? tc::make_desc_sw128(w0 + 16384, 8192, 1024) : tc::make_desc_sw128(w0 + 16384, 16, 1024);
                const uint32_t wstep = p.w_kn ? 128u : 2u;   // 16 K rows of an MN-major image = 2 048 B; 16 K columns of a K-major one = 32 B
                if constexpr (!PLAIN) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        tc::mma_ss(tmem, ah + 2 * ks, wh + wstep * ks, idesc, (kb > 0 || ks > 0));
                        tc::mma_ss_acc(tmem, al + 2 * ks, wh + wstep * ks, idesc);
                        tc::mma_ss_acc(tmem, ah + 2 * ks, wl + wstep * ks, idesc);
                    }
                } else {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) tc::mma_ss(tmem, ah + 2 * ks, wh + wstep * ks, idesc, (kb > 0 || ks > 0));
                }
                tc::mma_commit(&bar);
            }
            __syncwarp();
        }
        pending = true;
        // under this step's MMAs: the next A chunk starts its way from HBM into registers (unless an epilogue follows: its
        // registers are needed there) and the next step's weights are staged into the other buffer
        if (next_a && !last_k) load_a(ntile, nkb);
        if (nvalid && !w_resident) stage_w(nns, nkb, wb ^ 1);
        uint8_t* sOut = sW + (w_resident ? 1 : wb) * 32768;   // resident weights live in buffer 0 for the whole CTA
        if (last_k) {
            tc::mbar_wait(&bar, phase);
            phase ^= 1;
            pending = false;
            tc::tc_fence_after();
            // ---- epilogue of (tile, ns): tensor memory -> registers (thread = row) -> XOR-swizzled fp32 staging -> coalesced
            // 128-bit rows with bias / ReLU / dropout / aux mask / beta, one 64-column piece at a time.  Everything the copy-out
            // needs from global memory (bias, the aux mask or the old C: 8 rows per thread) is requested BEFORE the block-wide
            // barrier: a load inside the guarded row loop exposed its full latency per row (ncu: 63 % long-scoreboard stalls
            // in the dPre product, profiles/r02_ncu_split_v1_stalls.txt).
            const int64_t row0 = tile * TM;
            const int n0 = ns * NSMAX;
            for (int pc = 0; 64 * pc < nps; ++pc) {
                if (64 * pc + 32 * half < nps) {
                    uint32_t v[32];
                    tc::tmem_ld32(tmem + lane_base + 64 * pc + 32 * half, v);
                    tc::tmem_ld_wait();
                    const int r = wq * 32 + lane;
                    uint8_t* st = sOut + r * 256;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<uint4*>(st + (((8 * half + j) ^ (r & 15)) << 4)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                const int col = n0 + 64 * pc + 4 * c4;       // this thread's four columns in every row of the piece
                if constexpr (EPI != 0) {
                    // vectorised copy-out, specialised at compile time (the generic version below spent most of its ISSUE slots on
                    // flag tests, 64-bit index arithmetic and one dropout word per thread and row: 2 700 instructions per thread and
                    // step, 57 % issue-active - profiles/r02_ncu_split_v2_linear1.txt):
                    //   EPI 1  bias + ReLU + dropout (linear1)      EPI 2  aux mask (dPre)      EPI 3  bias / beta (everything else)
                    const bool col_ok = col < p.N;           // N % 4 == 0: whole float4 pieces
                    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (EPI != 2 && (p.flags & 1) && col_ok) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                    const bool use_old = (EPI == 3) && (p.beta != 0.0f);
                    float4 pre[(EPI == 1 || (PLAIN && EPI == 2)) ? 1 : 8];   // aux mask rows (EPI 2) or old C rows (EPI 3, beta)
                    uint2 prew[(PLAIN && EPI == 2) ? 8 : 1];   // plain mode: the saved hidden is bf16 (8 bytes per thread and row), kept as raw bits
                    if constexpr (PLAIN && EPI == 2) {
                        const __nv_bfloat16* src = static_cast<const __nv_bfloat16*>(p.aux) + (row0 + rr0) * p.ldaux + col;
                        const int64_t step = 16 * p.ldaux;
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            prew[u] = make_uint2(0u, 0u);
                            if (col_ok && row0 + rr0 + 16 * u < p.M) prew[u] = *reinterpret_cast<const uint2*>(src + u * step);
                        }
                    } else if (EPI == 2 || use_old) {
                        const float* src = (EPI == 2) ? static_cast<const float*>(p.aux) + (row0 + rr0) * p.ldaux + col
                                                      : static_cast<const float*>(p.C) + (row0 + rr0) * p.ldc + col;
                        const int64_t step = 16 * ((EPI == 2) ? p.ldaux : p.ldc);
#pragma unroll
                        for (int u = 0; u < ((EPI == 1 || (PLAIN && EPI == 2)) ? 1 : 8); ++u) {
                            pre[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (col_ok && row0 + rr0 + 16 * u < p.M) pre[u] = *reinterpret_cast<const float4*>(src + u * step);
                        }
                    }
                    // EPI 1 (N % 32 == 0): the piece holds 128 rows x two 32-element groups of the dropout stream = 32 keep words
                    // per warp (its 16 rows x 2 groups): lane L evaluates the word of (iteration L >> 2, row L >> 1 & 1, group L & 1)
                    // and the row loop fetches words with one shuffle
                    uint32_t my_word = 0xFFFFFFFFu;
                    if (EPI == 1 && p.thr) {
                        const int64_t wrow = p.rng_row0 + row0 + 2 * warp + ((lane >> 1) & 1) + 16 * (lane >> 2);
                        my_word = rng_keep_word_lo(p.keys, (uint64_t)wrow * (uint64_t)(p.N >> 5) + (uint64_t)((n0 + 64 * pc) >> 5) + (uint64_t)(lane & 1),
                                                   p.thr, p.low);
                    }
                    __syncthreads();
                    float* out = static_cast<float*>(p.C) + (row0 + rr0) * p.ldc + col;
                    __nv_bfloat16* out16 = static_cast<__nv_bfloat16*>(p.C) + (row0 + rr0) * p.ldc + col;
                    const int64_t ostep = 16 * p.ldc;
                    const int src_lane0 = 2 * (lane >> 4) + ((lane >> 3) & 1);
                    const float keep_scale = p.thr ? p.drop_scale : 1.0f;
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int rr = rr0 + 16 * u;
                        const bool live = (row0 + rr < p.M) && col_ok;
                        float4 x = *reinterpret_cast<const float4*>(sOut + rr * 256 + ((c4 ^ (rr & 15)) << 4));
                        if (EPI != 2) { x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w; }
                        if (EPI == 1) {
                            const uint32_t kw = __shfl_sync(0xffffffffu, my_word, 4 * u + src_lane0) >> (4 * (c4 & 7));
                            x.x = (kw & 1u) ? fmaxf(x.x, 0.0f) * keep_scale : 0.0f;
                            x.y = (kw & 2u) ? fmaxf(x.y, 0.0f) * keep_scale : 0.0f;
                            x.z = (kw & 4u) ? fmaxf(x.z, 0.0f) * keep_scale : 0.0f;
                            x.w = (kw & 8u) ? fmaxf(x.w, 0.0f) * keep_scale : 0.0f;
                        }
                        if constexpr (PLAIN && EPI == 2) {
                            // the hidden is >= 0 (after ReLU): live and kept <=> its bf16 bits are not (plus or minus) zero
                            const uint2 m = prew[u];
                            x.x = (m.x & 0x00007FFFu) ? x.x * p.aux_scale : 0.0f;
                            x.y = (m.x & 0x7FFF0000u) ? x.y * p.aux_scale : 0.0f;
                            x.z = (m.y & 0x00007FFFu) ? x.z * p.aux_scale : 0.0f;
                            x.w = (m.y & 0x7FFF0000u) ? x.w * p.aux_scale : 0.0f;
                        } else if (EPI == 2) {
                            x.x = pre[u].x > 0.0f ? x.x * p.aux_scale : 0.0f;
                            x.y = pre[u].y > 0.0f ? x.y * p.aux_scale : 0.0f;
                            x.z = pre[u].z > 0.0f ? x.z * p.aux_scale : 0.0f;
                            x.w = pre[u].w > 0.0f ? x.w * p.aux_scale : 0.0f;
                        }
                        if (use_old) {
                            constexpr bool one = (EPI == 1 || (PLAIN && EPI == 2));
                            x.x = fmaf(p.beta, pre[one ? 0 : u].x, x.x); x.y = fmaf(p.beta, pre[one ? 0 : u].y, x.y);
                            x.z = fmaf(p.beta, pre[one ? 0 : u].z, x.z); x.w = fmaf(p.beta, pre[one ? 0 : u].w, x.w);
                        }
                        if (live) {
                            if (PLAIN && p.c_bf16) *reinterpret_cast<uint2*>(out16 + u * ostep) = make_uint2(epi::cvt2(x.x, x.y), epi::cvt2(x.z, x.w));
                            else *reinterpret_cast<float4*>(out + u * ostep) = x;
                        }
                    }
                } else if (p.c_vec) {
                    const bool col_ok = col < p.N;           // N % 4 == 0: whole float4 pieces
                    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if ((p.flags & 1) && col_ok) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                    const bool use_aux = (p.flags & 8) != 0, use_old = (p.beta != 0.0f);
                    float4 pre[8];                           // aux mask rows, or the old C rows when there is no aux
                    if (use_aux || use_old) {
                        const float* src = use_aux ? static_cast<const float*>(p.aux) + (row0 + rr0) * p.ldaux + col
                                                   : static_cast<const float*>(p.C) + (row0 + rr0) * p.ldc + col;
                        const int64_t step = 16 * (use_aux ? p.ldaux : p.ldc);
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            pre[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (col_ok && row0 + rr0 + 16 * u < p.M) pre[u] = *reinterpret_cast<const float4*>(src + u * step);
                        }
                    }
                    __syncthreads();
                    uint64_t el = (uint64_t)(p.rng_row0 + row0 + rr0) * (uint64_t)p.N + (uint64_t)col;
                    float* out = static_cast<float*>(p.C) + (row0 + rr0) * p.ldc + col;
#pragma unroll
                    for (int u = 0; u < 8; ++u, el += 16ull * (uint64_t)p.N, out += 16 * p.ldc) {
                        const int rr = rr0 + 16 * u;
                        const bool live = (row0 + rr < p.M) && col_ok;
                        if (!live) continue;
                        float4 x = *reinterpret_cast<const float4*>(sOut + rr * 256 + ((c4 ^ (rr & 15)) << 4));
                        x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
                        if (p.flags & 2) {
                            x.x = fmaxf(x.x, 0.0f); x.y = fmaxf(x.y, 0.0f); x.z = fmaxf(x.z, 0.0f); x.w = fmaxf(x.w, 0.0f);
                        }
                        if (p.flags & 4) {
                            // el % 4 == 0: the four columns lie in one 32-element group of the stream
                            const uint32_t kw = rng_keep_word_lo(p.keys, el >> 5, p.thr, p.low) >> (el & 31);
                            x.x = (kw & 1u) ? x.x * p.drop_scale : 0.0f;
                            x.y = (kw & 2u) ? x.y * p.drop_scale : 0.0f;
                            x.z = (kw & 4u) ? x.z * p.drop_scale : 0.0f;
                            x.w = (kw & 8u) ? x.w * p.drop_scale : 0.0f;
                        }
                        if (use_aux) {
                            x.x = pre[u].x > 0.0f ? x.x * p.aux_scale : 0.0f;
                            x.y = pre[u].y > 0.0f ? x.y * p.aux_scale : 0.0f;
                            x.z = pre[u].z > 0.0f ? x.z * p.aux_scale : 0.0f;
                            x.w = pre[u].w > 0.0f ? x.w * p.aux_scale : 0.0f;
                        }
                        if (use_old) {
                            const float4 c = (use_aux && live) ? *reinterpret_cast<const float4*>(out) : pre[u];
                            x.x = fmaf(p.beta, c.x, x.x); x.y = fmaf(p.beta, c.y, x.y);
                            x.z = fmaf(p.beta, c.z, x.z); x.w = fmaf(p.beta, c.w, x.w);
                        }
                        if (live) *reinterpret_cast<float4*>(out) = x;
                    }
                } else {
                    __syncthreads();
                    // unaligned shapes (parity-sized feature dimensions): element by element
                    for (int u = 0; u < 8; ++u) {
                        const int rr = rr0 + 16 * u;
                        const int64_t row = row0 + rr;
                        if (row >= p.M) continue;
                        const float* sp = reinterpret_cast<const float*>(sOut + rr * 256 + ((c4 ^ (rr & 15)) << 4));
                        for (int j = 0; j < 4 && col + j < p.N; ++j) {
                            float y = sp[j];
                            float* o = static_cast<float*>(p.C) + row * p.ldc + col + j;
                            if (p.flags & 1) y += p.bias[col + j];
                            if (p.flags & 2) y = fmaxf(y, 0.0f);
                            if (p.flags & 4)
                                y *= rng_dropout_mult(p.keys, (uint64_t)(p.rng_row0 + row) * (uint64_t)p.N + (uint64_t)(col + j), p.thr, p.drop_scale);
                            if (p.flags & 8) y = (static_cast<const float*>(p.aux)[row * p.ldaux + col + j] > 0.0f) ? y * p.aux_scale : 0.0f;
                            if (p.beta != 0.0f) y = fmaf(p.beta, *o, y);
                            *o = y;
                        }
                    }
                }
                __syncthreads();                              // the staging is reused by the next piece
            }
            tc::tc_fence_before();
            if (next_a) load_a(ntile, nkb);                  // the epilogue's registers are dead now
        }
        tile = ntile; ns = nns; kb = nkb; valid = nvalid;
        if (!w_resident) wb ^= 1;
    }
    if (pending) tc::mbar_wait(&bar, phase);
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<NSMAX>(tmem);
}

// ---------------------------------------------------------------------------------------------------------------
// weight-gradient GEMM: dW[n1, n2] += sum_m A[m, n1] B[m, n2]; db[n1] += sum_m A[m, n1].  N1 <= 128, N2 <= 64 per launch.
// A step is 64 rows: both operands are MN-major images (rows = the contraction index), the accumulator is
// [128 A columns (lanes) x (64 B columns + a ones column for db)].
// ---------------------------------------------------------------------------------------------------------------
struct WgP {
    const float* A;
    const float* B;
    int64_t M, lda, ldb;
    int N1, N2;
    float* dW;
    int64_t s1, s2;       // dW element (n1, n2) lives at dW[n1 s1 + n2 s2] (a transposed destination is a stride swap)
    float* db;
    int a_vec, b_vec;
};

// shared memory: A hi (2 blocks x 8 KB) | A lo | B hi (8 KB) | ones (8 KB) | B lo (8 KB)
constexpr uint32_t kWgSmem = 16384 * 2 + 8192 * 3;
constexpr int WR = 64;      // rows per step

__global__ void __launch_bounds__(kThreads, 2) gemm_split_wgrad_kernel(const WgP p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((tc::smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* sAh = smem;
    uint8_t* sAl = smem + 16384;
    uint8_t* sBh = smem + 32768;
    uint8_t* sOnes = smem + 40960;
    uint8_t* sBl = smem + 49152;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc<128>(&tmem_slot);
    for (int e = tid; e < 8192 / 4; e += kThreads) reinterpret_cast<uint32_t*>(sOnes)[e] = 0x3F803F80u;
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int64_t n_steps = (p.M + WR - 1) / WR;
    float a[32], b[16];
    auto load = [&](int64_t s) {
        const int64_t r0 = s * WR;
        const int rv = (int)((p.M - r0 < WR) ? p.M - r0 : WR);
        blk_load<8>(a, p.A + r0 * p.lda, p.lda, rv, p.N1, 7, p.a_vec, 2048, tid);
        blk_load<4>(b, p.B + r0 * p.ldb, p.ldb, rv, p.N2, 6, p.b_vec, 1024, tid);
    };
    int64_t s = blockIdx.x;
    if (s < n_steps) load(s);
    bool pending = false;
    uint32_t phase = 0;
    int64_t it = 0;
    for (; s < n_steps; s += gridDim.x, ++it) {
        if (pending) {
            tc::mbar_wait(&bar, phase);
            phase ^= 1;
            pending = false;
        }
        blk_store<8>(sAh, sAl, a, 7, 8192u, p.a_vec, 2048, tid);
        blk_store<4>(sBh, sBl, b, 6, 0u, p.b_vec, 1024, tid);
        if (s + gridDim.x < n_steps) load(s + gridDim.x);
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        if (warp == 0) {
            if (tc::elect_one()) {
                const uint32_t idesc80 = tc::make_idesc(128, 80, 1, 1);     // [B hi | ones]: dW and db
                const uint32_t idesc64 = tc::make_idesc(128, 64, 1, 1);
                const uint32_t bh0 = tc::smem_u32(sBh);
                const uint64_t ah = tc::make_desc_sw128(tc::smem_u32(sAh), 8192, 1024), al = tc::make_desc_sw128(tc::smem_u32(sAl), 8192, 1024);
                const uint64_t bh = tc::make_desc_sw128(bh0, tc::smem_u32(sOnes) - bh0, 1024);
                const uint64_t bl = tc::make_desc_sw128(tc::smem_u32(sBl), 8192, 1024);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    tc::mma_ss(tmem, ah + 128 * ks, bh + 128 * ks, idesc80, (it > 0 || ks > 0));
                    tc::mma_ss_acc(tmem, al + 128 * ks, bh + 128 * ks, idesc80);
                    tc::mma_ss_acc(tmem, ah + 128 * ks, bl + 128 * ks, idesc64);
                }
                tc::mma_commit(&bar);
            }
            __syncwarp();
        }
        pending = true;
    }
    if (pending) {
        tc::mbar_wait(&bar, phase);
        tc::tc_fence_after();
        // flush: thread = A column (accumulator lane); warps 0-3 take B columns 0..31, warps 4-7 columns 32..63 and db
        const int wq = warp & 3, half = warp >> 2;
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
        const int n1 = wq * 32 + lane;
        uint32_t v[32];
        tc::tmem_ld32(tmem + lane_base + 32 * half, v);
        tc::tmem_ld_wait();
        if (n1 < p.N1) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (32 * half + j < p.N2) atomicAdd(p.dW + (int64_t)n1 * p.s1 + (int64_t)(32 * half + j) * p.s2, __uint_as_float(v[j]));
        }
        if (half == 1 && p.db) {
            uint32_t c16[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(c16[0]), "=r"(c16[1]), "=r"(c16[2]), "=r"(c16[3]), "=r"(c16[4]), "=r"(c16[5]), "=r"(c16[6]),
                           "=r"(c16[7]), "=r"(c16[8]), "=r"(c16[9]), "=r"(c16[10]), "=r"(c16[11]), "=r"(c16[12]), "=r"(c16[13]),
                           "=r"(c16[14]), "=r"(c16[15])
                         : "r"(tmem + lane_base + 64)
                         : "memory");
            tc::tmem_ld_wait();
            if (n1 < p.N1) atomicAdd(p.db + n1, __uint_as_float(c16[0]));
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<128>(tmem);
}

// Weights -> the swizzled bf16 images the rows kernel multiplies: block (ns, kb) = 32 KB, hi image in the first 16 KB, lo image (split
// mode) in the second, laid out exactly as the in-kernel staging writes them.  One CTA per block; a few microseconds per weight.
__global__ void __launch_bounds__(kThreads) gemm_split_pack_kernel(const float* __restrict__ W, int w_kn, int64_t ldw, int N, int K, int w_vec,
                                                                   uint8_t* __restrict__ out) {
    const int KB = (K + KC - 1) / KC;
    const int ns = blockIdx.x / KB, kb = blockIdx.x - ns * KB;
    const int tid = threadIdx.x;
    const int nw = (N - ns * NSMAX < NSMAX) ? N - ns * NSMAX : NSMAX;
    const int cv = (K - kb * KC < KC) ? K - kb * KC : KC;
    uint8_t* hi = out + (size_t)blockIdx.x * 32768;
    float w[32];
    // zero first: rows / columns the image does not cover must read as zero operands
    for (int e = tid; e < 32768 / 16; e += kThreads) reinterpret_cast<uint4*>(hi)[e] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    if (!w_kn) {
        const int groups = 16 * ((nw + 15) / 16 * 16);
        blk_load<8>(w, W + (int64_t)ns * NSMAX * ldw + (int64_t)kb * KC, ldw, nw, cv, 6, w_vec, groups, tid);
        blk_store<8>(hi, hi + 16384, w, 6, 0u, w_vec, groups, tid);
    } else {
        const int cl2 = (nw > 64) ? 7 : 6;
        const int groups = 16 * ((nw + 63) / 64 * 64);
        blk_load<8>(w, W + (int64_t)kb * KC * ldw + (int64_t)ns * NSMAX, ldw, cv, nw, cl2, w_vec, groups, tid);
        blk_store<8>(hi, hi + 16384, w, cl2, 8192u, w_vec, groups, tid);
    }
}

bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; }

}  // namespace

namespace {
// shared launcher of the split (fp32 operands, three products) and plain (bf16 operands, one product) rows GEMM
int launch_rows(const void* A, int a_bf16, int64_t M, int K, int64_t lda, const float* W, int w_kn, int64_t ldw, int N, const float* bias,
                int epi, uint64_t seed, uint32_t rng_stream, int thr, int64_t rng_row0, const void* aux, int aux_bf16, int64_t ldaux,
                float aux_scale, float beta, void* C, int c_bf16, int64_t ldc, int split, const void* packed_w, cudaStream_t stream) {
    if (!A || !W || !C || M < 0 || K < 1 || N < 1 || lda < K || ldc < N || ldw < (w_kn ? N : K)) return U2GNN_EINVAL;
    if ((epi & 1) && !bias) return U2GNN_EINVAL;
    if ((epi & 8) && (!aux || ldaux < N)) return U2GNN_EINVAL;
    if (epi & ~15) return U2GNN_EINVAL;
    if (thr < 0 || thr > 255) return U2GNN_EINVAL;
    if (split && (a_bf16 || c_bf16 || aux_bf16)) return U2GNN_EINVAL;      // the split exists to keep fp32 operands exact
    if (!split && !a_bf16) return U2GNN_EUNSUPPORTED;                       // plain mode streams bf16 rows (fp32 rows: u2gnn_gemm_tc_rows_ex)
    if (!split && (epi & 8) && !aux_bf16) return U2GNN_EUNSUPPORTED;        // ... and masks with a bf16 hidden
    if (c_bf16 && beta != 0.0f) return U2GNN_EINVAL;
    if (M == 0) return U2GNN_OK;
    if ((epi & 4) && thr == 0) epi &= ~4;
    RowsP p;
    p.A = A; p.a_bf16 = a_bf16; p.c_bf16 = c_bf16; p.aux_bf16 = aux_bf16; p.split = split;
    p.M = M; p.lda = lda; p.K = K; p.W = W; p.w_kn = w_kn; p.ldw = ldw; p.N = N;
    p.Wp = static_cast<const uint8_t*>(packed_w);
    if (p.Wp && !aligned16(p.Wp)) return U2GNN_EALIGN;
    p.bias = bias; p.flags = epi; p.keys = rng_keys(seed, rng_stream); p.thr = thr; p.low = rng_thr_low(thr);
    p.drop_scale = thr ? rng_keep_scale(thr) : 1.0f; p.rng_row0 = rng_row0;
    p.aux = aux; p.ldaux = ldaux; p.aux_scale = aux_scale; p.beta = beta; p.C = C; p.ldc = ldc;
    p.a_vec = aligned16(A) && (lda & (a_bf16 ? 7 : 3)) == 0 && (K & (a_bf16 ? 7 : 3)) == 0;
    p.w_vec = aligned16(W) && (ldw & 3) == 0 && ((w_kn ? N : K) & 3) == 0;
    p.c_vec = (reinterpret_cast<uintptr_t>(C) % (c_bf16 ? 8 : 16)) == 0 && (ldc & 3) == 0 && (N & 3) == 0 && (!(epi & 1) || aligned16(bias)) &&
              (!(epi & 8) || ((reinterpret_cast<uintptr_t>(aux) % (aux_bf16 ? 8 : 16)) == 0 && (ldaux & 3) == 0));
    if (a_bf16 && !p.a_vec) return U2GNN_EALIGN;
    const int64_t n_tiles = (M + TM - 1) / TM;
    const int64_t cap = (int64_t)U2GNN_NUM_SMS * 2;
    const int grid = (int)(n_tiles < cap ? n_tiles : cap);
    auto launch = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowsSmem);
        kern<<<grid, kThreads, kRowsSmem, stream>>>(p);
    };
    // epilogue specialisations for the aligned shapes of the throughput configurations; everything else is generic
    const int e = p.flags;
    const int kind = (p.c_vec && ((e & ~1) == (2 | 4) || (e & ~1) == 2) && beta == 0.0f && (N & 31) == 0) ? 1
                     : (p.c_vec && e == 8 && beta == 0.0f) ? 2 : (p.c_vec && (e & ~1) == 0) ? 3 : 0;
    if (kind == 0 && (c_bf16 || aux_bf16)) return U2GNN_EUNSUPPORTED;       // bf16 results / masks: the vectorised epilogues only
    if (split) {
        if (kind == 1) launch(gemm_split_rows_kernel<1, false>);
        else if (kind == 2) launch(gemm_split_rows_kernel<2, false>);
        else if (kind == 3) launch(gemm_split_rows_kernel<3, false>);
        else launch(gemm_split_rows_kernel<0, false>);
    } else {
        if (kind == 1) launch(gemm_split_rows_kernel<1, true>);
        else if (kind == 2) launch(gemm_split_rows_kernel<2, true>);
        else if (kind == 3) launch(gemm_split_rows_kernel<3, true>);
        else launch(gemm_split_rows_kernel<0, true>);
    }
    U2GNN_CHECK_LAUNCH();
}
}  // namespace

extern "C" int u2gnn_gemm_split_rows(const float* A, int64_t M, int K, int64_t lda, const float* W, int w_kn, int64_t ldw, int N,
                                     const float* bias, int epi, uint64_t seed, uint32_t rng_stream, int thr, int64_t rng_row0,
                                     const float* aux, int64_t ldaux, float aux_scale, float beta, float* C, int64_t ldc,
                                     const void* packed_w, u2gnn_stream_t stream) {
    return launch_rows(A, 0, M, K, lda, W, w_kn, ldw, N, bias, epi, seed, rng_stream, thr, rng_row0, aux, 0, ldaux, aux_scale, beta, C, 0, ldc, 1,
                       packed_w, as_stream(stream));
}

extern "C" size_t u2gnn_gemm_split_packed_bytes(int N, int K) {
    if (N < 1 || K < 1) return 0;
    return (size_t)((N + NSMAX - 1) / NSMAX) * (size_t)((K + KC - 1) / KC) * 32768;
}

extern "C" int u2gnn_gemm_split_pack(const float* W, int w_kn, int64_t ldw, int N, int K, void* packed, size_t packed_size,
                                     u2gnn_stream_t stream) {
    if (!W || !packed || N < 1 || K < 1 || ldw < (w_kn ? N : K)) return U2GNN_EINVAL;
    if (packed_size < u2gnn_gemm_split_packed_bytes(N, K)) return U2GNN_EWORKSPACE;
    if (!aligned16(packed)) return U2GNN_EALIGN;
    const int w_vec = aligned16(W) && (ldw & 3) == 0 && ((w_kn ? N : K) & 3) == 0;
    const int blocks = ((N + NSMAX - 1) / NSMAX) * ((K + KC - 1) / KC);
    gemm_split_pack_kernel<<<blocks, kThreads, 0, as_stream(stream)>>>(W, w_kn, ldw, N, K, w_vec, static_cast<uint8_t*>(packed));
    U2GNN_CHECK_LAUNCH();
}

// The same K-looping kernel with ONE bf16 product per k-step: A bf16 rows (any K that is a multiple of 8, any N), result fp32 or bf16,
// aux mask fp32 or bf16.  The FFN of 64 < d <= 128 (engine.ffn_wide_*) runs linear1 + ReLU + dropout, linear2, dH + mask and dy1 as one
// launch each (u2gnn_gemm_tc_rows_ex holds all of K in shared memory: K, N <= 256 per launch, separate elementwise passes).
extern "C" int u2gnn_gemm_tc_rows_kloop(const void* A, int64_t M, int K, int64_t lda, const float* W, int w_kn, int64_t ldw, int N,
                                        const float* bias, int epi, uint64_t seed, uint32_t rng_stream, int thr, int64_t rng_row0,
                                        const void* aux, int aux_bf16, int64_t ldaux, float aux_scale, float beta, void* C, int c_bf16,
                                        int64_t ldc, const void* packed_w, u2gnn_stream_t stream) {
    return launch_rows(A, 1, M, K, lda, W, w_kn, ldw, N, bias, epi, seed, rng_stream, thr, rng_row0, aux, aux_bf16, ldaux, aux_scale, beta, C,
                       c_bf16, ldc, 0, packed_w, as_stream(stream));
}

extern "C" int u2gnn_gemm_split_wgrad(const float* A, int64_t M, int N1, int64_t lda, const float* B, int N2, int64_t ldb, float* dW,
                                      int64_t ldw_n1, int64_t ldw_n2, float* db, u2gnn_stream_t stream) {
    if (!A || !B || !dW || M < 0 || N1 < 1 || N2 < 1 || lda < N1 || ldb < N2) return U2GNN_EINVAL;
    if (M == 0) return U2GNN_OK;
    cudaFuncSetAttribute(gemm_split_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgSmem);
    const int64_t n_steps = (M + WR - 1) / WR;
    const int64_t cap = (int64_t)U2GNN_NUM_SMS * 2;
    const int grid = (int)(n_steps < cap ? n_steps : cap);
    // slices of <= 128 A columns x <= 64 B columns, one launch each (the accumulator of a CTA is 128 lanes x 80 columns)
    for (int i = 0; i < N1; i += 128) {
        for (int j = 0; j < N2; j += 64) {
            WgP p;
            p.A = A + i; p.B = B + j; p.M = M; p.lda = lda; p.ldb = ldb;
            p.N1 = (N1 - i < 128) ? N1 - i : 128;
            p.N2 = (N2 - j < 64) ? N2 - j : 64;
            p.dW = dW + (int64_t)i * ldw_n1 + (int64_t)j * ldw_n2; p.s1 = ldw_n1; p.s2 = ldw_n2;
            p.db = (db && j == 0) ? db + i : nullptr;
            p.a_vec = aligned16(p.A) && (lda & 3) == 0 && (p.N1 & 3) == 0;
            p.b_vec = aligned16(p.B) && (ldb & 3) == 0 && (p.N2 & 3) == 0;
            gemm_split_wgrad_kernel<<<grid, kThreads, kWgSmem, as_stream(stream)>>>(p);
        }
    }
    U2GNN_CHECK_LAUNCH();
}
