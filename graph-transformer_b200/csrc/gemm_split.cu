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
    const float* A;
    int64_t M, lda;
    int K;
    const float* W;
    int w_kn;
    int64_t ldw;
    int N;
    const float* bias;
    int flags;            // 1 bias, 2 ReLU, 4 dropout, 8 aux mask (the u2gnn_sgemm epilogue bits)
    RngKeys keys;
    int thr, low;
    float drop_scale;
    int64_t rng_row0;
    const float* aux;
    int64_t ldaux;
    float aux_scale;
    float beta;
    float* C;
    int64_t ldc;
    int a_vec, w_vec, c_vec;
};

// shared memory: A hi | A lo (16 KB each) | W hi | W lo (16 KB each: <= 128 output columns x 64) | staging (32 KB)
constexpr uint32_t kRowsSmem = 4 * 16384 + 32768;

__global__ void __launch_bounds__(kThreads, 2) gemm_split_rows_kernel(const RowsP p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((tc::smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* sAh = smem;
    uint8_t* sAl = smem + 16384;
    uint8_t* sWh = smem + 32768;
    uint8_t* sWl = smem + 49152;
    uint8_t* sOut = smem + 65536;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
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
    const bool w_prefetch = (KB > 1);
    const int wq = warp & 3, half = warp >> 2;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;

    // geometry of a step
    auto a_src = [&](int64_t tile, int kb) { return p.A + tile * TM * p.lda + (int64_t)kb * KC; };
    auto a_rv = [&](int64_t tile) { const int64_t r = p.M - tile * TM; return (int)(r < TM ? r : TM); };
    auto k_cv = [&](int kb) { const int c = p.K - kb * KC; return c < KC ? c : KC; };
    auto n_w = [&](int ns) { const int c = p.N - ns * NSMAX; return c < NSMAX ? c : NSMAX; };
    // W block of (ns, kb): K-major image [nps rows x 64] (w_kn 0) or MN-major image [64 rows x npw cols] (w_kn 1)
    auto w_groups = [&](int ns) {
        const int nw = n_w(ns);
        return p.w_kn ? 16 * ((nw + 63) / 64 * 64) : 16 * ((nw + 15) / 16 * 16);
    };
    float a[32], w[32];
    auto load_a = [&](int64_t tile, int kb) { blk_load<8>(a, a_src(tile, kb), p.lda, a_rv(tile), k_cv(kb), 6, p.a_vec, 2048, tid); };
    auto load_w = [&](float (&dst)[32], int ns, int kb, int g0) {
        // groups [g0, g0 + 2048) of the block (one call covers <= 128 x 64 elements)
        const int nw = n_w(ns);
        if (!p.w_kn) {
            // rows = output columns; this call's groups start at row g0 / 16
            const int r0 = g0 >> 4;
            blk_load<8>(dst, p.W + ((int64_t)ns * NSMAX + r0) * p.ldw + (int64_t)kb * KC, p.ldw, nw - r0, k_cv(kb), 6, p.w_vec, w_groups(ns) - g0, tid);
        } else {
            const int cl2 = (nw > 64) ? 7 : 6;
            blk_load<8>(dst, p.W + (int64_t)kb * KC * p.ldw + (int64_t)ns * NSMAX, p.ldw, k_cv(kb), nw, cl2, p.w_vec, w_groups(ns), tid);
        }
    };
    auto store_w = [&](const float (&src)[32], int ns, int g0) {
        const int nw = n_w(ns);
        if (!p.w_kn) {
            const int r0 = g0 >> 4;
            blk_store<8>(sWh + r0 * 128, sWl + r0 * 128, src, 6, 0u, p.w_vec, w_groups(ns) - g0, tid);
        } else {
            blk_store<8>(sWh, sWl, src, (nw > 64) ? 7 : 6, 8192u, p.w_vec, w_groups(ns), tid);
        }
    };

    int64_t tile = blockIdx.x;
    int ns = 0, kb = 0;
    bool valid = tile < n_tiles, first = true, pending = false;
    uint32_t phase = 0;
    if (valid) {
        load_a(tile, 0);
        if (w_prefetch) load_w(w, 0, 0, 0);
    }
    while (valid) {
        const bool a_new = !(KB == 1 && ns > 0);             // K <= 64: the A tile stays staged across the output slices
        const bool w_new = !(w_resident && !first);          // one chunk, one slice: the weights stay staged for the whole CTA
        if (pending) {                                       // the previous step's MMAs have read the staged operands
            tc::mbar_wait(&bar, phase);
            phase ^= 1;
            pending = false;
        }
        if (a_new) blk_store<8>(sAh, sAl, a, 6, 0u, p.a_vec, 2048, tid);
        if (w_new) {
            if (w_prefetch) {
                store_w(w, ns, 0);
            } else {
                load_w(w, ns, kb, 0);
                store_w(w, ns, 0);
            }
        }
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
        auto prefetch = [&]() {
            if (!nvalid) return;
            if (!(KB == 1 && nns > 0)) load_a(ntile, nkb);
            if (w_prefetch) load_w(w, nns, nkb, 0);
        };
        if (!last_k) prefetch();                             // in flight under this step's MMAs
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        const int nw = n_w(ns);
        const int nps = (nw + 15) / 16 * 16;
        if (warp == 0) {
            if (tc::elect_one()) {
                const uint32_t idesc = tc::make_idesc(TM, nps, 0, p.w_kn);
                const uint64_t ah = tc::make_desc_sw128(tc::smem_u32(sAh), 16, 1024), al = tc::make_desc_sw128(tc::smem_u32(sAl), 16, 1024);
                const uint64_t wh = p.w_kn ? tc::make_desc_sw128(tc::smem_u32(sWh), 8192, 1024) : tc::make_desc_sw128(tc::smem_u32(sWh), 16, 1024);
                const uint64_t wl = p.w_kn ? tc::make_desc_sw128(tc::smem_u32(sWl), 8192, 1024) : tc::make_desc_sw128(tc::smem_u32(sWl), 16, 1024);
                const uint32_t wstep = p.w_kn ? 128u : 2u;   // 16 K rows of an MN-major image = 2 048 B; 16 K columns of a K-major one = 32 B
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    tc::mma_ss(tmem, ah + 2 * ks, wh + wstep * ks, idesc, (kb > 0 || ks > 0));
                    tc::mma_ss_acc(tmem, al + 2 * ks, wh + wstep * ks, idesc);
                    tc::mma_ss_acc(tmem, ah + 2 * ks, wl + wstep * ks, idesc);
                }
                tc::mma_commit(&bar);
            }
            __syncwarp();
        }
        pending = true;
        if (last_k) {
            tc::mbar_wait(&bar, phase);
            phase ^= 1;
            pending = false;
            tc::tc_fence_after();
            // ---- epilogue of (tile, ns): tensor memory -> registers (thread = row) -> XOR-swizzled fp32 staging -> coalesced
            // 128-bit rows with bias / ReLU / dropout / aux mask / beta, one 64-column piece at a time
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
                __syncthreads();
#pragma unroll 2
                for (int u = 0; u < 8; ++u) {
                    const int e = u * kThreads + tid;
                    const int rr = e >> 4, c4 = e & 15;
                    const int64_t row = row0 + rr;
                    const int col = n0 + 64 * pc + 4 * c4;
                    if (row >= p.M || col >= p.N) continue;
                    const float4 o = *reinterpret_cast<const float4*>(sOut + rr * 256 + ((c4 ^ (rr & 15)) << 4));
                    float x[4] = {o.x, o.y, o.z, o.w};
                    const int nval = (p.N - col < 4) ? p.N - col : 4;
                    float* out = p.C + row * p.ldc + col;
                    if (p.c_vec && nval == 4) {
                        if (p.flags & 1) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                            x[0] += b.x; x[1] += b.y; x[2] += b.z; x[3] += b.w;
                        }
                        if (p.flags & 2) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) x[j] = fmaxf(x[j], 0.0f);
                        }
                        if (p.flags & 4) {
                            const uint64_t el = (uint64_t)(p.rng_row0 + row) * (uint64_t)p.N + (uint64_t)col;     // N % 4 == 0: one 32-element group
                            const uint32_t kw = rng_keep_word_lo(p.keys, el >> 5, p.thr, p.low) >> (el & 31);
#pragma unroll
                            for (int j = 0; j < 4; ++j) x[j] = ((kw >> j) & 1u) ? x[j] * p.drop_scale : 0.0f;
                        }
                        if (p.flags & 8) {
                            const float4 m = __ldg(reinterpret_cast<const float4*>(p.aux + row * p.ldaux + col));
                            x[0] = m.x > 0.0f ? x[0] * p.aux_scale : 0.0f;
                            x[1] = m.y > 0.0f ? x[1] * p.aux_scale : 0.0f;
                            x[2] = m.z > 0.0f ? x[2] * p.aux_scale : 0.0f;
                            x[3] = m.w > 0.0f ? x[3] * p.aux_scale : 0.0f;
                        }
                        if (p.beta != 0.0f) {
                            const float4 c = *reinterpret_cast<const float4*>(out);
                            x[0] = fmaf(p.beta, c.x, x[0]); x[1] = fmaf(p.beta, c.y, x[1]);
                            x[2] = fmaf(p.beta, c.z, x[2]); x[3] = fmaf(p.beta, c.w, x[3]);
                        }
                        *reinterpret_cast<float4*>(out) = make_float4(x[0], x[1], x[2], x[3]);
                    } else {
                        for (int j = 0; j < nval; ++j) {
                            float y = x[j];
                            if (p.flags & 1) y += p.bias[col + j];
                            if (p.flags & 2) y = fmaxf(y, 0.0f);
                            if (p.flags & 4)
                                y *= rng_dropout_mult(p.keys, (uint64_t)(p.rng_row0 + row) * (uint64_t)p.N + (uint64_t)(col + j), p.thr, p.drop_scale);
                            if (p.flags & 8) y = (p.aux[row * p.ldaux + col + j] > 0.0f) ? y * p.aux_scale : 0.0f;
                            if (p.beta != 0.0f) y = fmaf(p.beta, out[j], y);
                            out[j] = y;
                        }
                    }
                }
                __syncthreads();                              // the staging is reused by the next piece
            }
            tc::tc_fence_before();
            prefetch();                                      // the epilogue's registers are dead now
        }
        tile = ntile; ns = nns; kb = nkb; valid = nvalid; first = false;
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

bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; }

}  // namespace

extern "C" int u2gnn_gemm_split_rows(const float* A, int64_t M, int K, int64_t lda, const float* W, int w_kn, int64_t ldw, int N,
                                     const float* bias, int epi, uint64_t seed, uint32_t rng_stream, int thr, int64_t rng_row0,
                                     const float* aux, int64_t ldaux, float aux_scale, float beta, float* C, int64_t ldc,
                                     u2gnn_stream_t stream) {
    if (!A || !W || !C || M < 0 || K < 1 || N < 1 || lda < K || ldc < N || ldw < (w_kn ? N : K)) return U2GNN_EINVAL;
    if ((epi & 1) && !bias) return U2GNN_EINVAL;
    if ((epi & 8) && (!aux || ldaux < N)) return U2GNN_EINVAL;
    if (epi & ~15) return U2GNN_EINVAL;
    if (thr < 0 || thr > 255) return U2GNN_EINVAL;
    if (M == 0) return U2GNN_OK;
    if ((epi & 4) && thr == 0) epi &= ~4;
    RowsP p;
    p.A = A; p.M = M; p.lda = lda; p.K = K; p.W = W; p.w_kn = w_kn; p.ldw = ldw; p.N = N;
    p.bias = bias; p.flags = epi; p.keys = rng_keys(seed, rng_stream); p.thr = thr; p.low = rng_thr_low(thr);
    p.drop_scale = thr ? rng_keep_scale(thr) : 1.0f; p.rng_row0 = rng_row0;
    p.aux = aux; p.ldaux = ldaux; p.aux_scale = aux_scale; p.beta = beta; p.C = C; p.ldc = ldc;
    p.a_vec = aligned16(A) && (lda & 3) == 0 && (K & 3) == 0;
    p.w_vec = aligned16(W) && (ldw & 3) == 0 && ((w_kn ? N : K) & 3) == 0;
    p.c_vec = aligned16(C) && (ldc & 3) == 0 && (N & 3) == 0 && (!(epi & 1) || aligned16(bias)) &&
              (!(epi & 8) || (aligned16(aux) && (ldaux & 3) == 0));
    cudaFuncSetAttribute(gemm_split_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowsSmem);
    const int64_t n_tiles = (M + TM - 1) / TM;
    const int64_t cap = (int64_t)U2GNN_NUM_SMS * 2;
    gemm_split_rows_kernel<<<(int)(n_tiles < cap ? n_tiles : cap), kThreads, kRowsSmem, as_stream(stream)>>>(p);
    U2GNN_CHECK_LAUNCH();
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
