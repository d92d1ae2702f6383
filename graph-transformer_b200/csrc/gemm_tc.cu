// bf16 tensor-core GEMMs for the attention-block projections of the bf16 mode (QKV, out-proj, their input
// gradients and weight gradients).  The matrices are tall and skinny (M = rows of the batch, K and N in
// {64, 128, 192}), so both kernels are HBM-bound: fp32 rows are read once with 128-bit loads, rounded to bf16
// into swizzled shared-memory tiles, multiplied with tcgen05.mma (accumulators in tensor memory) and written /
// accumulated in fp32.  Phase-serial CTAs (load -> MMA -> store); several CTAs per SM overlap the phases.
//
//   u2gnn_gemm_tc_rows   C[M, N] = A[M, K] W^T (+ bias)        W given as [N, K] (w_kn = 0) or [K, N] (w_kn = 1)
//   u2gnn_gemm_tc_wgrad  dW[N1, N2] += A[M, N1]^T B[M, N2],  db[N1] += colsum(A)     (N2 = 64 padded)
// Replaces the F.linear calls inside nn.MultiheadAttention (torch/nn/functional.py multi_head_attention_forward:
// in_proj / out_proj) and their autograd in the bf16 mode; the fp32 mode keeps u2gnn_sgemm.
#include "common.cuh"
#include "tc_common.cuh"
#include "ffn_epi.cuh"
#include "rng.cuh"

namespace {

constexpr int TM = 128;
constexpr int kThreads = 256;

// [rows x K] block (row stride lda in ELEMENTS; fp32, or bf16 when src_bf16) -> K/64 swizzled bf16 tiles of [128 x 64];
// rows >= M and cols >= K_true zero.  A bf16 source is what the upstream kernel already rounded (bit-identical to
// rounding here) at half the HBM bytes.
template <int NT, int KPC = 0>      // KPC: padded K known at compile time (index arithmetic without runtime divisions), 0 = runtime
__device__ __forceinline__ void stage_rows_bf16(uint8_t* tiles, const void* __restrict__ src_, int src_bf16, int64_t row0, int64_t M,
                                                int K_true, int KP_rt, int64_t lda, int tid) {
    const int KP = KPC ? KPC : KP_rt;
    const int c4n = KP / 4;                                   // 4-element groups per padded row
    if ((K_true & 3) == 0 && (lda & 3) == 0) {
        if (src_bf16 && (K_true & 7) == 0 && (lda & 7) == 0) {
            // 16-byte pieces (8 bf16 = one swizzle chunk), every load of the tile in flight before the first store
            const __nv_bfloat16* src = static_cast<const __nv_bfloat16*>(src_);
            const int c8n = KP / 8;
            for (int base = 0; base < TM * c8n; base += NT * 8) {
                uint4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = base + u * NT + tid;
                    const int r = e / c8n, c8 = e - r * c8n;
                    v[u] = make_uint4(0u, 0u, 0u, 0u);
                    if (e < TM * c8n && row0 + r < M && 8 * c8 < K_true) v[u] = __ldg(reinterpret_cast<const uint4*>(src + (row0 + r) * lda) + c8);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = base + u * NT + tid;
                    if (e >= TM * c8n) continue;
                    const int r = e / c8n, c8 = e - r * c8n;
                    *reinterpret_cast<uint4*>(tiles + (c8 >> 3) * 16384 + tc::sw128_chunk(r, c8 & 7)) = v[u];
                }
            }
            return;
        }
        if (src_bf16) {
            const __nv_bfloat16* src = static_cast<const __nv_bfloat16*>(src_);
            for (int base = 0; base < TM * c4n; base += NT * 8) {
                uint2 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = base + u * NT + tid;
                    const int r = e / c4n, c4 = e - r * c4n;
                    v[u] = make_uint2(0u, 0u);
                    if (e < TM * c4n && row0 + r < M && 4 * c4 < K_true) v[u] = __ldg(reinterpret_cast<const uint2*>(src + (row0 + r) * lda) + c4);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = base + u * NT + tid;
                    if (e >= TM * c4n) continue;
                    const int r = e / c4n, c4 = e - r * c4n;
                    const int col = 4 * c4;
                    *reinterpret_cast<uint2*>(tiles + (col >> 6) * 16384 + tc::sw128_offset(r, col & 63)) = v[u];
                }
            }
            return;
        }
        const float* src = static_cast<const float*>(src_);
        for (int base = 0; base < TM * c4n; base += NT * 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = base + u * NT + tid;
                const int r = e / c4n, c4 = e - r * c4n;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (e < TM * c4n && row0 + r < M && 4 * c4 < K_true) v[u] = __ldg(reinterpret_cast<const float4*>(src + (row0 + r) * lda) + c4);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = base + u * NT + tid;
                if (e >= TM * c4n) continue;
                const int r = e / c4n, c4 = e - r * c4n;
                const int col = 4 * c4;
                uint2 w;
                w.x = epi::cvt2(v[u].x, v[u].y);
                w.y = epi::cvt2(v[u].z, v[u].w);
                *reinterpret_cast<uint2*>(tiles + (col >> 6) * 16384 + tc::sw128_offset(r, col & 63)) = w;
            }
        }
    } else {
        for (int e = tid; e < TM * KP; e += NT) {
            const int r = e / KP, k = e - r * KP;
            __nv_bfloat16 v = __float2bfloat16(0.0f);
            if (row0 + r < M && k < K_true)
                v = src_bf16 ? static_cast<const __nv_bfloat16*>(src_)[(row0 + r) * lda + k]
                             : __float2bfloat16(static_cast<const float*>(src_)[(row0 + r) * lda + k]);
            *reinterpret_cast<__nv_bfloat16*>(tiles + (k >> 6) * 16384 + tc::sw128_offset(r, k & 63)) = v;
        }
    }
}

// fp32 rows idx[row0 + r] of a 64-column table -> one swizzled bf16 [128 x 64] tile (the gather of F.embedding fused into the GEMM
// operand load); rows >= M and indices outside the table are zero
// the 8 table rows thread `tid` of a 256-thread CTA reads for the tile at row0 (-1 past M): loaded one tile AHEAD of the rows
// themselves so that the index -> row dependency is not a second exposed memory latency per tile
// (32-bit: a 64-column fp32 table with 2^31 rows would be 550 GB; out-of-range indices become -1)
__device__ __forceinline__ void gather_idx8(int (&rows)[8], const int64_t* __restrict__ idx, int64_t n_table, int64_t row0, int64_t M, int tid) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int64_t r = row0 + ((u * kThreads + tid) >> 4);
        const int64_t v = (r < M) ? __ldg(idx + r) : -1;
        rows[u] = (v >= 0 && v < n_table) ? (int)v : -1;
    }
}
template <int NT>
__device__ __forceinline__ void stage_rows_gather64(uint8_t* tile, const float* __restrict__ table, const int (&rows)[8], int64_t ld, int tid) {
    static_assert(NT == kThreads, "one batch of 8 float4 per thread");
    for (int base = 0; base < TM * 16; base += NT * 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = base + u * NT + tid;
            const int c4 = e & 15;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rows[u] >= 0) v[u] = __ldg(reinterpret_cast<const float4*>(table + (int64_t)rows[u] * ld) + c4);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = base + u * NT + tid;
            if (e >= TM * 16) continue;
            const int r = e >> 4, c4 = e & 15;
            uint2 w;
            w.x = epi::cvt2(v[u].x, v[u].y);
            w.y = epi::cvt2(v[u].z, v[u].w);
            *reinterpret_cast<uint2*>(tile + tc::sw128_offset(r, c4 * 4)) = w;
        }
    }
}

// bf16 row block -> swizzled tiles with cp.async (16-byte chunk = one swizzle chunk, no registers, asynchronous): issue only,
// the caller commits / waits.  Rows >= M and chunks >= K_true are zero-filled (src-size 0).  Needs K_true, lda multiples of 8
// and a 16-byte aligned base (rows_async_ok).
__device__ __forceinline__ bool rows_async_ok(const void* src, int src_bf16, int K_true, int64_t lda) {
    return src_bf16 && (K_true & 7) == 0 && (lda & 7) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
}
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NT>
__device__ __forceinline__ void stage_rows_async(uint8_t* tiles, const void* __restrict__ src_, int64_t row0, int64_t M, int K_true,
                                                 int KP, int64_t lda, int tid) {
    const __nv_bfloat16* src = static_cast<const __nv_bfloat16*>(src_);
    const uint32_t base = tc::smem_u32(tiles);
    // 8 lanes = the 8 chunks (128 contiguous bytes) of one row of one 64-column tile
    const int c = tid & 7;
    for (int t = 0; t < KP / 64; ++t) {
        const bool col_ok = 64 * t + 8 * c < K_true;
        for (int r = tid >> 3; r < TM; r += NT / 8) {
            const bool ok = col_ok && row0 + r < M;
            const void* g = ok ? static_cast<const void*>(src + (row0 + r) * lda + 64 * t + 8 * c) : src_;
            cp_async16(base + (uint32_t)t * 16384u + tc::sw128_chunk(r, c), g, ok ? 16u : 0u);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// rows GEMM
// ---------------------------------------------------------------------------------------------------------------
struct RowsParams {
    const void* A;
    int a_bf16, c_bf16;   // A / C stored as bf16 row-major instead of fp32 (lda / ldc in elements)
    int64_t M, lda;
    int K, KP;            // true / padded (multiple of 64) inner size
    const float* W;
    int w_kn;             // 0: W[N][K]   1: W[K][N]
    int N, NP;            // true / padded (multiple of 16) output size
    const float* bias;
    void* C;
    int64_t ldc;
    float beta;           // C = result + beta * C  (0 or 1; fp32 C only)
    // fused residual + dropout + LayerNorm epilogue (LN kernel variant only; N == 64):
    //   z = res + dropout(A W^T + bias);  y = LayerNorm(z) * gamma + ln_beta;  stats = (mean, rstd)
    const float* res;
    int64_t ldres;
    const int64_t* res_idx;   // optional: residual row of output row r = res[res_idx[r]] (gather fused; res_rows = rows of that table)
    int64_t res_rows;
    RngKeys keys;
    int thr, low;
    float scale;
    const float* gamma;
    const float* ln_beta;
    float* z;
    float* y;
    float* stats;
    uint8_t* y_img;       // optional: y also as bf16 swizzled [128 x 64] tile images (the FFN backward's operand format)
};

constexpr float kLnEps = 1e-5f;

__device__ __forceinline__ float group16_sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <bool LN, int MINB>        // MINB: CTAs per SM the register budget is sized for (wide outputs are limited to 2 by tensor memory / smem)
__global__ void __launch_bounds__(kThreads, MINB) gemm_tc_rows_kernel(const RowsParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // offset on the shared-window address: keeps LDS / STS (a uintptr_t round trip makes every access generic)
    const int kt = p.KP / 64;
    uint8_t* sA = smem;                                     // kt tiles of [128 x 64]
    uint8_t* sB = smem + kt * 16384;                        // kt tiles of [NP x 64] (NP*128 B each)
    uint8_t* sOut = sB + kt * p.NP * 128;                   // fp32 staging of one 64-column output piece: [128 rows x 272 B]
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::fence_barrier_init();
    }
    const int tcols = (p.NP <= 64) ? 64 : (p.NP <= 128 ? 128 : 256);
    if (warp == 0) {
        if (tcols == 64) tc::tmem_alloc<64>(&tmem_slot);
        else if (tcols == 128) tc::tmem_alloc<128>(&tmem_slot);
        else tc::tmem_alloc<256>(&tmem_slot);
    }
    // weights -> K-major image: element (n, k) of the [NP x KP] operand
    for (int e = tid; e < p.NP * p.KP; e += kThreads) {
        const int n = e / p.KP, k = e - n * p.KP;
        float w = 0.0f;
        if (n < p.N && k < p.K) w = p.w_kn ? p.W[(size_t)k * p.N + n] : p.W[(size_t)n * p.K + k];
        *reinterpret_cast<__nv_bfloat16*>(sB + (k >> 6) * (p.NP * 128) + tc::sw128_offset(n, k & 63)) = __float2bfloat16(w);
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = tc::make_idesc(TM, p.NP, 0, 0);
    const uint64_t a_desc = tc::make_desc_sw128(tc::smem_u32(sA), 16, 1024);
    const uint64_t b_desc = tc::make_desc_sw128(tc::smem_u32(sB), 16, 1024);
    const int64_t n_tiles = (p.M + TM - 1) / TM;
    const int wq = warp & 3, half = warp >> 2;              // lane quarter / column half for the epilogue
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    uint32_t phase = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * TM;
        int rrows[8];                                           // fused gather of the residual: its table rows, requested before the operand tile
        if (LN && p.res_idx) gather_idx8(rrows, p.res_idx, p.res_rows, row0, p.M, tid);
        if (p.KP == 64) stage_rows_bf16<kThreads, 64>(sA, p.A, p.a_bf16, row0, p.M, p.K, 64, p.lda, tid);
        else if (p.KP == 192) stage_rows_bf16<kThreads, 192>(sA, p.A, p.a_bf16, row0, p.M, p.K, 192, p.lda, tid);
        else stage_rows_bf16<kThreads>(sA, p.A, p.a_bf16, row0, p.M, p.K, p.KP, p.lda, tid);
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        if (warp == 0) {
            if (tc::elect_one()) {
                for (int ks = 0; ks < p.KP / 16; ++ks) {
                    const uint32_t ko = (uint32_t)((ks >> 2) * 1024 + (ks & 3) * 2);
                    const uint32_t bo = (uint32_t)((ks >> 2) * (p.NP * 8) + (ks & 3) * 2);
                    tc::mma_ss(tmem, a_desc + ko, b_desc + bo, idesc, ks > 0);
                }
                tc::mma_commit(&bar);
            }
            __syncwarp();
        }
        // rows the coalesced phase of this thread adds in (16 lanes x float4 per row), requested while the MMA runs:
        // the residual (LN variant) or the old C of the first 64-column piece (beta != 0)
        const bool vec_ok = ((p.ldc & 3) == 0) && ((p.N & 3) == 0);
        const bool pre_old = !LN && p.beta != 0.0f && !p.c_bf16 && vec_ok;
        float4 pre[8];
        if (LN || pre_old) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = u * kThreads + tid;
                const int64_t row = row0 + (e >> 4);
                pre[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if constexpr (LN) {
                    if (row < p.M) {
                        int64_t rrow = row;
                        if (p.res_idx) {
                            rrow = rrows[u];
                            if (rrow < 0) continue;                           // (the forward kernel that gathered x has flagged it)
                        }
                        pre[u] = __ldg(reinterpret_cast<const float4*>(p.res + rrow * p.ldres) + (e & 15));
                    }
                } else {
                    if (row < p.M && 4 * (e & 15) < p.N) pre[u] = *(reinterpret_cast<const float4*>(static_cast<const float*>(p.C) + row * p.ldc) + (e & 15));
                }
            }
        }
        tc::mbar_wait(&bar, phase);
        phase ^= 1;
        tc::tc_fence_after();
        // epilogue, one 64-column piece at a time: tensor memory -> registers (thread = row of the lane quarter, the two
        // warps of a quarter take 32 columns each) -> padded shared-memory staging -> COALESCED 128-bit global stores
        // (consecutive threads = consecutive 16-byte pieces of a row).  Writing rows straight from the thread-per-row
        // registers touched 32 cache lines per store instruction and was the largest cost of this kernel.
        // The copy-out loads every staged piece (and the bias, once) BEFORE the guarded arithmetic / stores: guarded
        // loads cannot be hoisted by the compiler and made each of the 4-8 iterations wait for its own round trip.
        const bool bf16_vec = !LN && p.c_bf16 && vec_ok && (p.ldc & 7) == 0 && (p.N & 7) == 0;
        for (int c0 = 0; c0 < p.NP; c0 += 64) {
            if (c0 + 32 * half < p.NP) {
                uint32_t v[32];
                tc::tmem_ld32(tmem + lane_base + c0 + 32 * half, v);
                tc::tmem_ld_wait();
                uint8_t* srow = sOut + (wq * 32 + lane) * 272;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    // 16-byte piece J = 8 half + j of the row.  bf16 copy-out: a thread converts pieces (2c, 2c+1), so they are
                    // staged at positions (c, 8 + c) and each quarter-warp load covers 128 contiguous bytes (no conflicts)
                    const int J = 8 * half + j;
                    const int pos = bf16_vec ? ((J >> 1) + 8 * (J & 1)) : J;
                    *reinterpret_cast<float4*>(srow + 16 * pos) = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                              __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                }
            }
            __syncthreads();
            if constexpr (LN) {
                // N == 64: 16 consecutive lanes hold one row as float4 pieces - the layout of ln_fwd_vec_kernel<16>, same
                // arithmetic in the same order (layernorm.cu), so z / y / stats are what GEMM + LayerNorm kernels produce
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma) + (tid & 15));
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.ln_beta) + (tid & 15));
                const float4 bias4 = __ldg(reinterpret_cast<const float4*>(p.bias) + (tid & 15));
                float4 st[4];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if ((u & 3) == 0) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) st[q] = *reinterpret_cast<const float4*>(sOut + (((u + q) * kThreads + tid) >> 4) * 272 + 16 * (tid & 15));
                    }
                    const int e = u * kThreads + tid;
                    const int rr = e >> 4, l = e & 15;
                    const int64_t row = row0 + rr;
                    const bool ok = row < p.M;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok) {
                        v = st[u & 3];
                        v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
                        if (p.thr) {
                            const uint64_t el = (uint64_t)(row * 64 + 4 * l);
                            const uint32_t w = rng_keep_word_lo(p.keys, el >> 5, p.thr, p.low) >> (el & 31);
                            v.x = (w & 1u) ? v.x * p.scale : 0.0f;
                            v.y = (w & 2u) ? v.y * p.scale : 0.0f;
                            v.z = (w & 4u) ? v.z * p.scale : 0.0f;
                            v.w = (w & 8u) ? v.w * p.scale : 0.0f;
                        }
                        v.x += pre[u].x; v.y += pre[u].y; v.z += pre[u].z; v.w += pre[u].w;
                        reinterpret_cast<float4*>(p.z + row * 64)[l] = v;
                    }
                    const float mean = group16_sum((v.x + v.y) + (v.z + v.w)) * (1.0f / 64.0f);
                    const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
                    const float rstd = rsqrtf(group16_sum((dx * dx + dy * dy) + (dz * dz + dw * dw)) * (1.0f / 64.0f) + kLnEps);
                    if (ok) {
                        const float4 yv = make_float4(dx * rstd * g4.x + b4.x, dy * rstd * g4.y + b4.y, dz * rstd * g4.z + b4.z, dw * rstd * g4.w + b4.w);
                        reinterpret_cast<float4*>(p.y + row * 64)[l] = yv;
                        if (p.y_img) {      // the tile image the FFN backward bulk-copies: 128-byte rows, 16-byte chunks XOR (row & 7)
                            uint2 w;
                            w.x = tc::pack_bf16(yv.x, yv.y);
                            w.y = tc::pack_bf16(yv.z, yv.w);
                            *reinterpret_cast<uint2*>(p.y_img + (size_t)(row >> 7) * 16384 + (size_t)(row & 127) * 128 +
                                                      ((((l >> 1) ^ (int)(row & 7)) << 4) | ((l & 1) << 3))) = w;
                        }
                        if (l == 0) *reinterpret_cast<float2*>(p.stats + 2 * row) = make_float2(mean, rstd);
                    }
                }
                __syncthreads();
                continue;
            }
            if (bf16_vec) {
                // bf16 result (rounded once here: what every consumer would do on load), 8 columns = 16 bytes per thread
                const int c8 = tid & 7;
                const int col = c0 + 8 * c8;
                float4 bb0 = make_float4(0.f, 0.f, 0.f, 0.f), bb1 = bb0;
                if (p.bias && col < p.N) {
                    bb0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                    bb1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
                }
                float4 o0[4], o1[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int rr = (u * kThreads + tid) >> 3;
                    o0[u] = *reinterpret_cast<const float4*>(sOut + rr * 272 + 16 * c8);
                    o1[u] = *reinterpret_cast<const float4*>(sOut + rr * 272 + 128 + 16 * c8);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t row = row0 + ((u * kThreads + tid) >> 3);
                    if (row < p.M && col < p.N) {
                        uint4 w;
                        w.x = epi::cvt2(o0[u].x + bb0.x, o0[u].y + bb0.y); w.y = epi::cvt2(o0[u].z + bb0.z, o0[u].w + bb0.w);
                        w.z = epi::cvt2(o1[u].x + bb1.x, o1[u].y + bb1.y); w.w = epi::cvt2(o1[u].z + bb1.z, o1[u].w + bb1.w);
                        *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.C) + row * p.ldc + col) = w;
                    }
                }
                __syncthreads();
                continue;
            }
            if (!p.c_bf16 && vec_ok) {
                // fp32 result, 4 columns = 16 bytes per thread, 16 lanes per row
                const int c4 = tid & 15;
                const int col = c0 + 4 * c4;
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias && col < p.N) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                float4 o[4];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if ((u & 3) == 0) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) o[q] = *reinterpret_cast<const float4*>(sOut + (((u + q) * kThreads + tid) >> 4) * 272 + 16 * c4);
                    }
                    const int64_t row = row0 + ((u * kThreads + tid) >> 4);
                    if (row >= p.M || col >= p.N) continue;
                    float* out = static_cast<float*>(p.C) + row * p.ldc + col;
                    float4 r = o[u & 3];
                    r.x += b4.x; r.y += b4.y; r.z += b4.z; r.w += b4.w;
                    if (p.beta != 0.0f) {
                        const float4 old = (pre_old && c0 == 0) ? pre[u] : *reinterpret_cast<const float4*>(out);
                        r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
                    }
                    *reinterpret_cast<float4*>(out) = r;
                }
                __syncthreads();
                continue;
            }
            // unaligned shapes (parity-sized feature dimensions): scalar tail handling
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = u * kThreads + tid;
                const int rr = e >> 4, c4 = e & 15;
                const int64_t row = row0 + rr;
                const int col = c0 + 4 * c4;
                if (row >= p.M || col >= p.N) continue;
                const float4 o = *reinterpret_cast<const float4*>(sOut + rr * 272 + 16 * c4);
                const float ov[4] = {o.x, o.y, o.z, o.w};
                if (p.c_bf16) {                              // rounded once here: what every consumer would do on load
                    __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(p.C) + row * p.ldc + col;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (col + j < p.N) ob[j] = __float2bfloat16(ov[j] + (p.bias ? p.bias[col + j] : 0.0f));
                    continue;
                }
                float* out = static_cast<float*>(p.C) + row * p.ldc + col;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (col + j < p.N) {
                        float x = ov[j] + (p.bias ? p.bias[col + j] : 0.0f);
                        if (p.beta != 0.0f) x += out[j];
                        out[j] = x;
                    }
            }
            __syncthreads();                                 // staging is reused by the next piece / tile
        }
        tc::tc_fence_before();
        __syncthreads();                                     // TMEM and sA are reused by the next tile
    }
    __syncthreads();
    if (warp == 0) {
        if (tcols == 64) tc::tmem_dealloc<64>(tmem);
        else if (tcols == 128) tc::tmem_dealloc<128>(tmem);
        else tc::tmem_dealloc<256>(tmem);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// weight-gradient GEMM:  dW[N1, N2] += A^T B,  db[N1] += colsum(A).   N1 <= 256 (64-column groups), N2 <= 64.
// TMEM: accumulator g covers A columns [128 g, 128 g + 128): 80 columns each (64 for dW + ones column for db).
// ---------------------------------------------------------------------------------------------------------------
struct WgradParams {
    const void* A;
    const void* B;
    int a_bf16, b_bf16;   // operands stored as bf16 row-major
    int64_t M, lda, ldb;
    int N1, N2;
    float* dW;            // [N1, N2]
    float* db;            // [N1] or null
    const int64_t* b_idx; // optional (fp32 B, N2 = 64): B row r = B[b_idx[r]] of a table with b_rows rows
    int64_t b_rows;
};

__global__ void __launch_bounds__(kThreads, 2) gemm_tc_wgrad_kernel(const WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // offset on the shared-window address: keeps LDS / STS (a uintptr_t round trip makes every access generic)
    const int ga = (p.N1 + 63) / 64;                         // 64-column groups of A (1..4)
    const int nacc = (ga + 1) / 2;                           // accumulators of 128 A-columns
    // per stage: ga A tiles, 1 B tile ; then ones tile and zero tile
    const uint32_t stage_bytes = (uint32_t)(ga + 1) * 16384;
    uint8_t* sOnes = smem + 2 * stage_bytes;
    uint8_t* sZero = sOnes + 16384;
    __shared__ uint64_t bar_mma[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        tc::mbar_init(&bar_mma[0], 1);
        tc::mbar_init(&bar_mma[1], 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc<256>(&tmem_slot);
    for (int e = tid; e < 16384 / 4; e += kThreads) {
        reinterpret_cast<uint32_t*>(sOnes)[e] = 0x3F803F80u;
        reinterpret_cast<uint32_t*>(sZero)[e] = 0u;
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = tc::make_idesc(128, 80, 1, 1);    // M = 128 A-columns, N = [B | ones], both MN-major
    const int64_t n_tiles = (p.M + TM - 1) / TM;
    // Two stages, loads one tile ahead of the MMAs.  bf16 operands travel with cp.async (asynchronous, no registers) straight
    // into the swizzled tiles; an fp32 operand is converted through registers AFTER the current tile's MMAs were issued, so
    // its load latency overlaps them and the cp.async traffic of the same tile.  (The first version loaded, converted and
    // stored each operand synchronously: one CTA per SM for the wide in_proj gradient exposed three DRAM round trips per
    // tile - 37 % of the HBM bandwidth.)
    const bool a_async = rows_async_ok(p.A, p.a_bf16, p.N1, p.lda);
    const bool b_async = !p.b_idx && rows_async_ok(p.B, p.b_bf16, p.N2, p.ldb);
    auto issue_async = [&](int64_t tile, uint8_t* st) {
        if (a_async) stage_rows_async<kThreads>(st, p.A, tile * TM, p.M, p.N1, ga * 64, p.lda, tid);
        if (b_async) stage_rows_async<kThreads>(st + ga * 16384, p.B, tile * TM, p.M, p.N2, 64, p.ldb, tid);
    };
    int grows[8];                                           // fused gather: table rows of the tile the next finish_sync stages
    auto finish_sync = [&](int64_t tile, uint8_t* st) {
        if (!a_async) stage_rows_bf16<kThreads>(st, p.A, p.a_bf16, tile * TM, p.M, p.N1, ga * 64, p.lda, tid);
        if (p.b_idx) {
            stage_rows_gather64<kThreads>(st + ga * 16384, static_cast<const float*>(p.B), grows, p.ldb, tid);
            gather_idx8(grows, p.b_idx, p.b_rows, (tile + gridDim.x) * TM, p.M, tid);       // for the tile after this one
        } else if (!b_async) stage_rows_bf16<kThreads>(st + ga * 16384, p.B, p.b_bf16, tile * TM, p.M, p.N2, 64, p.ldb, tid);
    };
    int64_t it = 0;
    if ((int64_t)blockIdx.x < n_tiles) {
        if (p.b_idx) gather_idx8(grows, p.b_idx, p.b_rows, (int64_t)blockIdx.x * TM, p.M, tid);
        issue_async(blockIdx.x, smem);
        cp_async_commit();
        finish_sync(blockIdx.x, smem);
    }
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int s = (int)(it & 1);
        uint8_t* st = smem + s * stage_bytes;
        uint8_t* st_next = smem + (s ^ 1) * stage_bytes;
        const int64_t next = tile + gridDim.x;
        if (next < n_tiles) {
            if (it >= 1) tc::mbar_wait(&bar_mma[s ^ 1], (uint32_t)(((it - 1) >> 1) & 1));   // MMAs of tile it-1 have read that stage
            issue_async(next, st_next);
        }
        cp_async_commit();
        cp_async_wait<1>();                                   // this thread's chunks of tile `it` have landed
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        if (warp == 0) {
            if (tc::elect_one()) {
                const uint32_t a0 = tc::smem_u32(st), b0 = tc::smem_u32(st + ga * 16384);
                const uint64_t bd = tc::make_desc_sw128(b0, tc::smem_u32(sOnes) - b0, 1024);       // [B | ones]
                for (int g = 0; g < nacc; ++g) {
                    const uint32_t t0 = a0 + (uint32_t)(2 * g) * 16384;
                    const bool second = (2 * g + 1 < ga);
                    const uint32_t lbo = second ? 16384u : (tc::smem_u32(sZero) - t0);
                    const uint64_t ad = tc::make_desc_sw128(t0, lbo, 1024);
                    for (int ks = 0; ks < 8; ++ks) tc::mma_ss(tmem + 80 * g, ad + 128 * ks, bd + 128 * ks, idesc, (it > 0 || ks > 0));
                }
                tc::mma_commit(&bar_mma[s]);
            }
            __syncwarp();
        }
        if (next < n_tiles) finish_sync(next, st_next);
    }
    cp_async_wait<0>();
    // drain: wait for the last one / two commits (commit k on a stage completes phase k & 1)
    if (it >= 1) {
        const int s1 = (int)((it - 1) & 1);
        tc::mbar_wait(&bar_mma[s1], (uint32_t)(((it - 1) >> 1) & 1));
        if (it >= 2) tc::mbar_wait(&bar_mma[s1 ^ 1], (uint32_t)(((it - 2) >> 1) & 1));
        tc::tc_fence_after();
        // flush: thread = A column (row of the accumulator); 128 threads per accumulator pass
        const int wq = warp & 3;
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
        for (int g = warp >> 2; g < nacc; g += 2) {
            const int n1 = 128 * g + wq * 32 + lane;
            uint32_t v[32];
            for (int c0 = 0; c0 < 64; c0 += 32) {
                tc::tmem_ld32(tmem + lane_base + 80 * g + c0, v);
                tc::tmem_ld_wait();
                if (n1 < p.N1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j < p.N2) atomicAdd(p.dW + (size_t)n1 * p.N2 + c0 + j, __uint_as_float(v[j]));
                }
            }
            uint32_t b16[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(b16[0]), "=r"(b16[1]), "=r"(b16[2]), "=r"(b16[3]), "=r"(b16[4]), "=r"(b16[5]), "=r"(b16[6]),
                           "=r"(b16[7]), "=r"(b16[8]), "=r"(b16[9]), "=r"(b16[10]), "=r"(b16[11]), "=r"(b16[12]), "=r"(b16[13]),
                           "=r"(b16[14]), "=r"(b16[15])
                         : "r"(tmem + lane_base + 80 * g + 64)
                         : "memory");
            tc::tmem_ld_wait();
            if (p.db && n1 < p.N1) atomicAdd(p.db + n1, __uint_as_float(b16[0]));
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

// ---------------------------------------------------------------------------------------------------------------
// Backward of one projection in ONE pass over its output gradient A[M, N1] (bf16, N1 = 64 ga):
//     C[M, 64]   = A W (+ beta C)          input gradient   (W = the layer's weight [N1, 64], i.e. [K][N])
//     dW[N1, 64] += A^T B,  db[N1] += colsum(A)             weight / bias gradient (B = the layer's input rows)
// The rows GEMM reads the staged A tiles K-major and the weight-gradient GEMM reads the SAME tiles MN-major, so A crosses
// HBM once instead of twice (in_proj: 384 of the 1 536 bytes per row the two separate kernels moved).  NS-stage ring,
// loads NS-1 tiles ahead: A (and a bf16 B) via cp.async, an fp32 B through registers under the MMAs; the epilogue of tile i
// (tensor memory -> staging -> coalesced stores, old C prefetched) runs while the loads of the next tiles are in flight.
// TMEM: weight-gradient accumulators [0, 80 nacc), rows accumulator [192, 256).
// ---------------------------------------------------------------------------------------------------------------
struct BwdParams {
    const void* A;        // [M, N1] bf16
    const void* B;        // [M, 64] fp32 or bf16
    int b_bf16, c_bf16;
    int64_t M, lda, ldb, ldc;
    int N1;
    const float* W;       // [N1, 64]
    void* C;              // [M, 64] fp32 (beta 0 / 1) or bf16 (beta 0)
    float beta;
    float* dW;            // [N1, 64]
    float* db;            // [N1] or null
    const int64_t* b_idx; // optional (fp32 B): B row r = B[b_idx[r]] of a table with b_rows rows
    int64_t b_rows;
};

template <int NS>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_dgrad_wgrad_kernel(const BwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int ga = p.N1 / 64;
    const int nacc = (ga + 1) / 2;
    const uint32_t stage_bytes = (uint32_t)(ga + 1) * 16384;
    uint8_t* sW = smem + NS * stage_bytes;                    // ga tiles of [64 (n) x 64 (k)] K-major: element (n, k) = W[k][n]
    uint8_t* sOut = sW + ga * 8192;                           // [128 rows x 272 B] fp32 staging
    uint8_t* sOnes = sOut + 128 * 272;
    uint8_t* sZero = sOnes + 16384;
    __shared__ uint64_t bar_mma[NS];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) tc::mbar_init(&bar_mma[i], 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc<256>(&tmem_slot);
    for (int e = tid; e < 16384 / 4; e += kThreads) {
        reinterpret_cast<uint32_t*>(sOnes)[e] = 0x3F803F80u;
        reinterpret_cast<uint32_t*>(sZero)[e] = 0u;
    }
    for (int e = tid; e < p.N1 * 64; e += kThreads) {
        const int k = e >> 6, n = e & 63;                     // consecutive threads read consecutive floats of W
        *reinterpret_cast<__nv_bfloat16*>(sW + (k >> 6) * 8192 + tc::sw128_offset(n, k & 63)) = __float2bfloat16(p.W[e]);
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t tmem_rows = tmem + 192;
    const uint32_t idesc_w = tc::make_idesc(128, 80, 1, 1);   // weight gradient: M = 128 A-columns, N = [B | ones], both MN-major
    const uint32_t idesc_r = tc::make_idesc(TM, 64, 0, 0);    // rows: K-major A tiles and weights
    const uint64_t w_desc = tc::make_desc_sw128(tc::smem_u32(sW), 16, 1024);
    const int64_t n_tiles = (p.M + TM - 1) / TM;
    const bool b_async = !p.b_idx && rows_async_ok(p.B, p.b_bf16, 64, p.ldb);
    auto issue_async = [&](int64_t tile, uint8_t* st) {
        stage_rows_async<kThreads>(st, p.A, tile * TM, p.M, p.N1, p.N1, p.lda, tid);
        if (b_async) stage_rows_async<kThreads>(st + ga * 16384, p.B, tile * TM, p.M, 64, 64, p.ldb, tid);
    };
    auto finish_sync = [&](int64_t tile, uint8_t* st) {
        if (p.b_idx) {
            int grows[8];
            gather_idx8(grows, p.b_idx, p.b_rows, tile * TM, p.M, tid);
            stage_rows_gather64<kThreads>(st + ga * 16384, static_cast<const float*>(p.B), grows, p.ldb, tid);
        } else if (!b_async) stage_rows_bf16<kThreads, 64>(st + ga * 16384, p.B, p.b_bf16, tile * TM, p.M, 64, 64, p.ldb, tid);
    };
    // prologue: tiles 0 .. NS-2 of this CTA
#pragma unroll
    for (int j = 0; j < NS - 1; ++j) {
        const int64_t t = blockIdx.x + (int64_t)j * gridDim.x;
        if (t < n_tiles) issue_async(t, smem + j * stage_bytes);
        cp_async_commit();
        if (t < n_tiles) finish_sync(t, smem + j * stage_bytes);
    }
    const int wq = warp & 3, half = warp >> 2;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    int64_t it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int s = (int)(it % NS), sp = (int)((it + NS - 1) % NS);
        uint8_t* st = smem + s * stage_bytes;
        uint8_t* st_pre = smem + sp * stage_bytes;
        const int64_t row0 = tile * TM;
        const int64_t nxt = tile + (int64_t)(NS - 1) * gridDim.x;
        // stage sp was read by the MMAs of tile it-1; the epilogue of that tile has already waited for them
        if (nxt < n_tiles) issue_async(nxt, st_pre);
        cp_async_commit();
        cp_async_wait<NS - 1>();                              // this thread's chunks of tile `it` have landed
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        if (warp == 0) {
            if (tc::elect_one()) {
                const uint32_t a0 = tc::smem_u32(st), b0 = tc::smem_u32(st + ga * 16384);
                const uint64_t a_desc = tc::make_desc_sw128(a0, 16, 1024);
                for (int ks = 0; ks < p.N1 / 16; ++ks)
                    tc::mma_ss(tmem_rows, a_desc + (uint32_t)((ks >> 2) * 1024 + (ks & 3) * 2), w_desc + (uint32_t)((ks >> 2) * 512 + (ks & 3) * 2),
                               idesc_r, ks > 0);
                const uint64_t bd = tc::make_desc_sw128(b0, tc::smem_u32(sOnes) - b0, 1024);       // [B | ones]
                for (int g = 0; g < nacc; ++g) {
                    const uint32_t t0 = a0 + (uint32_t)(2 * g) * 16384;
                    const bool second = (2 * g + 1 < ga);
                    const uint32_t lbo = second ? 16384u : (tc::smem_u32(sZero) - t0);
                    const uint64_t ad = tc::make_desc_sw128(t0, lbo, 1024);
                    for (int ks = 0; ks < 8; ++ks) tc::mma_ss(tmem + 80 * g, ad + 128 * ks, bd + 128 * ks, idesc_w, (it > 0 || ks > 0));
                }
                tc::mma_commit(&bar_mma[s]);
            }
            __syncwarp();
        }
        // old C rows of this tile (beta), requested before the fp32 operand of the prefetched tile and the MMA wait
        float4 pre[8];
        const bool use_old = p.beta != 0.0f && !p.c_bf16;
        if (use_old) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = u * kThreads + tid;
                const int64_t row = row0 + (e >> 4);
                pre[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < p.M) pre[u] = *(reinterpret_cast<const float4*>(static_cast<const float*>(p.C) + row * p.ldc) + (e & 15));
            }
        }
        if (nxt < n_tiles) finish_sync(nxt, st_pre);
        tc::mbar_wait(&bar_mma[s], (uint32_t)((it / NS) & 1));
        tc::tc_fence_after();
        {
            uint32_t v[32];
            tc::tmem_ld32(tmem_rows + lane_base + 32 * half, v);
            tc::tmem_ld_wait();
            uint8_t* srow = sOut + (wq * 32 + lane) * 272;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int J = 8 * half + j;
                const int pos = p.c_bf16 ? ((J >> 1) + 8 * (J & 1)) : J;      // bf16 copy-out: pieces (2c, 2c+1) at (c, 8 + c)
                *reinterpret_cast<float4*>(srow + 16 * pos) = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                          __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
            }
        }
        tc::tc_fence_before();
        __syncthreads();
        if (p.c_bf16) {
            const int c8 = tid & 7;
            float4 o0[4], o1[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int rr = (u * kThreads + tid) >> 3;
                o0[u] = *reinterpret_cast<const float4*>(sOut + rr * 272 + 16 * c8);
                o1[u] = *reinterpret_cast<const float4*>(sOut + rr * 272 + 128 + 16 * c8);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t row = row0 + ((u * kThreads + tid) >> 3);
                if (row < p.M) {
                    uint4 w;
                    w.x = epi::cvt2(o0[u].x, o0[u].y); w.y = epi::cvt2(o0[u].z, o0[u].w);
                    w.z = epi::cvt2(o1[u].x, o1[u].y); w.w = epi::cvt2(o1[u].z, o1[u].w);
                    *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.C) + row * p.ldc + 8 * c8) = w;
                }
            }
        } else {
            const int c4 = tid & 15;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int rr = (u * kThreads + tid) >> 4;
                const int64_t row = row0 + rr;
                float4 r = *reinterpret_cast<const float4*>(sOut + rr * 272 + 16 * c4);
                if (row >= p.M) continue;
                if (use_old) { r.x += pre[u].x; r.y += pre[u].y; r.z += pre[u].z; r.w += pre[u].w; }
                *(reinterpret_cast<float4*>(static_cast<float*>(p.C) + row * p.ldc) + c4) = r;
            }
        }
        __syncthreads();                                      // staging reused by the next tile
    }
    cp_async_wait<0>();
    if (it >= 1) {
        // every tile's MMAs were waited for by its epilogue: the accumulators are final
        tc::tc_fence_after();
        for (int g = half; g < nacc; g += 2) {
            const int n1 = 128 * g + wq * 32 + lane;
            uint32_t v[32];
            for (int c0 = 0; c0 < 64; c0 += 32) {
                tc::tmem_ld32(tmem + lane_base + 80 * g + c0, v);
                tc::tmem_ld_wait();
                if (n1 < p.N1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) atomicAdd(p.dW + (size_t)n1 * 64 + c0 + j, __uint_as_float(v[j]));
                }
            }
            uint32_t b16[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(b16[0]), "=r"(b16[1]), "=r"(b16[2]), "=r"(b16[3]), "=r"(b16[4]), "=r"(b16[5]), "=r"(b16[6]),
                           "=r"(b16[7]), "=r"(b16[8]), "=r"(b16[9]), "=r"(b16[10]), "=r"(b16[11]), "=r"(b16[12]), "=r"(b16[13]),
                           "=r"(b16[14]), "=r"(b16[15])
                         : "r"(tmem + lane_base + 80 * g + 64)
                         : "memory");
            tc::tmem_ld_wait();
            if (p.db && n1 < p.N1) atomicAdd(p.db + n1, __uint_as_float(b16[0]));
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

}  // namespace

// (Round 1 also had a warp-specialised persistent version of the rows kernel - 8 loader warps, 4 epilogue warps, 1 MMA warp per SM.
// Measured on B200 over 40 projection launches of a 65 536-node step: 9.95 ms against 8.67 ms for this phase-serial kernel
// (bf16 K = 192 with beta: 637 us per launch against ~220 us): 128 epilogue threads and 64 KB of loads in flight per SM are not
// enough; three phase-serial CTAs of 256 threads per SM overlap better.  Removed in round 2.)
extern "C" int u2gnn_gemm_tc_rows_ex(const void* A, int a_bf16, int64_t M, int K, int64_t lda, const float* W, int w_kn, int N,
                                     const float* bias, float beta, void* C, int c_bf16, int64_t ldc, u2gnn_stream_t stream) {
    if (!A || !W || !C || M < 0 || K < 1 || N < 1 || lda < K || ldc < N) return U2GNN_EINVAL;
    if (c_bf16 && beta != 0.0f) return U2GNN_EINVAL;
    if (K > 256 || N > 256) return U2GNN_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(C)) % 16) return U2GNN_EALIGN;
    if (M == 0) return U2GNN_OK;
    RowsParams p;
    p.A = A; p.a_bf16 = a_bf16; p.c_bf16 = c_bf16; p.M = M; p.lda = lda; p.K = K; p.KP = (K + 63) / 64 * 64;
    p.W = W; p.w_kn = w_kn; p.N = N; p.NP = (N + 15) / 16 * 16;
    p.bias = bias; p.C = C; p.ldc = ldc; p.beta = beta;
    p.res = nullptr; p.res_idx = nullptr; p.res_rows = 0;
    const int kt = p.KP / 64;
    const size_t smem = 1024 + (size_t)kt * 16384 + (size_t)kt * p.NP * 128 + 128 * 272;
    if (smem > 227 * 1024) return U2GNN_EUNSUPPORTED;
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    const int tmem_cols = (p.NP <= 64) ? 64 : (p.NP <= 128 ? 128 : 256);
    if (per_sm > 512 / tmem_cols) per_sm = 512 / tmem_cols;   // tensor-memory columns per CTA
    if (per_sm > 3) per_sm = 3;                              // __launch_bounds__(256, 3)
    if (per_sm < 1) per_sm = 1;
    const int64_t n_tiles = (M + TM - 1) / TM;
    const int64_t cap = (int64_t)U2GNN_NUM_SMS * per_sm;
    const int grid = (int)(n_tiles < cap ? n_tiles : cap);
    if (per_sm >= 3) {
        cudaFuncSetAttribute(gemm_tc_rows_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        gemm_tc_rows_kernel<false, 3><<<grid, kThreads, smem, as_stream(stream)>>>(p);
    } else {
        cudaFuncSetAttribute(gemm_tc_rows_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        gemm_tc_rows_kernel<false, 2><<<grid, kThreads, smem, as_stream(stream)>>>(p);
    }
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_gemm_tc_rows_ln(const void* A, int a_bf16, int64_t M, int K, int64_t lda, const float* W, int w_kn,
                                     const float* bias, const float* res, int64_t ldres, const int64_t* res_idx, int64_t res_rows,
                                     uint64_t seed, uint32_t rng_stream,
                                     int thr, const float* gamma, const float* beta, float* z, float* y, float* stats,
                                     void* y_img, u2gnn_stream_t stream) {
    constexpr int N = 64;
    if (!A || !W || !bias || !res || !gamma || !beta || !z || !y || !stats || M < 0 || K < 1 || lda < K || ldres < N ||
        thr < 0 || thr > 255)
        return U2GNN_EINVAL;
    if (K > 256) return U2GNN_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(z) |
         reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) |
         reinterpret_cast<uintptr_t>(bias)) % 16 || reinterpret_cast<uintptr_t>(stats) % 8 || (ldres & 3))
        return U2GNN_EALIGN;
    if (M == 0) return U2GNN_OK;
    RowsParams p;
    p.A = A; p.a_bf16 = a_bf16; p.c_bf16 = 0; p.M = M; p.lda = lda; p.K = K; p.KP = (K + 63) / 64 * 64;
    p.W = W; p.w_kn = w_kn; p.N = N; p.NP = N;
    p.bias = bias; p.C = nullptr; p.ldc = N; p.beta = 0.0f;
    p.res = res; p.ldres = ldres; p.res_idx = res_idx; p.res_rows = res_rows;
    p.keys = rng_keys(seed, rng_stream); p.thr = thr; p.low = rng_thr_low(thr);
    p.scale = thr ? rng_keep_scale(thr) : 1.0f;
    p.gamma = gamma; p.ln_beta = beta; p.z = z; p.y = y; p.stats = stats;
    p.y_img = static_cast<uint8_t*>(y_img);
    const int kt = p.KP / 64;
    const size_t smem = 1024 + (size_t)kt * 16384 + (size_t)kt * p.NP * 128 + 128 * 272;
    if (smem > 227 * 1024) return U2GNN_EUNSUPPORTED;
    cudaFuncSetAttribute(gemm_tc_rows_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm > 3) per_sm = 3;                              // __launch_bounds__(256, 3)
    if (per_sm < 1) per_sm = 1;
    const int64_t n_tiles = (M + TM - 1) / TM;
    const int64_t cap = (int64_t)U2GNN_NUM_SMS * per_sm;
    gemm_tc_rows_kernel<true, 3><<<(int)(n_tiles < cap ? n_tiles : cap), kThreads, smem, as_stream(stream)>>>(p);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_gemm_tc_rows(const float* A, int64_t M, int K, int64_t lda, const float* W, int w_kn, int N,
                                  const float* bias, float beta, float* C, int64_t ldc, u2gnn_stream_t stream) {
    return u2gnn_gemm_tc_rows_ex(A, 0, M, K, lda, W, w_kn, N, bias, beta, C, 0, ldc, stream);
}

extern "C" int u2gnn_gemm_tc_wgrad_ex(const void* A, int a_bf16, int64_t M, int N1, int64_t lda, const void* B, int b_bf16, int N2,
                                      int64_t ldb, const int64_t* b_idx, int64_t b_rows, float* dW, float* db, u2gnn_stream_t stream) {
    if (!A || !B || !dW || M < 0 || N1 < 1 || N2 < 1 || lda < N1 || ldb < N2) return U2GNN_EINVAL;
    if (b_idx && (b_bf16 || N2 != 64 || (ldb & 3) || b_rows < 1)) return U2GNN_EUNSUPPORTED;
    if (N1 > 256 || N2 > 64) return U2GNN_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) % 16) return U2GNN_EALIGN;
    if (M == 0) return U2GNN_OK;
    WgradParams p;
    p.A = A; p.B = B; p.a_bf16 = a_bf16; p.b_bf16 = b_bf16; p.M = M; p.lda = lda; p.ldb = ldb; p.N1 = N1; p.N2 = N2; p.dW = dW; p.db = db;
    p.b_idx = b_idx; p.b_rows = b_rows;
    const int ga = (N1 + 63) / 64;
    const size_t smem = 1024 + (size_t)2 * (ga + 1) * 16384 + 2 * 16384;
    if (smem > 227 * 1024) return U2GNN_EUNSUPPORTED;
    cudaFuncSetAttribute(gemm_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t n_tiles = (M + TM - 1) / TM;
    const int64_t cap = (int64_t)U2GNN_NUM_SMS * ((smem + 1024 <= 113 * 1024) ? 2 : 1);   // narrow A: two CTAs per SM overlap load and MMA phases
    gemm_tc_wgrad_kernel<<<(int)(n_tiles < cap ? n_tiles : cap), kThreads, smem, as_stream(stream)>>>(p);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_gemm_tc_wgrad(const float* A, int64_t M, int N1, int64_t lda, const float* B, int N2, int64_t ldb,
                                   float* dW, float* db, u2gnn_stream_t stream) {
    return u2gnn_gemm_tc_wgrad_ex(A, 0, M, N1, lda, B, 0, N2, ldb, nullptr, 0, dW, db, stream);
}

extern "C" int u2gnn_gemm_tc_dgrad_wgrad(const void* A, int64_t M, int N1, int64_t lda, const void* B, int b_bf16, int64_t ldb,
                                         const int64_t* b_idx, int64_t b_rows, const float* W, void* C, int c_bf16, int64_t ldc,
                                         float beta, float* dW, float* db, u2gnn_stream_t stream) {
    if (!A || !B || !W || !C || !dW || M < 0 || N1 < 64 || lda < N1 || ldb < 64 || ldc < 64) return U2GNN_EINVAL;
    if (b_idx && (b_bf16 || b_rows < 1)) return U2GNN_EUNSUPPORTED;
    if (c_bf16 && beta != 0.0f) return U2GNN_EINVAL;
    if ((N1 & 63) || N1 > 256) return U2GNN_EUNSUPPORTED;
    if ((lda & 7) || (ldc & 7) || (ldb & (b_bf16 ? 7 : 3))) return U2GNN_EALIGN;
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) % 16) return U2GNN_EALIGN;
    if (M == 0) return U2GNN_OK;
    BwdParams p;
    p.A = A; p.B = B; p.b_bf16 = b_bf16; p.c_bf16 = c_bf16; p.M = M; p.lda = lda; p.ldb = ldb; p.ldc = ldc; p.N1 = N1;
    p.W = W; p.C = C; p.beta = beta; p.dW = dW; p.db = db; p.b_idx = b_idx; p.b_rows = b_rows;
    const int ga = N1 / 64;
    const size_t fixed = 1024 + (size_t)ga * 8192 + 128 * 272 + 2 * 16384;
    const size_t stage = (size_t)(ga + 1) * 16384;
    const int64_t n_tiles = (M + TM - 1) / TM;
    const int grid = (int)(n_tiles < U2GNN_NUM_SMS ? n_tiles : U2GNN_NUM_SMS);
    auto launch = [&](auto kern, int ns) -> int {
        const size_t smem = fixed + ns * stage;
        if (smem > 227 * 1024) return U2GNN_EUNSUPPORTED;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, kThreads, smem, as_stream(stream)>>>(p);
        return U2GNN_OK;
    };
    int rc;
    const size_t cap = 227 * 1024 - 1024;                    // dynamic limit minus room for the static barriers
    if (fixed + 4 * stage <= cap) rc = launch(gemm_tc_dgrad_wgrad_kernel<4>, 4);
    else if (fixed + 3 * stage <= cap) rc = launch(gemm_tc_dgrad_wgrad_kernel<3>, 3);
    else rc = launch(gemm_tc_dgrad_wgrad_kernel<2>, 2);
    if (rc != U2GNN_OK) return rc;
    U2GNN_CHECK_LAUNCH();
}

// (Round 1 also had a variant of this kernel with the LayerNorm1 backward evaluated while the operand tile is staged, so that the
// dropout-masked gradient was never stored.  Parity-green but slower as measured on B200 - 66.06 against 65.48 ms per step: one
// CTA of 256 threads per SM doing the LayerNorm arithmetic in its load phase is less efficient than the 24-warp LayerNorm pass
// followed by this cp.async-fed kernel.  Removed in round 2.)
