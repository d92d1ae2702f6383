// Short-sequence self-attention, thread-per-row formulation (attn_axis="neighbors", every query row live).
//
// seqattn.cu's warp-per-node kernels are latency-bound (one accumulator chain per lane, 17 of 32 lanes busy).
// Here a CTA stages K and V (and, in the backward, Q and the incoming gradient) of a few whole node sequences in
// shared memory and every thread owns one query row: the q.k and p.v products are 64-wide register FMAs fed by
// broadcast LDS.128, softmax runs inside the thread, and the key-side gradients (dk, dv) are a second phase with
// one thread per key row reading the per-node ds / p matrices from shared memory.  Same arithmetic and dropout
// stream as seqattn.cu (nn.MultiheadAttention, one head); fp32 throughout.
#include "common.cuh"
#include "rng.cuh"

namespace {

struct AttnRng {
    RngKeys keys;
    int thr;
    float scale;
};

constexpr int PP = 33;   // pitch of the per-row score / probability scratch (S <= 32)

template <int D>
__device__ __forceinline__ float dot_row(const float (&a)[D], const float* __restrict__ b) {
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll
    for (int c = 0; c < D; c += 4) {
        const float4 v = *reinterpret_cast<const float4*>(b + c);
        acc0 = fmaf(a[c], v.x, acc0);
        acc1 = fmaf(a[c + 1], v.y, acc1);
        acc2 = fmaf(a[c + 2], v.z, acc2);
        acc3 = fmaf(a[c + 3], v.w, acc3);
    }
    return (acc0 + acc1) + (acc2 + acc3);
}

template <int D>
__device__ __forceinline__ void axpy_row(float (&acc)[D], float s, const float* __restrict__ b) {
#pragma unroll
    for (int c = 0; c < D; c += 4) {
        const float4 v = *reinterpret_cast<const float4*>(b + c);
        acc[c] = fmaf(s, v.x, acc[c]);
        acc[c + 1] = fmaf(s, v.y, acc[c + 1]);
        acc[c + 2] = fmaf(s, v.z, acc[c + 2]);
        acc[c + 3] = fmaf(s, v.w, acc[c + 3]);
    }
}

// ---- element-type helpers for the last-timestep kernels: qkv / dqkv either fp32 or bf16 (BF) row-major
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
// four consecutive elements starting at element offset `off` (multiple of 4)
template <bool BF>
__device__ __forceinline__ float4 ld4(const void* __restrict__ base, int64_t off) {
    if (BF) {
        const uint2 w = __ldg(reinterpret_cast<const uint2*>(static_cast<const uint16_t*>(base) + off));
        return make_float4(bf_lo(w.x), bf_hi(w.x), bf_lo(w.y), bf_hi(w.y));
    }
    return __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(base) + off));
}
template <bool BF>
__device__ __forceinline__ void st4(void* __restrict__ base, int64_t off, float4 v) {
    if (BF) *reinterpret_cast<uint2*>(static_cast<uint16_t*>(base) + off) = make_uint2(pack_bf2(v.x, v.y), pack_bf2(v.z, v.w));
    else *reinterpret_cast<float4*>(static_cast<float*>(base) + off) = v;
}
template <int D, bool BF>
__device__ __forceinline__ float dot_row_t(const float (&a)[D], const void* __restrict__ base, int64_t off) {
    if (!BF) return dot_row<D>(a, static_cast<const float*>(base) + off);
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll
    for (int c = 0; c < D; c += 8) {
        const uint4 w = *reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(base) + off + c);
        acc0 = fmaf(a[c], bf_lo(w.x), acc0);     acc1 = fmaf(a[c + 1], bf_hi(w.x), acc1);
        acc2 = fmaf(a[c + 2], bf_lo(w.y), acc2); acc3 = fmaf(a[c + 3], bf_hi(w.y), acc3);
        acc0 = fmaf(a[c + 4], bf_lo(w.z), acc0); acc1 = fmaf(a[c + 5], bf_hi(w.z), acc1);
        acc2 = fmaf(a[c + 6], bf_lo(w.w), acc2); acc3 = fmaf(a[c + 7], bf_hi(w.w), acc3);
    }
    return (acc0 + acc1) + (acc2 + acc3);
}
template <int D, bool BF>
__device__ __forceinline__ void axpy_row_t(float (&acc)[D], float s, const void* __restrict__ base, int64_t off) {
    if (!BF) {
        axpy_row<D>(acc, s, static_cast<const float*>(base) + off);
        return;
    }
#pragma unroll
    for (int c = 0; c < D; c += 8) {
        const uint4 w = *reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(base) + off + c);
        acc[c] = fmaf(s, bf_lo(w.x), acc[c]);         acc[c + 1] = fmaf(s, bf_hi(w.x), acc[c + 1]);
        acc[c + 2] = fmaf(s, bf_lo(w.y), acc[c + 2]); acc[c + 3] = fmaf(s, bf_hi(w.y), acc[c + 3]);
        acc[c + 4] = fmaf(s, bf_lo(w.z), acc[c + 4]); acc[c + 5] = fmaf(s, bf_hi(w.z), acc[c + 5]);
        acc[c + 6] = fmaf(s, bf_lo(w.w), acc[c + 6]); acc[c + 7] = fmaf(s, bf_hi(w.w), acc[c + 7]);
    }
}

// cooperative copy of `rows` rows of width D (fp32) from a strided global matrix into padded shared memory
template <int D, int NT>
__device__ __forceinline__ void stage_rows(float* dst, const float* __restrict__ src, int64_t src_ld, int rows, int tid) {
    constexpr int P = D + 4;
    for (int e = tid; e < rows * (D / 4); e += NT) {
        const int r = e / (D / 4), c4 = e % (D / 4);
        *reinterpret_cast<float4*>(dst + r * P + 4 * c4) = __ldg(reinterpret_cast<const float4*>(src + r * src_ld) + c4);
    }
}

template <int D, int NT>
__global__ void __launch_bounds__(NT) seqattn_rows_fwd_kernel(const float* __restrict__ qkv, int64_t B, int S, AttnRng rng,
                                                              float* __restrict__ ctx, int NB) {
    extern __shared__ __align__(16) float sm[];
    constexpr int P = D + 4;
    const int max_rows = NB * S;
    float* Ks = sm;
    float* Vs = Ks + max_rows * P;
    float* Ps = Vs + max_rows * P;          // [max_rows][PP]
    const int tid = threadIdx.x;
    const float qscale = sqrtf(1.0f / (float)D);
    const int64_t n_tiles = (B + NB - 1) / NB;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t node0 = tile * NB;
        const int nodes = (int)((B - node0 < NB) ? B - node0 : NB);
        const int rows = nodes * S;
        const float* base = qkv + node0 * S * 3 * D;
        __syncthreads();
        stage_rows<D, NT>(Ks, base + D, 3 * D, rows, tid);
        stage_rows<D, NT>(Vs, base + 2 * D, 3 * D, rows, tid);
        float q[D];
        if (tid < rows) {
#pragma unroll
            for (int c = 0; c < D; c += 4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(base + (int64_t)tid * 3 * D + c));
                q[c] = v.x * qscale; q[c + 1] = v.y * qscale; q[c + 2] = v.z * qscale; q[c + 3] = v.w * qscale;
            }
        }
        __syncthreads();
        if (tid < rows) {
            const int node = tid / S, i = tid - node * S;
            const float* kn = Ks + node * S * P;
            const float* vn = Vs + node * S * P;
            float* pr = Ps + tid * PP;
            float m = -INFINITY;
            for (int j = 0; j < S; ++j) {
                const float s = dot_row<D>(q, kn + j * P);
                pr[j] = s;
                m = fmaxf(m, s);
            }
            float sum = 0.f;
            for (int j = 0; j < S; ++j) {
                const float e = expf(pr[j] - m);
                pr[j] = e;
                sum += e;
            }
            const float inv = 1.0f / sum;
            const uint64_t ebase = (uint64_t)((node0 + node) * S + i) * (uint64_t)S;
            float acc[D];
#pragma unroll
            for (int c = 0; c < D; ++c) acc[c] = 0.f;
            for (int j = 0; j < S; ++j) {
                const float pd = (pr[j] * inv) * rng_dropout_mult(rng.keys, ebase + (uint64_t)j, rng.thr, rng.scale);
                axpy_row<D>(acc, pd, vn + j * P);
            }
            float4* out = reinterpret_cast<float4*>(ctx + ((node0 * S) + tid) * (int64_t)D);
#pragma unroll
            for (int c = 0; c < D; c += 4) out[c >> 2] = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
        }
    }
}

template <int D, int NT>
__global__ void __launch_bounds__(NT) seqattn_rows_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ dctx,
                                                              int64_t B, int S, AttnRng rng, float* __restrict__ dqkv,
                                                              int NB) {
    extern __shared__ __align__(16) float sm[];
    constexpr int P = D + 4;
    const int max_rows = NB * S;
    float* Qs = sm;
    float* Ks = Qs + max_rows * P;
    float* Vs = Ks + max_rows * P;
    float* Gs = Vs + max_rows * P;
    float* Ps = Gs + max_rows * P;          // probabilities, then dropped probabilities  [max_rows][PP]
    float* Ds = Ps + max_rows * PP;         // dP, then dS                                [max_rows][PP]
    const int tid = threadIdx.x;
    const float qscale = sqrtf(1.0f / (float)D);
    const int64_t n_tiles = (B + NB - 1) / NB;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t node0 = tile * NB;
        const int nodes = (int)((B - node0 < NB) ? B - node0 : NB);
        const int rows = nodes * S;
        const float* base = qkv + node0 * S * 3 * D;
        __syncthreads();
        stage_rows<D, NT>(Qs, base, 3 * D, rows, tid);
        stage_rows<D, NT>(Ks, base + D, 3 * D, rows, tid);
        stage_rows<D, NT>(Vs, base + 2 * D, 3 * D, rows, tid);
        stage_rows<D, NT>(Gs, dctx + node0 * S * D, D, rows, tid);
        __syncthreads();
        float* gout = dqkv + (node0 * S + tid) * (int64_t)(3 * D);
        const int node = tid / S, i = tid - node * S;
        // ---- phase A: thread = query row.  p, dp, ds and dq
        if (tid < rows) {
            const float* kn = Ks + node * S * P;
            const float* vn = Vs + node * S * P;
            float* pr = Ps + tid * PP;
            float* dr = Ds + tid * PP;
            float a[D];
#pragma unroll
            for (int c = 0; c < D; ++c) a[c] = Qs[tid * P + c] * qscale;
            float m = -INFINITY;
            for (int j = 0; j < S; ++j) {
                const float s = dot_row<D>(a, kn + j * P);
                pr[j] = s;
                m = fmaxf(m, s);
            }
            float sum = 0.f;
            for (int j = 0; j < S; ++j) {
                const float e = expf(pr[j] - m);
                pr[j] = e;
                sum += e;
            }
            const float inv = 1.0f / sum;
#pragma unroll
            for (int c = 0; c < D; ++c) a[c] = Gs[tid * P + c];          // a <- incoming gradient row
            const uint64_t ebase = (uint64_t)((node0 + node) * S + i) * (uint64_t)S;
            float tsum = 0.f;
            for (int j = 0; j < S; ++j) {
                const float mult = rng_dropout_mult(rng.keys, ebase + (uint64_t)j, rng.thr, rng.scale);
                const float p = pr[j] * inv;
                const float dp = dot_row<D>(a, vn + j * P) * mult;
                dr[j] = dp;
                tsum = fmaf(p, dp, tsum);
                pr[j] = p;
            }
#pragma unroll
            for (int c = 0; c < D; ++c) a[c] = 0.f;                       // a <- dq accumulator
            for (int j = 0; j < S; ++j) {
                const float p = pr[j];
                const float ds = p * (dr[j] - tsum);
                dr[j] = ds;
                pr[j] = p * rng_dropout_mult(rng.keys, ebase + (uint64_t)j, rng.thr, rng.scale);
                axpy_row<D>(a, ds, kn + j * P);
            }
            float4* o = reinterpret_cast<float4*>(gout);
#pragma unroll
            for (int c = 0; c < D; c += 4)
                o[c >> 2] = make_float4(a[c] * qscale, a[c + 1] * qscale, a[c + 2] * qscale, a[c + 3] * qscale);
        }
        __syncthreads();
        // ---- phase B: thread = key row j of its node.  dk_j = scale * sum_i ds_ij q_i ; dv_j = sum_i p~_ij g_i
        if (tid < rows) {
            const int j = i;
            float dk[D];
#pragma unroll
            for (int c = 0; c < D; ++c) dk[c] = 0.f;
            for (int ii = 0; ii < S; ++ii) axpy_row<D>(dk, Ds[(node * S + ii) * PP + j], Qs + (node * S + ii) * P);
            float4* o = reinterpret_cast<float4*>(gout + D);
#pragma unroll
            for (int c = 0; c < D; c += 4)
                o[c >> 2] = make_float4(dk[c] * qscale, dk[c + 1] * qscale, dk[c + 2] * qscale, dk[c + 3] * qscale);
#pragma unroll
            for (int c = 0; c < D; ++c) dk[c] = 0.f;                      // reuse as dv
            for (int ii = 0; ii < S; ++ii) axpy_row<D>(dk, Ps[(node * S + ii) * PP + j], Gs + (node * S + ii) * P);
            o = reinterpret_cast<float4*>(gout + 2 * D);
#pragma unroll
            for (int c = 0; c < D; c += 4) o[c >> 2] = make_float4(dk[c], dk[c + 1], dk[c + 2], dk[c + 3]);
        }
    }
}


// ---- last timestep of a U2GNN layer (dead-row elimination): only query position 0 of every node is live.
//      One thread per node; K / V rows are read straight from global memory (each thread streams its node's
//      contiguous [S, 3D] block), scores live in a per-thread shared-memory scratch row.
template <int D, int NT, bool BF>      // BF: qkv stored as bf16 (rounded once by the projection that produced it)
__global__ void __launch_bounds__(NT) seqattn_last_fwd_kernel(const void* __restrict__ qkv, int64_t B, int S, AttnRng rng,
                                                              float* __restrict__ ctx) {
    __shared__ float Ps[NT * PP];
    float* pr = Ps + threadIdx.x * PP;
    const float qscale = sqrtf(1.0f / (float)D);
    for (int64_t b = (int64_t)blockIdx.x * NT + threadIdx.x; b < B; b += (int64_t)gridDim.x * NT) {
        const int64_t base = b * S * 3 * D;                    // element offset of the node's [S, 3D] block
        float a[D];
#pragma unroll
        for (int c = 0; c < D; c += 4) {
            const float4 v = ld4<BF>(qkv, base + c);
            a[c] = v.x * qscale; a[c + 1] = v.y * qscale; a[c + 2] = v.z * qscale; a[c + 3] = v.w * qscale;
        }
        float m = -INFINITY;
        for (int j = 0; j < S; ++j) {
            const float s = dot_row_t<D, BF>(a, qkv, base + (int64_t)j * 3 * D + D);
            pr[j] = s;
            m = fmaxf(m, s);
        }
        float sum = 0.f;
        for (int j = 0; j < S; ++j) {
            const float e = expf(pr[j] - m);
            pr[j] = e;
            sum += e;
        }
        const float inv = 1.0f / sum;
#pragma unroll
        for (int c = 0; c < D; ++c) a[c] = 0.f;
        for (int j = 0; j < S; ++j) {
            const float pd = (pr[j] * inv) * rng_dropout_mult(rng.keys, (uint64_t)b * (uint64_t)S + (uint64_t)j, rng.thr, rng.scale);
            axpy_row_t<D, BF>(a, pd, qkv, base + (int64_t)j * 3 * D + 2 * D);
        }
        float4* out = reinterpret_cast<float4*>(ctx + b * D);
#pragma unroll
        for (int c = 0; c < D; c += 4) out[c >> 2] = make_float4(a[c], a[c + 1], a[c + 2], a[c + 3]);
    }
}

// Backward of the dead-row-eliminated last timestep (only position 0 of every node has a query).  Half a warp per node,
// lane = four feature columns: every K / V row is ONE coalesced 256-byte read per half-warp and every dK / dV / dQ row
// one coalesced 256-byte write (the thread-per-node version wrote each node's 13 KB from a single thread and reached
// 14 % of the HBM roofline).  Scores and dP are reduced across the 16 lanes with shuffles and kept in registers.
template <int D, int NT, bool BF>      // BF: qkv AND dqkv stored as bf16 (their other producers / consumers are tensor-core GEMMs)
__global__ void __launch_bounds__(NT) seqattn_last_bwd_kernel(const void* __restrict__ qkv, const float* __restrict__ dctx,
                                                              int64_t B, int S, AttnRng rng, int low, void* __restrict__ dqkv) {
    static_assert(D == 64 || D == 32 || D == 16 || D == 8 || D == 4, "lane = 4 feature columns");
    constexpr int LPN = D / 4;                                        // lanes per node (16: half a warp, 8: a quarter)
    constexpr int NPW = 32 / LPN;                                     // nodes per warp
    const int hl = threadIdx.x % LPN;
    const int64_t hw = ((int64_t)blockIdx.x * NT + threadIdx.x) / LPN, n_hw = ((int64_t)gridDim.x * NT) / LPN;
    const float qscale = sqrtf(1.0f / (float)D);
    auto reduce16 = [](float v) {
#pragma unroll
        for (int o = LPN / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    };
    const int64_t B_pad = (B + NPW - 1) / NPW * NPW;                  // the node groups of a warp iterate together (shuffles)
    for (int64_t b0 = hw; b0 < B_pad; b0 += n_hw) {
        const bool live = b0 < B;
        const int64_t b = live ? b0 : B - 1;
        const int64_t base = b * S * 3 * D;                    // element offset of the node's [S, 3D] block (qkv and dqkv)
        float4 q = ld4<BF>(qkv, base + 4 * hl);
        q.x *= qscale; q.y *= qscale; q.z *= qscale; q.w *= qscale;
        const float4 g = __ldg(reinterpret_cast<const float4*>(dctx + b * D) + hl);
        float pr[32], dr[32];
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            pr[j] = -INFINITY;
            dr[j] = 0.f;
            if (j < S) {
                const float4 k = ld4<BF>(qkv, base + (int64_t)j * 3 * D + D + 4 * hl);
                const float4 v = ld4<BF>(qkv, base + (int64_t)j * 3 * D + 2 * D + 4 * hl);
                pr[j] = reduce16(q.x * k.x + q.y * k.y + q.z * k.z + q.w * k.w);
                dr[j] = reduce16(g.x * v.x + g.y * v.y + g.z * v.z + g.w * v.w);
                m = fmaxf(m, pr[j]);
            }
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            pr[j] = (j < S) ? expf(pr[j] - m) : 0.f;
            sum += pr[j];
        }
        const float inv = 1.0f / sum;
        // keep bits of elements b*S .. b*S + S - 1 of the probability tensor: at most two words of the stream
        const uint64_t ebase = (uint64_t)b * (uint64_t)S;
        uint32_t keep_bits = 0xFFFFFFFFu;
        if (rng.thr) {
            const uint32_t w0 = rng_keep_word_lo(rng.keys, ebase >> 5, rng.thr, low);
            const uint32_t w1 = rng_keep_word_lo(rng.keys, (ebase >> 5) + 1, rng.thr, low);
            const uint32_t sh = (uint32_t)(ebase & 31);
            keep_bits = sh ? ((w0 >> sh) | (w1 << (32 - sh))) : w0;
        }
        const float dscale = rng.thr ? rng.scale : 1.0f;
        float tsum = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float mult = ((keep_bits >> j) & 1u) ? dscale : 0.0f;
            pr[j] *= inv;
            dr[j] *= mult;
            tsum = fmaf(pr[j], dr[j], tsum);
        }
        float4 dq = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (j < S) {
                const float ds = pr[j] * (dr[j] - tsum);
                const float pd = pr[j] * (((keep_bits >> j) & 1u) ? dscale : 0.0f);
                const float4 k = ld4<BF>(qkv, base + (int64_t)j * 3 * D + D + 4 * hl);
                dq.x = fmaf(ds, k.x, dq.x); dq.y = fmaf(ds, k.y, dq.y); dq.z = fmaf(ds, k.z, dq.z); dq.w = fmaf(ds, k.w, dq.w);
                if (live) {
                    const int64_t row = base + (int64_t)j * 3 * D + 4 * hl;
                    if (j > 0) st4<BF>(dqkv, row, make_float4(0.f, 0.f, 0.f, 0.f));                        // rows without a query
                    st4<BF>(dqkv, row + D, make_float4(ds * q.x, ds * q.y, ds * q.z, ds * q.w));          // q already carries sqrt(1/d)
                    st4<BF>(dqkv, row + 2 * D, make_float4(pd * g.x, pd * g.y, pd * g.z, pd * g.w));
                }
            }
        }
        if (live) st4<BF>(dqkv, base + 4 * hl, make_float4(dq.x * qscale, dq.y * qscale, dq.z * qscale, dq.w * qscale));
    }
}

template <int D, bool BF = false>
int launch_last_fwd(const void* qkv, int64_t B, int S, AttnRng rng, float* ctx, cudaStream_t st) {
    constexpr int NT = 128;
    const int64_t blocks = (B + NT - 1) / NT;
    const int grid = (int)(blocks < (int64_t)U2GNN_NUM_SMS * 8 ? blocks : (int64_t)U2GNN_NUM_SMS * 8);
    seqattn_last_fwd_kernel<D, NT, BF><<<grid, NT, 0, st>>>(qkv, B, S, rng, ctx);
    return 1;
}
template <int D, bool BF = false>
int launch_last_bwd(const void* qkv, const float* dctx, int64_t B, int S, AttnRng rng, void* dqkv, cudaStream_t st) {
    constexpr int NT = 256;
    const int64_t blocks = (B * (D / 4) + NT - 1) / NT;
    const int grid = (int)(blocks < (int64_t)U2GNN_NUM_SMS * 8 ? blocks : (int64_t)U2GNN_NUM_SMS * 8);
    seqattn_last_bwd_kernel<D, NT, BF><<<grid, NT, 0, st>>>(qkv, dctx, B, S, rng, rng_thr_low(rng.thr), dqkv);
    return 1;
}

template <int D>
int launch_fwd(const float* qkv, int64_t B, int S, AttnRng rng, float* ctx, cudaStream_t st) {
    constexpr int NT = 128;
    const int NB = NT / S;
    const size_t smem = ((size_t)2 * NB * S * (D + 4) + (size_t)NB * S * PP) * sizeof(float);
    auto k = seqattn_rows_fwd_kernel<D, NT>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t tiles = (B + NB - 1) / NB;
    const int grid = (int)(tiles < (int64_t)U2GNN_NUM_SMS * 2 ? tiles : (int64_t)U2GNN_NUM_SMS * 2);
    k<<<grid, NT, smem, st>>>(qkv, B, S, rng, ctx, NB);
    return 1;
}

template <int D>
int launch_bwd(const float* qkv, const float* dctx, int64_t B, int S, AttnRng rng, float* dqkv, cudaStream_t st) {
    constexpr int NT = 64;
    const int NB = NT / S;
    const size_t smem = ((size_t)4 * NB * S * (D + 4) + (size_t)2 * NB * S * PP) * sizeof(float);
    auto k = seqattn_rows_bwd_kernel<D, NT>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t tiles = (B + NB - 1) / NB;
    const int grid = (int)(tiles < (int64_t)U2GNN_NUM_SMS * 3 ? tiles : (int64_t)U2GNN_NUM_SMS * 3);
    k<<<grid, NT, smem, st>>>(qkv, dctx, B, S, rng, dqkv, NB);
    return 1;
}

AttnRng make_rng(uint64_t seed, uint32_t stream, int thr) {
    AttnRng r;
    r.keys = rng_keys(seed, stream);
    r.thr = thr;
    r.scale = thr ? rng_keep_scale(thr) : 1.0f;
    return r;
}

bool aligned16(const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; }

}  // namespace

// Internal dispatch used by u2gnn_seqattn_fwd / u2gnn_seqattn_bwd (seqattn.cu): returns 1 when the
// thread-per-row kernels handled the call, 0 when the caller must use the warp-per-node kernels.
int seqattn_rows_try_fwd(const float* qkv, int64_t B, int S, int Sq, int d, uint64_t seed, uint32_t rng_stream, int thr,
                         float* ctx, cudaStream_t st) {
    if (S < 2 || !aligned16(qkv) || !aligned16(ctx)) return 0;
    const AttnRng rng = make_rng(seed, rng_stream, thr);
    if (Sq == 1) {
        if (d == 64) return launch_last_fwd<64>(qkv, B, S, rng, ctx, st);
        if (d == 32) return launch_last_fwd<32>(qkv, B, S, rng, ctx, st);
        if (d == 16) return launch_last_fwd<16>(qkv, B, S, rng, ctx, st);
        if (d == 8) return launch_last_fwd<8>(qkv, B, S, rng, ctx, st);
        if (d == 4) return launch_last_fwd<4>(qkv, B, S, rng, ctx, st);
        return 0;
    }
    if (d == 64) return launch_fwd<64>(qkv, B, S, rng, ctx, st);
    if (d == 32) return launch_fwd<32>(qkv, B, S, rng, ctx, st);
    // small feature sizes of the unsupervised configurations (configs[1] / configs[3]: d = 4): a row is one float4
    if (d == 16) return launch_fwd<16>(qkv, B, S, rng, ctx, st);
    if (d == 8) return launch_fwd<8>(qkv, B, S, rng, ctx, st);
    if (d == 4) return launch_fwd<4>(qkv, B, S, rng, ctx, st);
    // padded feature sizes of engine.py's attention block for 64 < d <= 128 (configs[2], d = 65 -> 68): q | k | v blocks
    // zero-padded to a multiple of 4 floats so that every row is 16-byte aligned; the sqrt(1/d) scale is folded into W_q
    if (d == 68) return launch_fwd<68>(qkv, B, S, rng, ctx, st);
    if (d == 80) return launch_fwd<80>(qkv, B, S, rng, ctx, st);
    if (d == 96) return launch_fwd<96>(qkv, B, S, rng, ctx, st);
    if (d == 112) return launch_fwd<112>(qkv, B, S, rng, ctx, st);
    if (d == 128) return launch_fwd<128>(qkv, B, S, rng, ctx, st);
    return 0;
}

int seqattn_rows_try_bwd(const float* qkv, const float* dctx, int64_t B, int S, int Sq, int d, uint64_t seed,
                         uint32_t rng_stream, int thr, float* dqkv, cudaStream_t st) {
    if (S < 2 || !aligned16(qkv) || !aligned16(dctx) || !aligned16(dqkv)) return 0;
    const AttnRng rng = make_rng(seed, rng_stream, thr);
    if (Sq == 1) {
        if (d == 64) return launch_last_bwd<64>(qkv, dctx, B, S, rng, dqkv, st);
        if (d == 32) return launch_last_bwd<32>(qkv, dctx, B, S, rng, dqkv, st);
        if (d == 16) return launch_last_bwd<16>(qkv, dctx, B, S, rng, dqkv, st);
        if (d == 8) return launch_last_bwd<8>(qkv, dctx, B, S, rng, dqkv, st);
        if (d == 4) return launch_last_bwd<4>(qkv, dctx, B, S, rng, dqkv, st);
        return 0;
    }
    if (d == 64) return launch_bwd<64>(qkv, dctx, B, S, rng, dqkv, st);
    if (d == 32) return launch_bwd<32>(qkv, dctx, B, S, rng, dqkv, st);
    if (d == 16) return launch_bwd<16>(qkv, dctx, B, S, rng, dqkv, st);
    if (d == 8) return launch_bwd<8>(qkv, dctx, B, S, rng, dqkv, st);
    if (d == 4) return launch_bwd<4>(qkv, dctx, B, S, rng, dqkv, st);
    if (d == 68) return launch_bwd<68>(qkv, dctx, B, S, rng, dqkv, st);
    if (d == 80) return launch_bwd<80>(qkv, dctx, B, S, rng, dqkv, st);
    if (d == 96) return launch_bwd<96>(qkv, dctx, B, S, rng, dqkv, st);
    if (d == 112) return launch_bwd<112>(qkv, dctx, B, S, rng, dqkv, st);
    if (d == 128) return launch_bwd<128>(qkv, dctx, B, S, rng, dqkv, st);
    return 0;
}

// Last-timestep attention (only query position 0 of every node live) with qkv / dqkv optionally stored as bf16: in the bf16
// mode their producer (in_proj GEMM) and consumers (in_proj weight / input gradient GEMMs) are tensor-core kernels, so the
// [N*S, 3d] tensors cross HBM at half the bytes.  ctx / dctx ([N, d], one row per node) stay fp32.
extern "C" int u2gnn_seqattn_last_fwd_ex(const void* qkv, int qkv_bf16, int64_t B, int S, int d, uint64_t seed, uint32_t rng_stream,
                                         int thr, float* ctx, u2gnn_stream_t stream) {
    if (!qkv || !ctx || B < 0 || S < 2 || S > 32 || thr < 0 || thr > 255) return U2GNN_EINVAL;
    if (d != 64 && d != 32) return U2GNN_EUNSUPPORTED;
    if (!aligned16(qkv) || !aligned16(ctx)) return U2GNN_EALIGN;
    if (B == 0) return U2GNN_OK;
    const AttnRng rng = make_rng(seed, rng_stream, thr);
    cudaStream_t st = as_stream(stream);
    if (d == 64) qkv_bf16 ? launch_last_fwd<64, true>(qkv, B, S, rng, ctx, st) : launch_last_fwd<64, false>(qkv, B, S, rng, ctx, st);
    else qkv_bf16 ? launch_last_fwd<32, true>(qkv, B, S, rng, ctx, st) : launch_last_fwd<32, false>(qkv, B, S, rng, ctx, st);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_seqattn_last_bwd_ex(const void* qkv, const float* dctx, int io_bf16, int64_t B, int S, int d, uint64_t seed,
                                         uint32_t rng_stream, int thr, void* dqkv, u2gnn_stream_t stream) {
    if (!qkv || !dctx || !dqkv || B < 0 || S < 2 || S > 32 || thr < 0 || thr > 255) return U2GNN_EINVAL;
    if (d != 64 && d != 32) return U2GNN_EUNSUPPORTED;
    if (!aligned16(qkv) || !aligned16(dctx) || !aligned16(dqkv)) return U2GNN_EALIGN;
    if (B == 0) return U2GNN_OK;
    const AttnRng rng = make_rng(seed, rng_stream, thr);
    cudaStream_t st = as_stream(stream);
    if (d == 64) io_bf16 ? launch_last_bwd<64, true>(qkv, dctx, B, S, rng, dqkv, st) : launch_last_bwd<64, false>(qkv, dctx, B, S, rng, dqkv, st);
    else io_bf16 ? launch_last_bwd<32, true>(qkv, dctx, B, S, rng, dqkv, st) : launch_last_bwd<32, false>(qkv, dctx, B, S, rng, dqkv, st);
    U2GNN_CHECK_LAUNCH();
}
