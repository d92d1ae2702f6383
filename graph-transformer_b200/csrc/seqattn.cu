// Softmax self-attention pieces of the fp32 path (nn.MultiheadAttention with one head;
// torch/nn/functional.py multi_head_attention_forward: q*sqrt(1/d), softmax, dropout, @v).
//   * short sequences (attn_axis="neighbors", S = k+1 <= 32): one warp per node sequence, scores
//     and probabilities live in registers (lane j <-> key j), reductions are warp shuffles
//   * long sequences (attn_axis="nodes"): row softmax forward/backward around SGEMMs
#include "common.cuh"
#include "rng.cuh"

namespace {

__device__ __forceinline__ int pitch_of(int d) { return d | 1; }  // odd pitch: lane j reads row j conflict-free

struct AttnRng {
    RngKeys keys;
    int thr;
    float scale;
};

__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// q, k, v of sequence b -> this warp's shared memory with cp.async (LDGSTS): every element of the
// node's contiguous [S, 3d] block is in flight at once, so the global latency is paid once per node
__device__ __forceinline__ void load_qkv(const float* __restrict__ qkv, int64_t b, int S, int Sq, int d,
                                         float* qs, float* ks, float* vs, int lane) {
    const int P = pitch_of(d);
    const float* base = qkv + b * (int64_t)S * 3 * d;
    const int total = S * 3 * d;
    for (int e = lane; e < total; e += 32) {
        const int r = e / (3 * d), rem = e - r * 3 * d;
        const int part = rem / d, c = rem - part * d;
        if (part == 0) {
            if (r < Sq) cp_async4(qs + r * P + c, base + e);
        } else {
            cp_async4((part == 1 ? ks : vs) + r * P + c, base + e);
        }
    }
}

// probabilities of query i for key `lane`: returns p (un-dropped) and writes the dropout multiplier
__device__ __forceinline__ float attn_probs(const float* qs, const float* ks, int i, int S, int d, int lane, float qscale,
                                            const AttnRng& rng, uint64_t elem_base, float* mult_out) {
    const int P = pitch_of(d);
    float s = -INFINITY;
    if (lane < S) {
        float acc = 0.0f;
        const float* qr = qs + i * P;
        const float* kr = ks + lane * P;
        for (int c = 0; c < d; ++c) acc = fmaf(qr[c], kr[c], acc);
        s = acc * qscale;
    }
    const float m = warp_max(s);
    const float e = (lane < S) ? expf(s - m) : 0.0f;
    const float sum = warp_sum(e);
    *mult_out = (lane < S) ? rng_dropout_mult(rng.keys, elem_base + (uint64_t)lane, rng.thr, rng.scale) : 0.0f;
    return e / sum;
}

__global__ void __launch_bounds__(256) seqattn_fwd_kernel(const float* __restrict__ qkv, int64_t B, int S, int Sq, int d,
                                                          AttnRng rng, float* __restrict__ ctx, int warps) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int P = pitch_of(d);
    float* qs = sm + (size_t)w * (Sq + 2 * S) * P;
    float* ks = qs + Sq * P;
    float* vs = ks + S * P;
    const float qscale = sqrtf(1.0f / (float)d);
    for (int64_t b = (int64_t)blockIdx.x * warps + w; b < B; b += (int64_t)gridDim.x * warps) {
        __syncwarp();
        load_qkv(qkv, b, S, Sq, d, qs, ks, vs, lane);
        cp_async_wait_all();
        __syncwarp();
        for (int i = 0; i < Sq; ++i) {
            float mult;
            const float p = attn_probs(qs, ks, i, S, d, lane, qscale, rng, (uint64_t)(b * Sq + i) * (uint64_t)S, &mult);
            const float pd = p * mult;
            float* out = ctx + (b * Sq + i) * (int64_t)d;
            for (int c0 = 0; c0 < d; c0 += 32) {
                const int c = c0 + lane;
                float acc = 0.0f;
                for (int j = 0; j < S; ++j) {
                    const float pj = __shfl_sync(0xffffffffu, pd, j);
                    if (c < d) acc = fmaf(pj, vs[j * P + c], acc);
                }
                if (c < d) out[c] = acc;
            }
        }
    }
}

__global__ void __launch_bounds__(256) seqattn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ dctx,
                                                          int64_t B, int S, int Sq, int d, AttnRng rng,
                                                          float* __restrict__ dqkv, int warps) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int P = pitch_of(d);
    const size_t per_warp = (size_t)(2 * Sq + 4 * S) * P;
    float* qs = sm + (size_t)w * per_warp;
    float* ks = qs + Sq * P;
    float* vs = ks + S * P;
    float* dk = vs + S * P;
    float* dv = dk + S * P;
    float* dall = dv + S * P;      // [Sq][P] gradient rows
    const float qscale = sqrtf(1.0f / (float)d);
    for (int64_t b = (int64_t)blockIdx.x * warps + w; b < B; b += (int64_t)gridDim.x * warps) {
        __syncwarp();
        load_qkv(qkv, b, S, Sq, d, qs, ks, vs, lane);
        {   // all Sq gradient rows of this node (contiguous [Sq, d]) in the same async batch
            const float* gsrc = dctx + b * (int64_t)Sq * d;
            for (int e = lane; e < Sq * d; e += 32) cp_async4(dall + (e / d) * P + (e % d), gsrc + e);
        }
        for (int e = lane; e < S * P; e += 32) {
            dk[e] = 0.0f;
            dv[e] = 0.0f;
        }
        cp_async_wait_all();
        __syncwarp();
        float* gbase = dqkv + b * (int64_t)S * 3 * d;
        for (int i = 0; i < S; ++i) {
            if (i >= Sq) {  // rows without a query (dead-row-eliminated timestep): dq = 0
                for (int c = lane; c < d; c += 32) gbase[(int64_t)i * 3 * d + c] = 0.0f;
                continue;
            }
            const float* drow = dall + i * P;
            float mult;
            const float p = attn_probs(qs, ks, i, S, d, lane, qscale, rng, (uint64_t)(b * Sq + i) * (uint64_t)S, &mult);
            const float pd = p * mult;
            float dpt = 0.0f;
            if (lane < S) {
                const float* vr = vs + lane * P;
                for (int c = 0; c < d; ++c) dpt = fmaf(drow[c], vr[c], dpt);
            }
            const float dp = dpt * mult;
            const float tsum = warp_sum((lane < S) ? p * dp : 0.0f);
            const float ds = (lane < S) ? p * (dp - tsum) : 0.0f;
            for (int c0 = 0; c0 < d; c0 += 32) {
                const int c = c0 + lane;
                float dq = 0.0f;
                const float qv = (c < d) ? qs[i * P + c] * qscale : 0.0f;
                const float dr = (c < d) ? drow[c] : 0.0f;
                for (int j = 0; j < S; ++j) {
                    const float dsj = __shfl_sync(0xffffffffu, ds, j);
                    const float pdj = __shfl_sync(0xffffffffu, pd, j);
                    if (c < d) {
                        dq = fmaf(dsj, ks[j * P + c], dq);
                        dk[j * P + c] = fmaf(dsj, qv, dk[j * P + c]);
                        dv[j * P + c] = fmaf(pdj, dr, dv[j * P + c]);
                    }
                }
                if (c < d) gbase[(int64_t)i * 3 * d + c] = dq * qscale;
            }
        }
        __syncwarp();
        for (int j = 0; j < S; ++j)
            for (int c = lane; c < d; c += 32) {
                gbase[(int64_t)j * 3 * d + d + c] = dk[j * P + c];
                gbase[(int64_t)j * 3 * d + 2 * d + c] = dv[j * P + c];
            }
    }
}

// ---- long-sequence row softmax (one warp per row) ----
__global__ void __launch_bounds__(256) softmax_rows_fwd_kernel(float* __restrict__ scores, int64_t M, int64_t N,
                                                               float* __restrict__ pd, AttnRng rng) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < M; r += nwarps) {
        float* row = scores + r * N;
        float m = -INFINITY;
        for (int64_t c = lane; c < N; c += 32) m = fmaxf(m, row[c]);
        m = warp_max(m);
        float sum = 0.0f;
        for (int64_t c = lane; c < N; c += 32) sum += expf(row[c] - m);
        sum = warp_sum(sum);
        for (int64_t c = lane; c < N; c += 32) {
            const float p = expf(row[c] - m) / sum;
            row[c] = p;
            pd[r * N + c] = p * rng_dropout_mult(rng.keys, (uint64_t)(r * N + c), rng.thr, rng.scale);
        }
    }
}

__global__ void __launch_bounds__(256) softmax_rows_bwd_kernel(const float* __restrict__ probs, float* __restrict__ dp,
                                                               int64_t M, int64_t N, AttnRng rng) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < M; r += nwarps) {
        const float* p = probs + r * N;
        float* g = dp + r * N;
        float t = 0.0f;
        for (int64_t c = lane; c < N; c += 32) {
            const float gv = g[c] * rng_dropout_mult(rng.keys, (uint64_t)(r * N + c), rng.thr, rng.scale);
            g[c] = gv;
            t = fmaf(p[c], gv, t);
        }
        t = warp_sum(t);
        for (int64_t c = lane; c < N; c += 32) g[c] = p[c] * (g[c] - t);
    }
}

AttnRng make_rng(uint64_t seed, uint32_t stream, int thr) {
    AttnRng r;
    r.keys = rng_keys(seed, stream);
    r.thr = thr;
    r.scale = thr ? rng_keep_scale(thr) : 1.0f;
    return r;
}

int pick_warps(size_t floats_per_warp) {
    int w = 8;
    while (w > 1 && (size_t)w * floats_per_warp * sizeof(float) > 110 * 1024) w >>= 1;   // two CTAs per SM
    return w;
}

}  // namespace

// thread-per-row kernels for the common shapes (seqattn_rows.cu)
int seqattn_rows_try_fwd(const float* qkv, int64_t B, int S, int Sq, int d, uint64_t seed, uint32_t rng_stream, int thr,
                         float* ctx, cudaStream_t st);
int seqattn_rows_try_bwd(const float* qkv, const float* dctx, int64_t B, int S, int Sq, int d, uint64_t seed,
                         uint32_t rng_stream, int thr, float* dqkv, cudaStream_t st);

extern "C" int u2gnn_seqattn_fwd(const float* qkv, int64_t B, int S, int Sq, int d, uint64_t seed, uint32_t rng_stream,
                                 int thr, float* ctx, u2gnn_stream_t stream) {
    if (!qkv || !ctx || B < 0 || d <= 0) return U2GNN_EINVAL;
    if (S < 1 || S > 32 || (Sq != S && Sq != 1) || thr < 0 || thr > 255) return U2GNN_EINVAL;
    if (B == 0) return U2GNN_OK;
    if (seqattn_rows_try_fwd(qkv, B, S, Sq, d, seed, rng_stream, thr, ctx, as_stream(stream))) { U2GNN_CHECK_LAUNCH(); }
    const int P = d | 1;
    const size_t per_warp = (size_t)(Sq + 2 * S) * P;
    const int warps = pick_warps(per_warp);
    const size_t smem = (size_t)warps * per_warp * sizeof(float);
    if (smem > 200 * 1024) return U2GNN_EUNSUPPORTED;
    cudaFuncSetAttribute(seqattn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    seqattn_fwd_kernel<<<grid_for(B, warps, 2), warps * 32, smem, as_stream(stream)>>>(qkv, B, S, Sq, d,
                                                                                      make_rng(seed, rng_stream, thr),
                                                                                      ctx, warps);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_seqattn_bwd(const float* qkv, const float* dctx, int64_t B, int S, int Sq, int d, uint64_t seed,
                                 uint32_t rng_stream, int thr, float* dqkv, u2gnn_stream_t stream) {
    if (!qkv || !dctx || !dqkv || B < 0 || d <= 0) return U2GNN_EINVAL;
    if (S < 1 || S > 32 || (Sq != S && Sq != 1) || thr < 0 || thr > 255) return U2GNN_EINVAL;
    if (B == 0) return U2GNN_OK;
    if (seqattn_rows_try_bwd(qkv, dctx, B, S, Sq, d, seed, rng_stream, thr, dqkv, as_stream(stream))) { U2GNN_CHECK_LAUNCH(); }
    const int P = d | 1;
    const size_t per_warp = (size_t)(2 * Sq + 4 * S) * P;
    const int warps = pick_warps(per_warp);
    const size_t smem = (size_t)warps * per_warp * sizeof(float);
    if (smem > 200 * 1024) return U2GNN_EUNSUPPORTED;
    cudaFuncSetAttribute(seqattn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    seqattn_bwd_kernel<<<grid_for(B, warps, 2), warps * 32, smem, as_stream(stream)>>>(qkv, dctx, B, S, Sq, d,
                                                                                      make_rng(seed, rng_stream, thr),
                                                                                      dqkv, warps);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_softmax_rows_fwd(float* scores, int64_t M, int64_t N, float* probs_dropped, uint64_t seed,
                                      uint32_t rng_stream, int thr, u2gnn_stream_t stream) {
    if (!scores || !probs_dropped || M < 0 || N <= 0 || thr < 0 || thr > 255) return U2GNN_EINVAL;
    if (M == 0) return U2GNN_OK;
    softmax_rows_fwd_kernel<<<grid_for(M, 8, 8), 256, 0, as_stream(stream)>>>(scores, M, N, probs_dropped,
                                                                             make_rng(seed, rng_stream, thr));
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_softmax_rows_bwd(const float* probs, float* dprobs_inout, int64_t M, int64_t N, uint64_t seed,
                                      uint32_t rng_stream, int thr, u2gnn_stream_t stream) {
    if (!probs || !dprobs_inout || M < 0 || N <= 0 || thr < 0 || thr > 255) return U2GNN_EINVAL;
    if (M == 0) return U2GNN_OK;
    softmax_rows_bwd_kernel<<<grid_for(M, 8, 8), 256, 0, as_stream(stream)>>>(probs, dprobs_inout, M, N,
                                                                             make_rng(seed, rng_stream, thr));
    U2GNN_CHECK_LAUNCH();
}
