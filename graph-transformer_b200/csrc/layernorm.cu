// Residual + dropout + LayerNorm (post-norm blocks of nn.TransformerEncoderLayer,
// torch/nn/modules/transformer.py:944-958; LayerNorm eps 1e-5, biased variance) and the small
// elementwise helpers of the fp32 path.  One warp per row; rows are independent; HBM-bound.
#include "common.cuh"
#include "rng.cuh"

namespace {

constexpr float kLnEps = 1e-5f;
constexpr int kMaxSlots = 8;  // feature sizes up to 256 keep per-lane partial sums in registers

struct DropRng {
    RngKeys keys;
    int thr;
    float scale;
};

__global__ void __launch_bounds__(256) add_dropout_ln_fwd_kernel(const float* __restrict__ res,
                                                                 const float* __restrict__ a, int64_t M, int d,
                                                                 DropRng rng, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, float* __restrict__ z,
                                                                 float* __restrict__ y, float* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float inv_d = 1.0f / (float)d;
    for (int64_t r = warp; r < M; r += nwarps) {
        float zv[kMaxSlots];
        float sum = 0.0f;
#pragma unroll
        for (int s = 0; s < kMaxSlots; ++s) {
            const int c = lane + 32 * s;
            float v = 0.0f;
            if (c < d) {
                v = a[r * d + c] * rng_dropout_mult(rng.keys, (uint64_t)(r * d + c), rng.thr, rng.scale);
                if (res) v += res[r * d + c];
                z[r * d + c] = v;
            }
            zv[s] = v;
            sum += v;
        }
        const float mean = warp_sum(sum) * inv_d;
        float sq = 0.0f;
#pragma unroll
        for (int s = 0; s < kMaxSlots; ++s) {
            const int c = lane + 32 * s;
            const float t = (c < d) ? zv[s] - mean : 0.0f;
            sq = fmaf(t, t, sq);
        }
        const float rstd = rsqrtf(warp_sum(sq) * inv_d + kLnEps);
#pragma unroll
        for (int s = 0; s < kMaxSlots; ++s) {
            const int c = lane + 32 * s;
            if (c < d) y[r * d + c] = (zv[s] - mean) * rstd * gamma[c] + beta[c];
        }
        if (lane == 0 && stats) {
            stats[2 * r] = mean;
            stats[2 * r + 1] = rstd;
        }
    }
}

__global__ void __launch_bounds__(256) add_dropout_ln_bwd_kernel(const float* __restrict__ dy,
                                                                 const float* __restrict__ z,
                                                                 const float* __restrict__ stats, int64_t M, int d,
                                                                 const float* __restrict__ gamma, DropRng rng,
                                                                 float* __restrict__ dz, float* __restrict__ da,
                                                                 float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float inv_d = 1.0f / (float)d;
    float accg[kMaxSlots], accb[kMaxSlots];
#pragma unroll
    for (int s = 0; s < kMaxSlots; ++s) accg[s] = accb[s] = 0.0f;
    for (int64_t r = warp; r < M; r += nwarps) {
        const float mean = stats[2 * r], rstd = stats[2 * r + 1];
        float xh[kMaxSlots], dxh[kMaxSlots];
        float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
        for (int s = 0; s < kMaxSlots; ++s) {
            const int c = lane + 32 * s;
            xh[s] = dxh[s] = 0.0f;
            if (c < d) {
                const float g = dy[r * d + c];
                xh[s] = (z[r * d + c] - mean) * rstd;
                dxh[s] = g * gamma[c];
                accg[s] = fmaf(g, xh[s], accg[s]);
                accb[s] += g;
            }
            s1 += dxh[s];
            s2 = fmaf(dxh[s], xh[s], s2);
        }
        const float m1 = warp_sum(s1) * inv_d, m2 = warp_sum(s2) * inv_d;
#pragma unroll
        for (int s = 0; s < kMaxSlots; ++s) {
            const int c = lane + 32 * s;
            if (c < d) {
                const float g = rstd * (dxh[s] - m1 - xh[s] * m2);
                dz[r * d + c] = g;
                if (da) da[r * d + c] = g * rng_dropout_mult(rng.keys, (uint64_t)(r * d + c), rng.thr, rng.scale);
            }
        }
    }
    // per-warp partial sums -> global (few atomics: one per column per warp)
#pragma unroll
    for (int s = 0; s < kMaxSlots; ++s) {
        const int c = lane + 32 * s;
        if (c < d) {
            if (dgamma) atomicAdd(dgamma + c, accg[s]);
            if (dbeta) atomicAdd(dbeta + c, accb[s]);
        }
    }
}

__global__ void __launch_bounds__(256) ln_apply_kernel(const float* __restrict__ z, const float* __restrict__ stats,
                                                       int64_t M, int d, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float* __restrict__ y) {
    const int64_t total = M * d;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / d;
        const int c = (int)(e - r * d);
        y[e] = (z[e] - stats[2 * r]) * stats[2 * r + 1] * gamma[c] + beta[c];
    }
}

__global__ void __launch_bounds__(256) dropout_apply_kernel(const float* __restrict__ x, int64_t n, DropRng rng,
                                                            float* __restrict__ y) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        y[e] = x[e] * rng_dropout_mult(rng.keys, (uint64_t)e, rng.thr, rng.scale);
}

__global__ void __launch_bounds__(256) axpy_kernel(float alpha, const float* __restrict__ x, float* __restrict__ y,
                                                   int64_t n) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        y[e] = fmaf(alpha, x[e], y[e]);
}

__global__ void __launch_bounds__(256) copy_rows_kernel(const float* __restrict__ src, int64_t ld_src,
                                                        float* __restrict__ dst, int64_t ld_dst, int64_t rows, int d,
                                                        int accumulate) {
    const int64_t total = rows * d;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / d;
        const int c = (int)(e - r * d);
        const float v = src[r * ld_src + c];
        float* o = dst + r * ld_dst + c;
        *o = accumulate ? *o + v : v;
    }
}

DropRng make_rng(uint64_t seed, uint32_t stream, int thr) {
    DropRng r;
    r.keys = rng_keys(seed, stream);
    r.thr = thr;
    r.scale = thr ? rng_keep_scale(thr) : 1.0f;
    return r;
}

}  // namespace

extern "C" int u2gnn_add_dropout_ln_fwd(const float* res, const float* a, int64_t M, int d, uint64_t seed,
                                        uint32_t rng_stream, int thr, const float* gamma, const float* beta, float* z,
                                        float* y, float* stats, u2gnn_stream_t stream) {
    if (!a || !gamma || !beta || !z || !y || M < 0 || d <= 0 || thr < 0 || thr > 255) return U2GNN_EINVAL;
    if (d > 32 * kMaxSlots) return U2GNN_EUNSUPPORTED;
    if (M == 0) return U2GNN_OK;
    add_dropout_ln_fwd_kernel<<<grid_for(M, 8, 8), 256, 0, as_stream(stream)>>>(res, a, M, d, make_rng(seed, rng_stream, thr),
                                                                               gamma, beta, z, y, stats);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_add_dropout_ln_bwd(const float* dy, const float* z, const float* stats, int64_t M, int d,
                                        const float* gamma, uint64_t seed, uint32_t rng_stream, int thr, float* dz,
                                        float* da, float* dgamma, float* dbeta, u2gnn_stream_t stream) {
    if (!dy || !z || !stats || !gamma || !dz || M < 0 || d <= 0 || thr < 0 || thr > 255) return U2GNN_EINVAL;
    if (d > 32 * kMaxSlots) return U2GNN_EUNSUPPORTED;
    if (M == 0) return U2GNN_OK;
    // fewer, fatter warps than the forward: each warp flushes d atomics at the end
    add_dropout_ln_bwd_kernel<<<grid_for(M, 64, 2), 256, 0, as_stream(stream)>>>(dy, z, stats, M, d, gamma,
                                                                                make_rng(seed, rng_stream, thr), dz, da,
                                                                                dgamma, dbeta);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_ln_apply(const float* z, const float* stats, int64_t M, int d, const float* gamma,
                              const float* beta, float* y, u2gnn_stream_t stream) {
    if (!z || !stats || !gamma || !beta || !y || M < 0 || d <= 0) return U2GNN_EINVAL;
    if (M == 0) return U2GNN_OK;
    ln_apply_kernel<<<grid_for(M * d, 256, 8), 256, 0, as_stream(stream)>>>(z, stats, M, d, gamma, beta, y);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_dropout_apply(const float* x, int64_t numel, uint64_t seed, uint32_t rng_stream, int thr, float* y,
                                   u2gnn_stream_t stream) {
    if (!x || !y || numel < 0 || thr < 0 || thr > 255) return U2GNN_EINVAL;
    if (numel == 0) return U2GNN_OK;
    dropout_apply_kernel<<<grid_for(numel, 256, 8), 256, 0, as_stream(stream)>>>(x, numel, make_rng(seed, rng_stream, thr), y);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_axpy(float alpha, const float* x, float* y, int64_t numel, u2gnn_stream_t stream) {
    if (!x || !y || numel < 0) return U2GNN_EINVAL;
    if (numel == 0) return U2GNN_OK;
    axpy_kernel<<<grid_for(numel, 256, 8), 256, 0, as_stream(stream)>>>(alpha, x, y, numel);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_copy_rows(const float* src, int64_t ld_src, float* dst, int64_t ld_dst, int64_t rows, int d,
                               int accumulate, u2gnn_stream_t stream) {
    if (!src || !dst || rows < 0 || d <= 0) return U2GNN_EINVAL;
    if (rows == 0) return U2GNN_OK;
    copy_rows_kernel<<<grid_for(rows * d, 256, 8), 256, 0, as_stream(stream)>>>(src, ld_src, dst, ld_dst, rows, d, accumulate);
    U2GNN_CHECK_LAUNCH();
}
