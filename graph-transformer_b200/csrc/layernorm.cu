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


// ---- vectorised variants (d = 4 * LPR, LPR lanes per row, 32 / LPR rows per warp): 128-bit accesses, one dropout
//      keep word per float4 (its 4 elements share a 32-element group because d % 4 == 0)
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float4 drop4(const DropRng& rng, int low, uint64_t e, float4 v) {
    if (rng.thr == 0) return v;
    const uint32_t w = rng_keep_word_lo(rng.keys, e >> 5, rng.thr, low) >> (e & 31);
    v.x = (w & 1u) ? v.x * rng.scale : 0.0f;
    v.y = (w & 2u) ? v.y * rng.scale : 0.0f;
    v.z = (w & 4u) ? v.z * rng.scale : 0.0f;
    v.w = (w & 8u) ? v.w * rng.scale : 0.0f;
    return v;
}

template <int LPR>
__global__ void __launch_bounds__(256) ln_fwd_vec_kernel(const float* __restrict__ res, const float* __restrict__ a, int64_t M,
                                                         DropRng rng, int low, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float* __restrict__ z,
                                                         float* __restrict__ y, float* __restrict__ stats) {
    constexpr int D = 4 * LPR, RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane / LPR, l = lane % LPR;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float4 g4 = reinterpret_cast<const float4*>(gamma)[l], b4 = reinterpret_cast<const float4*>(beta)[l];
    const float inv_d = 1.0f / (float)D;
    for (int64_t r0 = warp * RPW; r0 < M; r0 += nwarps * RPW) {
        const int64_t r = r0 + sub;
        const bool ok = r < M;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) {
            v = drop4(rng, low, (uint64_t)(r * D + 4 * l), __ldg(reinterpret_cast<const float4*>(a + r * D) + l));
            if (res) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(res + r * D) + l);
                v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
            }
            reinterpret_cast<float4*>(z + r * D)[l] = v;
        }
        const float mean = group_sum<LPR>((v.x + v.y) + (v.z + v.w)) * inv_d;
        const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
        const float rstd = rsqrtf(group_sum<LPR>((dx * dx + dy * dy) + (dz * dz + dw * dw)) * inv_d + kLnEps);
        if (ok) {
            reinterpret_cast<float4*>(y + r * D)[l] = make_float4(dx * rstd * g4.x + b4.x, dy * rstd * g4.y + b4.y,
                                                                  dz * rstd * g4.z + b4.z, dw * rstd * g4.w + b4.w);
            if (l == 0 && stats) {
                stats[2 * r] = mean;
                stats[2 * r + 1] = rstd;
            }
        }
    }
}

template <int LPR>
__global__ void __launch_bounds__(256, 3) ln_bwd_vec_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                         const float* __restrict__ stats, int64_t M,
                                                         const float* __restrict__ gamma, DropRng rng, int low,
                                                         float* __restrict__ dz, void* __restrict__ da_, int da_bf16,
                                                         float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                         float* __restrict__ dasum) {
    constexpr int D = 4 * LPR, RPW = 32 / LPR;
    float* da = static_cast<float*>(da_);
    float4 as = make_float4(0.f, 0.f, 0.f, 0.f);
    const int lane = threadIdx.x & 31, sub = lane / LPR, l = lane % LPR;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float4 g4 = reinterpret_cast<const float4*>(gamma)[l];
    const float inv_d = 1.0f / (float)D;
    float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = make_float4(0.f, 0.f, 0.f, 0.f);
    // two row groups per iteration: the loads of both (dy, z, stats) are in flight together (one group per iteration kept
    // 32 KB outstanding per SM at 32 resident warps - below what the HBM latency x bandwidth product needs)
    constexpr int UNR = 4;
    for (int64_t r0 = warp * (UNR * RPW); r0 < M; r0 += nwarps * (UNR * RPW)) {
        float4 gk[UNR], xk[UNR];
        float mk[UNR], sk[UNR];
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
            const int64_t r = r0 + k * RPW + sub;
            gk[k] = xk[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            mk[k] = sk[k] = 0.f;
            if (r < M) {
                gk[k] = __ldg(reinterpret_cast<const float4*>(dy + r * D) + l);
                xk[k] = __ldg(reinterpret_cast<const float4*>(z + r * D) + l);
                mk[k] = stats[2 * r];
                sk[k] = stats[2 * r + 1];
            }
        }
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
            const int64_t r = r0 + k * RPW + sub;
            const bool ok = r < M;
            const float4 g = gk[k], x = xk[k];
            const float mean = mk[k], rstd = sk[k];
            const float4 xh = make_float4((x.x - mean) * rstd, (x.y - mean) * rstd, (x.z - mean) * rstd, (x.w - mean) * rstd);
            const float4 dh = make_float4(g.x * g4.x, g.y * g4.y, g.z * g4.z, g.w * g4.w);
            ag.x = fmaf(g.x, xh.x, ag.x); ag.y = fmaf(g.y, xh.y, ag.y); ag.z = fmaf(g.z, xh.z, ag.z); ag.w = fmaf(g.w, xh.w, ag.w);
            ab.x += g.x; ab.y += g.y; ab.z += g.z; ab.w += g.w;
            const float m1 = group_sum<LPR>((dh.x + dh.y) + (dh.z + dh.w)) * inv_d;
            const float m2 = group_sum<LPR>((dh.x * xh.x + dh.y * xh.y) + (dh.z * xh.z + dh.w * xh.w)) * inv_d;
            if (ok) {
                const float4 o = make_float4(rstd * (dh.x - m1 - xh.x * m2), rstd * (dh.y - m1 - xh.y * m2),
                                             rstd * (dh.z - m1 - xh.z * m2), rstd * (dh.w - m1 - xh.w * m2));
                reinterpret_cast<float4*>(dz + r * D)[l] = o;
                if (da || dasum) {
                    const float4 m = drop4(rng, low, (uint64_t)(r * D + 4 * l), o);
                    if (!da) {
                    } else if (da_bf16) {      // consumers are tensor-core kernels that round on load: round once here, half the bytes
                        uint2 w;
                        w.x = cvt_bf16x2(m.x, m.y);
                        w.y = cvt_bf16x2(m.z, m.w);
                        if (da_bf16 == 2)      // D == 64: swizzled [128 x 64] tile images (16-byte chunks XOR (row & 7)), the FFN backward's operand format
                            *reinterpret_cast<uint2*>(static_cast<uint8_t*>(da_) + (size_t)(r >> 7) * 16384 + (size_t)(r & 127) * 128 +
                                                      ((((l >> 1) ^ (int)(r & 7)) << 4) | ((l & 1) << 3))) = w;
                        else
                            reinterpret_cast<uint2*>(static_cast<uint16_t*>(da_) + r * D)[l] = w;
                    } else {
                        reinterpret_cast<float4*>(da + r * D)[l] = m;
                    }
                    as.x += m.x; as.y += m.y; as.z += m.z; as.w += m.w;
                }
            }
        }
    }
    // reduce the row groups of the warp, then one atomic per column per warp
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
        ag.x += __shfl_xor_sync(0xffffffffu, ag.x, o); ag.y += __shfl_xor_sync(0xffffffffu, ag.y, o);
        ag.z += __shfl_xor_sync(0xffffffffu, ag.z, o); ag.w += __shfl_xor_sync(0xffffffffu, ag.w, o);
        ab.x += __shfl_xor_sync(0xffffffffu, ab.x, o); ab.y += __shfl_xor_sync(0xffffffffu, ab.y, o);
        ab.z += __shfl_xor_sync(0xffffffffu, ab.z, o); ab.w += __shfl_xor_sync(0xffffffffu, ab.w, o);
        if (dasum) {
            as.x += __shfl_xor_sync(0xffffffffu, as.x, o); as.y += __shfl_xor_sync(0xffffffffu, as.y, o);
            as.z += __shfl_xor_sync(0xffffffffu, as.z, o); as.w += __shfl_xor_sync(0xffffffffu, as.w, o);
        }
    }
    if (sub == 0) {
        if (dasum) { atomicAdd(dasum + 4 * l, as.x); atomicAdd(dasum + 4 * l + 1, as.y); atomicAdd(dasum + 4 * l + 2, as.z); atomicAdd(dasum + 4 * l + 3, as.w); }
        if (dgamma) { atomicAdd(dgamma + 4 * l, ag.x); atomicAdd(dgamma + 4 * l + 1, ag.y); atomicAdd(dgamma + 4 * l + 2, ag.z); atomicAdd(dgamma + 4 * l + 3, ag.w); }
        if (dbeta) { atomicAdd(dbeta + 4 * l, ab.x); atomicAdd(dbeta + 4 * l + 1, ab.y); atomicAdd(dbeta + 4 * l + 2, ab.z); atomicAdd(dbeta + 4 * l + 3, ab.w); }
    }
}

bool vec_ok(int d, const void* a, const void* b, const void* c, const void* e, const void* f) {
    if (d != 4 && d != 8 && d != 16 && d != 32 && d != 64 && d != 128) return false;   // LPR = d / 4 lanes per row (d = 4: one lane per row, 32 rows per warp)
    const uintptr_t m = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
                        reinterpret_cast<uintptr_t>(e) | reinterpret_cast<uintptr_t>(f);
    return (m % 16) == 0;
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
    if (vec_ok(d, res, a, z, y, gamma) && reinterpret_cast<uintptr_t>(beta) % 16 == 0) {
        const DropRng rng = make_rng(seed, rng_stream, thr);
        const int low = rng_thr_low(thr);
        const int rpb = 8 * (128 / d);                       // rows per 256-thread block
        const int grid = grid_for(M, rpb, 8);
        cudaStream_t st = as_stream(stream);
        if (d == 4) ln_fwd_vec_kernel<1><<<grid, 256, 0, st>>>(res, a, M, rng, low, gamma, beta, z, y, stats);
        else if (d == 8) ln_fwd_vec_kernel<2><<<grid, 256, 0, st>>>(res, a, M, rng, low, gamma, beta, z, y, stats);
        else if (d == 16) ln_fwd_vec_kernel<4><<<grid, 256, 0, st>>>(res, a, M, rng, low, gamma, beta, z, y, stats);
        else if (d == 32) ln_fwd_vec_kernel<8><<<grid, 256, 0, st>>>(res, a, M, rng, low, gamma, beta, z, y, stats);
        else if (d == 64) ln_fwd_vec_kernel<16><<<grid, 256, 0, st>>>(res, a, M, rng, low, gamma, beta, z, y, stats);
        else ln_fwd_vec_kernel<32><<<grid, 256, 0, st>>>(res, a, M, rng, low, gamma, beta, z, y, stats);
        U2GNN_CHECK_LAUNCH();
    }
    add_dropout_ln_fwd_kernel<<<grid_for(M, 8, 8), 256, 0, as_stream(stream)>>>(res, a, M, d, make_rng(seed, rng_stream, thr),
                                                                               gamma, beta, z, y, stats);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_add_dropout_ln_bwd_ex(const float* dy, const float* z, const float* stats, int64_t M, int d,
                                           const float* gamma, uint64_t seed, uint32_t rng_stream, int thr, float* dz,
                                           void* da, int da_bf16, float* dgamma, float* dbeta, float* dasum,
                                           u2gnn_stream_t stream) {
    if (!dy || !z || !stats || !gamma || !dz || M < 0 || d <= 0 || thr < 0 || thr > 255) return U2GNN_EINVAL;
    if (da_bf16 && !da) return U2GNN_EINVAL;
    if (da_bf16 < 0 || da_bf16 > 2 || (da_bf16 == 2 && d != 64)) return U2GNN_EUNSUPPORTED;
    if (d > 32 * kMaxSlots) return U2GNN_EUNSUPPORTED;
    if (M == 0) return U2GNN_OK;
    if (vec_ok(d, dy, z, dz, da, gamma)) {
        const DropRng rng = make_rng(seed, rng_stream, thr);
        const int low = rng_thr_low(thr);
        const int grid = grid_for(M, 64 * (128 / d), 3);
        cudaStream_t st = as_stream(stream);
        if (d == 4) ln_bwd_vec_kernel<1><<<grid, 256, 0, st>>>(dy, z, stats, M, gamma, rng, low, dz, da, da_bf16, dgamma, dbeta, dasum);
        else if (d == 8) ln_bwd_vec_kernel<2><<<grid, 256, 0, st>>>(dy, z, stats, M, gamma, rng, low, dz, da, da_bf16, dgamma, dbeta, dasum);
        else if (d == 16) ln_bwd_vec_kernel<4><<<grid, 256, 0, st>>>(dy, z, stats, M, gamma, rng, low, dz, da, da_bf16, dgamma, dbeta, dasum);
        else if (d == 32) ln_bwd_vec_kernel<8><<<grid, 256, 0, st>>>(dy, z, stats, M, gamma, rng, low, dz, da, da_bf16, dgamma, dbeta, dasum);
        else if (d == 64) ln_bwd_vec_kernel<16><<<grid, 256, 0, st>>>(dy, z, stats, M, gamma, rng, low, dz, da, da_bf16, dgamma, dbeta, dasum);
        else ln_bwd_vec_kernel<32><<<grid, 256, 0, st>>>(dy, z, stats, M, gamma, rng, low, dz, da, da_bf16, dgamma, dbeta, dasum);
        U2GNN_CHECK_LAUNCH();
    }
    if (da_bf16 || dasum) return U2GNN_EUNSUPPORTED;            // bf16 da / fused column sum: vectorised shapes only
    // fewer, fatter warps than the forward: each warp flushes d atomics at the end (six blocks per SM: with two the kernel ran at
    // 1.4 TB/s on the d = 65 rows of configs[2] - one warp per row has only its own few loads in flight)
    add_dropout_ln_bwd_kernel<<<grid_for(M, 32, 6), 256, 0, as_stream(stream)>>>(dy, z, stats, M, d, gamma,
                                                                                make_rng(seed, rng_stream, thr), dz,
                                                                                static_cast<float*>(da), dgamma, dbeta);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_add_dropout_ln_bwd(const float* dy, const float* z, const float* stats, int64_t M, int d,
                                        const float* gamma, uint64_t seed, uint32_t rng_stream, int thr, float* dz,
                                        float* da, float* dgamma, float* dbeta, u2gnn_stream_t stream) {
    return u2gnn_add_dropout_ln_bwd_ex(dy, z, stats, M, d, gamma, seed, rng_stream, thr, dz, da, 0, dgamma, dbeta, nullptr, stream);
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
