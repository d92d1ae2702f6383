// K8: global-norm gradient clip + Adam over one flat parameter arena
// (torch.nn.utils.clip_grad_norm_(params, 0.5) + torch.optim.Adam.step(),
// train_pytorch_U2GNN_Sup.py:145,160-161).  All parameters (and, in the unsupervised model, the
// dense [V, D] class table) live in one contiguous fp32 buffer, so the whole optimiser step is two
// launches instead of ~6 per tensor.  Pure streaming: 28 B/parameter (read p,g,m,v; write p,m,v).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ sumsq) {
    __shared__ float red[8];
    float acc = 0.0f;
    const int64_t n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(g4 + i);
        acc = fmaf(v.x, v.x, acc);
        acc = fmaf(v.y, v.y, acc);
        acc = fmaf(v.z, v.z, acc);
        acc = fmaf(v.w, v.w, acc);
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc = fmaf(g[i], g[i], acc);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = (threadIdx.x < 8) ? red[threadIdx.x] : 0.0f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(sumsq, v);
    }
}

__global__ void __launch_bounds__(256) clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                        const float* __restrict__ sumsq, float max_norm, float lr,
                                                        float beta1, float beta2, float eps, float bc1, float bc2_sqrt) {
    const float total = sqrtf(sumsq[0]);
    float coef = max_norm / (total + 1e-6f);
    coef = coef < 1.0f ? coef : 1.0f;
    const float step = lr / bc1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i] * coef;
        const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = p[i] - step * (mi / denom);
    }
}

}  // namespace

extern "C" int u2gnn_grad_sqnorm(const float* g, int64_t n, float* sumsq, u2gnn_stream_t stream) {
    if (!g || !sumsq || n < 0) return U2GNN_EINVAL;
    if (reinterpret_cast<uintptr_t>(g) % 16 != 0) return U2GNN_EALIGN;
    if (n == 0) return U2GNN_OK;
    sqnorm_kernel<<<grid_for(n, 1024, 4), 256, 0, as_stream(stream)>>>(g, n, sumsq);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, const float* sumsq,
                               float max_norm, float lr, float beta1, float beta2, float eps, int64_t step,
                               u2gnn_stream_t stream) {
    if (!p || !g || !m || !v || !sumsq || n < 0 || step < 1) return U2GNN_EINVAL;
    if (n == 0) return U2GNN_OK;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    clip_adam_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(p, g, m, v, n, sumsq, max_norm, lr, beta1, beta2,
                                                                        eps, (float)bc1, (float)sqrt(bc2));
    U2GNN_CHECK_LAUNCH();
}
