// K5: classifier head (dropout on pooled graph embeddings + Linear, pytorch_U2GNN_Sup.py:42-44)
// and the label-smoothed soft cross-entropy (pytorch_U2GNN_Sup.py:48-59,
// train_pytorch_U2GNN_Sup.py:140-142).  G x C is tiny (C = 2..5): launch-bound, one thread per
// output element.
#include "common.cuh"
#include "rng.cuh"

namespace {

struct DropRng {
    RngKeys keys;
    int thr;
    float scale;
};

__global__ void head_fwd_kernel(const float* __restrict__ ge, int64_t G, int d, const float* __restrict__ W,
                                const float* __restrict__ b, int C, DropRng rng, float* __restrict__ scores,
                                int accumulate) {
    const int64_t total = G * C;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = e / C;
        const int c = (int)(e - g * C);
        float acc = b[c];
        for (int k = 0; k < d; ++k)
            acc = fmaf(ge[g * d + k] * rng_dropout_mult(rng.keys, (uint64_t)(g * d + k), rng.thr, rng.scale), W[c * d + k], acc);
        scores[e] = accumulate ? scores[e] + acc : acc;
    }
}

// one thread per graph: log-softmax over C classes, soft targets, gradient of the mean loss
__global__ void soft_ce_kernel(const float* __restrict__ scores, const int64_t* __restrict__ labels, int64_t G, int C,
                               float smoothing, float inv_g_total, float* __restrict__ loss,
                               float* __restrict__ dscores) {
    float local = 0.0f;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < G; g += (int64_t)gridDim.x * blockDim.x) {
        const float* s = scores + g * C;
        float m = -INFINITY;
        for (int c = 0; c < C; ++c) m = fmaxf(m, s[c]);
        float sum = 0.0f;
        for (int c = 0; c < C; ++c) sum += expf(s[c] - m);
        const float lse = m + logf(sum);
        const int64_t y = labels[g];
        const float off = smoothing / (float)(C - 1), on = 1.0f - smoothing;
        float tsum = 0.0f, l = 0.0f;
        for (int c = 0; c < C; ++c) {
            const float t = (c == y) ? on : off;
            tsum += t;
            l -= t * (s[c] - lse);
        }
        for (int c = 0; c < C; ++c) {
            const float t = (c == y) ? on : off;
            dscores[g * C + c] = (expf(s[c] - lse) * tsum - t) * inv_g_total;
        }
        local += l;
    }
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0 && local != 0.0f) atomicAdd(loss, local * inv_g_total);
}

__global__ void head_bwd_kernel(const float* __restrict__ dscores, const float* __restrict__ ge, int64_t G, int d,
                                const float* __restrict__ W, int C, DropRng rng, float* __restrict__ dge) {
    // phase A (threads over G*d): dge = (dscores @ W) * mask
    const int64_t total = G * d;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = e / d;
        const int k = (int)(e - g * d);
        float acc = 0.0f;
        for (int c = 0; c < C; ++c) acc = fmaf(dscores[g * C + c], W[c * d + k], acc);
        dge[e] = acc * rng_dropout_mult(rng.keys, (uint64_t)e, rng.thr, rng.scale);
    }
}

// phase B: dW[c,k] += sum_g dscores[g,c] * dropout(ge)[g,k],  db[c] += sum_g dscores[g,c].  The graphs are split into
// gridDim.y slices (one partial sum per slice, combined with one atomic per weight): a single serial loop over all
// graphs per weight took 2 ms at 4 K graphs.
__global__ void head_wgrad_kernel(const float* __restrict__ dscores, const float* __restrict__ ge, int64_t G, int d, int C,
                                  DropRng rng, float* __restrict__ dW, float* __restrict__ db) {
    const int64_t wt = (int64_t)C * d;
    const int64_t per = (G + gridDim.y - 1) / gridDim.y;
    const int64_t g0 = (int64_t)blockIdx.y * per, g1 = (g0 + per < G) ? g0 + per : G;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < wt; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e / d), k = (int)(e % d);
        float acc = 0.0f, bacc = 0.0f;
        for (int64_t g = g0; g < g1; ++g) {
            const float ds = dscores[g * C + c];
            acc = fmaf(ds, ge[g * d + k] * rng_dropout_mult(rng.keys, (uint64_t)(g * d + k), rng.thr, rng.scale), acc);
            bacc += ds;
        }
        atomicAdd(dW + e, acc);
        if (k == 0) atomicAdd(db + c, bacc);
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

extern "C" int u2gnn_head_fwd(const float* ge, int64_t G, int d, const float* W, const float* b, int C, uint64_t seed,
                              uint32_t rng_stream, int thr, float* scores, int accumulate, u2gnn_stream_t stream) {
    if (!ge || !W || !b || !scores || G < 0 || d <= 0 || C <= 0 || thr < 0 || thr > 255) return U2GNN_EINVAL;
    if (G == 0) return U2GNN_OK;
    head_fwd_kernel<<<grid_for(G * C, 128, 8), 128, 0, as_stream(stream)>>>(ge, G, d, W, b, C, make_rng(seed, rng_stream, thr),
                                                                           scores, accumulate);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_soft_ce_fwd_bwd(const float* scores, const int64_t* labels, int64_t G, int C, float smoothing,
                                     int64_t G_total, float* loss, float* dscores, u2gnn_stream_t stream) {
    if (!scores || !labels || !loss || !dscores || G < 0 || C < 2 || G_total < 1) return U2GNN_EINVAL;
    cudaMemsetAsync(loss, 0, sizeof(float), as_stream(stream));
    if (G == 0) return U2GNN_OK;
    soft_ce_kernel<<<grid_for(G, 128, 8), 128, 0, as_stream(stream)>>>(scores, labels, G, C, smoothing,
                                                                      1.0f / (float)G_total, loss, dscores);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_head_bwd(const float* dscores, const float* ge, int64_t G, int d, const float* W, int C,
                              uint64_t seed, uint32_t rng_stream, int thr, float* dW, float* db, float* dge,
                              u2gnn_stream_t stream) {
    if (!dscores || !ge || !W || !dW || !db || !dge || G < 0 || d <= 0 || C <= 0 || thr < 0 || thr > 255)
        return U2GNN_EINVAL;
    if (G == 0) return U2GNN_OK;
    head_bwd_kernel<<<grid_for(G * d, 128, 8), 128, 0, as_stream(stream)>>>(dscores, ge, G, d, W, C,
                                                                           make_rng(seed, rng_stream, thr), dge);
    const int64_t wt = (int64_t)C * d;
    int slices = (int)((G + 31) / 32);
    if (slices > 512) slices = 512;
    dim3 grid((unsigned)((wt + 127) / 128), (unsigned)slices);
    head_wgrad_kernel<<<grid, 128, 0, as_stream(stream)>>>(dscores, ge, G, d, C, make_rng(seed, rng_stream, thr), dW, db);
    U2GNN_CHECK_LAUNCH();
}
