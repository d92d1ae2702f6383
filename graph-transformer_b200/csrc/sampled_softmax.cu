// K7: fused sampled-softmax loss (sampled_softmax.py:36-56): gather of the true and sampled class
// rows, the [N, ns] logit GEMV block, exp / sum / log and the backward, without materialising the
// [N, ns] logits in HBM.  One warp per node; the sampled rows W[ids] are staged once per CTA in
// shared memory (chunked when ns*D does not fit) and reused by every node of the CTA.
// Faithful to the reference: no max-subtraction, no log-Q correction, true class only in the
// denominator if it was sampled.
#include "common.cuh"

namespace {

constexpr int kChunkFloats = 20 * 1024;  // shared-memory budget (floats) per staged chunk

__device__ __forceinline__ int pitch_of(int D) { return D | 1; }

// samp != nullptr: the sampled rows were gathered by the caller (row-sharded table: every rank holds all ns rows in one
// [ns, D] buffer) and ids are not dereferenced into W
__device__ __forceinline__ void stage_rows(const float* __restrict__ W, int64_t V, const int64_t* __restrict__ ids,
                                           const float* __restrict__ samp, int s0, int sc, int D, float* ws) {
    const int P = pitch_of(D);
    for (int e = threadIdx.x; e < sc * D; e += blockDim.x) {
        const int s = e / D, c = e - s * D;
        if (samp) {
            ws[s * P + c] = __ldg(samp + (size_t)(s0 + s) * D + c);
        } else {
            const int64_t id = ids[s0 + s];
            ws[s * P + c] = (id >= 0 && id < V) ? __ldg(W + id * D + c) : 0.0f;
        }
    }
}

// TF = the TF model's variant (tf.nn.sampled_softmax_loss with its defaults, U2GNN_tf/model_U2GNN_Unsup_multi.py:54-58):
//   t_i = x_i.W[y_i] + b[y_i] - log true_q[i],   l_is = x_i.W[s] + b[s] - log samp_q[s]  (-inf where s == y_i: accidental hit),
//   loss_i = log(exp(t_i) + sum_s exp(l_is)) - t_i = log(1 + sum_s exp(l_is - t_i));   denom_out = 1 + sum_s exp(l_is - t_i).
// Shifting by the true logit keeps the exponentials bounded without a second pass.
struct TfArgs {
    const float* bias;     // [V]
    const float* true_q;   // [N]  expected count of each label
    const float* samp_q;   // [ns] expected count of each sampled id
    float* dbias;          // [V]  backward only
};
// rows gathered by the caller + where their gradient goes + the device error word (bit 0: a label outside [0, V))
struct SsExtra {
    const float* samp;     // [ns, D] or null
    float* dsamp;          // [ns, D] or null (backward)
    int* err;              // or null
};

template <bool TF>
__global__ void __launch_bounds__(256) ss_fwd_kernel(const float* __restrict__ x, const int64_t* __restrict__ labels,
                                                     int64_t N, int D, const float* __restrict__ W, int64_t V,
                                                     const int64_t* __restrict__ ids, int ns, int chunk,
                                                     float* __restrict__ loss, float* __restrict__ denom_out,
                                                     int64_t nodes_per_block, TfArgs tf, SsExtra ex) {
    extern __shared__ float sm[];
    const int P = pitch_of(D);
    float* ws = sm;                                   // [chunk][P]
    float* xs = sm + (size_t)chunk * P;               // [warps][D]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, warps = blockDim.x >> 5;
    float* xw = xs + (size_t)w * D;
    float* offs = xs + (size_t)warps * D;             // TF: [chunk] b[id] - log samp_q
    int* sid = reinterpret_cast<int*>(offs + chunk);  // TF: [chunk] sampled ids (V < 2^31) for the accidental-hit test
    const int64_t n0 = (int64_t)blockIdx.x * nodes_per_block;
    const int64_t n1 = (n0 + nodes_per_block < N) ? n0 + nodes_per_block : N;
    for (int s0 = 0; s0 < ns; s0 += chunk) {
        const int sc = (ns - s0 < chunk) ? ns - s0 : chunk;
        __syncthreads();
        stage_rows(W, V, ids, ex.samp, s0, sc, D, ws);
        if (TF) {
            for (int s = threadIdx.x; s < sc; s += blockDim.x) {
                const int64_t id = ids[s0 + s];
                const bool ok = id >= 0 && id < V;
                offs[s] = ok ? __ldg(tf.bias + id) - logf(tf.samp_q[s0 + s]) : -INFINITY;
                sid[s] = ok ? (int)id : -1;
            }
        }
        __syncthreads();
        for (int64_t i = n0 + w; i < n1; i += warps) {
            __syncwarp();
            for (int c = lane; c < D; c += 32) xw[c] = x[i * D + c];
            __syncwarp();
            float t = 0.0f;
            int y = -2;
            if (TF) {
                const int64_t yl = labels[i];
                if (yl < 0 || yl >= V) continue;              // reported below (loss 0, error word)
                y = (int)yl;
                float dot = 0.0f;
                for (int c = lane; c < D; c += 32) dot = fmaf(xw[c], __ldg(W + yl * D + c), dot);
                t = warp_sum(dot) + __ldg(tf.bias + yl) - logf(tf.true_q[i]);
            }
            float part = 0.0f;
            for (int s = lane; s < sc; s += 32) {
                float dot = 0.0f;
                const float* wr = ws + s * P;
                for (int c = 0; c < D; ++c) dot = fmaf(xw[c], wr[c], dot);
                if (TF) part += (sid[s] == y) ? 0.0f : expf(dot + offs[s] - t);
                else part += expf(dot);
            }
            part = warp_sum(part);
            if (lane == 0) denom_out[i] = (s0 == 0) ? part : denom_out[i] + part;
        }
    }
    __syncthreads();
    // true logits and the loss
    for (int64_t i = n0 + w; i < n1; i += warps) {
        const int64_t y = labels[i];
        if (y < 0 || y >= V) {
            // the reference's index_select raises here; the kernel reports through the error word and yields loss 0
            if (lane == 0) {
                loss[i] = 0.0f;
                denom_out[i] = 1.0f;
                if (ex.err) atomicOr(ex.err, 1);
            }
            continue;
        }
        if (TF) {
            if (lane == 0) {
                const float dn = 1.0f + denom_out[i];
                denom_out[i] = dn;
                loss[i] = logf(dn);
            }
            continue;
        }
        float dot = 0.0f;
        for (int c = lane; c < D; c += 32) dot = fmaf(x[i * D + c], __ldg(W + y * D + c), dot);
        dot = warp_sum(dot);
        if (lane == 0) loss[i] = -logf(expf(dot) / denom_out[i]);
    }
}

template <bool TF>
__global__ void __launch_bounds__(256) ss_bwd_kernel(const float* __restrict__ dloss, const float* __restrict__ x,
                                                     const int64_t* __restrict__ labels, int64_t N, int D,
                                                     const float* __restrict__ W, int64_t V,
                                                     const int64_t* __restrict__ ids, int ns, int chunk,
                                                     const float* __restrict__ denom, float* __restrict__ dx,
                                                     float* __restrict__ dW, int64_t nodes_per_block, TfArgs tf, SsExtra ex) {
    extern __shared__ float sm[];
    const int P = pitch_of(D);
    float* ws = sm;                                   // [chunk][P]   sampled rows
    float* dws = ws + (size_t)chunk * P;              // [chunk][P]   their gradient, accumulated per CTA
    float* xs = dws + (size_t)chunk * P;              // [warps][D]
    float* es = xs + (size_t)(blockDim.x >> 5) * D;   // [warps][chunk] per-sample coefficients
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, warps = blockDim.x >> 5;
    float* xw = xs + (size_t)w * D;
    float* ew = es + (size_t)w * chunk;
    float* offs = es + (size_t)warps * chunk;         // TF: [chunk] b[id] - log samp_q
    float* dbs = offs + chunk;                        // TF: [chunk] bias gradient of the staged ids, accumulated per CTA
    int* sid = reinterpret_cast<int*>(dbs + chunk);   // TF: [chunk] sampled ids
    const int64_t n0 = (int64_t)blockIdx.x * nodes_per_block;
    const int64_t n1 = (n0 + nodes_per_block < N) ? n0 + nodes_per_block : N;

    // true-class term: dx = g * W[y];  dW[y] += g * x   (g = -dloss; TF: g = (p_true - 1) dloss, p_true = 1 / denom, db[y] += g)
    for (int64_t i = n0 + w; i < n1; i += warps) {
        const int64_t y = labels[i];
        if (y < 0 || y >= V) {                                // invalid label: no gradient, error word set
            for (int c = lane; c < D; c += 32) dx[i * D + c] = 0.0f;
            if (lane == 0 && ex.err) atomicOr(ex.err, 1);
            continue;
        }
        const float g = TF ? (1.0f / denom[i] - 1.0f) * dloss[i] : -dloss[i];
        for (int c = lane; c < D; c += 32) {
            dx[i * D + c] = g * __ldg(W + y * D + c);
            atomicAdd(dW + y * D + c, g * x[i * D + c]);
        }
        if (TF && lane == 0) atomicAdd(tf.dbias + y, g);
    }
    for (int s0 = 0; s0 < ns; s0 += chunk) {
        const int sc = (ns - s0 < chunk) ? ns - s0 : chunk;
        __syncthreads();
        stage_rows(W, V, ids, ex.samp, s0, sc, D, ws);
        for (int e = threadIdx.x; e < sc * P; e += blockDim.x) dws[e] = 0.0f;
        if (TF) {
            for (int s = threadIdx.x; s < sc; s += blockDim.x) {
                const int64_t id = ids[s0 + s];
                const bool ok = id >= 0 && id < V;
                offs[s] = ok ? __ldg(tf.bias + id) - logf(tf.samp_q[s0 + s]) : -INFINITY;
                sid[s] = ok ? (int)id : -1;
                dbs[s] = 0.0f;
            }
        }
        __syncthreads();
        for (int64_t i = n0 + w; i < n1; i += warps) {
            const int64_t yl = labels[i];
            if (yl < 0 || yl >= V) continue;                  // invalid label (reported above): contributes nothing
            __syncwarp();
            for (int c = lane; c < D; c += 32) xw[c] = x[i * D + c];
            __syncwarp();
            float t = 0.0f;
            int y = -2;
            if (TF) {
                y = (int)yl;
                float dot = 0.0f;
                for (int c = lane; c < D; c += 32) dot = fmaf(xw[c], __ldg(W + yl * D + c), dot);
                t = warp_sum(dot) + __ldg(tf.bias + yl) - logf(tf.true_q[i]);
            }
            const float coef = dloss[i] / denom[i];
            for (int s = lane; s < sc; s += 32) {
                float dot = 0.0f;
                const float* wr = ws + s * P;
                for (int c = 0; c < D; ++c) dot = fmaf(xw[c], wr[c], dot);
                if (TF) {
                    const float e = (sid[s] == y) ? 0.0f : coef * expf(dot + offs[s] - t);
                    ew[s] = e;
                    atomicAdd(dbs + s, e);
                } else {
                    ew[s] = coef * expf(dot);
                }
            }
            __syncwarp();
            // dx[c] += sum_s e_s W_s[c];  dWs[s][c] += e_s x[c]
            for (int c0 = 0; c0 < D; c0 += 32) {
                const int c = c0 + lane;
                if (c < D) {
                    float acc = 0.0f;
                    const float xc = xw[c];
                    for (int s = 0; s < sc; ++s) {
                        const float e = ew[s];
                        acc = fmaf(e, ws[s * P + c], acc);
                        atomicAdd(dws + s * P + c, e * xc);
                    }
                    dx[i * D + c] += acc;
                }
            }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < sc * D; e += blockDim.x) {
            const int s = e / D, c = e - s * D;
            if (ex.dsamp) {
                atomicAdd(ex.dsamp + (size_t)(s0 + s) * D + c, dws[s * P + c]);
            } else {
                const int64_t id = ids[s0 + s];
                if (id >= 0 && id < V) atomicAdd(dW + id * D + c, dws[s * P + c]);
            }
        }
        if (TF) {
            for (int s = threadIdx.x; s < sc; s += blockDim.x)
                if (sid[s] >= 0) atomicAdd(tf.dbias + sid[s], dbs[s]);
        }
    }
}

// Backward for small D (the unsupervised configs: D = d * L = 4 .. 16), reference loss only.  The generic kernel above gives
// one LANE per column, so at D = 4 only 4 of 32 lanes work and every (node, sample, column) costs a shared-memory atomic
// (8.0 ms per 258 K-node step of bench.py cfg4, 45 % of the step).  Here a lane owns SPL samples of the staged chunk: it
// forms e_s = coef * exp(x . w_s) for them, keeps the sampled rows' gradient in REGISTERS across all nodes of its warp
// (acc[SPL][D]) and contributes to dx through one warp reduction per node; the gradient leaves with one global atomic per
// (warp, sample, column).
template <int D, int SPL>
__global__ void __launch_bounds__(256) ss_bwd_small_kernel(const float* __restrict__ dloss, const float* __restrict__ x,
                                                           const int64_t* __restrict__ labels, int64_t N, const float* __restrict__ W,
                                                           int64_t V, const int64_t* __restrict__ ids, int ns,
                                                           const float* __restrict__ denom, float* __restrict__ dx,
                                                           float* __restrict__ dW, int64_t nodes_per_block, SsExtra ex) {
    constexpr int P = D | 1, CHUNK = 32 * SPL;
    __shared__ float ws[CHUNK * P];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const int64_t n0 = (int64_t)blockIdx.x * nodes_per_block;
    const int64_t n1 = (n0 + nodes_per_block < N) ? n0 + nodes_per_block : N;
    for (int s0 = 0; s0 < ns; s0 += CHUNK) {
        const int sc = (ns - s0 < CHUNK) ? ns - s0 : CHUNK;
        __syncthreads();
        stage_rows(W, V, ids, ex.samp, s0, sc, D, ws);
        __syncthreads();
        float acc[SPL][D];
#pragma unroll
        for (int j = 0; j < SPL; ++j)
#pragma unroll
            for (int c = 0; c < D; ++c) acc[j][c] = 0.0f;
        for (int64_t i = n0 + w; i < n1; i += warps) {
            const int64_t y = labels[i];
            const bool bad = y < 0 || y >= V;                 // reported through the error word, contributes nothing
            float xr[D], dxp[D];
#pragma unroll
            for (int c = 0; c < D; ++c) {
                xr[c] = x[i * D + c];
                dxp[c] = 0.0f;
            }
            const float coef = bad ? 0.0f : dloss[i] / denom[i];
#pragma unroll
            for (int j = 0; j < SPL; ++j) {
                const int sl = lane + 32 * j;
                if (sl < sc) {
                    const float* wr = ws + sl * P;
                    float dot = 0.0f;
#pragma unroll
                    for (int c = 0; c < D; ++c) dot = fmaf(xr[c], wr[c], dot);
                    const float e = coef * expf(dot);
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        dxp[c] = fmaf(e, wr[c], dxp[c]);
                        acc[j][c] = fmaf(e, xr[c], acc[j][c]);
                    }
                }
            }
            float dxl = 0.0f, xl = 0.0f;                     // this lane's column (static indexing keeps the arrays in registers)
#pragma unroll
            for (int c = 0; c < D; ++c) {
                const float t = warp_sum(dxp[c]);
                if (lane == c) {
                    dxl = t;
                    xl = xr[c];
                }
            }
            if (s0 == 0) {
                // true-class term (first chunk only): dx = g W[y], dW[y] += g x, g = -dloss
                if (bad) {
                    if (lane == 0 && ex.err) atomicOr(ex.err, 1);
                    if (lane < D) dx[i * D + lane] = 0.0f;
                } else if (lane < D) {
                    const float g = -dloss[i];
                    dx[i * D + lane] = g * __ldg(W + y * D + lane) + dxl;
                    atomicAdd(dW + y * D + lane, g * xl);
                }
            } else if (!bad && lane < D) {
                dx[i * D + lane] += dxl;
            }
        }
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
            const int sl = lane + 32 * j;
            if (sl < sc) {
                if (ex.dsamp) {
#pragma unroll
                    for (int c = 0; c < D; ++c) atomicAdd(ex.dsamp + (size_t)(s0 + sl) * D + c, acc[j][c]);
                } else {
                    const int64_t id = ids[s0 + sl];
                    if (id >= 0 && id < V) {
#pragma unroll
                        for (int c = 0; c < D; ++c) atomicAdd(dW + id * D + c, acc[j][c]);
                    }
                }
            }
        }
    }
}

int pick_chunk(int ns, int D, int copies) {
    const int P = D | 1;
    int chunk = kChunkFloats / (P * copies);
    if (chunk > ns) chunk = ns;
    if (chunk < 1) chunk = 1;
    return chunk;
}

}  // namespace

namespace {

int launch_ss_fwd(bool is_tf, const float* x, const int64_t* labels, int64_t N, int D, const float* W, int64_t V, const int64_t* ids,
                  int ns, float* loss, float* denom, TfArgs tf, SsExtra ex, cudaStream_t st) {
    if (D > 1024) return U2GNN_EUNSUPPORTED;
    if (N == 0) return U2GNN_OK;
    const int chunk = pick_chunk(ns, D, 1);
    const int threads = 256;
    const size_t smem = ((size_t)chunk * (D | 1) + (size_t)(threads / 32) * D + (size_t)2 * chunk) * sizeof(float);
    const int grid = grid_for(N, 64, 2);
    const int64_t npb = ceil_div64(N, grid);
    const unsigned blocks = (unsigned)ceil_div64(N, npb);
    if (is_tf) {
        cudaFuncSetAttribute(ss_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ss_fwd_kernel<true><<<blocks, threads, smem, st>>>(x, labels, N, D, W, V, ids, ns, chunk, loss, denom, npb, tf, ex);
    } else {
        cudaFuncSetAttribute(ss_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ss_fwd_kernel<false><<<blocks, threads, smem, st>>>(x, labels, N, D, W, V, ids, ns, chunk, loss, denom, npb, tf, ex);
    }
    U2GNN_CHECK_LAUNCH();
}

int launch_ss_bwd(bool is_tf, const float* dloss, const float* x, const int64_t* labels, int64_t N, int D, const float* W, int64_t V,
                  const int64_t* ids, int ns, const float* denom, float* dx, float* dW, TfArgs tf, SsExtra ex, cudaStream_t st) {
    if (D > 1024) return U2GNN_EUNSUPPORTED;
    if (N == 0) return U2GNN_OK;
    if (!is_tf && (D == 4 || D == 8 || D == 16)) {
        const int grid = grid_for(N, 64, 2);
        const int64_t npb = ceil_div64(N, grid);
        const unsigned blocks = (unsigned)ceil_div64(N, npb);
        if (D == 4) ss_bwd_small_kernel<4, 16><<<blocks, 256, 0, st>>>(dloss, x, labels, N, W, V, ids, ns, denom, dx, dW, npb, ex);
        else if (D == 8) ss_bwd_small_kernel<8, 8><<<blocks, 256, 0, st>>>(dloss, x, labels, N, W, V, ids, ns, denom, dx, dW, npb, ex);
        else ss_bwd_small_kernel<16, 4><<<blocks, 256, 0, st>>>(dloss, x, labels, N, W, V, ids, ns, denom, dx, dW, npb, ex);
        U2GNN_CHECK_LAUNCH();
    }
    const int threads = 256, warps = threads / 32;
    int chunk = pick_chunk(ns, D, 2);
    // es needs warps*chunk floats on top of the two staged copies (+ 3 chunk for the TF offsets / bias gradients / ids)
    auto floats = [&](int ch) { return (size_t)2 * ch * (D | 1) + (size_t)warps * (D + ch) + (size_t)3 * ch; };
    while (chunk > 1 && floats(chunk) > (size_t)(48 * 1024)) chunk /= 2;
    const size_t smem = floats(chunk) * sizeof(float);
    const int grid = grid_for(N, 64, 1);
    const int64_t npb = ceil_div64(N, grid);
    const unsigned blocks = (unsigned)ceil_div64(N, npb);
    if (is_tf) {
        cudaFuncSetAttribute(ss_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ss_bwd_kernel<true><<<blocks, threads, smem, st>>>(dloss, x, labels, N, D, W, V, ids, ns, chunk, denom, dx, dW, npb, tf, ex);
    } else {
        cudaFuncSetAttribute(ss_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ss_bwd_kernel<false><<<blocks, threads, smem, st>>>(dloss, x, labels, N, D, W, V, ids, ns, chunk, denom, dx, dW, npb, tf, ex);
    }
    U2GNN_CHECK_LAUNCH();
}

}  // namespace

extern "C" int u2gnn_sampled_softmax_fwd(const float* x, const int64_t* labels, int64_t N, int D, const float* W,
                                         int64_t V, const int64_t* ids, int ns, const float* samp_rows, float* loss, float* denom,
                                         int* err, u2gnn_stream_t stream) {
    if (!x || !labels || !W || !ids || !loss || !denom || N < 0 || D <= 0 || V <= 0 || ns <= 0) return U2GNN_EINVAL;
    return launch_ss_fwd(false, x, labels, N, D, W, V, ids, ns, loss, denom, TfArgs{nullptr, nullptr, nullptr, nullptr},
                         SsExtra{samp_rows, nullptr, err}, as_stream(stream));
}

extern "C" int u2gnn_sampled_softmax_bwd(const float* dloss, const float* x, const int64_t* labels, int64_t N, int D,
                                         const float* W, int64_t V, const int64_t* ids, int ns, const float* samp_rows,
                                         const float* denom, float* dx, float* dW, float* dsamp, int* err, u2gnn_stream_t stream) {
    if (!dloss || !x || !labels || !W || !ids || !denom || !dx || !dW || N < 0 || D <= 0 || V <= 0 || ns <= 0)
        return U2GNN_EINVAL;
    if ((samp_rows == nullptr) != (dsamp == nullptr)) return U2GNN_EINVAL;     // pre-gathered rows take their gradient back the same way
    return launch_ss_bwd(false, dloss, x, labels, N, D, W, V, ids, ns, denom, dx, dW, TfArgs{nullptr, nullptr, nullptr, nullptr},
                         SsExtra{samp_rows, dsamp, err}, as_stream(stream));
}

// the TF model's loss (bias, log-Q correction, accidental hits removed, label inside the softmax)
extern "C" int u2gnn_sampled_softmax_tf_fwd(const float* x, const int64_t* labels, int64_t N, int D, const float* W, const float* bias,
                                            int64_t V, const int64_t* ids, int ns, const float* true_q, const float* samp_q,
                                            float* loss, float* denom, int* err, u2gnn_stream_t stream) {
    if (!x || !labels || !W || !bias || !ids || !true_q || !samp_q || !loss || !denom || N < 0 || D <= 0 || V <= 0 || ns <= 0)
        return U2GNN_EINVAL;
    if (V > 2147483647LL) return U2GNN_EUNSUPPORTED;
    return launch_ss_fwd(true, x, labels, N, D, W, V, ids, ns, loss, denom, TfArgs{bias, true_q, samp_q, nullptr},
                         SsExtra{nullptr, nullptr, err}, as_stream(stream));
}

extern "C" int u2gnn_sampled_softmax_tf_bwd(const float* dloss, const float* x, const int64_t* labels, int64_t N, int D,
                                            const float* W, const float* bias, int64_t V, const int64_t* ids, int ns,
                                            const float* true_q, const float* samp_q, const float* denom, float* dx, float* dW,
                                            float* dbias, int* err, u2gnn_stream_t stream) {
    if (!dloss || !x || !labels || !W || !bias || !ids || !true_q || !samp_q || !denom || !dx || !dW || !dbias || N < 0 || D <= 0 ||
        V <= 0 || ns <= 0)
        return U2GNN_EINVAL;
    if (V > 2147483647LL) return U2GNN_EUNSUPPORTED;
    return launch_ss_bwd(true, dloss, x, labels, N, D, W, V, ids, ns, denom, dx, dW, TfArgs{bias, true_q, samp_q, dbias},
                         SsExtra{nullptr, nullptr, err}, as_stream(stream));
}
