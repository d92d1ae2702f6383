// K6: device-side log-uniform candidate sampler, bit-compatible with the reference's sequential
// host sampler (Log_Uniform_Sampler.cpp:57-71 driven by std::default_random_engine(1111) and
// std::uniform_real_distribution<double>, i.e. libstdc++ minstd_rand0 + generate_canonical<double,53>).
//
// The reference draws one value at a time until `size` distinct ids exist.  minstd_rand0 is a pure
// multiplicative LCG (x <- 16807 x mod 2^31-1), so draw t can be recomputed independently from the
// start state by modular exponentiation: g1 = x0 * a^(2t+1), g2 = x0 * a^(2t+2).  A batch of draws
// is generated in parallel, "first occurrence in stream order" is resolved with a hash table of
// minimum draw indices, and an ordered prefix sum finds the draw at which the reference would have
// stopped (its num_tries).  The id SET, the try count and the advanced engine state equal the
// reference's; ids are emitted in first-occurrence order (the reference's list order is libstdc++
// bucket order, which carries no meaning: the loss sums over the set).
#include "common.cuh"

namespace {

constexpr uint32_t kMod = 2147483647u;   // 2^31 - 1
constexpr uint32_t kMul = 16807u;
constexpr int kThreads = 1024;
constexpr int kBatch = 2 * kThreads;     // draws per round, two consecutive draws per thread

__device__ __forceinline__ uint32_t mulmod(uint32_t a, uint32_t b) {
    return (uint32_t)(((uint64_t)a * (uint64_t)b) % (uint64_t)kMod);
}
__device__ uint32_t powmod(uint32_t base, uint64_t e) {
    uint32_t r = 1u;
    while (e) {
        if (e & 1ull) r = mulmod(r, base);
        base = mulmod(base, base);
        e >>= 1;
    }
    return r;
}

// generate_canonical<double,53> over minstd_rand0 (two engine outputs), then the reference's
// value = lround(exp(x * log N)) - 1.  Explicit _rn intrinsics: no FMA contraction.
__device__ __forceinline__ int64_t draw_value(uint32_t g1, uint32_t g2, double log_n) {
    const double R = 2147483646.0;
    double sum = __dmul_rn((double)(g1 - 1u), 1.0);
    sum = __dadd_rn(sum, __dmul_rn((double)(g2 - 1u), R));
    double x = __ddiv_rn(sum, __dmul_rn(R, R));
    if (x >= 1.0) x = 0.99999999999999988898;  // nextafter(1, 0)
    return llround(exp(__dmul_rn(x, log_n))) - 1;
}

__device__ __forceinline__ uint32_t hash_slot(uint32_t v, uint32_t mask) { return (v * 0x9E3779B1u) & mask; }

__global__ void __launch_bounds__(kThreads) logu_sample_kernel(double log_n, int64_t size, uint32_t* __restrict__ state,
                                                               int64_t* __restrict__ out_ids,
                                                               int32_t* __restrict__ out_tries,
                                                               uint32_t* __restrict__ keys, uint32_t* __restrict__ minidx,
                                                               uint32_t cap_mask, const uint32_t* __restrict__ exclude) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t s_got, s_stop;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (uint32_t i = tid; i <= cap_mask; i += kThreads) {
        keys[i] = 0xFFFFFFFFu;
        minidx[i] = 0xFFFFFFFFu;
    }
    if (tid == 0) {
        s_got = 0u;
        s_stop = 0xFFFFFFFFu;
    }
    __syncthreads();
    const uint32_t x0 = state[0];
    for (uint32_t base = 0;; base += kBatch) {
        // ---- two consecutive draws per thread
        uint32_t t0 = base + 2u * tid;
        uint32_t a = mulmod(x0, powmod(kMul, 2ull * t0 + 1ull));
        uint32_t g[4];
        g[0] = a;
        g[1] = mulmod(g[0], kMul);
        g[2] = mulmod(g[1], kMul);
        g[3] = mulmod(g[2], kMul);
        uint32_t v[2], slot[2];
        bool skip[2];
        v[0] = (uint32_t)draw_value(g[0], g[1], log_n);
        v[1] = (uint32_t)draw_value(g[2], g[3], log_n);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            // sample_unique (Log_Uniform_Sampler.cpp:73-88): a draw that hits a label consumes the stream and is dropped
            skip[k] = exclude && ((exclude[v[k] >> 5] >> (v[k] & 31u)) & 1u);
            slot[k] = 0u;
            if (skip[k]) continue;
            uint32_t s = hash_slot(v[k], cap_mask);
            while (true) {
                uint32_t prev = atomicCAS(keys + s, 0xFFFFFFFFu, v[k]);
                if (prev == 0xFFFFFFFFu || prev == v[k]) break;
                s = (s + 1u) & cap_mask;
            }
            atomicMin(minidx + s, t0 + k);
            slot[k] = s;
        }
        __syncthreads();
        // ---- first occurrence flags and ordered prefix sum
        const uint32_t f0 = (!skip[0] && minidx[slot[0]] == t0) ? 1u : 0u;
        const uint32_t f1 = (!skip[1] && minidx[slot[1]] == t0 + 1u) ? 1u : 0u;
        uint32_t incl = f0 + f1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        const uint32_t got = s_got;
        const uint32_t excl = got + (wid ? warp_tot[wid - 1] : 0u) + incl - (f0 + f1);
        const uint32_t batch_total = warp_tot[31];
        // rank of each flagged draw among accepted ids; the draw with rank size-1 ends the stream
        if (f0 && excl < (uint32_t)size) {
            out_ids[excl] = (int64_t)v[0];
            if (excl == (uint32_t)size - 1u) s_stop = t0;
        }
        if (f1 && excl + f0 < (uint32_t)size) {
            out_ids[excl + f0] = (int64_t)v[1];
            if (excl + f0 == (uint32_t)size - 1u) s_stop = t0 + 1u;
        }
        __syncthreads();
        if (s_stop != 0xFFFFFFFFu) break;
        if (tid == 0) s_got = got + batch_total;
        __syncthreads();
    }
    if (tid == 0) {
        const uint32_t tries = s_stop + 1u;
        out_tries[0] = (int32_t)tries;
        state[0] = mulmod(x0, powmod(kMul, 2ull * tries));
    }
}

__global__ void expected_count_kernel(double log_np1, const int32_t* __restrict__ tries, const int64_t* __restrict__ ids,
                                      int64_t n, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double idx = (double)ids[i];
        const float p = (float)((log(idx + 2.0) - log(idx + 1.0)) / log_np1);       // Log_Uniform_Sampler.cpp:14
        out[i] = (float)(-expm1((double)tries[0] * log1p((double)(-p))));           // :28 (float prob, double math)
    }
}

// bit v of the map = id v is excluded (one of the labels)
__global__ void exclude_bitmap_kernel(const int64_t* __restrict__ labels, int64_t n, int64_t range_max, uint32_t* __restrict__ map) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = labels[i];
        if (v >= 0 && v < range_max) atomicOr(map + (v >> 5), 1u << (v & 31));
    }
}

uint32_t table_capacity(int64_t size) {
    uint64_t need = 4ull * (uint64_t)(size + kBatch);
    uint64_t cap = 1024;
    while (cap < need) cap <<= 1;
    return (uint32_t)cap;
}

}  // namespace

extern "C" size_t u2gnn_logu_sample_workspace_bytes(int64_t size) {
    if (size < 0) return 0;
    return (size_t)table_capacity(size) * 2 * sizeof(uint32_t);
}

extern "C" int u2gnn_logu_sample(int64_t range_max, int64_t size, uint32_t* state_inout, int64_t* out_ids,
                                 int32_t* out_tries, void* workspace, size_t workspace_bytes, u2gnn_stream_t stream) {
    if (!state_inout || !out_ids || !out_tries || !workspace) return U2GNN_EINVAL;
    if (range_max < 1 || range_max >= (1ll << 31) || size < 1) return U2GNN_EINVAL;
    if (size > range_max) return U2GNN_EINVAL;  // the reference would loop forever
    if (size > (1ll << 26)) return U2GNN_EUNSUPPORTED;
    if (workspace_bytes < u2gnn_logu_sample_workspace_bytes(size)) return U2GNN_EWORKSPACE;
    const uint32_t cap = table_capacity(size);
    uint32_t* keys = static_cast<uint32_t*>(workspace);
    uint32_t* minidx = keys + cap;
    // log(N) is evaluated on the host with the same libm call as the reference (:60)
    const double log_n = log((double)range_max);
    logu_sample_kernel<<<1, kThreads, 0, as_stream(stream)>>>(log_n, size, state_inout, out_ids, out_tries, keys, minidx,
                                                              cap - 1u, nullptr);
    U2GNN_CHECK_LAUNCH();
}

extern "C" size_t u2gnn_logu_sample_unique_workspace_bytes(int64_t range_max, int64_t size) {
    if (size < 0 || range_max < 1) return 0;
    return u2gnn_logu_sample_workspace_bytes(size) + (size_t)((range_max + 31) / 32) * sizeof(uint32_t) + sizeof(int32_t);
}

extern "C" int u2gnn_logu_sample_unique(int64_t range_max, int64_t size, const int64_t* labels, int64_t n_labels,
                                        uint32_t* state_inout, int64_t* out_ids, void* workspace, size_t workspace_bytes,
                                        u2gnn_stream_t stream) {
    if (!state_inout || !out_ids || !workspace || (!labels && n_labels > 0) || n_labels < 0) return U2GNN_EINVAL;
    if (range_max < 1 || range_max >= (1ll << 31) || size < 1) return U2GNN_EINVAL;
    if (size + n_labels > range_max) return U2GNN_EINVAL;  // fewer than `size` ids might remain: the reference would loop forever
    if (size > (1ll << 26)) return U2GNN_EUNSUPPORTED;
    if (workspace_bytes < u2gnn_logu_sample_unique_workspace_bytes(range_max, size)) return U2GNN_EWORKSPACE;
    const uint32_t cap = table_capacity(size);
    uint32_t* keys = static_cast<uint32_t*>(workspace);
    uint32_t* minidx = keys + cap;
    uint32_t* map = minidx + cap;
    int32_t* tries = reinterpret_cast<int32_t*>(map + (range_max + 31) / 32);
    cudaStream_t st = as_stream(stream);
    if (cudaMemsetAsync(map, 0, (size_t)((range_max + 31) / 32) * sizeof(uint32_t), st) != cudaSuccess) return U2GNN_ELAUNCH;
    if (n_labels > 0) exclude_bitmap_kernel<<<grid_for(n_labels, 256, 4), 256, 0, st>>>(labels, n_labels, range_max, map);
    logu_sample_kernel<<<1, kThreads, 0, st>>>(log((double)range_max), size, state_inout, out_ids, tries, keys, minidx, cap - 1u, map);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_logu_expected_count(int64_t range_max, const int32_t* tries, const int64_t* ids, int64_t n,
                                         float* out, u2gnn_stream_t stream) {
    if (!tries || !ids || !out || n < 0 || range_max < 1) return U2GNN_EINVAL;
    if (n == 0) return U2GNN_OK;
    expected_count_kernel<<<grid_for(n, 256, 4), 256, 0, as_stream(stream)>>>(log((double)range_max + 1.0), tries, ids, n, out);
    U2GNN_CHECK_LAUNCH();
}
