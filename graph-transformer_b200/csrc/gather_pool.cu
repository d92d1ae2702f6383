// K1 (row gather / scatter-add) and K4 (CSR segmented sum-pool).  Pure data movement: HBM-bound.
// Algorithmic bytes per gathered row: 8 (int64 index) + 4d (read) + 4d (write).
#include "common.cuh"

namespace {

// One row is moved by a group of `lanes_per_row` threads with 128-bit accesses when d % 4 == 0
// (VEC = 4), otherwise with scalar accesses.  Rows are independent; the index is read once per
// group and broadcast by the hardware.
template <int VEC>
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ table, int64_t n_table, int d,
                                                          const int64_t* __restrict__ idx, int64_t n_idx,
                                                          int64_t idx_stride, float* __restrict__ out,
                                                          int lanes_per_row, int* __restrict__ err) {
    const int rows_per_block = blockDim.x / lanes_per_row;
    const int lane = threadIdx.x % lanes_per_row;
    const int sub = threadIdx.x / lanes_per_row;
    const int dv = d / VEC;
    // (four rows per thread in flight, `gridDim.x * rows_per_block` apart, was measured: 97 us against 69 us per 1.1 M rows -
    // the row stores of a warp then spread over four distant regions; one row per iteration stays)
    for (int64_t row = (int64_t)blockIdx.x * rows_per_block + sub; row < n_idx;
         row += (int64_t)gridDim.x * rows_per_block) {
        int64_t src = __ldg(idx + row * idx_stride);
        if (src < 0 || src >= n_table) {          // F.embedding raises here: zero row + error word (never uninitialised output)
            if (VEC == 4) {
                float4* o = reinterpret_cast<float4*>(out + row * d);
                for (int c = lane; c < dv; c += lanes_per_row) o[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                float* o = out + row * d;
                for (int c = lane; c < d; c += lanes_per_row) o[c] = 0.0f;
            }
            if (err && lane == 0) atomicOr(err, 2);
            continue;
        }
        if (VEC == 4) {
            const float4* s = reinterpret_cast<const float4*>(table + src * d);
            float4* o = reinterpret_cast<float4*>(out + row * d);
            for (int c = lane; c < dv; c += lanes_per_row) o[c] = __ldg(s + c);
        } else {
            const float* s = table + src * d;
            float* o = out + row * d;
            for (int c = lane; c < d; c += lanes_per_row) o[c] = __ldg(s + c);
        }
    }
}

__global__ void __launch_bounds__(256) scatter_add_rows_kernel(const float* __restrict__ grad, int64_t n_idx, int d,
                                                               const int64_t* __restrict__ idx, int64_t idx_stride,
                                                               float* __restrict__ dst, int64_t n_dst) {
    const int64_t total = n_idx * d;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t row = e / d;
        int c = (int)(e - row * d);
        int64_t t = __ldg(idx + row * idx_stride);
        if (t < 0 || t >= n_dst) continue;
        atomicAdd(dst + t * d + c, grad[e]);
    }
}

// ---- deterministic transpose of the index list: counting sort by destination row ----
__global__ void idx_count_kernel(const int64_t* __restrict__ idx, int64_t n_idx, int64_t idx_stride, int64_t n_dst,
                                 unsigned long long* __restrict__ counts) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_idx; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t t = idx[i * idx_stride];
        if (t >= 0 && t < n_dst) atomicAdd(counts + t, 1ull);
    }
}

// single-block exclusive scan: chunks of blockDim with a carried running total
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(const unsigned long long* __restrict__ counts, int64_t n,
                                                              int64_t* __restrict__ rowptr,
                                                              unsigned long long* __restrict__ cursor) {
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0ull;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int64_t base = 0; base < n; base += blockDim.x) {
        int64_t i = base + threadIdx.x;
        unsigned long long v = (i < n) ? counts[i] : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            unsigned long long w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        unsigned long long prefix = carry + (wid ? warp_tot[wid - 1] : 0ull) + incl - v;
        if (i < n) {
            rowptr[i] = (int64_t)prefix;
            cursor[i] = prefix;
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) rowptr[n] = (int64_t)carry;
}

__global__ void idx_fill_kernel(const int64_t* __restrict__ idx, int64_t n_idx, int64_t idx_stride, int64_t n_dst,
                                unsigned long long* __restrict__ cursor, int64_t* __restrict__ t_pos) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_idx; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t t = idx[i * idx_stride];
        if (t >= 0 && t < n_dst) {
            unsigned long long slot = atomicAdd(cursor + t, 1ull);
            t_pos[slot] = i;
        }
    }
}

// the fill order inside a bucket is arbitrary; sorting each bucket by source position makes the
// later summation order (and hence the fp32 result) run-to-run identical
__global__ void idx_sort_buckets_kernel(const int64_t* __restrict__ rowptr, int64_t n_dst, int64_t* __restrict__ t_pos) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_dst; r += (int64_t)gridDim.x * blockDim.x) {
        int64_t b = rowptr[r], e = rowptr[r + 1];
        for (int64_t i = b + 1; i < e; ++i) {
            int64_t v = t_pos[i], j = i - 1;
            while (j >= b && t_pos[j] > v) {
                t_pos[j + 1] = t_pos[j];
                --j;
            }
            t_pos[j + 1] = v;
        }
    }
}

__global__ void __launch_bounds__(256) scatter_det_kernel(const float* __restrict__ grad, int d,
                                                          const int64_t* __restrict__ rowptr,
                                                          const int64_t* __restrict__ t_pos, float* __restrict__ dst,
                                                          int64_t n_dst, int accumulate, int lanes_per_row) {
    const int rows_per_block = blockDim.x / lanes_per_row;
    const int lane = threadIdx.x % lanes_per_row, sub = threadIdx.x / lanes_per_row;
    for (int64_t r = (int64_t)blockIdx.x * rows_per_block + sub; r < n_dst; r += (int64_t)gridDim.x * rows_per_block) {
        const int64_t b = rowptr[r], e = rowptr[r + 1];
        for (int c = lane; c < d; c += lanes_per_row) {
            float acc = accumulate ? dst[r * d + c] : 0.0f;
            for (int64_t j = b; j < e; ++j) acc += grad[t_pos[j] * d + c];
            dst[r * d + c] = acc;
        }
    }
}

__global__ void rowptr_from_coo_kernel(const int64_t* __restrict__ rows, int64_t nnz, int64_t G,
                                       int64_t* __restrict__ rowptr) {
    // rows is non-decreasing: rowptr[g] = first position with rows[pos] >= g (binary search per g)
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g <= G; g += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = nnz;
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (rows[mid] < g) lo = mid + 1; else hi = mid;
        }
        rowptr[g] = lo;
    }
}

// one thread group per graph; lanes stride over the feature dimension, nodes are accumulated in
// ascending order (the order of the reference's CPU COO spmm)
__global__ void __launch_bounds__(256) segment_sum_kernel(const float* __restrict__ x, int d,
                                                          const int64_t* __restrict__ rowptr, int64_t G,
                                                          float* __restrict__ out, int lanes_per_row) {
    const int rows_per_block = blockDim.x / lanes_per_row;
    const int lane = threadIdx.x % lanes_per_row, sub = threadIdx.x / lanes_per_row;
    for (int64_t g = (int64_t)blockIdx.x * rows_per_block + sub; g < G; g += (int64_t)gridDim.x * rows_per_block) {
        const int64_t b = rowptr[g], e = rowptr[g + 1];
        for (int c = lane; c < d; c += lanes_per_row) {
            float acc = 0.0f;
            int64_t n = b;
            for (; n + 4 <= e; n += 4) {  // 4 independent loads in flight, summed in order
                float v0 = __ldg(x + n * d + c), v1 = __ldg(x + (n + 1) * d + c);
                float v2 = __ldg(x + (n + 2) * d + c), v3 = __ldg(x + (n + 3) * d + c);
                acc = (((acc + v0) + v1) + v2) + v3;
            }
            for (; n < e; ++n) acc += __ldg(x + n * d + c);
            out[g * d + c] = acc;
        }
    }
}

// d % 4 == 0: 128-bit loads, d / 4 lanes per graph, EIGHT rows in flight per lane (the scalar kernel kept 4 x 4 bytes per lane in
// flight: 2.3 TB/s on 262 144 x 64 rows, latency-bound).  Every column is still summed in ascending node order, so the result
// is bit-identical to the scalar kernel and to the reference's CPU COO spmm.
__global__ void __launch_bounds__(256) segment_sum_vec_kernel(const float* __restrict__ x, int dv, const int64_t* __restrict__ rowptr,
                                                              int64_t G, float* __restrict__ out) {
    const int per_block = blockDim.x / dv;
    const int lane = threadIdx.x % dv, sub = threadIdx.x / dv;
    if (sub >= per_block) return;
    const float4* xv = reinterpret_cast<const float4*>(x);
    for (int64_t g = (int64_t)blockIdx.x * per_block + sub; g < G; g += (int64_t)gridDim.x * per_block) {
        const int64_t b = rowptr[g], e = rowptr[g + 1];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int64_t n = b;
        // sixteen rows in flight per lane, added in ascending order (the result stays bit-identical to the sequential sum): a graph
        // of ~60 nodes is four dependent DRAM round trips instead of eight - the kernel is latency-bound at 25 us per 67 MB
        for (; n + 16 <= e; n += 16) {
            float4 v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = __ldg(xv + (n + u) * dv + lane);
#pragma unroll
            for (int u = 0; u < 16; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
        for (; n + 8 <= e; n += 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(xv + (n + u) * dv + lane);
#pragma unroll
            for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
        for (; n < e; ++n) {
            const float4 v = __ldg(xv + n * dv + lane);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        reinterpret_cast<float4*>(out)[g * dv + lane] = acc;
    }
}

__global__ void __launch_bounds__(256) segment_bcast_kernel(const float* __restrict__ gout, int d,
                                                            const int64_t* __restrict__ rowptr, int64_t G,
                                                            float* __restrict__ gx, int64_t n, int accumulate) {
    // thread per (node, feature); the graph of a node is found by binary search in rowptr
    const int64_t total = n * d;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t node = e / d;
        int c = (int)(e - node * d);
        int64_t lo = 0, hi = G;  // largest g with rowptr[g] <= node
        while (hi - lo > 1) {
            int64_t mid = (lo + hi) >> 1;
            if (rowptr[mid] <= node) lo = mid; else hi = mid;
        }
        float v = gout[lo * d + c];
        gx[e] = accumulate ? gx[e] + v : v;
    }
}

// same, 128-bit pieces (d % 4 == 0): a quarter of the threads, searches and memory instructions
__global__ void __launch_bounds__(256) segment_bcast_vec_kernel(const float* __restrict__ gout, int dv,
                                                                const int64_t* __restrict__ rowptr, int64_t G,
                                                                float* __restrict__ gx, int64_t n, int accumulate) {
    const int64_t total = n * dv;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t node = e / dv;
        const int c = (int)(e - node * dv);
        int64_t lo = 0, hi = G;  // largest g with rowptr[g] <= node
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(rowptr + mid) <= node) lo = mid; else hi = mid;
        }
        float4 v = __ldg(reinterpret_cast<const float4*>(gout + lo * dv * 4) + c);
        float4* o = reinterpret_cast<float4*>(gx) + e;
        if (accumulate) {
            const float4 old = *o;
            v.x = old.x + v.x; v.y = old.y + v.y; v.z = old.z + v.z; v.w = old.w + v.w;
        }
        *o = v;
    }
}

int pick_lanes(int d_units) {
    int l = 1;
    while (l < 32 && l < d_units) l <<= 1;
    return l;
}

}  // namespace

extern "C" int u2gnn_gather_rows(const float* table, int64_t n_table, int d, const int64_t* idx, int64_t n_idx,
                                 int64_t idx_stride, float* out, int* err, u2gnn_stream_t stream) {
    if (!table || !idx || !out || d <= 0 || n_idx < 0 || idx_stride < 1) return U2GNN_EINVAL;
    if (n_idx == 0) return U2GNN_OK;
    const bool vec = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(out)) % 16 == 0);
    const int lanes = pick_lanes(vec ? d / 4 : d);
    const int grid = grid_for(n_idx, 256 / lanes, 8);
    if (vec)
        gather_rows_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(table, n_table, d, idx, n_idx, idx_stride, out, lanes, err);
    else
        gather_rows_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(table, n_table, d, idx, n_idx, idx_stride, out, lanes, err);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_scatter_add_rows(const float* grad, int64_t n_idx, int d, const int64_t* idx, int64_t idx_stride,
                                      float* dst, int64_t n_dst, u2gnn_stream_t stream) {
    if (!grad || !idx || !dst || d <= 0 || n_idx < 0) return U2GNN_EINVAL;
    if (n_idx == 0) return U2GNN_OK;
    scatter_add_rows_kernel<<<grid_for(n_idx * d, 256, 8), 256, 0, as_stream(stream)>>>(grad, n_idx, d, idx, idx_stride,
                                                                                       dst, n_dst);
    U2GNN_CHECK_LAUNCH();
}

extern "C" size_t u2gnn_index_transpose_workspace_bytes(int64_t n_idx, int64_t n_dst) {
    (void)n_idx;
    return (size_t)(2 * (n_dst + 1)) * sizeof(unsigned long long);
}

extern "C" int u2gnn_index_transpose_build(const int64_t* idx, int64_t n_idx, int64_t idx_stride, int64_t n_dst,
                                           int64_t* t_rowptr, int64_t* t_pos, void* workspace, size_t workspace_bytes,
                                           u2gnn_stream_t stream) {
    if (!idx || !t_rowptr || !t_pos || !workspace || n_dst <= 0) return U2GNN_EINVAL;
    if (workspace_bytes < u2gnn_index_transpose_workspace_bytes(n_idx, n_dst)) return U2GNN_EWORKSPACE;
    cudaStream_t s = as_stream(stream);
    unsigned long long* counts = static_cast<unsigned long long*>(workspace);
    unsigned long long* cursor = counts + (n_dst + 1);
    cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * (size_t)(n_dst + 1), s);
    if (n_idx > 0) idx_count_kernel<<<grid_for(n_idx, 256, 8), 256, 0, s>>>(idx, n_idx, idx_stride, n_dst, counts);
    exclusive_scan_kernel<<<1, 1024, 0, s>>>(counts, n_dst, t_rowptr, cursor);
    if (n_idx > 0) {
        idx_fill_kernel<<<grid_for(n_idx, 256, 8), 256, 0, s>>>(idx, n_idx, idx_stride, n_dst, cursor, t_pos);
        idx_sort_buckets_kernel<<<grid_for(n_dst, 256, 8), 256, 0, s>>>(t_rowptr, n_dst, t_pos);
    }
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_scatter_add_rows_det(const float* grad, int d, const int64_t* t_rowptr, const int64_t* t_pos,
                                          float* dst, int64_t n_dst, int accumulate, u2gnn_stream_t stream) {
    if (!grad || !t_rowptr || !t_pos || !dst || d <= 0) return U2GNN_EINVAL;
    if (n_dst == 0) return U2GNN_OK;
    const int lanes = pick_lanes(d);
    scatter_det_kernel<<<grid_for(n_dst, 256 / lanes, 8), 256, 0, as_stream(stream)>>>(grad, d, t_rowptr, t_pos, dst,
                                                                                      n_dst, accumulate, lanes);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_rowptr_from_coo(const int64_t* coo_rows, int64_t nnz, int64_t num_graphs, int64_t* rowptr,
                                     u2gnn_stream_t stream) {
    if (!coo_rows || !rowptr || num_graphs < 0 || nnz < 0) return U2GNN_EINVAL;
    rowptr_from_coo_kernel<<<grid_for(num_graphs + 1, 256, 4), 256, 0, as_stream(stream)>>>(coo_rows, nnz, num_graphs,
                                                                                           rowptr);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_segment_sum(const float* x, int64_t n, int d, const int64_t* rowptr, int64_t num_graphs,
                                 float* out, u2gnn_stream_t stream) {
    (void)n;
    if (!x || !rowptr || !out || d <= 0 || num_graphs < 0) return U2GNN_EINVAL;
    if (num_graphs == 0) return U2GNN_OK;
    if ((d & 3) == 0 && d <= 1024 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
        const int dv = d / 4;
        segment_sum_vec_kernel<<<grid_for(num_graphs, 256 / dv, 8), 256, 0, as_stream(stream)>>>(x, dv, rowptr, num_graphs, out);
        U2GNN_CHECK_LAUNCH();
    }
    const int lanes = pick_lanes(d);
    segment_sum_kernel<<<grid_for(num_graphs, 256 / lanes, 8), 256, 0, as_stream(stream)>>>(x, d, rowptr, num_graphs,
                                                                                           out, lanes);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_segment_sum_bwd(const float* grad_out, int64_t num_graphs, int d, const int64_t* rowptr,
                                     float* grad_x, int64_t n, int accumulate, u2gnn_stream_t stream) {
    if (!grad_out || !rowptr || !grad_x || d <= 0 || num_graphs <= 0) return U2GNN_EINVAL;
    if (n == 0) return U2GNN_OK;
    if ((d & 3) == 0 && ((reinterpret_cast<uintptr_t>(grad_out) | reinterpret_cast<uintptr_t>(grad_x)) & 15) == 0)
        segment_bcast_vec_kernel<<<grid_for(n * (d / 4), 256, 8), 256, 0, as_stream(stream)>>>(grad_out, d / 4, rowptr, num_graphs,
                                                                                              grad_x, n, accumulate);
    else
        segment_bcast_kernel<<<grid_for(n * d, 256, 8), 256, 0, as_stream(stream)>>>(grad_out, d, rowptr, num_graphs,
                                                                                    grad_x, n, accumulate);
    U2GNN_CHECK_LAUNCH();
}
