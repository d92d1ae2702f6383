// Device-side batch builder: replaces the host loop of get_batch_data (train_pytorch_U2GNN_Sup.py:91-119,
// train_pytorch_U2GNN_UnSup.py:96-128) for a batch given as a list of selected graphs of a dataset held in HBM as one
// CSR adjacency.  One thread per (batch node, slot):
//     input_x[b, 0]   = b                                   (batch-local id of the node itself)
//     input_x[b, j>0] = batch-local id of a neighbour drawn uniformly WITH replacement from the node's CSR row
//                       (an isolated node repeats itself, reference :108-112)
//     node_global[b]  = dataset-wide node id  (X_concat = gather_rows(X_all, node_global); input_y of the unsupervised
//                       script, :96-99)
// Neighbours never leave the node's graph, so the batch-local id is  nbr - graph_start[g] + batch_off[g].
// The draw for (b, j) is a pure function of (seed, stream, b * k + j - 1): r = rng_word(keys, e >> 0, plane 0) and
// idx = (r * deg) >> 32 — restated in oracle/u2gnn_oracle.py::sample_neighbors_device_stream, compared bit for bit.
// The reference's own index stream comes from numpy's MT19937 `choice` and is reproduced by the host builder
// (u2gnn_b200/data.py::build_batch), not here.
#include "common.cuh"
#include "rng.cuh"

namespace {

__global__ void __launch_bounds__(256) build_batch_kernel(const int64_t* __restrict__ g_rowptr, const int64_t* __restrict__ g_col,
                                                          const int64_t* __restrict__ graph_start,
                                                          const int64_t* __restrict__ batch_off, int64_t n_graphs, int k,
                                                          RngKeys keys, int64_t* __restrict__ input_x,
                                                          int64_t* __restrict__ node_global) {
    const int S = k + 1;
    const int64_t N = batch_off[n_graphs];
    const int64_t total = N * S;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = e / S;
        const int j = (int)(e - b * S);
        // graph of batch node b: last g with batch_off[g] <= b
        int64_t lo = 0, hi = n_graphs;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (batch_off[mid] <= b) lo = mid; else hi = mid;
        }
        const int64_t shift = batch_off[lo] - graph_start[lo];
        const int64_t v = b - shift;                           // dataset-wide id
        if (j == 0) {
            input_x[e] = b;
            if (node_global) node_global[b] = v;
            continue;
        }
        const int64_t r0 = g_rowptr[v], deg = g_rowptr[v + 1] - r0;
        if (deg <= 0) {
            input_x[e] = b;
            continue;
        }
        const uint64_t ctr = (uint64_t)b * (uint64_t)k + (uint64_t)(j - 1);
        const uint32_t r = rng_word(keys, ctr, 0u);
        const int64_t pick = (int64_t)(((uint64_t)r * (uint64_t)deg) >> 32);
        input_x[e] = g_col[r0 + pick] + shift;
    }
}

}  // namespace

extern "C" int u2gnn_build_batch(const int64_t* g_rowptr, const int64_t* g_col, const int64_t* graph_start,
                                 const int64_t* batch_off, int64_t n_graphs, int64_t n_nodes, int k, uint64_t seed,
                                 uint32_t rng_stream, int64_t* input_x, int64_t* node_global, u2gnn_stream_t stream) {
    if (!g_rowptr || !g_col || !graph_start || !batch_off || !input_x || n_graphs < 0 || n_nodes < 0 || k < 0) return U2GNN_EINVAL;
    if (n_graphs == 0 || n_nodes == 0) return U2GNN_OK;
    const int64_t total = n_nodes * (k + 1);
    build_batch_kernel<<<grid_for(total, 256, 8), 256, 0, as_stream(stream)>>>(g_rowptr, g_col, graph_start, batch_off, n_graphs, k,
                                                                              rng_keys(seed, rng_stream), input_x, node_global);
    U2GNN_CHECK_LAUNCH();
}
