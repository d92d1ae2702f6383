// Shared helpers for the u2gnn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/u2gnn_b200.h"

#define U2GNN_NUM_SMS 148

#define U2GNN_CHECK_LAUNCH()                                      \
    do {                                                          \
        if (cudaPeekAtLastError() != cudaSuccess) {               \
            (void)cudaGetLastError();                             \
            return U2GNN_ELAUNCH;                                 \
        }                                                         \
        return U2GNN_OK;                                          \
    } while (0)

static inline cudaStream_t as_stream(u2gnn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grid size for a grid-stride kernel: a multiple of the SM count, capped by the work available
static inline int grid_for(int64_t work_items, int per_block, int blocks_per_sm) {
    int64_t need = ceil_div64(work_items, per_block);
    int64_t cap = (int64_t)U2GNN_NUM_SMS * blocks_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
