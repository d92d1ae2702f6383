// Micro-probe of tcgen05.mma issue/execute rates (one CTA): cycles per MMA for a given N, operand source
// (SS / TS) and accumulator pattern (same accumulator vs rotating accumulators).  Used to size the fused
// kernels' tiles; not part of the hot path.
#include "../common.cuh"
#include "../tc_common.cuh"

namespace {
__global__ void __launch_bounds__(128) tc_probe_kernel(int N, int ts, int rotate, int count, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // offset on the shared-window address: keeps LDS / STS (a uintptr_t round trip makes every access generic)
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5;
    for (int e = t; e < 65536 / 4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = 0x3C003C00u;
    if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
    if (t == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 1 && rotate >= 2) {
        // whole-warp uniform issue: descriptors stay in uniform registers, k-steps are immediate adds
        const uint32_t idesc = (ts == 2) ? tc::make_idesc(128, N, 1, 1) : tc::make_idesc(128, N, 0, 0);
        const uint64_t ad = tc::make_desc_sw128(tc::smem_u32(smem), 16384, 1024);
        const uint64_t bd = tc::make_desc_sw128(tc::smem_u32(smem + 32768), 16384, 1024);
        const uint32_t at = tmem + 448;
        const long long t0 = clock64();
        for (int i = 0; i < count; i += 4) {
            if (tc::elect_one()) {
                if (ts == 2) {          // both operands MN-major: k-steps advance by 16 rows = 2048 B
                    tc::mma_ss_acc(tmem, ad, bd, idesc);
                    tc::mma_ss_acc(tmem, ad + 128, bd + 128, idesc);
                    tc::mma_ss_acc(tmem, ad + 256, bd + 256, idesc);
                    tc::mma_ss_acc(tmem, ad + 384, bd + 384, idesc);
                } else if (ts) {
                    tc::mma_ts_acc(tmem, at, bd, idesc);
                    tc::mma_ts_acc(tmem, at + 8, bd + 2, idesc);
                    tc::mma_ts_acc(tmem, at + 16, bd + 4, idesc);
                    tc::mma_ts_acc(tmem, at + 24, bd + 6, idesc);
                } else {
                    tc::mma_ss_acc(tmem, ad, bd, idesc);
                    tc::mma_ss_acc(tmem, ad + 2, bd + 2, idesc);
                    tc::mma_ss_acc(tmem, ad + 4, bd + 4, idesc);
                    tc::mma_ss_acc(tmem, ad + 6, bd + 6, idesc);
                }
            }
            __syncwarp();
        }
        const long long t1 = clock64();
        if (tc::elect_one()) tc::mma_commit(&bar);
        __syncwarp();
        tc::mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (t == 32) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    if (t == 0 && rotate < 2) {
        const uint32_t idesc = tc::make_idesc(128, N, 0, 0);
        const uint32_t a0 = tc::smem_u32(smem), b0 = tc::smem_u32(smem + 32768);
        const long long t0 = clock64();
        for (int i = 0; i < count; ++i) {
            const uint32_t d = rotate ? (uint32_t)((i & 1) * 256) : 0u;      // accumulator columns [0,N) or [256,256+N)
            const uint64_t bd = tc::make_desc_sw128(b0 + (i & 3) * 32, 16, 1024);
            if (ts) tc::mma_ts(tmem + d, tmem + 448 + (i & 3) * 8, bd, idesc, 1);
            else tc::mma_ss(tmem + d, tc::make_desc_sw128(a0 + (i & 3) * 32, 16, 1024), bd, idesc, 1);
        }
        const long long t1 = clock64();
        tc::mma_commit(&bar);
        tc::mbar_wait(&bar, 0);
        const long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

// TMEM -> register bandwidth: `warps` warps (warp % 4 = lane quarter) each issue `iters` x (tcgen05.ld.32x32b.x32 + wait)
// over rotating column offsets; out[0] = cycles of the slowest warp 0 measurement, out[1] = bytes moved by the CTA.
__global__ void __launch_bounds__(512) tmem_bw_kernel(int warps, int iters, int batch, long long* out) {
    __shared__ uint32_t tmem_slot;
    __shared__ long long t_max;
    const int t = threadIdx.x, warp = t >> 5;
    if (t == 0) t_max = 0;
    if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < warps) {
        for (int i = 0; i < iters; ++i) {
            uint32_t v[32], w[32];
            const uint32_t col = (uint32_t)(((i * 2 + (warp >> 2)) * 32) & 511);
            tc::tmem_ld32(tmem + lane_base + col, v);
            if (batch > 1) tc::tmem_ld32(tmem + lane_base + ((col + 32) & 511), w);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= v[j];
            if (batch > 1) {
#pragma unroll
                for (int j = 0; j < 32; ++j) acc ^= w[j];
            }
        }
    }
    const long long t1 = clock64();
    atomicMax((unsigned long long*)&t_max, (unsigned long long)(t1 - t0));
    __syncthreads();
    if (t == 0) {
        out[0] = t_max;
        out[1] = (long long)warps * iters * (batch > 1 ? 2 : 1) * 4096;
        out[2] = acc;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tmem);
}
}  // namespace

// ---------------------------------------------------------------------------------------------
// cta_group::2 rate probe: a CLUSTER of two CTAs (one TPC) executes M = 256 MMAs - A rows 0-127 from CTA 0's shared / tensor
// memory, rows 128-255 from CTA 1's, each CTA holding half of B's N columns - issued by one thread of the leader CTA.
// Measures the cycles per instruction as tc_probe_kernel does for cta_group::1 (operands are constants; results unused).
// ---------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mma2_ss_acc(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    asm volatile("tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma2_ts_acc(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc) {
    asm volatile("tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, 1;" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) tc_probe2_kernel(int N, int ts, int count, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5;
    const uint32_t rank = cluster_ctarank();
    for (int e = t; e < 65536 / 4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = 0x3C003C00u;
    if (t == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tc::smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (rank == 0 && warp == 1) {
        const uint32_t idesc = tc::make_idesc(256, N, 0, 0);
        const uint64_t ad = tc::make_desc_sw128(tc::smem_u32(smem), 16, 1024);
        const uint64_t bd = tc::make_desc_sw128(tc::smem_u32(smem + 32768), 16, 1024);
        const uint32_t at = tmem + 448;
        const long long t0 = clock64();
        for (int i = 0; i < count; i += 4) {
            if (tc::elect_one()) {
                if (ts) {
                    mma2_ts_acc(tmem, at, bd, idesc);
                    mma2_ts_acc(tmem, at + 8, bd + 2, idesc);
                    mma2_ts_acc(tmem, at + 16, bd + 4, idesc);
                    mma2_ts_acc(tmem, at + 24, bd + 6, idesc);
                } else {
                    mma2_ss_acc(tmem, ad, bd, idesc);
                    mma2_ss_acc(tmem, ad + 2, bd + 2, idesc);
                    mma2_ss_acc(tmem, ad + 4, bd + 4, idesc);
                    mma2_ss_acc(tmem, ad + 6, bd + 6, idesc);
                }
            }
            __syncwarp();
        }
        const long long t1 = clock64();
        if (tc::elect_one())
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(tc::smem_u32(&bar)),
                         "h"((uint16_t)3)
                         : "memory");
        __syncwarp();
        tc::mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (t == 32) { out[0] = t1 - t0; out[1] = t2 - t0; }
    } else if (rank == 1 && t == 0) {
        tc::mbar_wait(&bar, 0);                  // the multicast commit arrives on the peer's barrier too
    }
    tc::tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
}  // namespace

extern "C" int u2gnn_tc_probe2(int N, int ts, int count, long long* out, u2gnn_stream_t stream) {
    if (!out || N < 32 || N > 256 || N % 32 || count < 4 || ts < 0 || ts > 1) return U2GNN_EINVAL;
    const int smem = 65536 + 1024;
    cudaFuncSetAttribute(tc_probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    tc_probe2_kernel<<<2, 128, smem, as_stream(stream)>>>(N, ts, count, out);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_tmem_bw_probe(int warps, int iters, int batch, long long* out, u2gnn_stream_t stream) {
    if (!out || warps < 1 || warps > 16 || iters < 1) return U2GNN_EINVAL;
    tmem_bw_kernel<<<1, 512, 0, as_stream(stream)>>>(warps, iters, batch, out);
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_tc_probe(int N, int ts, int rotate, int count, long long* out, u2gnn_stream_t stream) {
    if (!out || N < 16 || N > 256 || N % 16 || count < 1 || ts < 0 || ts > 2) return U2GNN_EINVAL;
    const int smem = 65536 + 1024;
    cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    tc_probe_kernel<<<1, 128, smem, as_stream(stream)>>>(N, ts, rotate, count, out);
    U2GNN_CHECK_LAUNCH();
}

// ---------------------------------------------------------------------------------------------
// L2 reduction throughput probe: `groups` CTAs add into the SAME 32 KB tile at about the same time (the access pattern
// of split-K partial sums of a [128 x 64] fp32 tile), tile after tile.  mode 0: red.global.add.v4.f32 (16 B per op),
// mode 1: scalar atomicAdd, mode 2: plain 16-byte stores (bandwidth reference).
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(128) red_probe_kernel(float* __restrict__ buf, int64_t n_tiles, int groups, int mode) {
    const int slice = blockIdx.x / groups, n_slices = gridDim.x / groups;
    const float v = 1.0f;
    for (int64_t tile = slice; tile < n_tiles; tile += n_slices) {
        float* row = buf + tile * 8192 + (size_t)threadIdx.x * 64;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            // lanes of a warp cover consecutive 16-byte pieces: thread t handles piece (j * 128 + t) of the 2048 in the tile
            float* dst = buf + tile * 8192 + (size_t)(j * 128 + threadIdx.x) * 4;
            if (mode == 0) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
            } else if (mode == 1) {
                atomicAdd(dst, v); atomicAdd(dst + 1, v); atomicAdd(dst + 2, v); atomicAdd(dst + 3, v);
            } else if (mode == 2) {
                *reinterpret_cast<float4*>(dst) = make_float4(v, v, v, v);
            } else {
                // thread-per-row pattern (thread t owns row t: 16 pieces 256 B apart across the warp)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + 4 * j), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
            }
        }
    }
}

// bulk reductions (cp.reduce.async.bulk ... add.f32, issued by the copy engine instead of the LSU): the same split-K pattern.
// mode 4: 128 threads, one 256-byte row each per tile (rows staged at a 272-byte stride); mode 5: one 32 KB operation per tile;
// mode 6: ONE warp, 64-row halves through a single staging buffer with a read-wait between halves (the merged FFN backward's
// drain protocol: the wait is the time the staging buffer is blocked).
__device__ __forceinline__ void bulk_red_add_f32(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(tc::smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__global__ void __launch_bounds__(128) red_bulk_probe_kernel(float* __restrict__ buf, int64_t n_tiles, int groups, int mode) {
    extern __shared__ __align__(128) uint8_t stage[];
    for (int e = threadIdx.x; e < 128 * 272 / 4; e += 128) reinterpret_cast<float*>(stage)[e] = 1.0f;
    tc::fence_proxy_async();
    __syncthreads();
    const int slice = blockIdx.x / groups, n_slices = gridDim.x / groups;
    const int t = threadIdx.x;
    for (int64_t tile = slice; tile < n_tiles; tile += n_slices) {
        float* dst = buf + tile * 8192;
        if (mode == 4) {
            bulk_red_add_f32(dst + t * 64, stage + t * 272, 256);
            tc::bulk_commit();
            tc::bulk_wait_read<8>();
        } else if (mode == 5) {
            if (t == 0) {
                bulk_red_add_f32(dst, stage, 32768);
                tc::bulk_commit();
                tc::bulk_wait_read<8>();
            }
        } else if (t < 32) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                bulk_red_add_f32(dst + (h * 64 + t) * 64, stage + t * 272, 256);
                bulk_red_add_f32(dst + (h * 64 + 32 + t) * 64, stage + (32 + t) * 272, 256);
                tc::bulk_commit();
                tc::bulk_wait_read<0>();
                __syncwarp();
            }
        }
    }
    tc::bulk_wait_all<0>();
}
}  // namespace

extern "C" int u2gnn_red_probe(float* buf, int64_t n_tiles, int groups, int mode, u2gnn_stream_t stream) {
    if (!buf || n_tiles < 1 || groups < 1 || groups > U2GNN_NUM_SMS) return U2GNN_EINVAL;
    const int n_slices = U2GNN_NUM_SMS / groups;
    if (mode >= 4) {
        red_bulk_probe_kernel<<<groups * n_slices, 128, 128 * 272, as_stream(stream)>>>(buf, n_tiles, groups, mode);
        U2GNN_CHECK_LAUNCH();
    }
    red_probe_kernel<<<groups * n_slices, 128, 0, as_stream(stream)>>>(buf, n_tiles, groups, mode);
    U2GNN_CHECK_LAUNCH();
}
