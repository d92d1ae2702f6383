// Self-test of the tcgen05 plumbing in tc_common.cuh: one CTA computes a small GEMM through each
// operand path the fused kernels rely on, so descriptor / layout assumptions are pinned on the
// hardware by tests/test_gpu_tc.py before any fused kernel is trusted.
//   mode 0  C[128,N] = A[128,K] * B[N,K]^T     both operands K-major in shared memory (SS)
//   mode 1  C[128,N] = At[K,128]^T * Bt[K,N]   both operands MN-major in shared memory (SS)
//   mode 2  as mode 0 with A staged in tensor memory (TS)
//   mode 3  as mode 0 with B brought in by a 1-D bulk copy of a pre-swizzled image (UBLKCP + mbarrier tx)
//   mode 4  as mode 0 with an fp16 accumulator (idesc c_format = 0); C receives the RAW 32-bit tensor-memory columns
//   mode 5  as mode 2 with A staged in tensor memory as fp16 pairs (a_format = 0) against a bf16 B
#include "../common.cuh"
#include "../tc_common.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int kMaxK = 128;

__global__ void __launch_bounds__(128) tc_selftest_kernel(int mode, const float* __restrict__ A, const float* __restrict__ B,
                                                          float* __restrict__ C, int K, int N, uint8_t* scratch) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // offset on the shared-window address: keeps LDS / STS (a uintptr_t round trip makes every access generic)
    uint8_t* sA = smem;                        // up to 2 tiles of [128 x 64] bf16 (32 KB)
    uint8_t* sB = smem + 32768;                // up to 2 tiles (32 KB)
    __shared__ uint64_t bar_mma, bar_copy;
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5;

    if (warp == 0) tc::tmem_alloc<256>(&tmem_slot);
    if (t == 0) {
        tc::mbar_init(&bar_mma, 1);
        tc::mbar_init(&bar_copy, 1);
        tc::fence_barrier_init();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

    if (mode == 0 || mode == 2 || mode == 3 || mode == 4 || mode == 5) {
        // A: [128][K] fp32 row-major -> K/64 K-major tiles
        if (mode != 2 && mode != 5) {
            for (int e = t; e < 128 * K; e += 128) {
                int r = e / K, c = e % K;
                *reinterpret_cast<__nv_bfloat16*>(sA + (c >> 6) * 16384 + tc::sw128_offset(r, c & 63)) = __float2bfloat16(A[e]);
            }
        } else {
            // A row t -> TMEM lane t, bf16 pairs packed into 32-bit columns [128, 128 + K/2)
            for (int c0 = 0; c0 < K / 2; c0 += 32) {
                uint32_t v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float lo = A[t * K + 2 * (c0 + j)], hi = A[t * K + 2 * (c0 + j) + 1];
                    if (mode == 5) {
                        const __half2 h = __floats2half2_rn(lo, hi);
                        v[j] = *reinterpret_cast<const uint32_t*>(&h);
                    } else {
                        v[j] = tc::pack_bf16(lo, hi);
                    }
                }
                tc::tmem_st32(tmem + lane_base + 128 + c0, v);
            }
            tc::tmem_st_wait();
        }
        // B: [N][K] fp32 -> K-major tiles of [N x 64]
        uint8_t* dstB = (mode == 3) ? scratch : sB;
        for (int e = t; e < N * K; e += 128) {
            int r = e / K, c = e % K;
            *reinterpret_cast<__nv_bfloat16*>(dstB + (c >> 6) * (N * 128) + tc::sw128_offset(r, c & 63)) = __float2bfloat16(B[e]);
        }
        if (mode == 3) {
            __threadfence();
            __syncthreads();
            if (t == 0) {
                const uint32_t bytes = (uint32_t)(N * K * 2);
                tc::mbar_arrive_expect_tx(&bar_copy, bytes);
                tc::bulk_g2s(sB, scratch, bytes, &bar_copy);
            }
            tc::mbar_wait(&bar_copy, 0);
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        if (t == 0) {
            uint32_t idesc = tc::make_idesc(128, N, 0, 0);
            if (mode == 4) idesc &= ~(3u << 4);          // D format: fp16
            if (mode == 5) idesc &= ~(7u << 7);          // A format: fp16
            for (int ks = 0; ks < K / 16; ++ks) {
                const uint32_t koff = (uint32_t)((ks >> 2) * 16384 + (ks & 3) * 32);
                const uint32_t boff = (uint32_t)((ks >> 2) * (N * 128) + (ks & 3) * 32);
                const uint64_t bd = tc::make_desc_sw128(tc::smem_u32(sB) + boff, 16, 1024);
                if (mode == 2 || mode == 5) {
                    tc::mma_ts(tmem, tmem + 128 + ks * 8, bd, idesc, ks > 0);
                } else {
                    const uint64_t ad = tc::make_desc_sw128(tc::smem_u32(sA) + koff, 16, 1024);
                    tc::mma_ss(tmem, ad, bd, idesc, ks > 0);
                }
            }
            tc::mma_commit(&bar_mma);
        }
    } else {
        // mode 1: At[K][128], Bt[K][N] fp32; rows = K index; column tiles of 64 along M / N
        for (int e = t; e < K * 128; e += 128) {
            int r = e / 128, c = e % 128;
            *reinterpret_cast<__nv_bfloat16*>(sA + (c >> 6) * (K * 128) + tc::sw128_offset(r, c & 63)) = __float2bfloat16(A[e]);
        }
        for (int e = t; e < K * N; e += 128) {
            int r = e / N, c = e % N;
            *reinterpret_cast<__nv_bfloat16*>(sB + (c >> 6) * (K * 128) + tc::sw128_offset(r, c & 63)) = __float2bfloat16(B[e]);
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        if (t == 0) {
            const uint32_t idesc = tc::make_idesc(128, N, 1, 1);
            for (int ks = 0; ks < K / 16; ++ks) {
                const uint64_t ad = tc::make_desc_sw128(tc::smem_u32(sA) + ks * 2048, (uint32_t)(K * 128), 1024);
                const uint64_t bd = tc::make_desc_sw128(tc::smem_u32(sB) + ks * 2048, (uint32_t)(K * 128), 1024);
                tc::mma_ss(tmem, ad, bd, idesc, ks > 0);
            }
            tc::mma_commit(&bar_mma);
        }
    }
    tc::mbar_wait(&bar_mma, 0);
    tc::tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tc::tmem_ld32(tmem + lane_base + c0, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) C[t * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

}  // namespace

extern "C" int u2gnn_tc_selftest(int mode, const float* A, const float* B, float* C, int K, int N, void* scratch,
                                 u2gnn_stream_t stream) {
    if (!A || !B || !C || !scratch) return U2GNN_EINVAL;
    if (mode < 0 || mode > 5) return U2GNN_EINVAL;
    if ((N != 64 && N != 128) || K < 16 || K > kMaxK || K % 16) return U2GNN_EINVAL;
    if (mode != 1 && K % 64) return U2GNN_EINVAL;
    const int smem = 65536 + 1024;
    cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    tc_selftest_kernel<<<1, 128, smem, as_stream(stream)>>>(mode, A, B, C, K, N, static_cast<uint8_t*>(scratch));
    U2GNN_CHECK_LAUNCH();
}
