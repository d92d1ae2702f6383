// fp32 building blocks of the exact-precision (1e-4) path: a tiled CUDA-core SGEMM with fused
// epilogues (bias / ReLU / dropout / masked backward / split-K accumulation) and a column sum.
// The bf16 tensor-core path (ffn_tc.cu) is the throughput path; this one is the parity path and
// the on-device reference for it.
#include "common.cuh"
#include "rng.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

struct Epi {
    const float* bias;
    int flags;
    RngKeys keys;
    int64_t row0;
    int thr;
    float drop_scale;
    const float* aux;
    int64_t ldaux;
    float aux_scale;
};

// 256 threads, each owning a 4x4 block of the 64x64 output tile.  Operands are staged in shared
// memory K-major ([k][m] and [k][n]) so the inner loop is two conflict-free LDS.128 per 16 FMAs.
__global__ void __launch_bounds__(256) sgemm_kernel(int ta, int tb, int64_t M, int N, int64_t K, float alpha,
                                                    const float* __restrict__ A, int64_t lda,
                                                    const float* __restrict__ B, int64_t ldb, float beta,
                                                    float* __restrict__ C, int64_t ldc, Epi epi, int64_t k_per_slice,
                                                    int64_t m_base) {
    __shared__ __align__(16) float As[BK][BM + PAD];
    __shared__ __align__(16) float Bs[BK][BN + PAD];
    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;
    const int64_t m0 = m_base + (int64_t)blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;
    const int64_t kb = (int64_t)blockIdx.z * k_per_slice;
    const int64_t ke = (kb + k_per_slice < K) ? kb + k_per_slice : K;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    for (int64_t k0 = kb; k0 < ke; k0 += BK) {
        // ---- stage A tile: op(A)[m0..m0+63][k0..k0+15]
        if (!ta) {
            const int kk = t & 15;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int mm = (t >> 4) + 16 * i;
                const int64_t gm = m0 + mm, gk = k0 + kk;
                As[kk][mm] = (gm < M && gk < ke) ? __ldg(A + gm * lda + gk) : 0.0f;
            }
        } else {
            const int mm = t & 63;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int kk = (t >> 6) + 4 * i;
                const int64_t gm = m0 + mm, gk = k0 + kk;
                As[kk][mm] = (gm < M && gk < ke) ? __ldg(A + gk * lda + gm) : 0.0f;
            }
        }
        // ---- stage B tile: op(B)[k0..k0+15][n0..n0+63]
        if (!tb) {
            const int nn = t & 63;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int kk = (t >> 6) + 4 * i;
                const int64_t gk = k0 + kk;
                const int gn = n0 + nn;
                Bs[kk][nn] = (gn < N && gk < ke) ? __ldg(B + gk * ldb + gn) : 0.0f;
            }
        } else {
            const int kk = t & 15;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int nn = (t >> 4) + 16 * i;
                const int64_t gk = k0 + kk;
                const int gn = n0 + nn;
                Bs[kk][nn] = (gn < N && gk < ke) ? __ldg(B + (int64_t)gn * ldb + gk) : 0.0f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    const bool first_slice = (blockIdx.z == 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = alpha * acc[i][j];
            if (epi.flags & 16) {  // split-K accumulation: bias once, no nonlinear epilogue
                if ((epi.flags & 1) && first_slice) v += epi.bias[gn];
                atomicAdd(C + gm * ldc + gn, v);
                continue;
            }
            if (epi.flags & 1) v += epi.bias[gn];
            if (epi.flags & 2) v = fmaxf(v, 0.0f);
            if (epi.flags & 4) v *= rng_dropout_mult(epi.keys, (uint64_t)(epi.row0 + gm) * (uint64_t)N + (uint64_t)gn, epi.thr, epi.drop_scale);
            if (epi.flags & 8) v = (epi.aux[gm * epi.ldaux + gn] > 0.0f) ? v * epi.aux_scale : 0.0f;
            if (beta != 0.0f) v += beta * C[gm * ldc + gn];
            C[gm * ldc + gn] = v;
        }
    }
}

__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ A, int64_t M, int N, int64_t lda,
                                                     float* __restrict__ out, int64_t rows_per_block) {
    // block = 256 threads = 8 row-lanes x 32 columns per pass; partial sums reduced in shared memory
    __shared__ float red[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r1 = (r0 + rows_per_block < M) ? r0 + rows_per_block : M;
    const int col = blockIdx.x * 32 + cx;
    float acc = 0.0f;
    if (col < N)
        for (int64_t r = r0 + ry; r < r1; r += 8) acc += __ldg(A + r * lda + col);
    red[ry][cx] = acc;
    __syncthreads();
    if (ry == 0 && col < N) {
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i][cx];
        atomicAdd(out + col, s);
    }
}

}  // namespace

extern "C" int u2gnn_sgemm(int ta, int tb, int64_t M, int N, int64_t K, float alpha, const float* A, int64_t lda,
                           const float* B, int64_t ldb, float beta, float* C, int64_t ldc, const float* bias, int epi,
                           uint64_t seed, uint32_t rng_stream, int thr, int64_t rng_row0, const float* aux,
                           int64_t ldaux, float aux_scale, int splitk, u2gnn_stream_t stream) {
    if (!A || !B || !C || M < 0 || N <= 0 || K < 0) return U2GNN_EINVAL;
    if ((epi & 1) && !bias) return U2GNN_EINVAL;
    if ((epi & 8) && !aux) return U2GNN_EINVAL;
    if (thr < 0 || thr > 255) return U2GNN_EINVAL;
    if (M == 0) return U2GNN_OK;
    if (splitk < 1) splitk = 1;
    if (splitk > 1 && !(epi & 16)) return U2GNN_EINVAL;
    if ((epi & 16) && (epi & (2 | 4 | 8))) return U2GNN_EINVAL;
    if ((epi & 4) && thr == 0) epi &= ~4;
    int64_t k_per_slice = ceil_div64(ceil_div64(K, splitk), BK) * BK;
    if (k_per_slice < BK) k_per_slice = BK;
    const int slices = (int)ceil_div64(K > 0 ? K : 1, k_per_slice);
    Epi e;
    e.bias = bias;
    e.flags = epi;
    e.keys = rng_keys(seed, rng_stream);
    e.thr = thr;
    e.row0 = rng_row0;
    e.drop_scale = thr ? rng_keep_scale(thr) : 1.0f;
    e.aux = aux;
    e.ldaux = ldaux;
    e.aux_scale = aux_scale;
    const int64_t gy = ceil_div64(M, BM);
    if (gy > 65535LL * 32768LL) return U2GNN_EUNSUPPORTED;
    // gridDim.y is limited to 65535: fold very tall problems into several launches
    const int64_t max_y = 65535;
    for (int64_t y0 = 0; y0 < gy; y0 += max_y) {
        const int64_t ny = (gy - y0 < max_y) ? gy - y0 : max_y;
        dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)ny, (unsigned)slices);
        sgemm_kernel<<<grid, 256, 0, as_stream(stream)>>>(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, e,
                                                          k_per_slice, y0 * BM);
    }
    U2GNN_CHECK_LAUNCH();
}

extern "C" int u2gnn_colsum(const float* A, int64_t M, int N, int64_t lda, float* out, int accumulate,
                            u2gnn_stream_t stream) {
    if (!A || !out || N <= 0 || M < 0) return U2GNN_EINVAL;
    if (!accumulate) cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N, as_stream(stream));
    if (M == 0) return U2GNN_OK;
    int64_t row_blocks = ceil_div64(M, 512);
    const int64_t cap = (int64_t)U2GNN_NUM_SMS * 4;
    if (row_blocks > cap) row_blocks = cap;
    const int64_t rows_per_block = ceil_div64(M, row_blocks);
    dim3 grid((unsigned)((N + 31) / 32), (unsigned)ceil_div64(M, rows_per_block));
    colsum_kernel<<<grid, 256, 0, as_stream(stream)>>>(A, M, N, lda, out, rows_per_block);
    U2GNN_CHECK_LAUNCH();
}
