// Throughput of the packed bf16x2 epilogue instructions of the fused FFN kernels, measured the way the chunk epilogue runs them:
// 16 warps per CTA (4 per scheduler), one CTA per SM, 16 independent loop-carried register chains per thread, clock64 around the loop.
// Prints cycles per warp-instruction per scheduler for single ops, pairs of ops (do they share a pipe?) and whole per-thread
// epilogue sequences (today's and candidates).  Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/probe_epi_ops tools/probe_epi_ops.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 256
#define NCH 16

enum Op { CVT, CVT_RELU, FMA_RELU, SETGT, AND, PRMT, MIN2, IMAD, LOP3, SHL, MUL2, SETGT_F, NOPS_ };
static const char* OPN[] = {"cvt.rn.bf16x2.f32 (F2FP)", "cvt.rn.relu.bf16x2.f32", "fma.rn.relu.bf16x2 (HFMA2)", "set.gt.u32.bf16x2 (HSET2)", "and.b32 (LOP3)", "prmt.b32",
                            "min.bf16x2 (HMNMX2)", "mad.lo.u32 (IMAD)", "lop3 3-input", "shl.b32", "mul.rn.bf16x2 (HMUL2)", "set.gt.bf16x2.bf16x2 (HSET2 -> 1.0/0)"};

// every op: x = chain register (read and written), a / b = loop-invariant operands
template <int OP> __device__ __forceinline__ uint32_t apply(uint32_t x, uint32_t a, uint32_t b) {
    uint32_t d;
    if (OP == CVT) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(x)), "f"(__uint_as_float(a)));
    else if (OP == CVT_RELU) asm volatile("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(x)), "f"(__uint_as_float(a)));
    else if (OP == FMA_RELU) asm volatile("fma.rn.relu.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(a), "r"(b));
    else if (OP == SETGT) asm volatile("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(a));
    else if (OP == AND) asm volatile("and.b32 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(a));
    else if (OP == PRMT) asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(a), "r"(b));
    else if (OP == MIN2) asm volatile("min.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(a));
    else if (OP == IMAD) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(a), "r"(b));
    else if (OP == LOP3) asm volatile("lop3.b32 %0, %1, %2, %3, 0xf8;" : "=r"(d) : "r"(x), "r"(a), "r"(b));
    else if (OP == SHL) asm volatile("shl.b32 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(a));
    else if (OP == MUL2) asm volatile("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(a));
    else if (OP == SETGT_F) asm volatile("set.gt.bf16x2.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(a));
    return d;
}

struct Regs { uint32_t r[NCH], q[NCH], km[NCH], bw[NCH]; };

__device__ __forceinline__ void load(Regs& g, const float* in, uint32_t zero) {
    for (int i = 0; i < NCH; ++i) {
        g.r[i] = __float_as_uint(in[threadIdx.x * 32 + i]) | zero;
        g.q[i] = __float_as_uint(in[threadIdx.x * 32 + 16 + i]) | zero;
        g.km[i] = g.q[i] * 2654435761u;
        g.bw[i] = g.r[i] >> 7;
    }
}
__device__ __forceinline__ void finish(const Regs& g, uint32_t acc, uint32_t* out, long long* cyc, long long t0, long long t1) {
    uint32_t s = acc;
    for (int i = 0; i < NCH; ++i) s ^= g.r[i] ^ g.q[i] ^ g.km[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// two independent sets of 16 chains, one per op (A alone when B < 0)
template <int A, int B>
__global__ void __launch_bounds__(512, 1) pair_kernel(const float* __restrict__ in, uint32_t* __restrict__ out, long long* __restrict__ cyc, uint32_t one, uint32_t zero) {
    Regs g; load(g, in, zero);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            g.r[j] = apply<A>(g.r[j], g.bw[j], g.km[j]);
            if (B >= 0) g.q[j] = apply<(B >= 0 ? B : 0)>(g.q[j], g.bw[j], g.km[j]);
        }
    }
    const long long t1 = clock64();
    finish(g, 0, out, cyc, t0, t1);
}

// whole per-thread epilogue sequences over 16 register pairs; the "accumulator" inputs are the previous iteration's packed results
// (loop-carried, so nothing is hoisted); the keep word changes every iteration
template <int SEQ>
__global__ void __launch_bounds__(512, 1) seq_kernel(const float* __restrict__ in, uint32_t* __restrict__ out, long long* __restrict__ cyc, uint32_t one, uint32_t zero) {
    Regs g; load(g, in, zero);
    uint32_t acc = zero, kw = g.km[0];
    const uint32_t c7f = 0x7FFF7FFFu | zero, sel = 0xFDB9u | zero, k1 = 0x3F803F80u | zero;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        kw = apply<IMAD>(kw, g.bw[1], g.bw[2]);                     // stands in for the RNG word
        uint32_t sh[8], km[NCH];
        if (SEQ != 6) {
#pragma unroll
            for (int s = 0; s < 8; ++s) sh[s] = kw << s;            // IMAD.SHL / SHF, the compiler's choice as in the product
        }
        if (SEQ <= 3) {                                              // pair masks by PRMT with sign replication
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                const int s1 = 7 - ((2 * j) & 7), s2 = s1 - 1, b = j >> 2;
                const uint32_t lo = 8u | (uint32_t)b, hi = 8u | (uint32_t)(4 + b);
                km[j] = apply<PRMT>(sh[s1], sh[s2], ((hi << 12) | (hi << 8) | (lo << 4) | lo) | zero);
            }
        }
        if (SEQ == 0 || SEQ == 1 || SEQ == 2) {
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                uint32_t h = apply<FMA_RELU>(apply<CVT>(g.r[j], g.q[j], 0), one, g.bw[j]);
                h = apply<AND>(h, km[j], 0);
                g.r[j] = h;
                if (SEQ == 1) acc = apply<LOP3>(apply<SETGT>(h, zero, 0), (0x00010001u << j) | zero, acc);
            }
            if (SEQ == 2) {
#pragma unroll
                for (int j = 0; j < NCH; j += 2)
                    acc = apply<LOP3>(apply<PRMT>(apply<IMAD>(g.r[j], 1u | zero, c7f), apply<IMAD>(g.r[j + 1], 1u | zero, c7f), sel), (0x01010101u << (j >> 1)) | zero, acc);
            }
        } else if (SEQ == 3) {                                       // bias inside the GEMM: cvt.relu, and; emit as 2
#pragma unroll
            for (int j = 0; j < NCH; ++j) g.r[j] = apply<AND>(apply<CVT_RELU>(g.r[j], g.q[j], 0), km[j], 0);
#pragma unroll
            for (int j = 0; j < NCH; j += 2)
                acc = apply<LOP3>(apply<PRMT>(apply<IMAD>(g.r[j], 1u | zero, c7f), apply<IMAD>(g.r[j + 1], 1u | zero, c7f), sel), (0x01010101u << (j >> 1)) | zero, acc);
        } else if (SEQ == 4 || SEQ == 5) {                           // dropout by multiplication: keep factors 1.0 / 0 from HSET2 on isolated bits
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                const uint32_t bits = apply<AND>(sh[j & 7], (0x00010001u << (j >> 3)) | zero, 0);
                const uint32_t kf = apply<SETGT_F>(bits, zero, 0);                                  // 1.0 where the bit is set (as a positive denormal > 0)
                uint32_t h;
                if (SEQ == 4) h = apply<FMA_RELU>(apply<CVT>(g.r[j], g.q[j], 0), kf, apply<MUL2>(kf, g.bw[j], 0));
                else h = apply<MUL2>(apply<CVT_RELU>(g.r[j], g.q[j], 0), kf, 0);                    // bias inside the GEMM
                g.r[j] = h;
            }
#pragma unroll
            for (int j = 0; j < NCH; j += 2)
                acc = apply<LOP3>(apply<PRMT>(apply<IMAD>(g.r[j], 1u | zero, c7f), apply<IMAD>(g.r[j + 1], 1u | zero, c7f), sel), (0x01010101u << (j >> 1)) | zero, acc);
        } else if (SEQ == 6) {                                       // no dropout, emit today
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                const uint32_t h = apply<FMA_RELU>(apply<CVT>(g.r[j], g.q[j], 0), one, g.bw[j]);
                g.r[j] = h;
                acc = apply<LOP3>(apply<SETGT>(h, zero, 0), (0x00010001u << j) | zero, acc);
            }
        }
    }
    const long long t1 = clock64();
    finish(g, acc ^ kw, out, cyc, t0, t1);
}

static double mean_cycles(long long* cyc) {
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double m = 0;
    for (int i = 0; i < 148; ++i) m += (double)h[i];
    return m / 148 / ITERS;
}

template <int A, int B>
void run_pair(const float* in, uint32_t* out, long long* cyc) {
    for (int rep = 0; rep < 2; ++rep) { pair_kernel<A, B><<<148, 512>>>(in, out, cyc, 0x3F803F80u, 0u); cudaDeviceSynchronize(); }
    const double c = mean_cycles(cyc);
    if (B < 0) printf("%-42s alone: %6.1f cycles per 4 warps x 16 instr = %.2f cycles per warp-instruction per scheduler\n", OPN[A], c, c / 64);
    else printf("%-42s + %-42s %6.1f cycles per 4 warps x (16 + 16) instr (same pipe: ~256, different pipes: ~128)\n", OPN[A], OPN[B >= 0 ? B : 0], c);
}
template <int SEQ>
void run_seq(const float* in, uint32_t* out, long long* cyc, const char* name) {
    for (int rep = 0; rep < 2; ++rep) { seq_kernel<SEQ><<<148, 512>>>(in, out, cyc, 0x3F803F80u, 0u); cudaDeviceSynchronize(); }
    printf("SEQ %d %-110s %6.1f cycles per 4 warps x 32 hidden values  [%s]\n", SEQ, name, mean_cycles(cyc), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float* in; uint32_t* out; long long* cyc;
    cudaMalloc(&in, 512 * 32 * 4); cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    static float h[512 * 32];
    for (int i = 0; i < 512 * 32; ++i) h[i] = (float)((i * 2654435761u) >> 8) / 16777216.0f - 0.5f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run_pair<CVT, -1>(in, out, cyc); run_pair<CVT_RELU, -1>(in, out, cyc); run_pair<FMA_RELU, -1>(in, out, cyc); run_pair<SETGT, -1>(in, out, cyc);
    run_pair<AND, -1>(in, out, cyc); run_pair<PRMT, -1>(in, out, cyc); run_pair<MIN2, -1>(in, out, cyc); run_pair<IMAD, -1>(in, out, cyc);
    run_pair<LOP3, -1>(in, out, cyc); run_pair<SHL, -1>(in, out, cyc); run_pair<MUL2, -1>(in, out, cyc); run_pair<SETGT_F, -1>(in, out, cyc);
    run_pair<CVT, FMA_RELU>(in, out, cyc); run_pair<CVT, SETGT>(in, out, cyc); run_pair<CVT, AND>(in, out, cyc); run_pair<CVT, PRMT>(in, out, cyc);
    run_pair<CVT, IMAD>(in, out, cyc); run_pair<CVT, MIN2>(in, out, cyc);
    run_pair<SETGT, FMA_RELU>(in, out, cyc); run_pair<SETGT, AND>(in, out, cyc); run_pair<SETGT, IMAD>(in, out, cyc); run_pair<SETGT, MIN2>(in, out, cyc);
    run_pair<SETGT, PRMT>(in, out, cyc);
    run_pair<MIN2, AND>(in, out, cyc); run_pair<MIN2, FMA_RELU>(in, out, cyc); run_pair<MIN2, IMAD>(in, out, cyc);
    run_pair<FMA_RELU, AND>(in, out, cyc); run_pair<FMA_RELU, IMAD>(in, out, cyc); run_pair<FMA_RELU, PRMT>(in, out, cyc); run_pair<IMAD, AND>(in, out, cyc);
    run_pair<PRMT, AND>(in, out, cyc); run_pair<MUL2, FMA_RELU>(in, out, cyc); run_pair<MUL2, AND>(in, out, cyc); run_pair<SHL, AND>(in, out, cyc);
    run_seq<0>(in, out, cyc, "today, no emit: 8 shl + 16 prmt; 16 x (cvt, fma.relu, and)");
    run_seq<1>(in, out, cyc, "today, emit:    8 shl + 16 prmt; 16 x (cvt, fma.relu, and, set.gt, lop3)");
    run_seq<2>(in, out, cyc, "emit by flag bytes: 8 shl + 16 prmt; 16 x (cvt, fma.relu, and, mad); 8 x (prmt, lop3)");
    run_seq<3>(in, out, cyc, "bias in the GEMM + flag bytes: 8 shl + 16 prmt; 16 x (cvt.relu, and, mad); 8 x (prmt, lop3)");
    run_seq<4>(in, out, cyc, "dropout by multiplication: 8 shl; 16 x (and, set.gt->1.0, mul, cvt, fma.relu, mad); 8 x (prmt, lop3)");
    run_seq<5>(in, out, cyc, "bias in the GEMM + dropout by multiplication: 8 shl; 16 x (and, set.gt->1.0, cvt.relu, mul, mad); 8 x (prmt, lop3)");
    run_seq<6>(in, out, cyc, "no dropout, emit today: 16 x (cvt, fma.relu, set.gt, lop3)");
    return 0;
}
