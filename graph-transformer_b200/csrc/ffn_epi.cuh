// Packed bf16x2 epilogue arithmetic shared by the fused FFN kernels.  The epilogue warps are the
// bottleneck of those kernels (the tensor pipe produces 128x128 fp32 values per 256 cycles per tile), so
// every per-element step works on register pairs:
//   cvt.rn.bf16x2.f32      two fp32 accumulators -> one packed register          (F2FP.BF16.PACK_AB)
//   fma.rn.relu.bf16x2     relu(x * 1 + bias) on both halves                     (HFMA2.BF16_V2.RELU)
//   set.gt.u32.bf16x2      0xFFFF / 0 per half where h > 0                        (HSET2.BF16_V2.GT)
//   prmt (sign replicate)  dropout keep bits -> 16-bit lane masks, one PRMT per pair
// The hidden activation is therefore h = bf16(bf16(S) + bf16(b1)) clamped at 0 — the same expression in the
// forward, dgrad and wgrad kernels, so the ReLU mask is identical in all three.
#pragma once
#include <stdint.h>

namespace epi {

__device__ __forceinline__ uint32_t cvt2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// (Measured, tools/trace_ffn.py: replacing cvt.rn.bf16x2 by integer rounding, IADD + IADD + PRMT, is SLOWER - all three
// land on the same half-rate ALU pipe as F2FP, LOP3 and PRMT, which is the pipe that bounds the chunk epilogue.)
__device__ __forceinline__ uint32_t relu_bias2(uint32_t x2, uint32_t bias2) {
    uint32_t d;
    asm("fma.rn.relu.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(x2), "r"(0x3F803F80u), "r"(bias2));
    return d;
}
__device__ __forceinline__ uint32_t gt0_mask2(uint32_t h2) {
    uint32_t d;
    asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(h2), "r"(0u));
    return d;
}

// keep word (bit e = element e of a 32-element group kept) -> 16 pair masks (0xFFFF per kept element).
// Element 2j needs keep bit 2j at the sign position of a byte: shift left by s1 = 7 - (2j & 7) puts it at the
// msb of byte j/4; element 2j+1 uses s2 = s1 - 1.  One PRMT with sign replication builds the pair mask.
__device__ __forceinline__ void keep_masks16(uint32_t keep, uint32_t (&m)[16]) {
    uint32_t sh[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) sh[s] = keep << s;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int s1 = 7 - ((2 * j) & 7), s2 = s1 - 1, b = j >> 2;
        const uint32_t lo = 8u | (uint32_t)b, hi = 8u | (uint32_t)(4 + b);
        m[j] = prmt(sh[s1], sh[s2], (hi << 12) | (hi << 8) | (lo << 4) | lo);
    }
}

// "split-pair" word (bit j = element 2j, bit 16 + j = element 2j + 1: what one LOP3 per register pair builds from the
// set.gt masks of packed values) -> 16 pair masks (0xFFFF per set element).  The mask words the FFN kernels exchange use
// this bit order over the 32 hidden units of a group.
__device__ __forceinline__ void split_masks16(uint32_t w, uint32_t (&m)[16]) {
    uint32_t sh[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) sh[s] = w << s;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const uint32_t lo = 8u | (uint32_t)(j >> 3), hi = 8u | (uint32_t)(2 + (j >> 3));
        m[j] = prmt(sh[7 - (j & 7)], 0u, (hi << 12) | (hi << 8) | (lo << 4) | lo);
    }
}
// position of element e in a split-pair word, and its inverse
__device__ __forceinline__ int split_pos(int e) { return (e & 1) * 16 + (e >> 1); }
__device__ __forceinline__ int split_elem(int b) { return (b < 16) ? 2 * b : 2 * (b - 16) + 1; }

}  // namespace epi
