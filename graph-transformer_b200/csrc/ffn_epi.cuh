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
// set.gt masks of packed values) -> 16 pair masks (0xFFFF per set element).  The weight-gradient kernel keeps its own row bits in
// this order when it evaluates the masks itself; the mask words the FFN kernels EXCHANGE use the flag-word order above.
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
// ---- forward-kernel variant (round 2, tools/probe_epi_ops.cu): F2FP, HSET2, LOP3 and PRMT all issue on the ALU pipe (2 cycles per
// warp instruction per scheduler), HFMA2 / HMUL2 / IMAD on the FMA pipe.  The chunk epilogue of the forward kernel was ALU-pipe bound
// (5 ALU + 1 FMA instruction per register pair), so dropout is applied as a MULTIPLICATION there: the keep bits of a pair are isolated
// at bits 14 / 30 of a word - the bf16 pair (2.0 | 0.0, 2.0 | 0.0) - by one IMAD (FMA pipe) and one LOP3, and
//     tn = c * -1 + (-b1),   g = tn * (2 keep) + 0,   h' = relu(g * -1 + 0) = 2 relu(c + b1) keep      (three HFMA2)
// sign(g) is set exactly where the hidden value is live AND kept (t = +-0 gives g = +0), so the mask word is gathered from the sign
// bytes of two g registers by one PRMT + one LOP3 per TWO pairs.  2 h is exact in bf16; the forward's W2c image carries the 0.5.
__device__ __forceinline__ uint32_t fma2(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t fma_relu2(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("fma.rn.relu.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// keep word (bit e = element e kept) -> 16 factor pairs (2.0 kept, 0.0 dropped: 0x4000 / 0x0000 per half).  Pair j needs bit 2j at
// bit 14 and bit 2j + 1 at bit 30: x * (2^(14-2j) + 2^(29-2j)) puts them there, and the two shifted copies of x never overlap (no
// carries) when x holds at most 14 low bits - hence the four pre-masked words.
__host__ __device__ __forceinline__ void keep_factors16(uint32_t keep, uint32_t (&kp)[16]) {
    const uint32_t hi16 = keep >> 16;
    const uint32_t xa = keep & 0x3FFFu, xb = keep & 0xC000u, ya = hi16 & 0x3FFFu, yb = hi16 & 0xC000u;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t cj = (1u << (14 - 2 * j)) + (1u << (29 - 2 * j));
        kp[j] = ((j < 7 ? xa : xb) * cj) & 0x40004000u;
        kp[8 + j] = ((j < 7 ? ya : yb) * cj) & 0x40004000u;
    }
}
// "flag" word: the order in which the sign bytes of the g registers of pairs 2k, 2k + 1 land (prmt 0xFBD9: byte 0 = low half of
// pair 2k, byte 1 = low half of pair 2k + 1, bytes 2 / 3 = their high halves; bit k of each byte).  Element e = 2 j + half of a
// 32-element group sits at bit 16 half + 8 (j & 1) + (j >> 1): both halves of a pair are 16 bits apart, like a split-pair word.
// This is the bit order of the mask words the FFN kernels exchange.
__host__ __device__ __forceinline__ int flag_pos(int e) { return 16 * (e & 1) + 8 * ((e >> 1) & 1) + (e >> 2); }
__host__ __device__ __forceinline__ int flag_elem(int b) { return (b >> 4) + 2 * ((b >> 3) & 1) + 4 * (b & 7); }
__device__ __forceinline__ uint32_t flag_gather(uint32_t g_even, uint32_t g_odd) { return prmt(g_even, g_odd, 0xFBD9u); }
// flag word -> 16 pair masks (0xFFFF per set element)
__device__ __forceinline__ void flag_masks16(uint32_t w, uint32_t (&m)[16]) {
    uint32_t sh[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) sh[s] = w << s;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const uint32_t lo = 8u | (uint32_t)(j & 1), hi = 8u | (uint32_t)(2 + (j & 1));
        m[j] = prmt(sh[7 - (j >> 1)], 0u, (hi << 12) | (hi << 8) | (lo << 4) | lo);
    }
}

// position of element e in a split-pair word, and its inverse
__device__ __forceinline__ int split_pos(int e) { return (e & 1) * 16 + (e >> 1); }
__device__ __forceinline__ int split_elem(int b) { return (b < 16) ? 2 * b : 2 * (b - 16) + 1; }

}  // namespace epi
