// Engine dropout stream: counter-based, stateless, bit-sliced.
//
// The reference's dropout masks come from torch's global generator and cannot be reproduced
// (SURVEY.md "Hard parts"), so the engine defines its own stream; oracle/u2gnn_oracle.py restates
// it so train-mode runs can be compared with dropout switched on.
//
//   keys(seed, stream):  k0 = mix32(seed_lo ^ stream*0x9E3779B1), k1 = mix32(seed_hi + stream + 0x7F4A7C15)
//   word(g, plane)    =  mix32( mix32(lo32(g) ^ k0) + hi32(g)*0x9E3779B1 + k1 + plane*0x632BE5AB )
//   element e belongs to group g = e >> 5, bit e & 31.  Its 8-bit random number r has bit j equal
//   to that bit of word(g, j); the element is KEPT iff r >= thr, thr = round(p*256).  The keep-mask
//   for a whole group is evaluated bit-sliced (one comparator over 32 elements); planes below the
//   lowest set bit of thr never matter, so p = 0.5 costs one word per 32 elements.
#pragma once
#include <stdint.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

struct RngKeys {
    uint32_t k0, k1;
};

__host__ __device__ __forceinline__ uint32_t rng_mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7FEB352Du;
    x ^= x >> 15;
    x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}

__host__ __device__ __forceinline__ RngKeys rng_keys(uint64_t seed, uint32_t stream) {
    RngKeys k;
    k.k0 = rng_mix32((uint32_t)seed ^ (stream * 0x9E3779B1u));
    k.k1 = rng_mix32((uint32_t)(seed >> 32) + stream + 0x7F4A7C15u);
    return k;
}

__host__ __device__ __forceinline__ uint32_t rng_word(RngKeys k, uint64_t g, uint32_t plane) {
    uint32_t a = rng_mix32((uint32_t)g ^ k.k0);
    return rng_mix32(a + (uint32_t)(g >> 32) * 0x9E3779B1u + k.k1 + plane * 0x632BE5ABu);
}

// keep-mask (bit i set = element 32*g+i kept) for thr in [1,255]
__host__ __device__ __forceinline__ uint32_t rng_keep_word(RngKeys k, uint64_t g, int thr) {
    uint32_t gt = 0u, eq = 0xFFFFFFFFu;
    int low = 0;
    while (((thr >> low) & 1) == 0) ++low;
    for (int j = 7; j >= low; --j) {
        uint32_t w = rng_word(k, g, (uint32_t)j);
        if ((thr >> j) & 1) {
            eq &= w;
        } else {
            gt |= eq & w;
            eq &= ~w;
        }
    }
    return gt | eq;
}

// same result as rng_keep_word with the lowest set bit of thr (`low`) precomputed by the caller; thr = 128
// (p = 0.5, the encoder's hard-coded dropout) needs a single word
__host__ __device__ __forceinline__ int rng_thr_low(int thr) {
    int low = 0;
    while (thr && ((thr >> low) & 1) == 0) ++low;
    return low;
}
__device__ __forceinline__ uint32_t rng_keep_word_lo(RngKeys k, uint64_t g, int thr, int low) {
    if (low == 7) return rng_word(k, g, 7u);
    uint32_t gt = 0u, eq = 0xFFFFFFFFu;
    for (int j = 7; j >= low; --j) {
        const uint32_t w = rng_word(k, g, (uint32_t)j);
        if ((thr >> j) & 1) {
            eq &= w;
        } else {
            gt |= eq & w;
            eq &= ~w;
        }
    }
    return gt | eq;
}

__host__ __device__ __forceinline__ float rng_keep_scale(int thr) { return 256.0f / (256.0f - (float)thr); }

// scalar query: multiplier (0 or scale) for element e.  thr == 0 -> 1.
__device__ __forceinline__ float rng_dropout_mult(RngKeys k, uint64_t e, int thr, float scale) {
    if (thr == 0) return 1.0f;
    uint32_t w = rng_keep_word(k, e >> 5, thr);
    return ((w >> (e & 31)) & 1u) ? scale : 0.0f;
}
