// Counter-based dropout RNG.  The reference uses torch.nn.Dropout (Philox, conformer/feed_forward.py:51,53,
// attention.py:98,151, convolution.py:145); its stream cannot be matched bit for bit, so the CUDA path uses its own
// stateless generator keyed by (seed, element offset): the backward pass regenerates the forward mask instead of
// storing it.  One 32-bit hash serves two neighbouring elements (16 random bits each): an element is dropped when its
// 16 bits are below p * 65536 (keep probability 1 - p to within 2^-16); kept values are scaled by 1/(1-p) by the caller.
#pragma once
#include <stdint.h>

namespace sarssl {

// 2 multiplies + 3 xor-shifts on the low word of the pair index ("lowbias32"-style finaliser); the seed and the high word of the index
// enter through `seedmix`, which compilers hoist out of vector loops (it only changes every 2^33 elements).
__host__ __device__ __forceinline__ uint32_t hash_pair(unsigned long long seed, unsigned long long pair_idx) {
    const uint32_t seedmix = (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x9E3779B1u) ^ ((uint32_t)(pair_idx >> 32) * 0x85EBCA6Bu);
    uint32_t x = ((uint32_t)pair_idx * 0xCC9E2D51u) ^ seedmix;
    x ^= x >> 15; x *= 0x2C1B3C6Du;
    x ^= x >> 12; x *= 0x297A2D39u;
    x ^= x >> 15;
    return x;
}

__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) { return (uint32_t)(p * 65536.0f); }

// keep decision for the element at `idx`
__host__ __device__ __forceinline__ bool keep_mask(unsigned long long seed, unsigned long long idx, float p) {
    const uint32_t h = hash_pair(seed, idx >> 1);
    const uint32_t bits = (idx & 1ull) ? (h >> 16) : (h & 0xFFFFu);
    return bits >= drop_threshold(p);
}

// keep decisions for the aligned pair (2*pair_idx, 2*pair_idx + 1): bit 0 / bit 1 of the result
__host__ __device__ __forceinline__ uint32_t keep_pair(unsigned long long seed, unsigned long long pair_idx, uint32_t thr) {
    const uint32_t h = hash_pair(seed, pair_idx);
    return ((h & 0xFFFFu) >= thr ? 1u : 0u) | ((h >> 16) >= thr ? 2u : 0u);
}

}  // namespace sarssl
