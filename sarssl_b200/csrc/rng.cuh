// Counter-based dropout RNG.  The reference uses torch.nn.Dropout (Philox, conformer/feed_forward.py:51,53,
// attention.py:98,151, convolution.py:145); its stream cannot be matched bit for bit, so the CUDA path uses its own
// stateless generator keyed by (seed, element offset): the backward pass regenerates the forward mask instead of
// storing it.  keep probability = 1 - p; kept values are scaled by 1/(1-p) by the caller.
#pragma once
#include <stdint.h>

namespace sarssl {

__host__ __device__ __forceinline__ uint32_t mix_hash(unsigned long long seed, unsigned long long idx) {
    unsigned long long x = idx + seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;       // splitmix64 finaliser
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (uint32_t)(x >> 32);
}

__host__ __device__ __forceinline__ bool keep_mask(unsigned long long seed, unsigned long long idx, float p) {
    // uniform in [0,1) with 24 bits; drop when u < p
    return (float)(mix_hash(seed, idx) >> 8) * (1.0f / 16777216.0f) >= p;
}

}  // namespace sarssl
