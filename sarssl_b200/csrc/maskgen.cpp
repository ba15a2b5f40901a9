// Host-side mask index generator, bit-exact with CPython's `random` module (MT19937).
//
// Reference behaviour (common/utils_module.py:263-267,305-308): per batch item, in order,
//     mask_patch_idx[b] = random.sample(range(npatch), nmasked);  mask_ch_idx[b] = random.randint(0, nmic-1)
// using the process-global `random` state that common/utils.py:51-56 seeds once per epoch.  The algorithms
// restated here are CPython's (Lib/random.py, Modules/_randommodule.c - a third-party dependency of the
// reference, any CPython >= 3.2; checked against 3.12): init_by_array seeding, genrand_uint32,
// getrandbits(k <= 32) = genrand >> (32 - k), _randbelow_with_getrandbits (rejection), and sample()'s two
// branches (partial shuffle of a pool for small populations, rejection set otherwise).
// The state (624 words + position) is caller-owned so the Python layer can round-trip random.getstate()/setstate().
#include <stdint.h>
#include <string.h>
#include <vector>
#include <unordered_set>
#include "../../include/sarssl_b200.h"

namespace {
constexpr int N = 624, M = 397;

struct MT {
    uint32_t* mt;   // 624 words
    uint32_t& idx;  // position
    explicit MT(uint32_t* s) : mt(s), idx(s[624]) {}

    void init_genrand(uint32_t s) {
        mt[0] = s;
        for (int i = 1; i < N; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = N;
    }
    void init_by_array(const uint32_t* key, int klen) {
        init_genrand(19650218u);
        int i = 1, j = 0;
        for (int k = (N > klen ? N : klen); k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
            ++i; ++j;
            if (i >= N) { mt[0] = mt[N - 1]; i = 1; }
            if (j >= klen) j = 0;
        }
        for (int k = N - 1; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
            ++i;
            if (i >= N) { mt[0] = mt[N - 1]; i = 1; }
        }
        mt[0] = 0x80000000u;
    }
    uint32_t next() {
        static const uint32_t mag01[2] = {0u, 0x9908b0dfu};
        if (idx >= (uint32_t)N) {
            int kk;
            uint32_t y;
            for (kk = 0; kk < N - M; ++kk) {
                y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + M] ^ (y >> 1) ^ mag01[y & 1u];
            }
            for (; kk < N - 1; ++kk) {
                y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ mag01[y & 1u];
            }
            y = (mt[N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
            mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ mag01[y & 1u];
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    static int bit_length(uint32_t n) { int b = 0; while (n) { ++b; n >>= 1; } return b; }
    // Random._randbelow_with_getrandbits(n), 1 <= n < 2^31
    uint32_t randbelow(uint32_t n) {
        const int k = bit_length(n);
        uint32_t r = next() >> (32 - k);
        while (r >= n) r = next() >> (32 - k);
        return r;
    }
};

// 4 ** ceil(log(3k, 4)) : smallest power of four >= 3k (3k is never a power of four, so no rounding edge)
int64_t sample_setsize(int k) {
    int64_t setsize = 21;
    if (k > 5) {
        int64_t p = 1;
        while (p < 3 * (int64_t)k) p *= 4;
        setsize += p;
    }
    return setsize;
}
}  // namespace

extern "C" int sarssl_mt19937_seed_host(uint32_t* state_host, const uint32_t* key, int nkey) {
    if (!state_host || !key || nkey < 1) return SARSSL_ERR_ARG;
    MT g(state_host);
    g.init_by_array(key, nkey);
    return SARSSL_OK;
}

extern "C" int sarssl_mt19937_draw_masks_host(uint32_t* state_host, int nb, int npatch, int nmasked, int nmic,
                                              int64_t* patch_idx_host, int64_t* ch_idx_host, uint8_t* frame_flag_host) {
    if (!state_host || !patch_idx_host || !ch_idx_host) return SARSSL_ERR_ARG;
    if (nb < 0 || npatch < 1 || nmasked < 0 || nmasked > npatch || nmic < 1) return SARSSL_ERR_ARG;   // random.sample raises ValueError
    if (state_host[624] > 624u) return SARSSL_ERR_ARG;
    MT g(state_host);
    const bool use_pool = (int64_t)npatch <= sample_setsize(nmasked);
    std::vector<int32_t> pool;
    if (frame_flag_host) memset(frame_flag_host, 0, (size_t)nb * npatch);
    for (int b = 0; b < nb; ++b) {
        int64_t* out = patch_idx_host + (size_t)b * nmasked;
        if (use_pool) {
            pool.resize(npatch);
            for (int i = 0; i < npatch; ++i) pool[i] = i;
            for (int i = 0; i < nmasked; ++i) {
                const uint32_t j = g.randbelow((uint32_t)(npatch - i));
                out[i] = pool[j];
                pool[j] = pool[npatch - i - 1];
            }
        } else {
            std::unordered_set<uint32_t> seen;
            for (int i = 0; i < nmasked; ++i) {
                uint32_t j = g.randbelow((uint32_t)npatch);
                while (seen.count(j)) j = g.randbelow((uint32_t)npatch);
                seen.insert(j);
                out[i] = j;
            }
        }
        ch_idx_host[b] = g.randbelow((uint32_t)nmic);          // randint(0, nmic-1) == randrange(0, nmic)
        if (frame_flag_host)
            for (int i = 0; i < nmasked; ++i) frame_flag_host[(size_t)b * npatch + out[i]] = 1;
    }
    return SARSSL_OK;
}
