// Warp-per-transform 512-point complex FFT of the fused STFT front-end kernel (stft.cu).
//
// One warp owns one transform, 16 points per lane (512 = 16 x 32); no block-level barrier anywhere:
//   stage 1 (registers)   lane n2 holds z[32*n1 + n2], n1 = 0..15:  A[k1] = DFT16_{n1}(z) * W512^(n2*k1)
//   transpose (smem)      A[k1][n2] -> lane (k1, h) = k1 + 16*h reads n2 = 2*m + h, m = 0..15
//   stage 2 (registers)   E/O[k2'] = DFT16_m(.)  (h = 0: even n2, h = 1: odd n2), odd half times W32^k2'
//   combine (shuffles)    Z[k1 + 16*k2'] = E + O',  Z[k1 + 16*(k2'+16)] = E - O'   exchanged with lane ^ 16
// After the combine lane (k1, h) holds Z[k1 + 256*h + 16*s], s = 0..15; it is written to shared memory in natural bin order
// (wz_pos), from where lane l reads the bins k = 1 + l + 32*r it will store and their mirrors Z[512 - k] (the two real channels
// packed as z = x0 + i*x1 are separated with the mirrored bin): conflict-free, no selects, no special lanes.
// Complex additions run as packed f32x2 instructions on the device (one issue slot per complex add).
// __host__ __device__ so tests/host/host_fft_check.cpp can emulate the 32 lanes on the CPU.
#pragma once
#include "fft512.cuh"

namespace sarssl {

__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    unsigned long long ua, ub, ud;
    float2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(ud) : "l"(ua), "l"(ub));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(ud));
    return r;
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
    unsigned long long ua, ub, ud;
    float2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(ud) : "l"(ua), "l"(ub));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(ud));
    return r;
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
// a * (s, s) with one packed multiply
__host__ __device__ __forceinline__ float2 cscale2(float2 a, float2 s) {
#ifdef __CUDA_ARCH__
    unsigned long long ua, us, ud;
    float2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(us) : "f"(s.x), "f"(s.y));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(ud) : "l"(ua), "l"(us));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(ud));
    return r;
#else
    return make_float2(a.x * s.x, a.y * s.y);
#endif
}
// a * s + b, component-wise, one packed FMA
__host__ __device__ __forceinline__ float2 cfma2(float2 a, float2 s, float2 b) {
#ifdef __CUDA_ARCH__
    unsigned long long ua, us, ub, ud;
    float2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(us) : "f"(s.x), "f"(s.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(ud) : "l"(ua), "l"(us), "l"(ub));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(ud));
    return r;
#else
    return make_float2(a.x * s.x + b.x, a.y * s.y + b.y);
#endif
}

// 4-point forward DFT of (a, b, c, d) -> outputs at k = 0, 1, 2, 3.  The two outputs that need (b - d) * (-i) are formed with scalar
// adds (a packed add would first have to build the rotated pair).
__host__ __device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d) {
    const float2 s0 = cadd(a, c), s1 = csub(a, c), s2 = cadd(b, d), t = csub(b, d);
    a = cadd(s0, s2); c = csub(s0, s2);
    b = make_float2(s1.x + t.y, s1.y - t.x);
    d = make_float2(s1.x - t.y, s1.y + t.x);
}

// forward 16-point DFT in place: v[k] = sum_n v[n] * exp(-2*pi*i*n*k/16).  n = 4a + b, k = c + 4d.
__host__ __device__ __forceinline__ void dft16(float2 (&v)[16]) {
    // DFT4 over a for each b: v[4a + b] -> Y[b][c] stored at v[4c + b]
#pragma unroll
    for (int b = 0; b < 4; ++b) dft4(v[b], v[4 + b], v[8 + b], v[12 + b]);
    // twiddle Y[b][c] *= W16^(b*c)
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
    // W16^1 = (c1, -s1), W16^2 = (h, -h), W16^3 = (s1, -c1), W16^4 = -i, W16^6 = (-h, -h), W16^9 = (-c1, s1)
    v[4 * 1 + 1] = cmul(v[4 * 1 + 1], make_float2(c1, -s1));
    v[4 * 2 + 1] = cmul(v[4 * 2 + 1], make_float2(h, -h));
    v[4 * 3 + 1] = cmul(v[4 * 3 + 1], make_float2(s1, -c1));
    v[4 * 1 + 2] = cmul(v[4 * 1 + 2], make_float2(h, -h));
    v[4 * 2 + 2] = make_float2(v[4 * 2 + 2].y, -v[4 * 2 + 2].x);
    v[4 * 3 + 2] = cmul(v[4 * 3 + 2], make_float2(-h, -h));
    v[4 * 1 + 3] = cmul(v[4 * 1 + 3], make_float2(s1, -c1));
    v[4 * 2 + 3] = cmul(v[4 * 2 + 3], make_float2(-h, -h));
    v[4 * 3 + 3] = cmul(v[4 * 3 + 3], make_float2(-c1, s1));
    // DFT4 over b for each c: inputs v[4c + b] (b = 0..3) -> X[c + 4d] stored at v[4c + d]
#pragma unroll
    for (int c = 0; c < 4; ++c) dft4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    // now v[4c + d] = X[c + 4d]; bring to natural order v[k] = X[k] (swap v[4c+d] <-> v[c+4d]; register renaming, no instructions)
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int d = c + 1; d < 4; ++d) { const float2 t = v[4 * c + d]; v[4 * c + d] = v[4 * d + c]; v[4 * d + c] = t; }
}

constexpr int kWTransStride = 33;                         // float2 row stride of the [16][32] transpose buffer (2-way conflicts at worst)
constexpr int kWTransFloat2 = 16 * kWTransStride;         // 528 float2 = 4224 B per warp

// position of bin k in the natural-order Z buffer: the upper half is shifted by 16 entries (128 B) so that the stores of the two
// half-warps (k1 + 16*s and k1 + 256 + 16*s) fall into different banks.  Max 511 + 16 = 527 < kWTransFloat2.
__host__ __device__ __forceinline__ int wz_pos(int k) { return k + ((k >> 8) << 4); }

// stage-1 twiddles of lane n2: W512^(n2*k1), k1 = 1..15
struct WarpFftLane {
    float2 tw[16];
    __host__ __device__ void init(int lane) {
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) tw[k1] = unit512(lane * k1);
    }
};

// v[n1] = z[32*n1 + lane] on entry.  Writes the twiddled stage-1 result to the transpose buffer.
__host__ __device__ __forceinline__ void wfft_stage1(float2 (&v)[16], const WarpFftLane& c, float2* tb, int lane) {
    dft16(v);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) tb[k1 * kWTransStride + lane] = (k1 == 0) ? v[0] : cmul(v[k1], c.tw[k1]);
}

// after a warp sync: lane (k1 = lane & 15, h = lane >> 4) reads its 16 points, runs stage 2; on return v[s] = E or O' (odd half twiddled)
__host__ __device__ __forceinline__ void wfft_stage2(float2 (&v)[16], const float2* tb, int lane) {
    const int k1 = lane & 15, h = lane >> 4;
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = tb[k1 * kWTransStride + 2 * m + h];
    dft16(v);
    if (h) {                                              // O'[s] = O[s] * W32^s  (compile-time constants after unrolling)
        constexpr float kc[16] = {1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f, 0.70710678118654752440f,
                                  0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f, 0.0f, -0.19509032201612826785f,
                                  -0.38268343236508977173f, -0.55557023301960222474f, -0.70710678118654752440f, -0.83146961230254523708f,
                                  -0.92387953251128675613f, -0.98078528040323044913f};
        constexpr float ks[16] = {0.0f, -0.19509032201612826785f, -0.38268343236508977173f, -0.55557023301960222474f, -0.70710678118654752440f,
                                  -0.83146961230254523708f, -0.92387953251128675613f, -0.98078528040323044913f, -1.0f, -0.98078528040323044913f,
                                  -0.92387953251128675613f, -0.83146961230254523708f, -0.70710678118654752440f, -0.55557023301960222474f,
                                  -0.38268343236508977173f, -0.19509032201612826785f};
#pragma unroll
        for (int s = 1; s < 16; ++s) v[s] = cmul(v[s], make_float2(kc[s], ks[s]));
    }
}

// combine + natural-order store.  `mine`/`other` are this lane's and lane ^ 16's stage-2 results; sgn = (+1, +1) on lanes 0..15 and
// (-1, -1) on lanes 16..31:  Z = other + sgn * mine  (E + O' below bin 256, E - O' above).  Z[k1 + 256*h + 16*s] -> zb[wz_pos(.)].
__host__ __device__ __forceinline__ void wfft_store_z(const float2 (&mine)[16], const float2 (&other)[16], float2 sgn, float2* zb, int lane) {
    const int base = wz_pos((lane & 15) + 256 * (lane >> 4));
#pragma unroll
    for (int s = 0; s < 16; ++s) zb[base + 16 * s] = cfma2(mine[s], sgn, other[s]);
}

// un-normalised split of bin k with its mirror p = Z[512 - k]:  2*X_ch0 = Z + conj(P),  2*X_ch1 = (Z - conj(P)) / i
//   -> (2 re0, 2 re1, 2 im0, 2 im1) = (z.x + p.x, z.y + p.y, z.y - p.y, p.x - z.x); the factor 1/2 is folded into the clip scale.
__host__ __device__ __forceinline__ float4 wfft_split2(float2 z, float2 p) {
    const float2 re = cadd(z, p);
    return make_float4(re.x, re.y, z.y - p.y, p.x - z.x);
}

}  // namespace sarssl
