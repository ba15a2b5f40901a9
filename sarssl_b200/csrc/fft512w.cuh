// Warp-per-transform 512-point complex FFT (second-generation front-end kernel).
//
// One warp owns one transform, 16 points per lane (512 = 16 x 32); no block-level barrier anywhere:
//   stage 1 (registers)   lane n2 holds z[32*n1 + n2], n1 = 0..15:  A[k1] = DFT16_{n1}(z) * W512^(n2*k1)
//   transpose (smem)      A[k1][n2] -> lane (k1, h) = k1 + 16*h reads n2 = 2*m + h, m = 0..15
//   stage 2 (registers)   E/O[k2'] = DFT16_m(.)  (h = 0: even n2, h = 1: odd n2), odd half times W32^k2'
//   combine (shuffles)    Z[k1 + 16*k2'] = E + O',  Z[k1 + 16*(k2'+16)] = E - O'   exchanged with lane ^ 16
// After the combine lane (k1, 0) holds Z[k1 + 16*s] and lane (k1, 1) holds Z[k1 + 256 + 16*s], s = 0..15.
// The two real channels packed as z = x0 + i*x1 are separated with the mirrored bins Z[512 - k], which live in lane
// ((16 - k1) & 15, 1 - h) at slot 15 - s (k1 != 0); lanes 0 and 16 (k1 == 0) mirror into each other at slot (16 - s) & 15.
// __host__ __device__ so tests/host/host_fft_check.cpp can emulate the 32 lanes on the CPU.
#pragma once
#include "fft512.cuh"

namespace sarssl {

__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ float2 cmuli_neg(float2 a) { return make_float2(a.y, -a.x); }           // a * (-i)

// 4-point forward DFT of (a, b, c, d) -> outputs at k = 0, 1, 2, 3
__host__ __device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d) {
    const float2 s0 = cadd(a, c), s1 = csub(a, c), s2 = cadd(b, d), s3 = cmuli_neg(csub(b, d));
    a = cadd(s0, s2); c = csub(s0, s2); b = cadd(s1, s3); d = csub(s1, s3);
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
    v[4 * 2 + 2] = cmuli_neg(v[4 * 2 + 2]);
    v[4 * 3 + 2] = cmul(v[4 * 3 + 2], make_float2(-h, -h));
    v[4 * 1 + 3] = cmul(v[4 * 1 + 3], make_float2(s1, -c1));
    v[4 * 2 + 3] = cmul(v[4 * 2 + 3], make_float2(-h, -h));
    v[4 * 3 + 3] = cmul(v[4 * 3 + 3], make_float2(-c1, s1));
    // DFT4 over b for each c: inputs v[4c + b] (b = 0..3) -> X[c + 4d] stored at v[4c + d]
#pragma unroll
    for (int c = 0; c < 4; ++c) dft4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    // now v[4c + d] = X[c + 4d]; bring to natural order v[k] = X[k] (swap v[4c+d] <-> v[c+4d])
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int d = c + 1; d < 4; ++d) { const float2 t = v[4 * c + d]; v[4 * c + d] = v[4 * d + c]; v[4 * d + c] = t; }
}

constexpr int kWTransStride = 33;                         // float2 row stride of the [16][32] transpose buffer (2-way conflicts at worst)
constexpr int kWTransFloat2 = 16 * kWTransStride;         // 528 float2 = 4224 B per warp

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

#ifdef __CUDACC__
// combine: exchange with lane ^ 16.  On return lane (k1, h) holds Z[k1 + 256*h + 16*s] in v[s].
__device__ __forceinline__ void wfft_combine(float2 (&v)[16], int lane) {
    const bool hi = lane >= 16;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const float ox = __shfl_xor_sync(0xffffffffu, v[s].x, 16), oy = __shfl_xor_sync(0xffffffffu, v[s].y, 16);
        v[s] = hi ? make_float2(ox - v[s].x, oy - v[s].y) : make_float2(v[s].x + ox, v[s].y + oy);
    }
}

// split one bin: X_ch0 = (Z + conj(P))/2, X_ch1 = (Z - conj(P))/(2i) -> (re0, re1, im0, im1)
__device__ __forceinline__ float4 wfft_split(float2 z, float2 p) {
    return make_float4(0.5f * (z.x + p.x), 0.5f * (z.y + p.y), 0.5f * (z.y - p.y), 0.5f * (p.x - z.x));
}

// Balanced split.  Lane (k1, 0) produces bins k = k1 + 16*j, j = 0..7; lane (p, 1) produces the bins k = k1' + 16*(j + 8) of
// k1' = (16 - p) & 15.  One exchange of v[8..15] with lane ((16 - k1) & 15) + 16*(1 - h) supplies everything:
//   h = 0: Z = v[j],        mirror = partner slot 15 - j = o[7 - j]          (k1 == 0: slot 16 - j = o[8 - j]; j == 0: own v[0])
//   h = 1: Z = o[j],        mirror = own slot 15 - (j + 8) = v[7 - j]        (k1' == 0: own v[8 - j])
// out[j] = bin value, kout[j] = its bin index.  Lane 16 additionally owns the Nyquist bin 256 (returned in nyq).
__device__ __forceinline__ void wfft_split_all(const float2 (&v)[16], int lane, float4 (&out)[8], float4& nyq) {
    const int k1 = lane & 15, h = lane >> 4;
    const int partner = ((16 - k1) & 15) + 16 * (1 - h);
    float2 o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = make_float2(__shfl_sync(0xffffffffu, v[8 + j].x, partner), __shfl_sync(0xffffffffu, v[8 + j].y, partner));
    const bool z0 = k1 == 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float2 z, pm;
        if (h == 0) {
            z = v[j];
            const float2 a = o[7 - j], b = (j == 0) ? v[0] : o[8 - j];      // b only meaningful for k1 == 0
            pm = z0 ? b : a;
        } else {
            z = o[j];
            const float2 a = v[7 - j], b = v[8 - j];
            pm = z0 ? b : a;
        }
        out[j] = wfft_split(z, pm);
    }
    nyq = wfft_split(v[0], v[0]);                          // meaningful on lane 16 only: Z[256] mirrors itself
}
#endif


}  // namespace sarssl
