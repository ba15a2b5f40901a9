// 512-point complex FFT shared by the STFT front-end kernels (sm_100a).
//
// 64 threads cooperate on one transform, 8 points each, three radix-8 passes with two shared-memory
// exchanges (512 = 8*8*8).  The two microphone channels of a frame are packed as z = x_ch0 + i*x_ch1, so ONE
// complex transform yields both one-sided spectra (split step below).  Everything thread-invariant across
// frames (window samples, twiddles) lives in registers (FftLane), so the per-frame cost is the butterflies plus
// three conflict-free exchanges.
//
// The code is __host__ __device__ so tests/host_fft_check.cpp can emulate the 64 lanes sequentially on the CPU
// and check the index algebra against an O(N^2) DFT without a GPU.
#pragma once
#include <math.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
static inline float4 make_float4(float a, float b, float c, float d) { float4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
#endif

namespace sarssl {

constexpr int kFftN = 512;
constexpr int kFftLanes = 64;
// scratch floats per 64-lane group: re and im planes of max(8*72, 8*68, 8*64)
constexpr int kFftPlane = 8 * 72;
constexpr int kFftScratchFloats = 2 * kFftPlane;

__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// forward 8-point DFT in place: v[k] = sum_n v[n] * exp(-2*pi*i*n*k/8)
__host__ __device__ __forceinline__ void dft8(float2 (&v)[8]) {
    const float h = 0.70710678118654752440f;
    float2 a0 = make_float2(v[0].x + v[4].x, v[0].y + v[4].y), a4 = make_float2(v[0].x - v[4].x, v[0].y - v[4].y);
    float2 a1 = make_float2(v[1].x + v[5].x, v[1].y + v[5].y), a5 = make_float2(v[1].x - v[5].x, v[1].y - v[5].y);
    float2 a2 = make_float2(v[2].x + v[6].x, v[2].y + v[6].y), a6 = make_float2(v[2].x - v[6].x, v[2].y - v[6].y);
    float2 a3 = make_float2(v[3].x + v[7].x, v[3].y + v[7].y), a7 = make_float2(v[3].x - v[7].x, v[3].y - v[7].y);
    // twiddles on the odd half: W8^1 = (h,-h), W8^2 = -i, W8^3 = (-h,-h)
    a5 = make_float2(h * (a5.x + a5.y), h * (a5.y - a5.x));
    a6 = make_float2(a6.y, -a6.x);
    a7 = make_float2(h * (a7.y - a7.x), -h * (a7.x + a7.y));
    // even outputs: DFT4(a0..a3)
    float2 b0 = make_float2(a0.x + a2.x, a0.y + a2.y), b2 = make_float2(a0.x - a2.x, a0.y - a2.y);
    float2 b1 = make_float2(a1.x + a3.x, a1.y + a3.y), b3 = make_float2(a1.y - a3.y, a3.x - a1.x);   // (a1-a3)*(-i)
    v[0] = make_float2(b0.x + b1.x, b0.y + b1.y);
    v[4] = make_float2(b0.x - b1.x, b0.y - b1.y);
    v[2] = make_float2(b2.x + b3.x, b2.y + b3.y);
    v[6] = make_float2(b2.x - b3.x, b2.y - b3.y);
    // odd outputs: DFT4(a4..a7)
    float2 c0 = make_float2(a4.x + a6.x, a4.y + a6.y), c2 = make_float2(a4.x - a6.x, a4.y - a6.y);
    float2 c1 = make_float2(a5.x + a7.x, a5.y + a7.y), c3 = make_float2(a5.y - a7.y, a7.x - a5.x);
    v[1] = make_float2(c0.x + c1.x, c0.y + c1.y);
    v[5] = make_float2(c0.x - c1.x, c0.y - c1.y);
    v[3] = make_float2(c2.x + c3.x, c2.y + c3.y);
    v[7] = make_float2(c2.x - c3.x, c2.y - c3.y);
}

__host__ __device__ __forceinline__ float2 unit512(int j) {       // exp(-2*pi*i*j/512), accurate
#ifdef __CUDA_ARCH__
    float s, c;
    sincospif(-(float)(j & 511) / 256.0f, &s, &c);
    return make_float2(c, s);
#else
    double a = -2.0 * 3.14159265358979323846 * (double)(j & 511) / 512.0;
    return make_float2((float)cos(a), (float)sin(a));
#endif
}

// Per-lane constants (depend only on the lane id l = 0..63, not on the frame).
struct FftLane {
    float win[8];      // periodic Hann at n = l + 64*r
    float2 tw1[8];     // pass-1 twiddle W512^(8*n2*k1),        n2 = l>>3, index k1
    float2 tw2[8];     // pass-2 twiddle W512^(n3*(k1 + 8*k2)), k1 = l>>3, n3 = l&7, index k2
    __host__ __device__ void init(int l) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int n = l + 64 * r;
#ifdef __CUDA_ARCH__
            win[r] = 0.5f - 0.5f * cospif((float)n / 256.0f);
#else
            win[r] = (float)(0.5 - 0.5 * cos(2.0 * 3.14159265358979323846 * n / 512.0));
#endif
            tw1[r] = unit512(8 * (l >> 3) * r);
            tw2[r] = unit512((l & 7) * ((l >> 3) + 8 * r));
        }
    }
};

// The three passes are separate functions because a group barrier is needed between them; the caller owns
// the barrier (bar.sync on the device, a plain loop boundary in the host emulation).
// Index algebra (n = 64*n1 + 8*n2 + n3, k = k1 + 8*k2 + 64*k3):
//   pass 1, lane l = 8*n2+n3 : A[k1]  = DFT8_{n1}(z[l + 64*n1]) * W512^(8*n2*k1)      -> sre[k1*72 + l]
//   pass 2, lane l = 8*k1+n3 : B[k2]  = DFT8_{n2}(A[k1; n2, n3]) * W512^(n3*(k1+8*k2)) -> sre[n3*68 + k1 + 8*k2]
//   pass 3, lane t = k1+8*k2 : Z[t + 64*k3] = DFT8_{n3}(B[k1, k2; n3])
__host__ __device__ __forceinline__ void fft_pass1(float2 (&v)[8], const FftLane& c, float* sre, float* sim, int l) {
    dft8(v);
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) {
        float2 a = (k1 == 0) ? v[0] : cmul(v[k1], c.tw1[k1]);
        sre[k1 * 72 + l] = a.x;
        sim[k1 * 72 + l] = a.y;
    }
}

__host__ __device__ __forceinline__ void fft_pass2_load(float2 (&v)[8], const float* sre, const float* sim, int l) {
    const int k1 = l >> 3, n3 = l & 7;
#pragma unroll
    for (int n2 = 0; n2 < 8; ++n2) v[n2] = make_float2(sre[k1 * 72 + 8 * n2 + n3], sim[k1 * 72 + 8 * n2 + n3]);
}

__host__ __device__ __forceinline__ void fft_pass2_store(float2 (&v)[8], const FftLane& c, float* sre, float* sim, int l) {
    const int k1 = l >> 3, n3 = l & 7;
    dft8(v);
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) {
        float2 a = cmul(v[k2], c.tw2[k2]);
        sre[n3 * 68 + k1 + 8 * k2] = a.x;
        sim[n3 * 68 + k1 + 8 * k2] = a.y;
    }
}

__host__ __device__ __forceinline__ void fft_pass3(float2 (&v)[8], const float* sre, const float* sim, int t) {
#pragma unroll
    for (int n3 = 0; n3 < 8; ++n3) v[n3] = make_float2(sre[n3 * 68 + t], sim[n3 * 68 + t]);
    dft8(v);            // v[k3] = Z[t + 64*k3]
}

// Split step.  Lane t holds Z[t + 64*k3]; the mirrored bins Z[512 - k] live in lane (64 - t) % 64, slot 7 - k3
// (slot (8 - k3) % 8 for t == 0).  exch_store publishes Z, exch_load fetches the mirror into p[].
__host__ __device__ __forceinline__ void split_store(const float2 (&v)[8], float* sre, float* sim, int t) {
#pragma unroll
    for (int k3 = 0; k3 < 8; ++k3) { sre[k3 * 64 + t] = v[k3].x; sim[k3 * 64 + t] = v[k3].y; }
}

// After split_load: for k = t + 64*k3 (k3 = 0..4; k3 == 4 only meaningful for t == 0, i.e. k = 256)
//   X_ch0[k] = (Z[k] + conj(Z[512-k])) / 2,   X_ch1[k] = (Z[k] - conj(Z[512-k])) / (2i)
// returned as float4 (re0, re1, im0, im1) == the reference's patch layout [reim][mic] for one bin.
__host__ __device__ __forceinline__ float4 split_bin(const float2 (&v)[8], const float* sre, const float* sim, int t, int k3) {
    const int pt = (64 - t) & 63;
    const int ps = (t == 0) ? ((8 - k3) & 7) : (7 - k3);
    const float pr = sre[ps * 64 + pt], pi = sim[ps * 64 + pt];
    const float zr = v[k3].x, zi = v[k3].y;
    return make_float4(0.5f * (zr + pr), 0.5f * (zi + pi), 0.5f * (zi - pi), 0.5f * (pr - zr));
}

}  // namespace sarssl
