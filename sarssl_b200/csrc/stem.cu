// CNN patch-embedding stem (model.py:50-64) on channel-last images [B][H = frame][W = bin][C] - CUDA-core kernels.
//   1x1 conv 4 -> 64 with the spectral / spatial input masking of model.py:541,563 fused into the load,
//   3x3 conv 64 -> 64 as an implicit GEMM (K = 9 taps x 64 channels) with the previous layer's BatchNorm + ReLU applied
//   while the operand tile is loaded (zero padding stays zero), its data gradient (same kernel, mirrored weights) and
//   its weight gradient (split-K over pixels, fixed-order reduction), and the 1x1 conv 64 -> 4.
// These are the exact-fp32 / reference-check versions; the bf16 tensor-core versions live in conv_tc.cu.
#include "common.cuh"
#include "vec.cuh"

namespace sarssl {

// packed fp32 pairs (one issue slot, two FMAs).  Same FLOP rate as scalar FFMA (measured: scripts/micro/ffma2_rate.cu), but half the
// issue slots - what the pixel kernels below are short of.
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack_f32x2(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ unsigned long long fma_f32x2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long add_f32x2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// ---- input loader shared by the narrow (4-channel) kernels ---------------------------------------------------------
// mode 0: plain [P][4] tensor of type T.  mode 1 / 2: fp32 patches (re0, re1, im0, im1) with the spectral / spatial mask:
//   spectral: masked frame -> keep only the un-masked microphone; other frames -> keep only the masked microphone
//   spatial : masked frame -> zeros; other frames -> both microphones
// mode 3: fp32 patches, no masking (downstream fine-tuning branch)
// mode 4: spectral input of the frozen-encoder branch (model.py:622): masked frame -> only the un-masked microphone; other frames -> zeros
template <typename T>
__device__ __forceinline__ float4 load_narrow(const void* in, long long p, int mode, const uint8_t* flag, const int32_t* ch, int W, int H) {
    if (mode == 0) {
        const T* q = static_cast<const T*>(in) + p * 4;
        if (sizeof(T) == 2) {                                   // 4 x bf16 = one 8-byte load
            const uint2 u = *reinterpret_cast<const uint2*>(q);
            const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
            return make_float4(a.x, a.y, b.x, b.y);
        }
        return *reinterpret_cast<const float4*>(q);
    }
    float4 x = *reinterpret_cast<const float4*>(static_cast<const float*>(in) + p * 4);
    if (mode == 3) return x;                                    // un-masked fp32 patches (downstream branch, model.py:667-678)
    // (pixel counts stay far below 2^32: 32-bit divisions, a shift when W is a power of two)
    const unsigned up = (unsigned)p, uw = (unsigned)W;
    const unsigned row = ((uw & (uw - 1)) == 0) ? (up >> (31 - __clz((int)uw))) : up / uw;
    const bool pm = flag[row] != 0;
    const int mc = ch[row / (unsigned)H];
    if (mode == 1 || (mode == 4 && pm)) {
        const int keep = pm ? 1 - mc : mc;
        if (keep == 0) { x.y = 0.f; x.w = 0.f; } else { x.x = 0.f; x.z = 0.f; }
    } else if (mode == 4 || pm) x = make_float4(0.f, 0.f, 0.f, 0.f);
    return x;
}

// The mask of modes 1 / 2 depends only on the pixel's image row (frame): bit c set = channel c is zeroed.
__device__ __forceinline__ unsigned narrow_zero_bits(long long p, int mode, const uint8_t* flag, const int32_t* ch, int W, int H) {
    if (mode != 1 && mode != 2 && mode != 4) return 0u;
    const unsigned up = (unsigned)p, uw = (unsigned)W;
    const unsigned row = ((uw & (uw - 1)) == 0) ? (up >> (31 - __clz((int)uw))) : up / uw;
    const bool pm = flag[row] != 0;
    if (mode == 2) return pm ? 0xFu : 0u;
    if (mode == 4 && !pm) return 0xFu;
    const int mc = ch[row / (unsigned)H];
    const int keep = pm ? 1 - mc : mc;
    return keep == 0 ? 0xAu : 0x5u;                            // keep microphone 0: zero (re1, im1) = channels 1, 3
}
__device__ __forceinline__ float4 apply_zero_bits(float4 x, unsigned z) {
    if (z & 1u) x.x = 0.f;
    if (z & 2u) x.y = 0.f;
    if (z & 4u) x.z = 0.f;
    if (z & 8u) x.w = 0.f;
    return x;
}

// out[p][o] = sum_c W[o][c] * in[p][c],  o < 64.  8 threads per pixel, 8 channels each.
// With scale/shift (the BatchNorm of the conv output, batch statistics known beforehand from sarssl_stem_input_stats) the kernel
// writes relu(scale * (W x) + shift) directly: the pre-BatchNorm tensor is never stored.
template <typename T>
__global__ void __launch_bounds__(256, 3) pw_expand_kernel(const void* __restrict__ in, int mode, const uint8_t* __restrict__ flag,
                                                      const int32_t* __restrict__ ch, const float* __restrict__ Wt, const float* __restrict__ scale,
                                                      const float* __restrict__ shift, T* __restrict__ out, long long P, int W, int H) {
    __shared__ float ws[64][4];
    if (threadIdx.x < 256) ws[threadIdx.x >> 2][threadIdx.x & 3] = Wt[threadIdx.x] * (scale ? scale[threadIdx.x >> 2] : 1.f);
    __syncthreads();
    const int sub = threadIdx.x & 7;
    const bool bn = scale != nullptr;
    // channel pairs packed for fma.rn.f32x2: wp[jp][c] = (W[2jp][c], W[2jp+1][c]), shp[jp] = (shift[2jp], shift[2jp+1])
    unsigned long long wp[4][4], shp[4];
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
        const int o = sub * 8 + 2 * jp;
        shp[jp] = bn ? pack_f32x2(shift[o], shift[o + 1]) : 0ull;
#pragma unroll
        for (int c = 0; c < 4; ++c) wp[jp][c] = pack_f32x2(ws[o][c], ws[o + 1][c]);
    }
    // 4 pixels per thread per iteration (independent loads in flight)
    for (long long p0 = ((long long)blockIdx.x * 32 + (threadIdx.x >> 3)) * 4; p0 < P; p0 += (long long)gridDim.x * 128) {
        float4 x[4];
        if ((W & 3) == 0) {                // the 4 consecutive pixels share an image row: one mask lookup (p0 % 4 == 0)
            const unsigned zb = narrow_zero_bits(p0, mode, flag, ch, W, H);
            const int lm = (mode == 1 || mode == 2 || mode == 4) ? 3 : mode;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                x[u] = (p0 + u < P) ? apply_zero_bits(load_narrow<T>(in, p0 + u, lm, nullptr, nullptr, W, H), zb) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) x[u] = (p0 + u < P) ? load_narrow<T>(in, p0 + u, mode, flag, ch, W, H) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (p0 + u >= P) break;
            float o[8];
            const unsigned long long x0 = pack_f32x2(x[u].x, x[u].x), x1 = pack_f32x2(x[u].y, x[u].y), x2 = pack_f32x2(x[u].z, x[u].z),
                                     x3 = pack_f32x2(x[u].w, x[u].w);
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                unsigned long long a = fma_f32x2(wp[jp][0], x0, shp[jp]);
                a = fma_f32x2(wp[jp][1], x1, a);
                a = fma_f32x2(wp[jp][2], x2, a);
                a = fma_f32x2(wp[jp][3], x3, a);
                const float2 r = unpack_f32x2(a);
                o[2 * jp] = bn ? fmaxf(r.x, 0.f) : r.x;
                o[2 * jp + 1] = bn ? fmaxf(r.y, 0.f) : r.y;
            }
            Vec8<T>::store(out + (p0 + u) * 64 + sub * 8, o);
        }
    }
}

// out[p][c] = sum_o W[c][o] * f(in[p][o]),  f = relu(in*scale+shift) when stats given
template <typename T>
__global__ void __launch_bounds__(256) pw_reduce_kernel(const T* __restrict__ in, const float* __restrict__ scale, const float* __restrict__ shift,
                                                      const float* __restrict__ Wt, T* __restrict__ out, long long P) {
    __shared__ float ws[4][64];
    __shared__ float sc[64], sh[64];
    ws[threadIdx.x >> 6][threadIdx.x & 63] = Wt[threadIdx.x];
    if (threadIdx.x < 64) { sc[threadIdx.x] = scale ? scale[threadIdx.x] : 1.f; sh[threadIdx.x] = shift ? shift[threadIdx.x] : 0.f; }
    __syncthreads();
    const int sub = threadIdx.x & 7;
    const bool tr = scale != nullptr;
    float wr[4][8], scr[8], shr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        scr[j] = sc[sub * 8 + j]; shr[j] = sh[sub * 8 + j];
#pragma unroll
        for (int c = 0; c < 4; ++c) wr[c][j] = ws[c][sub * 8 + j];
    }
    // weights packed for fma.rn.f32x2: (W[0][j], W[1][j]) and (W[2][j], W[3][j]) - one issue slot feeds two accumulators
    unsigned long long w01[8], w23[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { w01[j] = pack_f32x2(wr[0][j], wr[1][j]); w23[j] = pack_f32x2(wr[2][j], wr[3][j]); }
    const long long per_iter = (long long)gridDim.x * 128;
    const long long iters = (P + per_iter - 1) / per_iter;
    for (long long it = 0; it < iters; ++it) {
        const long long p0 = (it * gridDim.x + blockIdx.x) * 128 + (threadIdx.x >> 3) * 4;      // 4 consecutive pixels per 8-thread group
        float qv[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (p0 + u < P) Vec8<T>::load(in + (p0 + u) * 64 + sub * 8, qv[u]);
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j) qv[u][j] = 0.f;
            }
        }
        // acc[u*2 + h] = this thread's 8-channel share of outputs (2h, 2h+1) of pixel u
        unsigned long long acc[8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            unsigned long long a01 = 0ull, a23 = 0ull;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float v = qv[u][j];
                if (tr) v = fmaxf(fmaf(v, scr[j], shr[j]), 0.f);
                const unsigned long long vv = pack_f32x2(v, v);
                a01 = fma_f32x2(w01[j], vv, a01);
                a23 = fma_f32x2(w23[j], vv, a23);
            }
            acc[u * 2] = a01; acc[u * 2 + 1] = a23;
        }
        // reduce-scatter over the 8 threads of the pixel group: every step halves the values a thread still owns (7 packed shuffles
        // instead of 48 scalar ones); thread `sub` ends with acc[0] = outputs (2*(sub&1), +1) of pixel sub>>1
#pragma unroll
        for (int step = 0; step < 3; ++step) {
            const int half = 4 >> step, bit = 4 >> step;
            const bool up = (sub & bit) != 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (i >= half) continue;
                const unsigned long long send = up ? acc[i] : acc[i + half], keep = up ? acc[i + half] : acc[i];
                const unsigned long long got = __shfl_xor_sync(0xffffffffu, send, bit);
                acc[i] = add_f32x2(keep, got);
            }
        }
        const int pu = sub >> 1;
        if (p0 + pu < P) {
            const float2 r = unpack_f32x2(acc[0]);
            T* o = out + (p0 + pu) * 4 + (sub & 1) * 2;
            o[0] = from_f32<T>(r.x); o[1] = from_f32<T>(r.y);
        }
    }
}

// partial[cta][o][c] = sum_{p in cta} f(wide[p][o]) * narrow[p][c]
template <typename T>
__global__ void __launch_bounds__(256) pw_wgrad_kernel(const T* __restrict__ wide, const float* __restrict__ scale, const float* __restrict__ shift,
                                                     const void* __restrict__ narrow, int mode, const uint8_t* __restrict__ flag,
                                                     const int32_t* __restrict__ ch, float* __restrict__ partials, long long P, int W, int H) {
    __shared__ float red[32][8][33];
    const int sub = threadIdx.x & 7, pl = threadIdx.x >> 3;
    const bool tr = scale != nullptr;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = tr ? scale[sub * 8 + j] : 1.f; sh[j] = tr ? shift[sub * 8 + j] : 0.f; }
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
    for (long long p0 = ((long long)blockIdx.x * 32 + pl) * 4; p0 < P; p0 += (long long)gridDim.x * 128) {
        float4 x[4];
        float qv[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (p0 + u < P) {
                x[u] = load_narrow<T>(narrow, p0 + u, mode, flag, ch, W, H);
                Vec8<T>::load(wide + (p0 + u) * 64 + sub * 8, qv[u]);
            } else {
                x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 8; ++j) qv[u][j] = 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float v = qv[u][j];
                if (tr) v = fmaxf(v * sc[j] + sh[j], 0.f);
                acc[j][0] = fmaf(v, x[u].x, acc[j][0]); acc[j][1] = fmaf(v, x[u].y, acc[j][1]);
                acc[j][2] = fmaf(v, x[u].z, acc[j][2]); acc[j][3] = fmaf(v, x[u].w, acc[j][3]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) red[pl][sub][j * 4 + c] = acc[j][c];
    __syncthreads();
    // 256 outputs (o, c): thread t sums over the 32 pixel lanes
    {
        const int o = threadIdx.x >> 2, c = threadIdx.x & 3;
        float s = 0.f;
        for (int q = 0; q < 32; ++q) s += red[q][o >> 3][(o & 7) * 4 + c];
        partials[(size_t)blockIdx.x * 256 + threadIdx.x] = s;
    }
}

// ---- bf16 pixel contractions on the warp-level tensor-core path (mma.sync m16n8k16, fp32 accumulate) -----------------------------
// D[o][n] = sum_p A[p][o] * B[p][n]: M = 64 wide channels (4 m-tiles), K = pixels, N = 8 = the 4 narrow channels twice:
// columns 0..3 hold bf16(x), columns 4..7 the rounding residual bf16(x - bf16(x)), so fp32 patch inputs keep 16 mantissa bits.
// Both operands are pixel-major in memory, i.e. K is the slow index: ldmatrix.trans delivers the fragments.  Every warp owns 16
// pixels of a 128-pixel tile and a private cp.async ring (no block barrier in the main loop).  The A fragments are built in
// registers from the wide tensor(s):
//   kWgPlain : A = f(wide), f = BatchNorm+ReLU when scale/shift are given (fp32 math, rounded to bf16 like the materialised activation)
//              -> out[256] = D                                                      (weight gradient of either 1x1 conv)
//   kWgHead  : A = g = wide * [wide2 > 0]  (gradient through the ReLU that produced wide2); besides G = D it accumulates
//              S[o] = sum_p g (a ones column), the 8x8 second moments of the narrow operand and its column sums
//              -> out[392] = G[64][4], S[64], M[8][8], Sx[8]                        (everything the first stem layer's backward needs)
//   kWgTail  : A0 = m = [scale*wide+shift > 0], A1 = m * wide
//              -> out[512] = H0[64][4], H1[64][4]                                   (everything the last stem layer's backward needs)
enum { kWgPlain = 0, kWgHead = 1, kWgTail = 2 };
template <int MODE> struct WgCfg {
    static constexpr int kWide = MODE == kWgHead ? 2 : 1;                      // wide tiles per stage
    static constexpr int kStages = MODE == kWgHead ? 3 : 4;
    static constexpr int kWarpStage = kWide * 16 * 128 + 16 * 16;              // wide 16 px x 128 B each (16-byte chunks XOR-swizzled by pixel) + narrow 16 px x 16 B
    static constexpr int kSmem = 8 * kStages * kWarpStage;                     // 73,728 B (plain, tail) / 104,448 B (head)
    static constexpr int kOut = MODE == kWgPlain ? 256 : (MODE == kWgHead ? 392 : 512);
};

__device__ __forceinline__ uint32_t bn_relu_bf16x2(uint32_t r, float sc, float sh) {
    const float lo = fmaxf(fmaf(__uint_as_float(r << 16), sc, sh), 0.f), hi = fmaxf(fmaf(__uint_as_float(r & 0xffff0000u), sc, sh), 0.f);
    uint32_t o;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o) : "f"(hi), "f"(lo));
    return o;
}
// 0xffff in every half of z that holds a positive bf16 (sign clear, magnitude non-zero)
__device__ __forceinline__ uint32_t positive_halves_bf16x2(uint32_t z) {
    const uint32_t nz = ((z & 0x7fff7fffu) + 0x7fff7fffu) & ~z & 0x80008000u;
    return (nz >> 15) * 0xffffu;
}
// [sc*y + sh > 0] for a bf16 y as one packed compare: (y ^ flip) > thr with flip = sign bit when sc < 0 and thr = the largest bf16 <=
// (-sh/sc) * sign(sc) (y is exactly a bf16, so comparing against the rounded-down threshold is the same predicate).  Returned packed
// twice (both halves of an A-fragment register belong to the same channel): .x = flip, .y = thr.
__device__ __forceinline__ uint2 relu_gate_bf16x2(float sc, float sh) {
    uint32_t flip = 0u;
    float thr;
    if (sc == 0.f) thr = sh > 0.f ? -3.0e38f : 3.0e38f;
    else { thr = -sh / sc; if (sc < 0.f) { thr = -thr; flip = 0x8000u; } }
    uint32_t b = __float_as_uint(thr);
    if (thr >= 0.f) b &= 0xffff0000u;                                  // towards zero = down
    else if (b & 0xffffu) b = (b & 0xffff0000u) + 0x10000u;            // negative: away from zero = down
    b >>= 16;
    return make_uint2(flip | (flip << 16), b | (b << 16));
}
// 0xffff in every half of y that passes the gate
__device__ __forceinline__ uint32_t active_halves_bf16x2(uint32_t y, uint2 gate) {
    const uint32_t v = y ^ gate.x;
    return __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&v), *reinterpret_cast<const __nv_bfloat162*>(&gate.y));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int MODE>
__global__ void __launch_bounds__(256) pw_contract_mma_kernel(const __nv_bfloat16* __restrict__ wide, const __nv_bfloat16* __restrict__ wide2,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            const void* __restrict__ narrow, int mode, const uint8_t* __restrict__ flag,
                                                            const int32_t* __restrict__ ch, float* __restrict__ partials, long long P, int W, int H) {
    using Cfg = WgCfg<MODE>;
    constexpr int S = Cfg::kStages, WS = Cfg::kWarpStage, NOFF = Cfg::kWide * 16 * 128;
    extern __shared__ __align__(128) uint8_t wg_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    uint8_t* ring = wg_smem + (size_t)warp * S * WS;
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    const bool tr = scale != nullptr;
    float sc[4][2], sh[4][2];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) { sc[mt][h] = tr ? scale[16 * mt + g + 8 * h] : 1.f; sh[mt][h] = tr ? shift[16 * mt + g + 8 * h] : 0.f; }
    uint2 gate[4][2];
    if (MODE == kWgTail) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) { gate[mt][0] = relu_gate_bf16x2(sc[mt][0], sh[mt][0]); gate[mt][1] = relu_gate_bf16x2(sc[mt][1], sh[mt][1]); }
    }
    float acc[4][4], acc2[MODE == kWgPlain ? 1 : 4][4], accM[4], accX[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) { acc[mt][q] = 0.f; if (MODE != kWgPlain) acc2[mt][q] = 0.f; }
        accM[q] = accX[q] = 0.f;
    }
    const int ntiles = (int)((P + 127) / 128);
    const int my_tiles = (ntiles > (int)blockIdx.x) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    // stage `it` (tile blockIdx.x + it * gridDim.x): cp.async the warp's 16 x 128 B of each wide tensor, return its narrow pixel (lanes 0..15) in registers
    auto issue = [&](int it, float4& nx) {
        nx = make_float4(0.f, 0.f, 0.f, 0.f);
        if (it < my_tiles) {
            const long long p0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * 128 + warp * 16;
            const uint32_t dst = ring_s + (uint32_t)(it % S) * WS;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int id = lane + 32 * j, px = id >> 3, chunk = id & 7;
                const bool ok = p0 + px < P;
                const long long off = ok ? (p0 + px) * 64 + chunk * 8 : 0;
                const uint32_t d = dst + px * 128 + ((chunk ^ (px & 7)) << 4);
                const int nbytes = ok ? 16 : 0;                                       // zero-fill beyond the last pixel
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(wide + off), "r"(nbytes) : "memory");
                if (MODE == kWgHead) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d + 16 * 128), "l"(wide2 + off), "r"(nbytes) : "memory");
            }
            if (lane < 16 && p0 + lane < P) nx = load_narrow<__nv_bfloat16>(narrow, p0 + lane, mode, flag, ch, W, H);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto put_narrow = [&](int it, const float4& nx) {                                  // [px][8] bf16: value, then rounding residual
        if (lane < 16) {
            const __nv_bfloat162 h0 = __floats2bfloat162_rn(nx.x, nx.y), h1 = __floats2bfloat162_rn(nx.z, nx.w);
            const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
            const __nv_bfloat162 l0 = __floats2bfloat162_rn(nx.x - f0.x, nx.y - f0.y), l1 = __floats2bfloat162_rn(nx.z - f1.x, nx.w - f1.y);
            uint4 v;
            v.x = *reinterpret_cast<const uint32_t*>(&h0); v.y = *reinterpret_cast<const uint32_t*>(&h1);
            v.z = *reinterpret_cast<const uint32_t*>(&l0); v.w = *reinterpret_cast<const uint32_t*>(&l1);
            *reinterpret_cast<uint4*>(ring + (size_t)(it % S) * WS + NOFF + lane * 16) = v;
        }
    };

    float4 nx;
#pragma unroll 1
    for (int s = 0; s < S - 1; ++s) { issue(s, nx); put_narrow(s, nx); }
    const uint32_t ones = (g == 0) ? 0x3f803f80u : 0u;                                 // B fragment of a matrix whose column 0 is all ones
#pragma unroll 1
    for (int it = 0; it < my_tiles; ++it) {
        issue(it + S - 1, nx);                                                         // slot (it - 1) % S: consumed last iteration
        asm volatile("cp.async.wait_group %0;" ::"n"(S - 1) : "memory");
        __syncwarp();
        const uint32_t st = ring_s + (uint32_t)(it % S) * WS;
        uint32_t b0, b1;
        {
            const uint32_t addr = st + NOFF + (lane & 15) * 16;
            asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(addr));
        }
        if (MODE == kWgHead) {                                                         // the B fragment of [p][n] is also the A fragment of its transpose (rows 0..7)
            mma_bf16_16816(accM, b0, 0u, b1, 0u, b0, b1);
            mma_bf16_16816(accX, b0, 0u, b1, 0u, ones, ones);
        }
        const int mi = lane >> 3, r = lane & 7, px = (mi >> 1) * 8 + r;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            uint32_t a0, a1, a2, a3;
            const uint32_t addr = st + px * 128 + (((2 * mt + (mi & 1)) ^ r) << 4);
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(addr));
            if (MODE == kWgPlain) {
                if (tr) {
                    a0 = bn_relu_bf16x2(a0, sc[mt][0], sh[mt][0]); a2 = bn_relu_bf16x2(a2, sc[mt][0], sh[mt][0]);
                    a1 = bn_relu_bf16x2(a1, sc[mt][1], sh[mt][1]); a3 = bn_relu_bf16x2(a3, sc[mt][1], sh[mt][1]);
                }
                mma_bf16_16816(acc[mt], a0, a1, a2, a3, b0, b1);
            } else if (MODE == kWgHead) {
                uint32_t z0, z1, z2, z3;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(z0), "=r"(z1), "=r"(z2), "=r"(z3) : "r"(addr + 16 * 128));
                a0 &= positive_halves_bf16x2(z0); a1 &= positive_halves_bf16x2(z1); a2 &= positive_halves_bf16x2(z2); a3 &= positive_halves_bf16x2(z3);
                mma_bf16_16816(acc[mt], a0, a1, a2, a3, b0, b1);
                mma_bf16_16816(acc2[mt], a0, a1, a2, a3, ones, ones);
            } else {
                const uint32_t m0 = active_halves_bf16x2(a0, gate[mt][0]), m2 = active_halves_bf16x2(a2, gate[mt][0]);
                const uint32_t m1 = active_halves_bf16x2(a1, gate[mt][1]), m3 = active_halves_bf16x2(a3, gate[mt][1]);
                mma_bf16_16816(acc[mt], m0 & 0x3f803f80u, m1 & 0x3f803f80u, m2 & 0x3f803f80u, m3 & 0x3f803f80u, b0, b1);
                mma_bf16_16816(acc2[mt], a0 & m0, a1 & m1, a2 & m2, a3 & m3, b0, b1);
            }
        }
        __syncwarp();                                                                  // every lane is done reading slot it % S ...
        put_narrow(it + S - 1, nx);                                                    // ... and slot (it - 1) % S, refilled here
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // D columns n = 2t, 2t+1 live in lane quad position t: add the residual columns (t + 2) to the value columns, then sum the 8 warps
    constexpr int NOUT = Cfg::kOut;
    float* red = reinterpret_cast<float*>(wg_smem) + warp * NOUT;                      // [8 warps][NOUT]
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int o = 16 * mt + g + 8 * (q >> 1), c = 2 * t + (q & 1);
            const float v = acc[mt][q] + __shfl_down_sync(0xffffffffu, acc[mt][q], 2);
            if (t < 2) red[o * 4 + c] = v;
            if (MODE == kWgTail) {
                const float v2 = acc2[mt][q] + __shfl_down_sync(0xffffffffu, acc2[mt][q], 2);
                if (t < 2) red[256 + o * 4 + c] = v2;
            }
            if (MODE == kWgHead && t == 0 && (q & 1) == 0) red[256 + o] = acc2[mt][q];
        }
    if (MODE == kWgHead) {
        red[320 + g * 8 + 2 * t] = accM[0]; red[320 + g * 8 + 2 * t + 1] = accM[1];
        if (t == 0) red[384 + g] = accX[0];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NOUT; i += 256) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += reinterpret_cast<const float*>(wg_smem)[w * NOUT + i];
        partials[(size_t)blockIdx.x * NOUT + i] = sum;
    }
}

template <int MODE>
static int launch_pw_contract(const void* wide, const void* wide2, const float* scale, const float* shift, const void* narrow, int mode,
                              const uint8_t* flag, const int32_t* ch, float* partials, size_t partial_floats, long long P, int W, int H,
                              cudaStream_t stream, int* grid_out) {
    using Cfg = WgCfg<MODE>;
    static bool attr_set = false;
    if (!attr_set) { SARSSL_CUDA(cudaFuncSetAttribute(pw_contract_mma_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem)); attr_set = true; }
    const long long need = (P + 127) / 128, cap = resident_ctas(pw_contract_mma_kernel<MODE>, 256, Cfg::kSmem);
    const int grid = (int)(need < cap ? need : cap);
    if (partial_floats < (size_t)grid * Cfg::kOut) { set_last_error("stem pixel contraction: workspace too small"); return SARSSL_ERR_WORKSPACE; }
    pw_contract_mma_kernel<MODE><<<grid, 256, Cfg::kSmem, stream>>>(static_cast<const __nv_bfloat16*>(wide), static_cast<const __nv_bfloat16*>(wide2), scale, shift,
                                                                     narrow, mode, flag, ch, partials, P, W, H);
    SARSSL_LAUNCH_CHECK();
    *grid_out = grid;
    return SARSSL_OK;
}

// ---- fused backward of the first stem layer (1x1 conv 4 -> 64, BatchNorm, ReLU): one pass over (dz, z) -----------------------------
// With y = W x, g = dz * [z > 0] and the BatchNorm backward dy = sc*g - c0 - c1*y (see bn_act_bwd_kernel) the weight gradient is
//   dW[o][c] = sum_p dy[p][o] x[p][c] = sc[o] G[o][c] - c0[o] Sx[c] - c1[o] (W Mxx)[o][c],
// and sum_p g*y = sum_c W[o][c] G[o][c]: nothing of size P x 64 is written (the input gradient of this layer is never needed).
// sums [392] = G[64][4], S[64], M[8][8], Sx[8] (value | residual halves still separate).  One block of 256 threads = (o, c).
__global__ void stem_head_bwd_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ w64x4, const float* __restrict__ stats,
                                              double inv_rows, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dw64x4) {
    const int o = threadIdx.x >> 2, c = threadIdx.x & 3;
    const double mean = stats[o], rstd = stats[64 + o], sc = stats[128 + o];
    double Sgy = 0.0, WM = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        Sgy += (double)w64x4[o * 4 + k] * (double)sums[o * 4 + k];
        const double mxx = (double)sums[320 + k * 8 + c] + (double)sums[320 + k * 8 + c + 4] + (double)sums[320 + (k + 4) * 8 + c] + (double)sums[320 + (k + 4) * 8 + c + 4];
        WM += (double)w64x4[o * 4 + k] * mxx;
    }
    const double S = sums[256 + o], Sx = (double)sums[384 + c] + (double)sums[384 + c + 4];
    const double s2 = rstd * (Sgy - mean * S);
    const double c1 = sc * rstd * s2 * inv_rows, c0 = sc * S * inv_rows - c1 * mean;
    dw64x4[o * 4 + c] += (float)(sc * (double)sums[o * 4 + c] - c0 * Sx - c1 * WM);
    if (c == 0) { dgamma[o] += (float)s2; dbeta[o] += (float)S; }
}

// ---- fused backward of the last stem layer pair (BatchNorm, ReLU, 1x1 conv 64 -> 4) -------------------------------------------------
// Upstream gradient dq [P][4] (narrow).  dz[p][o] = sum_c W[c][o] dq[p][c] is never materialised: with m = [sc*y+sh > 0],
//   H0[o][c] = sum_p m dq,  H1[o][c] = sum_p m y dq   give   dW[c][o] = sc H1 + sh H0   (z = m (sc y + sh)),
//   sum_p g = sum_c W[c][o] H0[o][c],  sum_p g y = sum_c W[c][o] H1[o][c]                (g = m dz),
// and a second pass writes dy = sc*g - c0 - c1*y recomputing dz from dq in registers.
__global__ void stem_tail_bwd_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ w4x64, const float* __restrict__ stats,
                                              double inv_rows, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dw4x64,
                                              float* __restrict__ coef) {
    const int o = threadIdx.x >> 2, c = threadIdx.x & 3;
    const double mean = stats[o], rstd = stats[64 + o], sc = stats[128 + o], sh = stats[192 + o];
    dw4x64[c * 64 + o] += (float)(sc * (double)sums[256 + o * 4 + c] + sh * (double)sums[o * 4 + c]);
    if (c == 0) {
        double S = 0.0, Sgy = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { S += (double)w4x64[k * 64 + o] * (double)sums[o * 4 + k]; Sgy += (double)w4x64[k * 64 + o] * (double)sums[256 + o * 4 + k]; }
        const double s2 = rstd * (Sgy - mean * S);
        const double c1 = sc * rstd * s2 * inv_rows;
        dgamma[o] += (float)s2; dbeta[o] += (float)S;
        coef[o] = (float)(sc * S * inv_rows - c1 * mean);        // c0
        coef[64 + o] = (float)c1;
    }
}

// dy[p][o] = sc[o] * [sc*y+sh > 0] * (sum_c W[c][o] dq[p][c]) - c0[o] - c1[o] * y[p][o].  8 threads per pixel, 8 channels each.
// The 8 per-channel constants stay in shared memory ([j][sub] order: the 8 sub-groups of a warp read 128 contiguous bytes), read
// once per 4 pixels, so the kernel keeps enough CTAs resident to cover the HBM latency.
__global__ void __launch_bounds__(256, 3) stem_tail_bwd_apply_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ dq,
                                                                const float* __restrict__ w4x64, const float* __restrict__ stats,
                                                                const float* __restrict__ coef, __nv_bfloat16* __restrict__ dy, long long P) {
    __shared__ float4 cw[64], cb[64];                   // (sc*W[0..3][o]) and (sc, sh, c0, c1) at index j*8 + sub, o = sub*8 + j
    if (threadIdx.x < 64) {
        const int o = threadIdx.x, i = (o & 7) * 8 + (o >> 3);
        const float sc = stats[128 + o];
        cw[i] = make_float4(w4x64[o] * sc, w4x64[64 + o] * sc, w4x64[128 + o] * sc, w4x64[192 + o] * sc);
        cb[i] = make_float4(sc, stats[192 + o], coef[o], coef[64 + o]);
    }
    __syncthreads();
    const int sub = threadIdx.x & 7;
    for (long long p0 = ((long long)blockIdx.x * 32 + (threadIdx.x >> 3)) * 4; p0 < P; p0 += (long long)gridDim.x * 128) {
        float4 x[4];
        float v[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long p = p0 + u < P ? p0 + u : P - 1;          // clamp: the tail pixels are loaded, not stored
            x[u] = load_narrow<__nv_bfloat16>(dq, p, 0, nullptr, nullptr, 0, 0);
            Vec8<__nv_bfloat16>::load(y + p * 64 + sub * 8, v[u]);
        }
        // the 4 upstream-gradient channels of every pixel packed with themselves, so one f32x2 FMA serves two of the 4 pixels
        unsigned long long xp[2][4];                                      // [pixel pair][c] = (x[2q][c], x[2q+1][c])
#pragma unroll
        for (int q2 = 0; q2 < 2; ++q2) {
            xp[q2][0] = pack_f32x2(x[2 * q2].x, x[2 * q2 + 1].x); xp[q2][1] = pack_f32x2(x[2 * q2].y, x[2 * q2 + 1].y);
            xp[q2][2] = pack_f32x2(x[2 * q2].z, x[2 * q2 + 1].z); xp[q2][3] = pack_f32x2(x[2 * q2].w, x[2 * q2 + 1].w);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            asm volatile("" ::: "memory");                                // keep the constant loads where they are used (register pressure)
            const float4 w = cw[j * 8 + sub], b = cb[j * 8 + sub];
            const unsigned long long w0 = pack_f32x2(w.x, w.x), w1 = pack_f32x2(w.y, w.y), w2 = pack_f32x2(w.z, w.z), w3 = pack_f32x2(w.w, w.w);
            const unsigned long long c1n = pack_f32x2(-b.w, -b.w), c0n = pack_f32x2(-b.z, -b.z);
#pragma unroll
            for (int q2 = 0; q2 < 2; ++q2) {
                unsigned long long dz = fma_f32x2(w0, xp[q2][0], 0ull);
                dz = fma_f32x2(w1, xp[q2][1], dz);
                dz = fma_f32x2(w2, xp[q2][2], dz);
                dz = fma_f32x2(w3, xp[q2][3], dz);
                const float ya = v[2 * q2][j], yb = v[2 * q2 + 1][j];
                const float2 d = unpack_f32x2(dz);
                const unsigned long long gs = pack_f32x2(fmaf(ya, b.x, b.y) > 0.f ? d.x : 0.f, fmaf(yb, b.x, b.y) > 0.f ? d.y : 0.f);
                const float2 r = unpack_f32x2(add_f32x2(gs, fma_f32x2(c1n, pack_f32x2(ya, yb), c0n)));      // gs - (c1*y + c0)
                v[2 * q2][j] = r.x; v[2 * q2 + 1][j] = r.y;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (p0 + u < P) Vec8<__nv_bfloat16>::store(dy + (p0 + u) * 64 + sub * 8, v[u]);
    }
}

// ---- batch statistics of y = W x without forming y: y is linear in the 4 input channels, so sum_p y and sum_p y^2 follow from the
// column sums Sx[4] and second moments Mxx[4][4] of the (masked) input.  partials [cta][14] = Sx, then the upper triangle of Mxx.
template <typename T>
__global__ void __launch_bounds__(256) narrow_moments_kernel(const void* __restrict__ in, int mode, const uint8_t* __restrict__ flag,
                                                           const int32_t* __restrict__ ch, float* __restrict__ partials, long long P, int W, int H) {
    float a[14];
#pragma unroll
    for (int i = 0; i < 14; ++i) a[i] = 0.f;
    for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < P; p += (long long)gridDim.x * 256) {
        const float4 x = load_narrow<T>(in, p, mode, flag, ch, W, H);
        a[0] += x.x; a[1] += x.y; a[2] += x.z; a[3] += x.w;
        a[4] = fmaf(x.x, x.x, a[4]); a[5] = fmaf(x.x, x.y, a[5]); a[6] = fmaf(x.x, x.z, a[6]); a[7] = fmaf(x.x, x.w, a[7]);
        a[8] = fmaf(x.y, x.y, a[8]); a[9] = fmaf(x.y, x.z, a[9]); a[10] = fmaf(x.y, x.w, a[10]);
        a[11] = fmaf(x.z, x.z, a[11]); a[12] = fmaf(x.z, x.w, a[12]); a[13] = fmaf(x.w, x.w, a[13]);
    }
    __shared__ float red[8][14];
#pragma unroll
    for (int i = 0; i < 14; ++i) {
        const float v = warp_sum(a[i]);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 14) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        partials[(size_t)blockIdx.x * 14 + threadIdx.x] = t;
    }
}

// one block of 64 threads (o): out[0][o] = sum_p y[p][o], out[1][o] = sum_p y[p][o]^2  (the [1][2][64] partial-sum layout of batchnorm_finalize)
__global__ void narrow_moments_finalize_kernel(const float* __restrict__ partials, int nparts, const float* __restrict__ w64x4, float* __restrict__ out) {
    __shared__ double m[14];
    if (threadIdx.x < 14) {
        double t = 0.0;
        for (int p = 0; p < nparts; ++p) t += (double)partials[(size_t)p * 14 + threadIdx.x];
        m[threadIdx.x] = t;
    }
    __syncthreads();
    const int o = threadIdx.x;
    const double w0 = w64x4[o * 4], w1 = w64x4[o * 4 + 1], w2 = w64x4[o * 4 + 2], w3 = w64x4[o * 4 + 3];
    out[o] = (float)(w0 * m[0] + w1 * m[1] + w2 * m[2] + w3 * m[3]);
    out[64 + o] = (float)(w0 * w0 * m[4] + w1 * w1 * m[8] + w2 * w2 * m[11] + w3 * w3 * m[13] +
                          2.0 * (w0 * w1 * m[5] + w0 * w2 * m[6] + w0 * w3 * m[7] + w1 * w2 * m[9] + w1 * w3 * m[10] + w2 * w3 * m[12]));
}

__global__ void stem_reduce_kernel(const float* __restrict__ partials, int nparts, int width, float* __restrict__ out, int accumulate) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= width) return;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int p = 0;
    for (; p + 3 < nparts; p += 4) {
        s0 += (double)partials[(size_t)p * width + w]; s1 += (double)partials[(size_t)(p + 1) * width + w];
        s2 += (double)partials[(size_t)(p + 2) * width + w]; s3 += (double)partials[(size_t)(p + 3) * width + w];
    }
    for (; p < nparts; ++p) s0 += (double)partials[(size_t)p * width + w];
    const double s = (s0 + s1) + (s2 + s3);
    out[w] = accumulate ? out[w] + (float)s : (float)s;
}

// ---- 3x3 conv as implicit GEMM: out[p][n] = sum_{tap, ci} f(in[p + tap][ci]) * Wp[n][tap*64 + ci] -------------------------------
constexpr int CM = 64, CN = 64, CK = 16;

template <typename T>
__global__ void __launch_bounds__(256) conv3x3_kernel(const T* __restrict__ in, const float* __restrict__ scale, const float* __restrict__ shift,
                                                    const T* __restrict__ Wp, T* __restrict__ out, long long P, int H, int W) {
    __shared__ __align__(16) float As[CK][CM + 4];
    __shared__ __align__(16) float Bs[CK][CN + 4];
    __shared__ int ph[CM], pw[CM];
    __shared__ long long pbase[CM];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long p0 = (long long)blockIdx.x * CM;
    if (tid < CM) {
        const long long p = p0 + tid;
        if (p < P) { pw[tid] = (int)(p % W); ph[tid] = (int)((p / W) % H); pbase[tid] = p; } else { pw[tid] = -100000; ph[tid] = -100000; pbase[tid] = 0; }
    }
    __syncthreads();
    const bool tr = scale != nullptr;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < 576; k0 += CK) {
        const int tap = k0 >> 6, dh = tap / 3 - 1, dw = tap % 3 - 1, c0 = k0 & 63;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + i * 256;
            const int kk = idx & 15, mm = idx >> 4;
            const int h = ph[mm] + dh, w = pw[mm] + dw;
            float v = 0.f;
            if (h >= 0 && h < H && w >= 0 && w < W) {
                const int ci = c0 + kk;
                v = to_f32(in[(pbase[mm] + (long long)dh * W + dw) * 64 + ci]);
                if (tr) v = fmaxf(v * scale[ci] + shift[ci], 0.f);
            }
            As[kk][mm] = v;
            Bs[kk][mm] = to_f32(Wp[(long long)mm * 576 + k0 + kk]);          // here mm plays the role of n
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < CK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long p = p0 + ty * 4 + i;
        if (p >= P) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) out[p * 64 + tx * 4 + j] = from_f32<T>(acc[i][j]);
    }
}

// weight gradient: partial[split][o][tap*64+ci] = sum_{p in split} dy[p][o] * f(z_in[p + tap][ci]);  grid (9 taps, splits)
template <typename T>
__global__ void __launch_bounds__(256) conv3x3_wgrad_kernel(const T* __restrict__ dy, const T* __restrict__ in, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, float* __restrict__ partials, long long P, int H, int W,
                                                          long long pix_per_split) {
    __shared__ __align__(16) float As[CK][CM + 4];     // [pixel kk][o]
    __shared__ __align__(16) float Bs[CK][CN + 4];     // [pixel kk][ci]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int tap = blockIdx.x, dh = tap / 3 - 1, dw = tap % 3 - 1;
    const long long pbeg = (long long)blockIdx.y * pix_per_split;
    long long pend = pbeg + pix_per_split;
    if (pend > P) pend = P;
    const bool tr = scale != nullptr;
    const int cc = tid & 63;
    const float sc = tr ? scale[cc] : 1.f, sh = tr ? shift[cc] : 0.f;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (long long k0 = pbeg; k0 < pend; k0 += CK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + i * 256;
            const int c = idx & 63, kk = idx >> 6;            // c == cc for every i
            const long long p = k0 + kk;
            float a = 0.f, b = 0.f;
            if (p < pend) {
                a = to_f32(dy[p * 64 + c]);
                const int w = (int)(p % W) + dw, h = (int)((p / W) % H) + dh;
                if (h >= 0 && h < H && w >= 0 && w < W) {
                    b = to_f32(in[(p + (long long)dh * W + dw) * 64 + c]);
                    if (tr) b = fmaxf(b * sc + sh, 0.f);
                }
            }
            As[kk][c] = a;
            Bs[kk][c] = b;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < CK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* dst = partials + (size_t)blockIdx.y * 64 * 576;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[(size_t)(ty * 4 + i) * 576 + tap * 64 + tx * 4 + j] = acc[i][j];
}

template <typename K>
static int pix_grid(K kernel, long long P) {
    long long g = (P + 127) / 128;
    const long long cap = resident_ctas(kernel, 256);           // exactly one resident wave
    return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace sarssl

using namespace sarssl;

static int stem_expand_impl(const void* in, int mode, const uint8_t* frame_flag, const int32_t* ch_idx, const float* weight64x4, const float* scale,
                            const float* shift, void* out, long long P, int W, int H, int dtype, cudaStream_t stream) {
    SARSSL_CHECK_ARG(in && weight64x4 && out && P > 0 && P < 0xFFFFFFFFLL, "stem_expand: bad arguments (P must be in (0, 2^32))");
    SARSSL_CHECK_ARG(mode == 0 || mode == 3 || (frame_flag && ch_idx && (mode == 1 || mode == 2 || mode == 4)), "stem_expand: mode %d needs masks", mode);
    if (dtype == SARSSL_F32) pw_expand_kernel<float><<<pix_grid(pw_expand_kernel<float>, P), 256, 0, stream>>>(in, mode, frame_flag, ch_idx, weight64x4, scale, shift, (float*)out, P, W, H);
    else if (dtype == SARSSL_BF16) pw_expand_kernel<__nv_bfloat16><<<pix_grid(pw_expand_kernel<__nv_bfloat16>, P), 256, 0, stream>>>(in, mode, frame_flag, ch_idx, weight64x4, scale, shift, (__nv_bfloat16*)out, P, W, H);
    else { set_last_error("stem_expand: bad dtype"); return SARSSL_ERR_ARG; }
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_stem_expand(const void* in, int mode, const uint8_t* frame_flag, const int32_t* ch_idx, const float* weight64x4, void* out,
                                  long long P, int W, int H, int dtype, cudaStream_t stream) {
    return stem_expand_impl(in, mode, frame_flag, ch_idx, weight64x4, nullptr, nullptr, out, P, W, H, dtype, stream);
}

// out = relu(scale * conv1x1(in) + shift): conv + BatchNorm + ReLU in one pass (the BatchNorm statistics come from sarssl_stem_input_stats)
extern "C" int sarssl_stem_expand_bn_relu(const void* in, int mode, const uint8_t* frame_flag, const int32_t* ch_idx, const float* weight64x4,
                                          const float* scale, const float* shift, void* out, long long P, int W, int H, int dtype, cudaStream_t stream) {
    SARSSL_CHECK_ARG(scale && shift, "stem_expand_bn_relu: scale / shift missing");
    return stem_expand_impl(in, mode, frame_flag, ch_idx, weight64x4, scale, shift, out, P, W, H, dtype, stream);
}

// sums[0][o] = sum_p y[p][o], sums[1][o] = sum_p y[p][o]^2 of y = conv1x1(in) (64 outputs) from the input's first and second moments
extern "C" int sarssl_stem_input_stats(const void* in, int mode, const uint8_t* frame_flag, const int32_t* ch_idx, const float* weight64x4, float* sums2x64,
                                       long long P, int W, int H, int dtype, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    SARSSL_CHECK_ARG(in && weight64x4 && sums2x64 && workspace && P > 0 && P < 0xFFFFFFFFLL, "stem_input_stats: bad arguments (P must be in (0, 2^32))");
    SARSSL_CHECK_ARG(mode == 0 || mode == 3 || (frame_flag && ch_idx && (mode == 1 || mode == 2 || mode == 4)), "stem_input_stats: mode %d needs masks", mode);
    SARSSL_CHECK_ARG(dtype == SARSSL_F32 || dtype == SARSSL_BF16, "stem_input_stats: bad dtype");
    long long g = (P + 255) / 256;
    const long long cap = dtype == SARSSL_F32 ? resident_ctas(narrow_moments_kernel<float>, 256) : resident_ctas(narrow_moments_kernel<__nv_bfloat16>, 256);
    const int grid = (int)(g < cap ? g : cap);
    if (workspace_bytes < (size_t)grid * 14 * sizeof(float)) { set_last_error("stem_input_stats: workspace too small"); return SARSSL_ERR_WORKSPACE; }
    float* partials = static_cast<float*>(workspace);
    if (dtype == SARSSL_F32) narrow_moments_kernel<float><<<grid, 256, 0, stream>>>(in, mode, frame_flag, ch_idx, partials, P, W, H);
    else narrow_moments_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(in, mode, frame_flag, ch_idx, partials, P, W, H);
    SARSSL_LAUNCH_CHECK();
    narrow_moments_finalize_kernel<<<1, 64, 0, stream>>>(partials, grid, weight64x4, sums2x64);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_stem_reduce(const void* in, const float* in_scale, const float* in_shift, const float* weight4x64, void* out, long long P,
                                  int dtype, cudaStream_t stream) {
    SARSSL_CHECK_ARG(in && weight4x64 && out && P > 0, "stem_reduce: bad arguments");
    if (dtype == SARSSL_F32) pw_reduce_kernel<float><<<pix_grid(pw_reduce_kernel<float>, P), 256, 0, stream>>>((const float*)in, in_scale, in_shift, weight4x64, (float*)out, P);
    else if (dtype == SARSSL_BF16) pw_reduce_kernel<__nv_bfloat16><<<pix_grid(pw_reduce_kernel<__nv_bfloat16>, P), 256, 0, stream>>>((const __nv_bfloat16*)in, in_scale, in_shift, weight4x64, (__nv_bfloat16*)out, P);
    else { set_last_error("stem_reduce: bad dtype"); return SARSSL_ERR_ARG; }
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" size_t sarssl_stem_workspace_bytes(void) {
    const size_t a = (size_t)sm_count() * 8 * 256 * sizeof(float);           // pw_wgrad partials
    const size_t b = (size_t)sm_count() * 2 * 64 * 576 * sizeof(float);      // conv3x3 wgrad partials
    return a > b ? a : b;
}

// dW[o][c] (64 x 4, row-major) (+)= sum_p f(wide[p][o]) * narrow[p][c]
extern "C" int sarssl_stem_pw_wgrad(const void* wide, const float* wide_scale, const float* wide_shift, const void* narrow, int mode,
                                    const uint8_t* frame_flag, const int32_t* ch_idx, float* dweight64x4, int accumulate, long long P, int W, int H,
                                    int dtype, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    SARSSL_CHECK_ARG(wide && narrow && dweight64x4 && workspace && P > 0 && P < 0xFFFFFFFFLL, "stem_pw_wgrad: bad arguments (P must be in (0, 2^32))");
    float* partials = static_cast<float*>(workspace);
    int grid = 0;
    if (dtype == SARSSL_BF16) {
        const int rc = launch_pw_contract<kWgPlain>(wide, nullptr, wide_scale, wide_shift, narrow, mode, frame_flag, ch_idx, partials,
                                                    workspace_bytes / sizeof(float), P, W, H, stream, &grid);
        if (rc != SARSSL_OK) return rc;
    } else if (dtype == SARSSL_F32) {
        grid = pix_grid(pw_wgrad_kernel<float>, P);
        if (workspace_bytes < (size_t)grid * 256 * sizeof(float)) { set_last_error("stem_pw_wgrad: workspace too small"); return SARSSL_ERR_WORKSPACE; }
        pw_wgrad_kernel<float><<<grid, 256, 0, stream>>>((const float*)wide, wide_scale, wide_shift, narrow, mode, frame_flag, ch_idx, partials, P, W, H);
        SARSSL_LAUNCH_CHECK();
    } else { set_last_error("stem_pw_wgrad: bad dtype"); return SARSSL_ERR_ARG; }
    stem_reduce_kernel<<<1, 256, 0, stream>>>(partials, grid, 256, dweight64x4, accumulate);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

// Backward of the first stem layer in one pass (bf16 activations only): see stem_head_bwd_finalize_kernel.
// dz = gradient w.r.t. the layer's ReLU output z [P][64]; stats = (mean, rstd, scale, shift) of its BatchNorm; narrow = the layer input.
extern "C" int sarssl_stem_head_bwd(const void* dz, const void* z, const float* stats, const void* narrow, int mode, const uint8_t* frame_flag,
                                    const int32_t* ch_idx, const float* weight64x4, float* dgamma, float* dbeta, float* dweight64x4,
                                    long long P, int W, int H, int dtype, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    SARSSL_CHECK_ARG(dz && z && stats && narrow && weight64x4 && dgamma && dbeta && dweight64x4 && workspace && P > 0 && P < 0xFFFFFFFFLL,
                     "stem_head_bwd: bad arguments (P must be in (0, 2^32))");
    SARSSL_CHECK_ARG(dtype == SARSSL_BF16, "stem_head_bwd: bf16 activations only");
    float* partials = static_cast<float*>(workspace);
    int grid = 0;
    const size_t floats = workspace_bytes / sizeof(float);
    if (floats < 512) { set_last_error("stem_head_bwd: workspace too small"); return SARSSL_ERR_WORKSPACE; }
    const int rc = launch_pw_contract<kWgHead>(dz, z, nullptr, nullptr, narrow, mode, frame_flag, ch_idx, partials + 512, floats - 512, P, W, H, stream, &grid);
    if (rc != SARSSL_OK) return rc;
    stem_reduce_kernel<<<2, 256, 0, stream>>>(partials + 512, grid, WgCfg<kWgHead>::kOut, partials, 0);
    SARSSL_LAUNCH_CHECK();
    stem_head_bwd_finalize_kernel<<<1, 256, 0, stream>>>(partials, weight64x4, stats, 1.0 / (double)P, dgamma, dbeta, dweight64x4);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

// Backward of BatchNorm + ReLU + 1x1 conv (64 -> 4) at the end of the stem (bf16 activations only): see stem_tail_bwd_finalize_kernel.
// y = pre-BatchNorm activation [P][64], stats = its (mean, rstd, scale, shift), dq = gradient w.r.t. the conv output [P][4]; writes dy [P][64].
extern "C" int sarssl_stem_tail_bwd(const void* y, const float* stats, const void* dq, const float* weight4x64, float* dgamma, float* dbeta,
                                    float* dweight4x64, void* dy, long long P, int dtype, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    SARSSL_CHECK_ARG(y && stats && dq && weight4x64 && dgamma && dbeta && dweight4x64 && dy && workspace && P > 0 && P < 0xFFFFFFFFLL,
                     "stem_tail_bwd: bad arguments (P must be in (0, 2^32))");
    SARSSL_CHECK_ARG(dtype == SARSSL_BF16, "stem_tail_bwd: bf16 activations only");
    float* partials = static_cast<float*>(workspace);           // [0, 512) reduced sums, [512, 640) c0 | c1, then the per-CTA partials
    int grid = 0;
    const size_t floats = workspace_bytes / sizeof(float);
    if (floats < 640) { set_last_error("stem_tail_bwd: workspace too small"); return SARSSL_ERR_WORKSPACE; }
    const int rc = launch_pw_contract<kWgTail>(y, nullptr, stats + 128, stats + 192, dq, 0, nullptr, nullptr, partials + 640, floats - 640, P, 0, 0, stream, &grid);
    if (rc != SARSSL_OK) return rc;
    stem_reduce_kernel<<<2, 256, 0, stream>>>(partials + 640, grid, WgCfg<kWgTail>::kOut, partials, 0);
    SARSSL_LAUNCH_CHECK();
    stem_tail_bwd_finalize_kernel<<<1, 256, 0, stream>>>(partials, weight4x64, stats, 1.0 / (double)P, dgamma, dbeta, dweight4x64, partials + 512);
    SARSSL_LAUNCH_CHECK();
    const int g2 = pix_grid(stem_tail_bwd_apply_kernel, P);
    stem_tail_bwd_apply_kernel<<<g2, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(y), static_cast<const __nv_bfloat16*>(dq), weight4x64, stats,
                                                      partials + 512, static_cast<__nv_bfloat16*>(dy), P);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

// in/out: [B][H][W][64]; weight_packed [64 n][9 taps][64 ci] of the activation dtype; optional BN+ReLU on the input
extern "C" int sarssl_conv3x3(const void* in, const float* in_scale, const float* in_shift, const void* weight_packed, void* out, int B, int H, int W,
                              int dtype, cudaStream_t stream) {
    SARSSL_CHECK_ARG(in && weight_packed && out && B > 0 && H > 0 && W > 0, "conv3x3: bad arguments");
    const long long P = (long long)B * H * W;
    const unsigned grid = (unsigned)((P + CM - 1) / CM);
    if (dtype == SARSSL_F32) conv3x3_kernel<float><<<grid, 256, 0, stream>>>((const float*)in, in_scale, in_shift, (const float*)weight_packed, (float*)out, P, H, W);
    else if (dtype == SARSSL_BF16) conv3x3_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)in, in_scale, in_shift, (const __nv_bfloat16*)weight_packed, (__nv_bfloat16*)out, P, H, W);
    else { set_last_error("conv3x3: bad dtype"); return SARSSL_ERR_ARG; }
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

// dweight_packed [64 o][9][64 ci] fp32 (+)= sum_p dy[p][o] * f(in[p + tap][ci])
extern "C" int sarssl_conv3x3_wgrad(const void* dy, const void* in, const float* in_scale, const float* in_shift, float* dweight_packed, int accumulate,
                                    int B, int H, int W, int dtype, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    SARSSL_CHECK_ARG(dy && in && dweight_packed && workspace && B > 0 && H > 0 && W > 0, "conv3x3_wgrad: bad arguments");
    const long long P = (long long)B * H * W;
    long long splits = (long long)sm_count() * 2 / 9;
    if (splits < 1) splits = 1;
    long long pps = (P + splits - 1) / splits;
    pps = (pps + CK - 1) / CK * CK;
    splits = (P + pps - 1) / pps;
    if (workspace_bytes < (size_t)splits * 64 * 576 * sizeof(float)) { set_last_error("conv3x3_wgrad: workspace too small"); return SARSSL_ERR_WORKSPACE; }
    float* partials = static_cast<float*>(workspace);
    dim3 grid(9, (unsigned)splits);
    if (dtype == SARSSL_F32) conv3x3_wgrad_kernel<float><<<grid, 256, 0, stream>>>((const float*)dy, (const float*)in, in_scale, in_shift, partials, P, H, W, pps);
    else if (dtype == SARSSL_BF16) conv3x3_wgrad_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)in, in_scale, in_shift, partials, P, H, W, pps);
    else { set_last_error("conv3x3_wgrad: bad dtype"); return SARSSL_ERR_ARG; }
    SARSSL_LAUNCH_CHECK();
    stem_reduce_kernel<<<(64 * 576 + 255) / 256, 256, 0, stream>>>(partials, (int)splits, 64 * 576, dweight_packed, accumulate);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}
