// Depthwise 1-D convolution over time for channel-last tokens [B][T][D] (Conformer conv module, convolution.py:140:
// nn.Conv1d(D, D, k=31, groups=D, padding=15, bias=False); weight (D, 1, K)).  Memory-bound.
//   forward      c[b][t][d] = sum_k w[d][k] * a[b][t + k - pad][d]
//   input grad   da[b][t][d] = sum_k w[d][k] * dc[b][t - k + pad][d]      (same kernel, taps mirrored)
//   weight grad  dw[d][k] = sum_{b,t} dc[b][t][d] * a[b][t + k - pad][d]  (per-CTA partials, fixed-order final sum)
// Tiles of 64 time steps x 64 channels are staged in shared memory with their halo; threads run along channels so
// global accesses are coalesced and shared-memory accesses conflict-free.
#include "common.cuh"
#include "vec.cuh"

namespace sarssl {

constexpr int kDwT = 64, kDwD = 64, kDwMaxK = 31, kDwRows = kDwT + kDwMaxK - 1, kDwPer = kDwT / 4;

// column -> position inside a 64-float tile row: columns 32..63 swap neighbouring float4 slots, so the eight 32-byte-strided float4
// stores a quarter warp issues while staging a row land in 8 different bank groups (they were 2-way conflicting); the compute loops
// read 32 consecutive columns per warp, which stays a permutation of 32 banks.
__device__ __forceinline__ int dw_col(int c) { return c ^ ((c >> 5) << 2); }

// rows [t0 - pad, t0 - pad + nrows) x channels [d0, d0 + 64) of `in` -> tile (fp32), zero outside the sequence / channel range.
// 16-byte loads when D % 8 == 0.
template <typename T>
__device__ __forceinline__ void dw_stage(const T* __restrict__ in, float (*tile)[kDwD], int nrows, int b, int t_first, int d0, int Tn, int D) {
    if ((D & 7) == 0) {
        const int ch = (threadIdx.x & 7) * 8, d = d0 + ch;
        const T* src = in + ((long long)b * Tn + t_first) * D + d;     // row r of the tile is src + r * D
        for (int r = threadIdx.x >> 3; r < nrows; r += 32) {
            const int t = t_first + r;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
            if (t >= 0 && t < Tn && d < D) Vec8<T>::load(src + (long long)r * D, v);
            *reinterpret_cast<float4*>(&tile[r][dw_col(ch)]) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(&tile[r][dw_col(ch + 4)]) = make_float4(v[4], v[5], v[6], v[7]);
        }
    } else {
        const int dl = threadIdx.x & 63, d = d0 + dl;
        for (int r = threadIdx.x >> 6; r < nrows; r += 4) {
            const int t = t_first + r;
            tile[r][dw_col(dl)] = (d < D && t >= 0 && t < Tn) ? to_f32(in[((long long)b * Tn + t) * D + d]) : 0.f;
        }
    }
}

// Thread (channel dl, time group tg) produces 16 consecutive outputs of its channel: the 31 taps live in registers and every staged
// input is read from shared memory once per thread (46 loads for 496 FMAs; the kernel is FMA-, not shared-memory-bound).
template <typename T>
__global__ void __launch_bounds__(256) dwconv_kernel(const T* __restrict__ in, const float* __restrict__ w, T* __restrict__ out, int B, int Tn, int D,
                                                   int K, int flip) {
    __shared__ __align__(16) float tile[kDwRows][kDwD];
    const int pad = (K - 1) / 2;
    const int d0 = blockIdx.x * kDwD, t0 = blockIdx.y * kDwT, b = blockIdx.z;
    const int dl = threadIdx.x & 63, tg = threadIdx.x >> 6;
    const int d = d0 + dl;
    __shared__ float ws[kDwD * kDwMaxK];
    dw_stage<T>(in, tile, kDwRows, b, t0 - pad, d0, Tn, D);          // all 94 rows (finite data or zeros) whatever K is
    // the 64 channels' taps are contiguous in global memory: coalesced copy to shared memory (a per-thread read of its own 31 taps
    // would touch 32 different sectors per instruction), then each thread picks its row (stride K is odd: conflict-free)
    {
        const float* wsrc = w + (long long)d0 * K;
        const int nw = min(kDwD, D - d0) * K;                        // taps of the channels that exist
        for (int i = threadIdx.x; i < kDwD * K; i += 256) ws[i] = i < nw ? wsrc[i] : 0.f;
    }
    __syncthreads();
    float wr[kDwMaxK];
#pragma unroll
    for (int k = 0; k < kDwMaxK; ++k) wr[k] = k < K ? ws[dl * K + (flip ? K - 1 - k : k)] : 0.f;
    if (d >= D) return;
    const int tb = tg * kDwPer;
    float acc[kDwPer];
#pragma unroll
    for (int o = 0; o < kDwPer; ++o) acc[o] = 0.f;
#pragma unroll
    for (int r = 0; r < kDwPer + kDwMaxK - 1; ++r) {
        const float x = tile[tb + r][dw_col(dl)];
#pragma unroll
        for (int o = 0; o < kDwPer; ++o)
            if (r - o >= 0 && r - o < kDwMaxK) acc[o] = fmaf(wr[r - o], x, acc[o]);
    }
    T* op = out + ((long long)b * Tn + (t0 + tb)) * D + d;          // one 64-bit address, then a constant stride
    const int nvalid = Tn - (t0 + tb);
#pragma unroll
    for (int o = 0; o < kDwPer; ++o)
        if (o < nvalid) op[(long long)o * D] = from_f32<T>(acc[o]);
}

// grid (D/64, nchunks); CTA loops over (b, t-tile) pairs chunk-strided; partials [nchunks][D][K].  Same register blocking as the
// forward kernel with the roles swapped: 16 output gradients in registers, inputs streamed once, 31 tap accumulators.
template <typename T>
__global__ void __launch_bounds__(256) dwconv_wgrad_kernel(const T* __restrict__ a, const T* __restrict__ dc, float* __restrict__ partials, int B,
                                                         int Tn, int D, int K) {
    // 40 KB: staged tiles; re-used for the cross-thread reduction at the end (4*31*64 floats fit)
    __shared__ __align__(16) float buf[(kDwRows + kDwT) * kDwD];
    float (*ta)[kDwD] = reinterpret_cast<float (*)[kDwD]>(buf);
    float (*tdc)[kDwD] = reinterpret_cast<float (*)[kDwD]>(buf + kDwRows * kDwD);
    float (*red)[kDwMaxK][kDwD] = reinterpret_cast<float (*)[kDwMaxK][kDwD]>(buf);
    const int pad = (K - 1) / 2;
    const int d0 = blockIdx.x * kDwD;
    const int dl = threadIdx.x & 63, tg = threadIdx.x >> 6;
    const int d = d0 + dl, tb = tg * kDwPer;
    float acc[kDwMaxK];
#pragma unroll
    for (int k = 0; k < kDwMaxK; ++k) acc[k] = 0.f;
    const int ntile = (Tn + kDwT - 1) / kDwT;
    const long long nwork = (long long)B * ntile;
    for (long long wk = blockIdx.y; wk < nwork; wk += gridDim.y) {
        const int b = (int)(wk / ntile), t0 = (int)(wk % ntile) * kDwT;
        __syncthreads();
        dw_stage<T>(a, ta, kDwRows, b, t0 - pad, d0, Tn, D);
        dw_stage<T>(dc, tdc, kDwT, b, t0, d0, Tn, D);
        __syncthreads();
        float g[kDwPer];
#pragma unroll
        for (int o = 0; o < kDwPer; ++o) g[o] = tdc[tb + o][dw_col(dl)];
#pragma unroll
        for (int r = 0; r < kDwPer + kDwMaxK - 1; ++r) {
            const float x = ta[tb + r][dw_col(dl)];
#pragma unroll
            for (int o = 0; o < kDwPer; ++o)
                if (r - o >= 0 && r - o < kDwMaxK) acc[r - o] = fmaf(g[o], x, acc[r - o]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kDwMaxK; ++k) red[tg][k][dl] = acc[k];
    __syncthreads();
    if (tg == 0 && d < D)
        for (int k = 0; k < K; ++k)
            partials[((size_t)blockIdx.y * D + d) * K + k] = red[0][k][dl] + red[1][k][dl] + red[2][k][dl] + red[3][k][dl];
}

__global__ void dw_reduce_kernel(const float* __restrict__ partials, int nparts, long long width, float* __restrict__ out) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= width) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += (double)partials[(size_t)p * width + w];
    out[w] += (float)s;
}

}  // namespace sarssl

using namespace sarssl;

extern "C" int sarssl_dwconv(const void* in, const float* weight, void* out, int B, int T_, int D, int K, int flip, int dtype, cudaStream_t stream) {
    SARSSL_CHECK_ARG(in && weight && out && B > 0 && T_ > 0 && D > 0, "dwconv: bad arguments");
    SARSSL_CHECK_ARG(K >= 1 && K <= kDwMaxK && (K & 1), "dwconv: kernel size %d not supported (odd, <= %d)", K, kDwMaxK);
    SARSSL_CHECK_ARG(B <= 65535, "dwconv: B too large");
    dim3 grid((D + kDwD - 1) / kDwD, (T_ + kDwT - 1) / kDwT, B);
    if (dtype == SARSSL_F32) dwconv_kernel<float><<<grid, 256, 0, stream>>>((const float*)in, weight, (float*)out, B, T_, D, K, flip);
    else if (dtype == SARSSL_BF16) dwconv_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)in, weight, (__nv_bfloat16*)out, B, T_, D, K, flip);
    else { set_last_error("dwconv: bad dtype"); return SARSSL_ERR_ARG; }
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" size_t sarssl_dwconv_wgrad_workspace_bytes(int D, int K) { return (size_t)sm_count() * 2 * D * K * sizeof(float); }

// dweight (D, K) fp32 is accumulated (+=)
extern "C" int sarssl_dwconv_wgrad(const void* a, const void* dc, float* dweight, int B, int T_, int D, int K, int dtype, void* workspace,
                                   size_t workspace_bytes, cudaStream_t stream) {
    SARSSL_CHECK_ARG(a && dc && dweight && workspace && B > 0 && T_ > 0 && D > 0, "dwconv_wgrad: bad arguments");
    SARSSL_CHECK_ARG(K >= 1 && K <= kDwMaxK && (K & 1), "dwconv_wgrad: kernel size %d not supported", K);
    const int ntile = (T_ + kDwT - 1) / kDwT;
    long long nchunk = (long long)B * ntile;
    const long long cap = (long long)sm_count() * 2;
    if (nchunk > cap) nchunk = cap;
    if (workspace_bytes < (size_t)nchunk * D * K * sizeof(float)) { set_last_error("dwconv_wgrad: workspace too small"); return SARSSL_ERR_WORKSPACE; }
    dim3 grid((D + kDwD - 1) / kDwD, (unsigned)nchunk);
    float* partials = static_cast<float*>(workspace);
    if (dtype == SARSSL_F32) dwconv_wgrad_kernel<float><<<grid, 256, 0, stream>>>((const float*)a, (const float*)dc, partials, B, T_, D, K);
    else if (dtype == SARSSL_BF16) dwconv_wgrad_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)dc, partials, B, T_, D, K);
    else { set_last_error("dwconv_wgrad: bad dtype"); return SARSSL_ERR_ARG; }
    SARSSL_LAUNCH_CHECK();
    const long long width = (long long)D * K;
    dw_reduce_kernel<<<(unsigned)((width + 255) / 256), 256, 0, stream>>>(partials, (int)nchunk, width, dweight);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}
