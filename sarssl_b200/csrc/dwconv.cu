// Depthwise 1-D convolution over time for channel-last tokens [B][T][D] (Conformer conv module, convolution.py:140:
// nn.Conv1d(D, D, k=31, groups=D, padding=15, bias=False); weight (D, 1, K)).  Memory-bound.
//   forward      c[b][t][d] = sum_k w[d][k] * a[b][t + k - pad][d]
//   input grad   da[b][t][d] = sum_k w[d][k] * dc[b][t - k + pad][d]      (same kernel, taps mirrored)
//   weight grad  dw[d][k] = sum_{b,t} dc[b][t][d] * a[b][t + k - pad][d]  (per-CTA partials, fixed-order final sum)
// Tiles of 64 time steps x 64 channels are staged in shared memory with their halo; threads run along channels so
// global accesses are coalesced and shared-memory accesses conflict-free.
#include "common.cuh"

namespace sarssl {

constexpr int kDwT = 64, kDwD = 64, kDwMaxK = 31;

template <typename T>
__global__ void __launch_bounds__(256) dwconv_kernel(const T* __restrict__ in, const float* __restrict__ w, T* __restrict__ out, int B, int Tn, int D,
                                                   int K, int flip) {
    __shared__ float tile[kDwT + kDwMaxK - 1][kDwD];
    __shared__ float ws[kDwMaxK][kDwD];
    const int pad = (K - 1) / 2;
    const int d0 = blockIdx.x * kDwD, t0 = blockIdx.y * kDwT, b = blockIdx.z;
    const int dl = threadIdx.x & 63, tg = threadIdx.x >> 6;
    const int d = d0 + dl;
    for (int k = tg; k < K; k += 4) ws[k][dl] = d < D ? w[(long long)d * K + (flip ? K - 1 - k : k)] : 0.f;
    for (int r = tg; r < kDwT + K - 1; r += 4) {
        const int t = t0 + r - pad;
        tile[r][dl] = (d < D && t >= 0 && t < Tn) ? to_f32(in[((long long)b * Tn + t) * D + d]) : 0.f;
    }
    __syncthreads();
    if (d >= D) return;
    for (int tt = tg; tt < kDwT; tt += 4) {
        const int t = t0 + tt;
        if (t >= Tn) break;
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc = fmaf(ws[k][dl], tile[tt + k][dl], acc);
        out[((long long)b * Tn + t) * D + d] = from_f32<T>(acc);
    }
}

// grid (D/64, nchunks); CTA loops over (b, t-tile) pairs chunk-strided; partials [nchunks][D][K]
template <typename T>
__global__ void __launch_bounds__(256) dwconv_wgrad_kernel(const T* __restrict__ a, const T* __restrict__ dc, float* __restrict__ partials, int B,
                                                         int Tn, int D, int K) {
    // 40 KB: staged tiles; re-used for the cross-thread reduction at the end (4*31*64 floats fit)
    __shared__ float buf[(kDwT + kDwMaxK - 1 + kDwT) * kDwD];
    float (*ta)[kDwD] = reinterpret_cast<float (*)[kDwD]>(buf);
    float (*tdc)[kDwD] = reinterpret_cast<float (*)[kDwD]>(buf + (kDwT + kDwMaxK - 1) * kDwD);
    float (*red)[kDwMaxK][kDwD] = reinterpret_cast<float (*)[kDwMaxK][kDwD]>(buf);
    const int pad = (K - 1) / 2;
    const int d0 = blockIdx.x * kDwD;
    const int dl = threadIdx.x & 63, tg = threadIdx.x >> 6;
    const int d = d0 + dl;
    float acc[kDwMaxK];
#pragma unroll
    for (int k = 0; k < kDwMaxK; ++k) acc[k] = 0.f;
    const int ntile = (Tn + kDwT - 1) / kDwT;
    const long long nwork = (long long)B * ntile;
    for (long long wk = blockIdx.y; wk < nwork; wk += gridDim.y) {
        const int b = (int)(wk / ntile), t0 = (int)(wk % ntile) * kDwT;
        __syncthreads();
        for (int r = tg; r < kDwT + K - 1; r += 4) {
            const int t = t0 + r - pad;
            ta[r][dl] = (d < D && t >= 0 && t < Tn) ? to_f32(a[((long long)b * Tn + t) * D + d]) : 0.f;
        }
        for (int r = tg; r < kDwT; r += 4) {
            const int t = t0 + r;
            tdc[r][dl] = (d < D && t < Tn) ? to_f32(dc[((long long)b * Tn + t) * D + d]) : 0.f;
        }
        __syncthreads();
        for (int tt = tg * (kDwT / 4); tt < (tg + 1) * (kDwT / 4); ++tt) {
            const float g = tdc[tt][dl];
#pragma unroll
            for (int k = 0; k < kDwMaxK; ++k)
                if (k < K) acc[k] = fmaf(g, ta[tt + k][dl], acc[k]);
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
