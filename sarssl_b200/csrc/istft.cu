// Inverse STFT (evaluation tail).   Reference: ISTFT.forward, common/utils_module.py:91-113
// (inv=False): torch.istft(center=False, no window) = per-frame irfft, rectangular synthesis window, overlap-add
// divided by the window envelope (number of frames covering a sample: 2 in the interior, 1 in the first and last
// half frame).  Output length (nt+1)*hop.
//
// The two channels of a pair are packed as Z = X_ch0 + i*X_ch1 (Hermitian-extended), so one 512-point complex
// transform inverts both: ifft(Z) = conj(fft(conj(Z)))/N, re = channel 0, im = channel 1.  Like c2r transforms, the
// imaginary parts of the DC and Nyquist bins are ignored.
#include "common.cuh"
#include "fft512.cuh"
#include "../../include/sarssl_b200.h"

namespace sarssl {

constexpr int kSegPerCta = 8;
constexpr int kIGroups = 4;
constexpr int kIThreads = kIGroups * kFftLanes;

__device__ __forceinline__ void igroup_bar(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(kFftLanes) : "memory"); }

// spec element (b, t, k, ch) at spec[b*sb + t*st + k*sk + ch*sc] (strides in complex elements)
// PATCH = true: `spec` is the patch layout float4 [b][t][256] = (re0, re1, im0, im1) of bins 1..256 with an implied zero DC bin
// (what pretrain_evaluate feeds the iSTFT, learner.py:581-590); nch == 2.
template <bool PATCH>
__global__ void __launch_bounds__(kIThreads) istft_kernel(const float2* __restrict__ spec, float* __restrict__ sig, int nb, int nt,
                                                        int nch, long long sb, long long st, long long sk, long long sc,
                                                        int npair, int cps) {
    __shared__ float2 acc[kSegPerCta * 256];
    __shared__ float scratch[kIGroups][kFftScratchFloats];
    const int tid = threadIdx.x, g = tid >> 6, l = tid & 63;
    int item = blockIdx.x;
    const int p = item % npair; item /= npair;
    const int sblk = item % cps;
    const int b = item / cps;
    const int s0 = sblk * kSegPerCta;                       // first output segment (256 samples each, nt+1 in total)
    const int nseg = min(kSegPerCta, nt + 1 - s0);
    const int c0 = 2 * p, c1 = 2 * p + 1;
    for (int i = tid; i < kSegPerCta * 256; i += kIThreads) acc[i] = make_float2(0.f, 0.f);
    __syncthreads();
    FftLane lane;
    lane.init(l);
    float* sre = scratch[g];
    float* sim = sre + kFftPlane;
    // frames s0-1 .. s0+nseg-1 touch these segments
    for (int fi = g; fi < nseg + 1; fi += kIGroups) {
        const int t = s0 - 1 + fi;
        if (t >= 0 && t < nt) {                             // uniform per group
            const float2* base = spec + (size_t)b * sb + (size_t)t * st;
            float2 v[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int n = l + 64 * r;                   // bin index 0..511 of conj(Z)
                const int k = n <= 256 ? n : 512 - n;
                float2 a, c;
                if (PATCH) {
                    if (k == 0) { a = make_float2(0.f, 0.f); c = a; }
                    else {
                        const float4 q = reinterpret_cast<const float4*>(spec)[((size_t)b * nt + t) * 256 + (k - 1)];
                        a = make_float2(q.x, q.z); c = make_float2(q.y, q.w);
                    }
                } else {
                    a = base[(size_t)k * sk + (size_t)c0 * sc];
                    c = c1 < nch ? base[(size_t)k * sk + (size_t)c1 * sc] : make_float2(0.f, 0.f);
                }
                if (k == 0 || k == 256) { a.y = 0.f; c.y = 0.f; }
                if (n > 256) { a.y = -a.y; c.y = -c.y; }    // Hermitian extension X[n] = conj(X[512-n])
                // Z = a + i*c ; feed conj(Z)
                v[r] = make_float2(a.x - c.y, -(a.y + c.x));
            }
            fft_pass1(v, lane, sre, sim, l);
            igroup_bar(g);
            fft_pass2_load(v, sre, sim, l);
            igroup_bar(g);
            fft_pass2_store(v, lane, sre, sim, l);
            igroup_bar(g);
            fft_pass3(v, sre, sim, l);
            igroup_bar(g);
#pragma unroll
            for (int k3 = 0; k3 < 8; ++k3) {
                const int n = l + 64 * k3;                  // time index inside the frame
                const int seg = fi - 1 + (n >> 8);          // local segment: first half -> t - s0, second half -> t - s0 + 1
                if (seg >= 0 && seg < nseg) {
                    float2* dst = &acc[seg * 256 + (n & 255)];
                    atomicAdd(&dst->x, v[k3].x * (1.0f / 512.0f));
                    atomicAdd(&dst->y, -v[k3].y * (1.0f / 512.0f));
                }
            }
        }
    }
    __syncthreads();
    const long long nsample = (long long)(nt + 1) * 256;
    for (int i = tid; i < nseg * 256; i += kIThreads) {
        const int s = s0 + (i >> 8);
        const float inv = (s == 0 || s == nt) ? 1.0f : 0.5f;
        const float2 a = acc[i];
        float* o = sig + ((size_t)b * nsample + (size_t)s0 * 256 + i) * nch;
        o[c0] = a.x * inv;
        if (c1 < nch) o[c1] = a.y * inv;
    }
}

}  // namespace sarssl

using namespace sarssl;

extern "C" int sarssl_istft(const float* spec, float* sig, int nb, int nt, int nch, long long stride_b, long long stride_t,
                            long long stride_k, long long stride_c, int win_len, int hop, int nfft, cudaStream_t stream) {
    SARSSL_CHECK_ARG(spec && sig, "istft: null pointer");
    SARSSL_CHECK_ARG(nb > 0 && nt > 0 && nch > 0, "istft: bad dims nb=%d nt=%d nch=%d", nb, nt, nch);
    if (win_len != 512 || nfft != 512 || hop != 256) {
        set_last_error("istft: only win_len = nfft = 512, hop = 256 is implemented; got %d/%d/%d", win_len, nfft, hop);
        return SARSSL_ERR_UNSUPPORTED;
    }
    const int npair = (nch + 1) / 2, cps = (nt + 1 + kSegPerCta - 1) / kSegPerCta;
    istft_kernel<false><<<nb * cps * npair, kIThreads, 0, stream>>>(reinterpret_cast<const float2*>(spec), sig, nb, nt, nch, stride_b, stride_t,
                                                                   stride_k, stride_c, npair, cps);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

// iSTFT of a patch-layout spectrogram (nb, nt, 256, 2, 2) with zero DC -> sig (nb, (nt+1)*256, 2)
extern "C" int sarssl_istft_patches(const float* patches, float* sig, int nb, int nt, cudaStream_t stream) {
    SARSSL_CHECK_ARG(patches && sig && nb > 0 && nt > 0, "istft_patches: bad arguments");
    SARSSL_CHECK_ARG(aligned16(patches), "istft_patches: patches must be 16-byte aligned");
    const int cps = (nt + 1 + kSegPerCta - 1) / kSegPerCta;
    istft_kernel<true><<<nb * cps, kIThreads, 0, stream>>>(reinterpret_cast<const float2*>(patches), sig, nb, nt, 2, 0, 0, 0, 0, 1, cps);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

namespace sarssl {
__global__ void __launch_bounds__(256) max_partial_kernel(const float* __restrict__ x, long long n, float* __restrict__ partials) {
    __shared__ float red[32];
    float m = -INFINITY;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) m = fmaxf(m, x[i]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < 8 ? red[threadIdx.x] : -INFINITY;
        m = warp_max(m);
        if (threadIdx.x == 0) partials[blockIdx.x] = m;
    }
}
__global__ void __launch_bounds__(256) scale_by_max_kernel(float* __restrict__ x, long long n, const float* __restrict__ partials, int nparts) {
    __shared__ float s_inv;
    if (threadIdx.x < 32) {
        float m = -INFINITY;
        for (int i = threadIdx.x; i < nparts; i += 32) m = fmaxf(m, partials[i]);
        m = warp_max(m);
        if (threadIdx.x == 0) s_inv = 1.0f / m;
    }
    __syncthreads();
    const float inv = s_inv;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) x[i] *= inv;
}
}  // namespace sarssl

// x /= max(x)  (pretrain_evaluate normalises the reconstructed signals by their global maximum, learner.py:584,590); workspace >= 4 KB
extern "C" int sarssl_normalize_by_max(float* x, long long n, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    SARSSL_CHECK_ARG(x && workspace && n > 0, "normalize_by_max: bad arguments");
    long long g = (n + 2047) / 2048;
    if (g > 1024) g = 1024;
    if (workspace_bytes < (size_t)g * sizeof(float)) { set_last_error("normalize_by_max: workspace too small"); return SARSSL_ERR_WORKSPACE; }
    float* partials = static_cast<float*>(workspace);
    sarssl::max_partial_kernel<<<(unsigned)g, 256, 0, stream>>>(x, n, partials);
    SARSSL_LAUNCH_CHECK();
    sarssl::scale_by_max_kernel<<<(unsigned)g, 256, 0, stream>>>(x, n, partials, (int)g);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}
