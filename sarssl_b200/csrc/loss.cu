// Masked cross-channel reconstruction loss, forward + backward in one pass (sm_100a, HBM-bound).
//
// Reference: SARSSL.forward target/pred selection (model.py:585-590) + gen_loss (model.py:721-747):
//   tar   = masked-channel spectrogram at the masked frames, other = the un-masked channel there,
//   pred  = masked-channel slice of the decoder output at the masked frames,
//   loss  = mean (pred - tar)^2,  diff = mean (tar - other)^2   over (item, masked frame, bin, re/im).
// The reference gathers three (nb, nmasked, 256, 2) tensors in a per-item Python loop and autograd produces a
// dense dpred.  Here rows (item, frame) of the patch layout [f][re/im][mic] are streamed once: masked rows read
// pred + target (one float4 / 4 x bf16 per bin), accumulate both sums and write dpred; un-masked rows only write
// zeros.  Per-CTA partial sums land in the workspace and the last CTA to finish adds them in a fixed order, so
// the result is deterministic.
#include "common.cuh"
#include "../../include/sarssl_b200.h"

namespace sarssl {

struct bf16x4 { __nv_bfloat162 lo, hi; };

__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
    const uint2 raw = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&raw.x), b = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
    const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void store4(float* p, float4 v) { st_stream_f4(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ void store4(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 raw;
    raw.x = *reinterpret_cast<unsigned*>(&a);
    raw.y = *reinterpret_cast<unsigned*>(&b);
    *reinterpret_cast<uint2*>(p) = raw;
}

constexpr int kLossThreads = 256;

// ws: [0] = arrival counter (uint), then 2 floats per CTA
template <typename T>
__global__ void __launch_bounds__(kLossThreads) masked_loss_kernel(const T* __restrict__ pred, const float* __restrict__ patches,
                                                                 const uint8_t* __restrict__ frame_flag, const int32_t* __restrict__ ch_idx,
                                                                 float* __restrict__ out2, T* __restrict__ dpred, int nrows, int nt, int nf,
                                                                 float inv_count, unsigned* counter, float* partials) {
    __shared__ float red[32];
    __shared__ bool last;
    float acc_loss = 0.f, acc_diff = 0.f;
    const float gscale = 2.0f * inv_count;
    for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
        const bool masked = frame_flag[row] != 0;
        const size_t base = (size_t)row * nf * 4;
        if (!masked) {
            if (dpred != nullptr)
                for (int f = threadIdx.x; f < nf; f += kLossThreads) store4(dpred + base + (size_t)f * 4, make_float4(0.f, 0.f, 0.f, 0.f));
            continue;
        }
        const int mc = ch_idx[row / nt];
        for (int f = threadIdx.x; f < nf; f += kLossThreads) {
            const float4 p = load4(pred + base + (size_t)f * 4);          // (re0, re1, im0, im1)
            const float4 x = __ldcs(reinterpret_cast<const float4*>(patches + base + (size_t)f * 4));
            const float tr = mc ? x.y : x.x, ti = mc ? x.w : x.z;         // masked-channel target
            const float orr = mc ? x.x : x.y, oi = mc ? x.z : x.w;        // other channel
            const float dr = (mc ? p.y : p.x) - tr, di = (mc ? p.w : p.z) - ti;
            acc_loss += dr * dr + di * di;
            acc_diff += (tr - orr) * (tr - orr) + (ti - oi) * (ti - oi);
            if (dpred != nullptr) {
                const float gr = gscale * dr, gi = gscale * di;
                store4(dpred + base + (size_t)f * 4, mc ? make_float4(0.f, gr, 0.f, gi) : make_float4(gr, 0.f, gi, 0.f));
            }
        }
    }
    const float sl = block_sum(acc_loss, red);
    const float sd = block_sum(acc_diff, red);
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x] = sl;
        partials[2 * blockIdx.x + 1] = sd;
        __threadfence();
        last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x < 32) {
        __threadfence();
        float a = 0.f, d = 0.f;
        for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) { a += __ldcg(&partials[2 * i]); d += __ldcg(&partials[2 * i + 1]); }
        a = warp_sum(a);
        d = warp_sum(d);
        if (threadIdx.x == 0) { out2[0] = a * inv_count; out2[1] = d * inv_count; }
    }
}

// dpred rows *= *g (upstream gradient of the scalar loss), masked rows only
template <typename T>
__global__ void scale_masked_rows_kernel(T* __restrict__ dpred, const uint8_t* __restrict__ frame_flag, const float* __restrict__ g,
                                         int nrows, int nf) {
    const float s = *g;
    for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
        if (!frame_flag[row]) continue;
        const size_t base = (size_t)row * nf * 4;
        for (int f = threadIdx.x; f < nf; f += blockDim.x) {
            float4 v = load4(dpred + base + (size_t)f * 4);
            v.x *= s; v.y *= s; v.z *= s; v.w *= s;
            store4(dpred + base + (size_t)f * 4, v);
        }
    }
}

// pretrain_evaluate metrics (learner.py:592-602) over patch-layout tensors: sums[0] = sum (pred-gt)^2 over everything,
// sums[1] = the same restricted to masked frame x masked channel (where the reference's dense mask is 0).  Per-CTA partials, fixed order.
__global__ void __launch_bounds__(256) eval_mse_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const uint8_t* __restrict__ frame_flag,
                                                     const int32_t* __restrict__ ch_idx, float* __restrict__ partials, int nrows, int nt, int nf) {
    __shared__ float red[32];
    float all = 0.f, msk = 0.f;
    for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
        const bool masked = frame_flag[row] != 0;
        const int mc = ch_idx[row / nt];
        const size_t base = (size_t)row * nf * 4;
        for (int f = threadIdx.x; f < nf; f += 256) {
            const float4 p = load4(pred + base + (size_t)f * 4), x = load4(gt + base + (size_t)f * 4);
            const float d0 = (p.x - x.x) * (p.x - x.x) + (p.z - x.z) * (p.z - x.z);      // mic 0: re, im
            const float d1 = (p.y - x.y) * (p.y - x.y) + (p.w - x.w) * (p.w - x.w);      // mic 1
            all += d0 + d1;
            if (masked) msk += mc ? d1 : d0;
        }
    }
    const float sa = block_sum(all, red), sm = block_sum(msk, red);
    if (threadIdx.x == 0) { partials[2 * blockIdx.x] = sa; partials[2 * blockIdx.x + 1] = sm; }
}
__global__ void eval_mse_finalize_kernel(const float* __restrict__ partials, int n, float* __restrict__ out2) {
    if (threadIdx.x < 2) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += (double)partials[2 * i + threadIdx.x];
        out2[threadIdx.x] = (float)s;
    }
}

static int loss_grid(int nrows) {
    const int cap = sm_count() * 8;
    return nrows < cap ? nrows : cap;
}

}  // namespace sarssl

using namespace sarssl;

extern "C" size_t sarssl_masked_loss_workspace_bytes(int nb, int nt) {
    (void)nb; (void)nt;
    return 256 + (size_t)sm_count() * 8 * 2 * sizeof(float);
}

extern "C" int sarssl_masked_loss(const void* pred, int pred_dtype, const float* patches, const uint8_t* frame_flag,
                                  const int32_t* ch_idx, float* out2, void* dpred, int nb, int nt, int nf, int nmasked, void* workspace,
                                  size_t workspace_bytes, cudaStream_t stream) {
    SARSSL_CHECK_ARG(pred && patches && frame_flag && ch_idx && out2 && workspace, "masked_loss: null pointer");
    SARSSL_CHECK_ARG(nb > 0 && nt > 0 && nf > 0 && nmasked > 0, "masked_loss: bad dims nb=%d nt=%d nf=%d nmasked=%d", nb, nt, nf, nmasked);
    SARSSL_CHECK_ARG(pred_dtype == SARSSL_F32 || pred_dtype == SARSSL_BF16, "masked_loss: bad dtype %d", pred_dtype);
    SARSSL_CHECK_ARG(aligned16(pred) && aligned16(patches) && (!dpred || aligned16(dpred)), "masked_loss: buffers must be 16-byte aligned");
    if (workspace_bytes < sarssl_masked_loss_workspace_bytes(nb, nt)) {
        set_last_error("masked_loss: workspace %zu < required %zu", workspace_bytes, sarssl_masked_loss_workspace_bytes(nb, nt));
        return SARSSL_ERR_WORKSPACE;
    }
    unsigned* counter = static_cast<unsigned*>(workspace);
    float* partials = reinterpret_cast<float*>(static_cast<unsigned char*>(workspace) + 256);
    SARSSL_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned), stream));
    const int nrows = nb * nt;
    const float inv_count = 1.0f / ((float)nb * (float)nmasked * (float)nf * 2.0f);
    const int grid = loss_grid(nrows);
    if (pred_dtype == SARSSL_F32)
        masked_loss_kernel<float><<<grid, kLossThreads, 0, stream>>>(static_cast<const float*>(pred), patches, frame_flag, ch_idx, out2,
                                                                     static_cast<float*>(dpred), nrows, nt, nf, inv_count, counter, partials);
    else
        masked_loss_kernel<__nv_bfloat16><<<grid, kLossThreads, 0, stream>>>(static_cast<const __nv_bfloat16*>(pred), patches, frame_flag,
                                                                             ch_idx, out2, static_cast<__nv_bfloat16*>(dpred), nrows, nt, nf,
                                                                             inv_count, counter, partials);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

// sums2[0] = sum over all elements of (pred - gt)^2; sums2[1] = the same over masked frame x masked channel only
extern "C" int sarssl_eval_mse_sums(const float* pred, const float* gt, const uint8_t* frame_flag, const int32_t* ch_idx, float* sums2, int nb, int nt,
                                    int nf, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    SARSSL_CHECK_ARG(pred && gt && frame_flag && ch_idx && sums2 && workspace && nb > 0 && nt > 0 && nf > 0, "eval_mse_sums: bad arguments");
    if (workspace_bytes < sarssl_masked_loss_workspace_bytes(nb, nt)) { set_last_error("eval_mse_sums: workspace too small"); return SARSSL_ERR_WORKSPACE; }
    const int nrows = nb * nt, grid = loss_grid(nrows);
    float* partials = reinterpret_cast<float*>(static_cast<unsigned char*>(workspace) + 256);
    eval_mse_kernel<<<grid, 256, 0, stream>>>(pred, gt, frame_flag, ch_idx, partials, nrows, nt, nf);
    SARSSL_LAUNCH_CHECK();
    eval_mse_finalize_kernel<<<1, 32, 0, stream>>>(partials, grid, sums2);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_scale_masked_rows(void* dpred, int dtype, const uint8_t* frame_flag, const float* gscale_dev, int nb, int nt, int nf,
                                        cudaStream_t stream) {
    SARSSL_CHECK_ARG(dpred && frame_flag && gscale_dev, "scale_masked_rows: null pointer");
    const int nrows = nb * nt, grid = loss_grid(nrows);
    if (dtype == SARSSL_F32)
        scale_masked_rows_kernel<float><<<grid, 256, 0, stream>>>(static_cast<float*>(dpred), frame_flag, gscale_dev, nrows, nf);
    else
        scale_masked_rows_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<__nv_bfloat16*>(dpred), frame_flag, gscale_dev, nrows, nf);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}
