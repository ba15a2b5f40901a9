// Layout helpers around the patch ("frame-major") layout used by every kernel of the path.
//   sarssl_to_patch_layout : reference input layout (nb, nmic=2, nf, nt, 2) -> patch layout (nb, nt, nf, 2[re/im], 2[mic])
//                            = x.permute(0,2,3,4,1) + PatchSplit (model.py:524-525, utils_module.py:196-205) in one pass.
//                            Only needed when a caller hands SARSSL.forward a tensor that is NOT already our front-end's
//                            output (whose storage is patch layout, exposed as a permuted view).
//   sarssl_expand_masks    : PatchMask.forward's three dense float masks (utils_module.py:258-270) from the compact
//                            representation (one flag per (item, frame) + one microphone id per item).  API mirror only;
//                            the fused model path never materialises them.
#include "common.cuh"

namespace sarssl {

// grid (nt/32, nf/32, nb), block (32, 8)
__global__ void to_patch_layout_kernel(const float2* __restrict__ x, float4* __restrict__ out, int nf, int nt) {
    __shared__ float2 tile[2][32][33];
    const int b = blockIdx.z, t0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
    for (int m = 0; m < 2; ++m)
        for (int i = threadIdx.y; i < 32; i += 8) {
            const int f = f0 + i, t = t0 + threadIdx.x;
            if (f < nf && t < nt) tile[m][i][threadIdx.x] = x[(((size_t)b * 2 + m) * nf + f) * nt + t];
        }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int t = t0 + i, f = f0 + threadIdx.x;
        if (f < nf && t < nt) {
            const float2 a = tile[0][threadIdx.x][i], c = tile[1][threadIdx.x][i];
            out[((size_t)b * nt + t) * nf + f] = make_float4(a.x, c.x, a.y, c.y);
        }
    }
}

__global__ void expand_masks_kernel(const uint8_t* __restrict__ frame_flag, const int32_t* __restrict__ ch_idx, float* __restrict__ mask,
                                    float* __restrict__ mask_patch, float* __restrict__ mask_ch, int npatch, int dpatch, int nmic,
                                    size_t total) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(i % nmic);
        const size_t row = i / ((size_t)dpatch * nmic);             // (item, patch)
        const int b = (int)(row / npatch);
        const bool pm = frame_flag[row] != 0, cm = ch_idx[b] == m;
        mask_patch[i] = pm ? 0.f : 1.f;
        mask_ch[i] = cm ? 0.f : 1.f;
        mask[i] = (pm && cm) ? 0.f : 1.f;
    }
}

}  // namespace sarssl

using namespace sarssl;

extern "C" int sarssl_to_patch_layout(const float* x, float* patches, int nb, int nf, int nt, cudaStream_t stream) {
    SARSSL_CHECK_ARG(x && patches && nb > 0 && nf > 0 && nt > 0, "to_patch_layout: bad arguments");
    SARSSL_CHECK_ARG(aligned16(x) && aligned16(patches), "to_patch_layout: buffers must be 16-byte aligned");
    SARSSL_CHECK_ARG(nb <= 65535, "to_patch_layout: nb=%d exceeds the grid z limit", nb);
    dim3 grid((nt + 31) / 32, (nf + 31) / 32, nb), block(32, 8);
    to_patch_layout_kernel<<<grid, block, 0, stream>>>(reinterpret_cast<const float2*>(x), reinterpret_cast<float4*>(patches), nf, nt);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_expand_masks(const uint8_t* frame_flag, const int32_t* ch_idx, float* mask, float* mask_patch, float* mask_ch, int nb,
                                   int npatch, int dpatch, int nmic, cudaStream_t stream) {
    SARSSL_CHECK_ARG(frame_flag && ch_idx && mask && mask_patch && mask_ch, "expand_masks: null pointer");
    const size_t total = (size_t)nb * npatch * dpatch * nmic;
    const int blocks = (int)((total + 255) / 256 < (size_t)sm_count() * 8 ? (total + 255) / 256 : (size_t)sm_count() * 8);
    expand_masks_kernel<<<blocks, 256, 0, stream>>>(frame_flag, ch_idx, mask, mask_patch, mask_ch, npatch, dpatch, nmic, total);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}
