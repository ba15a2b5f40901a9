/* sarssl_b200 - C ABI of the B200-native SAR-SSL pre-training hot path (libsarssl_b200.so, sm_100a).
 *
 * The reference (Audio-WestlakeU/SAR-SSL) is pure Python/PyTorch and has no FFI of its own; the
 * boundary it exposes is the Python class surface of code/model.py, code/learner.py and
 * code/common/utils_module.py.  Every entry point below names the reference call site it replaces (paths relative
 * to /root/reference/code); the Python mirror of that surface lives in sarssl_b200/*.py and reaches this library
 * through ctypes (see INTEGRATION.md for the binding a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers + sizes only; device pointers unless the name says `host`; the caller owns every buffer;
 *   - buffers are contiguous in the documented layout and 16-byte aligned;
 *   - kernels are enqueued on the caller's `stream` (cudaStream_t passed as void*), never synchronise the device
 *     (except the two *_host helpers that say so) and keep no state between calls;
 *   - return 0 on success, a negative SARSSL_ERR_* code or a positive cudaError_t otherwise; never throw;
 *     sarssl_last_error() returns a thread-local description of the last failure;
 *   - there is no CPU fallback: without a CUDA device every compute entry fails with the CUDA error.
 */
#ifndef SARSSL_B200_H
#define SARSSL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define SARSSL_OK 0
#define SARSSL_ERR_ARG (-1)
#define SARSSL_ERR_WORKSPACE (-2)
#define SARSSL_ERR_UNSUPPORTED (-3)
#define SARSSL_ERR_NCCL (-4)

/* dtype codes for tensors that may be stored as fp32 or bf16 (math is always fp32 / fp32-accumulate) */
#define SARSSL_F32 0
#define SARSSL_BF16 1

int sarssl_version(void);
const char* sarssl_last_error(void);

/* ------------------------------------------------------------------------------------------------------------
 * A1/A2  STFT front-end
 * ---------------------------------------------------------------------------------------------------------- */

/* floor((nsample - win_len) / hop + 1)                                   common/utils_module.py:59 */
int sarssl_stft_num_frames(long long nsample, int win_len, int hop);

/* STFT.forward                                                            common/utils_module.py:49-72
 *   sig  (nb, nsample, nch) f32, channel innermost
 *   spec (nb, nt, nfft/2+1, nch) complex64 interleaved (re, im) - frame-major storage; the Python wrapper returns
 *        the (nb, nf, nt, nch) permuted view the reference returns.
 *   Periodic Hann, center=False, onesided, un-normalised.  Only win_len = nfft = 512, hop = 256 (the values the
 *   hot path uses, run_pretrain.py:67-72) are implemented; anything else returns SARSSL_ERR_UNSUPPORTED. */
int sarssl_stft_spectrum(const float* sig, float* spec, int nb, long long nsample, int nch, int win_len, int hop, int nfft,
                         cudaStream_t stream);

/* STFTLearner.data_preprocess (ch_mode 'M', fre_used_ratio 1)             learner.py:525-572, utils_module.py:124-148
 *   sig     (nb, nsample, nch) f32
 *   patches (nb*(nch-1), nt, 256, 2 [re/im], 2 [mic 0, mic j]) f32 = X / (mean_{f,t} |X_mic0| + eps), bins 1..256.
 *           This is the reference's tensor (nb*(nch-1), 2, 256, nt, 2) stored in patch order, i.e. exactly
 *           PatchSplit's output vec_patch[b,t,f,r,m] (utils_module.py:196-205); the wrapper returns the permuted view.
 *   workspace: sarssl_stft_workspace_bytes(nb, nsample, nch, generic) bytes.  nch == 2 with an even nsample takes the
 *           fused single-kernel path; otherwise (or with force_generic != 0) the three-kernel generic path runs and
 *           needs the larger `generic` workspace (SARSSL_ERR_WORKSPACE tells the caller to retry with it). */
size_t sarssl_stft_workspace_bytes(int nb, long long nsample, int nch, int generic);
int sarssl_stft_frontend(const float* sig, float* patches, int nb, long long nsample, int nch, int win_len, int hop, int nfft,
                         float eps, int force_generic, void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* Synchronises `stream`; *flag_host != 0 means the fused kernel's clip rendezvous timed out (never expected). */
int sarssl_stft_frontend_error_flag(const void* workspace, int* flag_host, cudaStream_t stream);

/* ISTFT.forward (inv=False)                                               common/utils_module.py:91-113
 *   spec: complex64, element (b, t, k, ch) at spec[2*(b*stride_b + t*stride_t + k*stride_k + ch*stride_c)] (strides in
 *         complex elements, so both the reference's (nb, nf, nt, nch) layout and our frame-major one are accepted)
 *   sig  (nb, (nt+1)*hop, nch) f32
 *   rectangular synthesis window, overlap-add divided by the number of overlapping frames (torch.istft envelope). */
int sarssl_istft(const float* spec, float* sig, int nb, int nt, int nch, long long stride_b, long long stride_t, long long stride_k,
                 long long stride_c, int win_len, int hop, int nfft, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * A4  mask indices: CPython `random` (MT19937) bit-exact, host side       common/utils_module.py:255-272,305-308
 * ---------------------------------------------------------------------------------------------------------- */

/* state_host: 625 uint32 = random.getstate()[1] (624 words + position).  Equivalent of random.seed(int):
 * key = little-endian 32-bit words of |seed| (nkey >= 1). */
int sarssl_mt19937_seed_host(uint32_t* state_host, const uint32_t* key, int nkey);
/* Per item b in order: random.sample(range(npatch), nmasked) then random.randint(0, nmic-1); advances the state.
 *   patch_idx_host (nb, nmasked) int64, ch_idx_host (nb) int64,
 *   frame_flag_host (nb, npatch) uint8: 1 where the frame is masked (nullable). */
int sarssl_mt19937_draw_masks_host(uint32_t* state_host, int nb, int npatch, int nmasked, int nmic, int64_t* patch_idx_host,
                                   int64_t* ch_idx_host, uint8_t* frame_flag_host);

/* PatchMask.forward's dense masks (API mirror; the fused path keeps masks compact)   common/utils_module.py:255-272
 *   frame_flag (nb, npatch) uint8, ch_idx (nb) int32 -> mask, mask_patch, mask_ch: (nb, npatch, dpatch, nmic) f32, 0 = masked */
int sarssl_expand_masks(const uint8_t* frame_flag, const int32_t* ch_idx, float* mask, float* mask_patch, float* mask_ch, int nb,
                        int npatch, int dpatch, int nmic, cudaStream_t stream);

/* A3  x.permute + PatchSplit in one pass                                  model.py:524-525, utils_module.py:196-205
 *   x (nb, 2, nf, nt, 2) f32 contiguous (reference input layout) -> patches (nb, nt, nf, 2, 2) */
int sarssl_to_patch_layout(const float* x, float* patches, int nb, int nf, int nt, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * A12  masked cross-channel reconstruction loss, forward + backward fused   model.py:585-592,721-747
 * ---------------------------------------------------------------------------------------------------------- */

/*   pred    (nb, nt, nf*4) pred_dtype   decoder output, inner index f*4 + reim*2 + mic            (model.py:589)
 *   patches (nb, nt, nf*4) f32          targets in patch layout (front-end output)
 *   frame_flag (nb, nt) uint8 1 = masked frame; ch_idx (nb) int32 masked microphone
 *   out2    2 floats: loss = mean (pred - tar)^2, diff = mean (tar - tar_other)^2 over (item, masked frame, bin, re/im)
 *   dpred   (nb, nt, nf*4) pred_dtype or NULL: d loss / d pred (dense: zero off the masked frames / channel)
 *   workspace: sarssl_masked_loss_workspace_bytes(nb, nt) bytes (per-CTA partial sums; deterministic reduction). */
size_t sarssl_masked_loss_workspace_bytes(int nb, int nt);
int sarssl_masked_loss(const void* pred, int pred_dtype, const float* patches, const uint8_t* frame_flag, const int32_t* ch_idx,
                       float* out2, void* dpred, int nb, int nt, int nf, int nmasked, void* workspace, size_t workspace_bytes,
                       cudaStream_t stream);
/* dpred[masked rows] *= *gscale_dev  (upstream gradient of the scalar loss, read on the device: no host sync) */
int sarssl_scale_masked_rows(void* dpred, int dtype, const uint8_t* frame_flag, const float* gscale_dev, int nb, int nt, int nf,
                             cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SARSSL_B200_H */
