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
/* The variants of data_preprocess off the pre-training hot path (three-kernel generic route; workspace: generic = 1):
 *   all_pairs != 0: ch_mode 'MM' - every channel pair a < b in AddChToBatch's order (utils_module.py:136-143), nb * nch (nch - 1) / 2 output items;
 *   bins [first_bin, first_bin + nbins) of the 257: fre_used_ratio 1 -> (1, 256), 0.5 -> (0, 128) (learner.py:514-517).
 *   patches (items, nt, nbins, 2, 2). */
int sarssl_stft_frontend_ex(const float* sig, float* patches, int nb, long long nsample, int nch, int win_len, int hop, int nfft, float eps,
                            int all_pairs, int first_bin, int nbins, void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* Synchronises `stream`; *flag_host != 0 means the fused kernel's clip rendezvous timed out (never expected) in some launch
 * on this workspace: the flag is sticky (only a fresh zeroed workspace clears it), so one check per epoch covers every step.
 * The workspace must be zero-initialised when it is first handed to sarssl_stft_frontend. */
int sarssl_stft_frontend_error_flag(const void* workspace, int* flag_host, cudaStream_t stream);

/* ISTFT.forward (inv=False)                                               common/utils_module.py:91-113
 *   spec: complex64, element (b, t, k, ch) at spec[2*(b*stride_b + t*stride_t + k*stride_k + ch*stride_c)] (strides in
 *         complex elements, so both the reference's (nb, nf, nt, nch) layout and our frame-major one are accepted)
 *   sig  (nb, (nt+1)*hop, nch) f32
 *   rectangular synthesis window, overlap-add divided by the number of overlapping frames (torch.istft envelope). */
int sarssl_istft(const float* spec, float* sig, int nb, int nt, int nch, long long stride_b, long long stride_t, long long stride_k,
                 long long stride_c, int win_len, int hop, int nfft, cudaStream_t stream);

/* iSTFT of a patch-layout spectrogram (nb, nt, 256, 2, 2) = bins 1..256 with an implied zero DC bin - what
 * pretrain_evaluate inverts (learner.py:581-590) -> sig (nb, (nt+1)*256, 2) */
int sarssl_istft_patches(const float* patches, float* sig, int nb, int nt, cudaStream_t stream);

/* x /= max(x) over n floats (pretrain_evaluate, learner.py:584,590); workspace >= 4 KB */
int sarssl_normalize_by_max(float* x, long long n, void* workspace, size_t workspace_bytes, cudaStream_t stream);

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
/* pretrain_evaluate's error sums (learner.py:592-602) over patch-layout fp32 tensors (nb, nt, nf, 2, 2):
 *   sums2[0] = sum (pred - gt)^2 over all elements, sums2[1] = the same over masked frame x masked microphone.
 *   workspace: sarssl_masked_loss_workspace_bytes(nb, nt). */
int sarssl_eval_mse_sums(const float* pred, const float* gt, const uint8_t* frame_flag, const int32_t* ch_idx, float* sums2, int nb, int nt, int nf,
                         void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* dpred[masked rows] *= *gscale_dev  (upstream gradient of the scalar loss, read on the device: no host sync) */
int sarssl_scale_masked_rows(void* dpred, int dtype, const uint8_t* frame_flag, const float* gscale_dev, int nb, int nt, int nf,
                             cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * A7-A11  dense contractions (torch.nn.Linear / Conv1d(k=1) / matmul of the reference, e.g. conformer/feed_forward.py:47-54,
 * attention.py:72-103, convolution.py:136-146, model.py:297-301) and their gradients
 * ---------------------------------------------------------------------------------------------------------- */

#define SARSSL_ACT_NONE 0
#define SARSSL_ACT_RELU 1
#define SARSSL_ACT_SWISH 2

/* C[z][m][n] = resid[z][m][n] + beta * dropout( act( alpha * sum_k A[z][m][k] * B[z][n][k] + bias[n] ) ),  z = (z1, z2)
 *   element strides: A(m,k) at A + z1*sAb1 + z2*sAb2 + m*sAm + k*sAk; B(n,k) likewise; C/pre_out/resid row-major with ldc/ldr.
 *   ab_dtype / c_dtype: SARSSL_F32 or SARSSL_BF16 storage, fp32 accumulation.  bias is fp32 (nullable).
 *   pre_out (nullable, C's dtype/layout) receives the pre-activation value.  accumulate != 0: C += result.
 *   drop_p/drop_seed: epilogue dropout keyed by the C element offset.  a_drop_p/a_drop_seed: multiply A by the dropout mask
 *   (and 1/(1-p)) keyed by the A element offset - how backward re-applies a forward mask without storing it. */
typedef struct sarssl_gemm_args {
    const void* A; const void* B; void* C; void* pre_out; const void* resid; const float* bias;
    long long sAm, sAk, sAb1, sAb2;
    long long sBn, sBk, sBb1, sBb2;
    long long ldc, sCb1, sCb2, ldr;
    int M, N, K, nb1, nb2;
    int ab_dtype, c_dtype, act, accumulate;
    float alpha, beta;
    float drop_p; unsigned long long drop_seed;
    float a_drop_p; unsigned long long a_drop_seed;
    const unsigned long long* seed_dev;   /* nullable device word ADDED to both seeds inside the kernel: the per-step part of the dropout seed
                                           * when the step is replayed as a CUDA graph (launch arguments are frozen in a graph) */
} sarssl_gemm_args;
/* CUDA-core kernel (exact fp32 when ab_dtype is F32): any shape / stride. */
int sarssl_gemm(const sarssl_gemm_args* args, cudaStream_t stream);
/* Tensor-core kernel: tcgen05.mma with TMEM accumulators, TMA-staged 128B-swizzled operand tiles, warp-specialised
 * (TMA / MMA / epilogue).  Same argument block; requires bf16 operands, no batching, no A-side dropout, each operand K-major
 * (unit stride along K) or MN-major (unit stride along M / N) - forward = K x K, data gradient = K x MN, weight gradient =
 * MN x MN -, leading dimensions multiple of 8 elements, 16-byte aligned bases.  Returns SARSSL_ERR_UNSUPPORTED otherwise (call sarssl_gemm). */
int sarssl_gemm_tc(const sarssl_gemm_args* args, cudaStream_t stream);


/* ------------------------------------------------------------------------------------------------------------
 * A7-A10  memory-bound Conformer pieces.  `dtype` = storage type of the activation tensors (SARSSL_F32 / SARSSL_BF16);
 * parameters, statistics and parameter gradients are always fp32; gradients of parameters are ACCUMULATED (+=).
 * Reduction workspaces: sarssl_reduce_workspace_bytes(cols) bytes.
 * ---------------------------------------------------------------------------------------------------------- */
size_t sarssl_reduce_workspace_bytes(int cols);

/* nn.LayerNorm (eps 1e-5)        conformer/feed_forward.py:40, attention.py:139, convolution.py:137, Conformer.py:87
 *   x rows at stride ldx, out rows at stride ldo (lets the final LayerNorm write into the concatenated decoder input);
 *   mean/rstd (rows) fp32 are saved for the backward (nullable for inference). */
int sarssl_layernorm_fwd(const void* x, long long ldx, const float* gamma, const float* beta, void* out, long long ldo, float* mean,
                         float* rstd, int rows, int cols, float eps, int dtype, cudaStream_t stream);
/*   dx = add + dLN(dy)   (add nullable: the residual branch's gradient; dx/add contiguous [rows][cols]) */
int sarssl_layernorm_bwd(const void* dy, long long lddy, const void* x, long long ldx, const float* mean, const float* rstd,
                         const float* gamma, const void* add, void* dx, float* dgamma, float* dbeta, int rows, int cols, int dtype,
                         void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* out[c] (+)= sum_r x[r][c]      (bias gradients, u/v bias gradients) */
int sarssl_colsum(const void* x, long long ldx, float* out, int rows, int cols, int dtype, int accumulate, void* workspace,
                  size_t workspace_bytes, cudaStream_t stream);

/* nn.BatchNorm2d / nn.BatchNorm1d over channel-last data [rows][C]      model.py:52-62, conformer/convolution.py:141
 *   training != 0: batch statistics (biased variance) + running-stat update (momentum, unbiased variance, counter);
 *   training == 0: running statistics.  stats = 4*C floats: mean, rstd, scale = gamma*rstd, shift = beta - mean*scale. */
int sarssl_batchnorm_stats(const void* y, long long rows, int C, const float* gamma, const float* beta, float eps, float momentum,
                           float* running_mean, float* running_var, long long* num_batches_tracked, float* stats, int training, int dtype,
                           void* workspace, size_t workspace_bytes, cudaStream_t stream);
/*   the same from per-CTA partial sums [nparts][2][C] produced elsewhere (training mode only) */
int sarssl_batchnorm_finalize(const float* partials, int nparts, long long rows, int C, const float* gamma, const float* beta, float eps,
                              float momentum, float* running_mean, float* running_var, long long* num_batches_tracked, float* stats,
                              cudaStream_t stream);
/*   z = act(y*scale + shift), act: SARSSL_ACT_RELU (stem) / SARSSL_ACT_SWISH (conv module) / NONE */
int sarssl_batchnorm_act_fwd(const void* y, const float* stats, int act, void* z, long long rows, int C, int dtype, cudaStream_t stream);
/*   dz -> dy through the activation and the batch-statistics normalisation; dgamma/dbeta accumulated */
int sarssl_batchnorm_act_bwd(const void* dz, const void* y, const float* stats, int act, void* dy, float* dgamma, float* dbeta,
                             long long rows, int C, int dtype, void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* Swish / ReLU / GLU gradients and GLU forward                              conformer/activation.py:19-42
 *   swish_bwd: du = ds * dropout_mask(offset)/(1-p) * swish'(u)  (regenerates the FFN dropout mask, feed_forward.py:51) */
int sarssl_swish_bwd(const void* ds, const void* u, void* du, long long n, float drop_p, unsigned long long seed, const unsigned long long* seed_dev, int dtype,
                     cudaStream_t stream);
/*   dst = alpha * src * dropout_mask(offset)/(1-p): gradient of `x + alpha*Dropout(v)` w.r.t. v (modules.py:33 + the Dropout sites) */
int sarssl_scale_dropout(const void* src, void* dst, long long n, float alpha, float drop_p, unsigned long long seed, const unsigned long long* seed_dev, int dtype,
                         cudaStream_t stream);
int sarssl_relu_bwd(const void* dz, const void* z, void* dy, long long n, int dtype, cudaStream_t stream);
int sarssl_glu_fwd(const void* g, void* a, long long rows, int D, int dtype, cudaStream_t stream);
int sarssl_glu_bwd(const void* da, const void* g, void* dg, long long rows, int D, int dtype, cudaStream_t stream);

/* relative-position attention glue                                         conformer/attention.py:87-97,105-113
 *   add_head_bias: qu = q + u_bias, qv = q + v_bias (q = first D columns of rows with stride ld)
 *   attn_softmax_fwd: prob[b][h][i][:] = softmax_j((content[b][h][i][j] + shift(pos[h][b])[i][j]) * scale); attn_dropped
 *                     (nullable) = Dropout(prob) with the mask keyed by the element offset (attention.py:98)
 *   attn_softmax_bwd: dattn (in place) -> dscore; dpos[h][b] = inverse shift of dscore; dropout mask regenerated */
int sarssl_add_head_bias(const void* q, long long ld, const float* u_bias, const float* v_bias, void* qu, void* qv, long long rows, int D,
                         int dtype, cudaStream_t stream);
int sarssl_attn_softmax_fwd(const void* content, const void* pos, void* prob, void* attn_dropped, int B, int H, int T, float scale,
                            float drop_p, unsigned long long seed, const unsigned long long* seed_dev, int dtype, cudaStream_t stream);
int sarssl_attn_softmax_bwd(void* dattn_inout, const void* prob, void* dpos, int B, int H, int T, float scale, float drop_p,
                            unsigned long long seed, const unsigned long long* seed_dev, int dtype, cudaStream_t stream);
/* (seed_dev, nullable, everywhere it appears: a device word added to `seed` inside the kernel - see sarssl_gemm_args.seed_dev) */
int sarssl_add2(const void* a, long long lda, const void* b, long long ldb, void* out, long long ldo, long long rows, int cols, int dtype,
                cudaStream_t stream);

/* depthwise Conv1d over time, channel-last [B][T][D], weight (D, K) fp32, K odd <= 31      conformer/convolution.py:140
 *   flip = 0: forward; flip = 1: input gradient (mirrored taps).  wgrad accumulates into dweight. */
int sarssl_dwconv(const void* in, const float* weight, void* out, int B, int T, int D, int K, int flip, int dtype, cudaStream_t stream);
size_t sarssl_dwconv_wgrad_workspace_bytes(int D, int K);
int sarssl_dwconv_wgrad(const void* a, const void* dc, float* dweight, int B, int T, int D, int K, int dtype, void* workspace,
                        size_t workspace_bytes, cudaStream_t stream);

/* time-mean pooling of the downstream branch (torch.mean(embed, dim=1), model.py:705) and its gradient:
 *   pooled[b][d] = mean_t x[b][t][d] (fp32);   dx[b][t][d] = dpooled[b][d] / T      (x / dx rows at stride ldx) */
int sarssl_mean_pool_fwd(const void* x, long long ldx, float* pooled, int B, int T, int D, int dtype, cudaStream_t stream);
int sarssl_mean_pool_bwd(const float* dpooled, void* dx, long long ldx, int B, int T, int D, int dtype, cudaStream_t stream);

/* utilities: dtype cast, 4-d permuting copy (weight packing / gradient un-packing), fill */
int sarssl_cast(const void* src, int src_dtype, void* dst, int dst_dtype, long long n, cudaStream_t stream);
int sarssl_permute4(const void* src, int src_dtype, void* dst, int dst_dtype, const int* dims4_host, const long long* src_strides4_host,
                    int accumulate, cudaStream_t stream);
int sarssl_fill_f32(float* p, float value, long long n, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Input pipeline in front of the path (next row 4)                      dataset.py:142-178 (soundfile.read)
 * Host functions: RIFF/WAVE decode (PCM 8/16/24/32, float 32/64) to float32 [nsample][nch], scaled like libsndfile.
 * ---------------------------------------------------------------------------------------------------------- */
int sarssl_wav_info(const char* path, int* fs, int* nch, long long* nsample);
/* frames [first, first + count) -> out_host[count][nch]; frames past the end are zero-filled, *nread = real frames */
int sarssl_wav_read_f32(const char* path, long long first, long long count, float* out_host, long long* nread);
/* A whole batch in one call (the DataLoader workers + default_collate of run_pretrain.py:191-199): file i -> out_host[i][count][nch], decoded by
 * `nthreads` host threads.  Every file must have nch channels at fs Hz (fs <= 0: any) and, with exact != 0, exactly first + count frames (else frames
 * past the end are zero-filled).  On error *bad_index (nullable) is the first failing file. */
int sarssl_wav_read_batch_f32(const char* const* paths, int nfiles, long long first, long long count, int nch, int fs, int exact, float* out_host,
                              int nthreads, int* bad_index);

/* ------------------------------------------------------------------------------------------------------------
 * A6  CNN patch-embedding stem on channel-last images [B][H = frame][W = bin][C]      model.py:50-64,203-208
 * ---------------------------------------------------------------------------------------------------------- */
/* 1x1 conv 4 -> 64.  mode 0: `in` is [P][4] of `dtype`; mode 1 / 2: `in` is the fp32 patch tensor and the spectral /
 * spatial input masking of model.py:541 / :563 is applied on load (frame_flag (B,H) uint8, ch_idx (B) int32); mode 3: fp32 patches,
 * un-masked (downstream branch, model.py:676-678); mode 4: spectral input of the frozen-encoder branch (model.py:622: masked frames keep
 * their un-masked microphone, the other frames are zero). */
int sarssl_stem_expand(const void* in, int mode, const uint8_t* frame_flag, const int32_t* ch_idx, const float* weight64x4, void* out,
                       long long P, int W, int H, int dtype, cudaStream_t stream);
/* conv 4 -> 64 + BatchNorm + ReLU in one pass: out = relu(scale * conv(in) + shift).  The batch statistics of the conv output are
 * known beforehand: the conv is linear in 4 channels, so sarssl_stem_input_stats derives sum(y), sum(y^2) from the input's first and
 * second moments (one pass over the [P][4] input); feed sums2x64 to sarssl_batchnorm_finalize with nparts = 1. */
int sarssl_stem_expand_bn_relu(const void* in, int mode, const uint8_t* frame_flag, const int32_t* ch_idx, const float* weight64x4,
                               const float* scale, const float* shift, void* out, long long P, int W, int H, int dtype, cudaStream_t stream);
int sarssl_stem_input_stats(const void* in, int mode, const uint8_t* frame_flag, const int32_t* ch_idx, const float* weight64x4, float* sums2x64,
                            long long P, int W, int H, int dtype, void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* 1x1 conv 64 -> 4 with optional BatchNorm+ReLU of the input applied on load (in_scale/in_shift nullable) */
int sarssl_stem_reduce(const void* in, const float* in_scale, const float* in_shift, const float* weight4x64, void* out, long long P,
                       int dtype, cudaStream_t stream);
size_t sarssl_stem_workspace_bytes(void);
/* dW[o][c] (64x4) (+)= sum_p f(wide[p][o]) * narrow[p][c];  f = ReLU(BN) when wide_scale given; narrow loaded like stem_expand */
int sarssl_stem_pw_wgrad(const void* wide, const float* wide_scale, const float* wide_shift, const void* narrow, int mode,
                         const uint8_t* frame_flag, const int32_t* ch_idx, float* dweight64x4, int accumulate, long long P, int W, int H,
                         int dtype, void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* Fused backward of the first stem layer (1x1 conv 4 -> 64, BatchNorm, ReLU; model.py:51-53), bf16 activations only: ONE pass over
 * dz (gradient w.r.t. the ReLU output) and z (that output) yields dweight64x4 +=, dgamma +=, dbeta +=; no P x 64 tensor is written
 * (this layer's input gradient is never needed).  stats = (mean, rstd, scale, shift)[64] of its BatchNorm; narrow as stem_expand. */
int sarssl_stem_head_bwd(const void* dz, const void* z, const float* stats, const void* narrow, int mode, const uint8_t* frame_flag,
                         const int32_t* ch_idx, const float* weight64x4, float* dgamma, float* dbeta, float* dweight64x4,
                         long long P, int W, int H, int dtype, void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* Fused backward of BatchNorm + ReLU + 1x1 conv 64 -> 4 at the end of the stem (model.py:58-60), bf16 activations only.
 * y = pre-BatchNorm activation [P][64] with stats (mean, rstd, scale, shift)[64]; dq = gradient w.r.t. the conv output [P][4].
 * dweight4x64 +=, dgamma +=, dbeta +=, and dy [P][64] (gradient w.r.t. y) is written; the P x 64 gradient of the ReLU output
 * is recomputed in registers, never stored. */
int sarssl_stem_tail_bwd(const void* y, const float* stats, const void* dq, const float* weight4x64, float* dgamma, float* dbeta,
                         float* dweight4x64, void* dy, long long P, int dtype, void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* 3x3 conv 64 -> 64, padding 1, as implicit GEMM; weight_packed [n][tap = (dh+1)*3 + (dw+1)][ci] in `dtype`;
 * BatchNorm+ReLU of the INPUT applied on load when in_scale given (zero padding stays zero). */
int sarssl_conv3x3(const void* in, const float* in_scale, const float* in_shift, const void* weight_packed, void* out, int B, int H, int W,
                   int dtype, cudaStream_t stream);
int sarssl_conv3x3_wgrad(const void* dy, const void* in, const float* in_scale, const float* in_shift, float* dweight_packed, int accumulate,
                         int B, int H, int W, int dtype, void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* Tensor-core versions (bf16 only): tcgen05 implicit GEMM, TMA row boxes shared by the three horizontal taps, weights
 * resident in shared memory, persistent CTAs.  in_scale / in_shift (both nullable, 64 floats each): `in` is the PRE-BatchNorm tensor and
 * the operand is relu(in_scale * in + in_shift) (model.py:52-56), applied to the landed row boxes in shared memory by transform warps, so
 * the post-BatchNorm/ReLU activation never exists in HBM; null: `in` is used as it is (already activated input, or dy for the data
 * gradient, with the mirrored weight pack).  wgrad workspace: sarssl_conv3x3_wgrad_tc_workspace_bytes(). */
/*   bn_partials (nullable): sarssl_conv3x3_tc_grid(B,H,W) x 2 x 64 floats receive per-CTA sums of out / out^2 (fused BatchNorm batch
 *   statistics of the conv output); turn them into stats with sarssl_batchnorm_finalize. */
int sarssl_conv3x3_tc_grid(int B, int H, int W);
int sarssl_conv3x3_tc(const void* in, const void* weight_packed, void* out, float* bn_partials, const float* in_scale, const float* in_shift,
                      int B, int H, int W, cudaStream_t stream);
size_t sarssl_conv3x3_wgrad_tc_workspace_bytes(void);
int sarssl_conv3x3_wgrad_tc(const void* dy, const void* in, const float* in_scale, const float* in_shift, float* dweight_packed, int accumulate,
                            int B, int H, int W, void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * A14  optimizer: torch.optim.Adam(betas (0.9, 0.999), eps 1e-8, wd 0) over flat fp32 arenas      learner.py:83,111-113
 * ---------------------------------------------------------------------------------------------------------- */
/* g is multiplied by grad_scale first; param_bf16 (nullable) receives the refreshed bf16 copy; zero_grad != 0 clears g.
 * Hyper-parameters are doubles: 1 - beta, the bias corrections and lr / (1 - beta1^step) are formed in double on the host and rounded
 * to fp32 once, exactly as torch.optim.Adam does with its Python floats. */
int sarssl_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* param_bf16, long long n, int step, double lr,
                     double beta1, double beta2, double eps, float grad_scale, int zero_grad, const float* hyper_dev, cudaStream_t stream);
/* hyper_dev (nullable): two device floats {lr / (1 - beta1^step), sqrt(1 - beta2^step)} that replace the values derived from `step` and `lr`
 * (a step replayed as a CUDA graph cannot change launch arguments); sarssl_adam_hyper_host fills them on the host exactly as above. */
int sarssl_adam_hyper_host(float* hyper2_host, int step, double lr, double beta1, double beta2);

/* ------------------------------------------------------------------------------------------------------------
 * A15  data-parallel gradient exchange (replaces nn.DataParallel, learner.py:25-31): NCCL sum all-reduce of contiguous buckets
 * of the flat fp32 gradient arena; one communicator per process; the only process-global state of the library.
 * ---------------------------------------------------------------------------------------------------------- */
int sarssl_comm_unique_id_bytes(void);
int sarssl_comm_get_unique_id(void* id_host);                 /* rank 0; distribute the bytes to the other ranks out of band */
int sarssl_comm_init(int rank, int world, const void* id_host);
int sarssl_comm_world_size(void);
int sarssl_allreduce_sum_f32(float* buf, size_t n, cudaStream_t stream);
int sarssl_comm_destroy(void);

/* ------------------------------------------------------------------------------------------------------------
 * diagnostics
 * ---------------------------------------------------------------------------------------------------------- */
/* Hardware probe used by tests/test_tc_probe_gpu.py: one UMMA tile whose A descriptor starts row_off (0..7) 128-byte rows into
 * a TMA-written 128B-swizzled tile (mode 0: K-major A [136][64]; mode 1: MN-major A [72][128], B [64][64]); D f32 [128][64]. */
int sarssl_probe_umma_row_offset(const void* A, const void* B, float* D, int mode, int row_off, int use_base_offset, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SARSSL_B200_H */
