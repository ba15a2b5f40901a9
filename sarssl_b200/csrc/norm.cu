// Normalisation and reduction kernels of the Conformer / CNN-stem path (memory-bound, warp-shuffle reductions).
//   LayerNorm fwd/bwd            nn.LayerNorm in conformer/feed_forward.py:40, attention.py:139, convolution.py:137, Conformer.py:87
//   BatchNorm (batch statistics) nn.BatchNorm2d model.py:52-62, nn.BatchNorm1d convolution.py:141 - channel-last data [rows][C]
//   column sums                  bias / gain gradients
// All tensors are row-major [rows][cols] with fp32 or bf16 storage and fp32 math; every cross-CTA reduction goes
// through per-CTA partials that a second kernel adds in a fixed order (deterministic, final sums in double).
#include "common.cuh"

namespace sarssl {

constexpr int kLnMaxPerLane = 32;      // cols <= 1024

// ---------------------------------------------------------------- LayerNorm forward: one warp per row
template <typename T>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const T* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, T* __restrict__ out, long long ldo,
                                                   float* __restrict__ mean, float* __restrict__ rstd, int rows, int cols, float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    if (row >= rows) return;
    const T* xr = x + (long long)row * ldx;
    float v[kLnMaxPerLane];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
        const int c = lane + i * 32;
        v[i] = c < cols ? to_f32(xr[c]) : 0.f;
        s += v[i];
    }
    const float mu = warp_sum(s) / cols;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
        const int c = lane + i * 32;
        const float d = c < cols ? v[i] - mu : 0.f;
        q += d * d;
    }
    const float rs = rsqrtf(warp_sum(q) / cols + eps);
    if (lane == 0 && mean) { mean[row] = mu; rstd[row] = rs; }
    T* orow = out + (long long)row * ldo;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
        const int c = lane + i * 32;
        if (c < cols) orow[c] = from_f32<T>((v[i] - mu) * rs * gamma[c] + beta[c]);
    }
}

// ---------------------------------------------------------------- LayerNorm backward
// dx = add + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;  partial dgamma/dbeta per CTA
template <typename T>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const T* __restrict__ dy, long long lddy, const T* __restrict__ x, long long ldx,
                                                   const float* __restrict__ mean, const float* __restrict__ rstd,
                                                   const float* __restrict__ gamma, const T* __restrict__ add, T* __restrict__ dx,
                                                   float* __restrict__ partials, int rows, int cols) {
    __shared__ float sh[8][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float dg[kLnMaxPerLane], db[kLnMaxPerLane];
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) { dg[i] = 0.f; db[i] = 0.f; }
    for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
        const T* dyr = dy + (long long)row * lddy;
        const T* xr = x + (long long)row * ldx;
        const float mu = mean[row], rs = rstd[row];
        float g[kLnMaxPerLane], xh[kLnMaxPerLane];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < kLnMaxPerLane; ++i) {
            const int c = lane + i * 32;
            if (c < cols) {
                const float d = to_f32(dyr[c]);
                xh[i] = (to_f32(xr[c]) - mu) * rs;
                g[i] = d * gamma[c];
                dg[i] += d * xh[i];
                db[i] += d;
                s1 += g[i];
                s2 += g[i] * xh[i];
            } else { g[i] = 0.f; xh[i] = 0.f; }
        }
        s1 = warp_sum(s1) / cols;
        s2 = warp_sum(s2) / cols;
#pragma unroll
        for (int i = 0; i < kLnMaxPerLane; ++i) {
            const int c = lane + i * 32;
            if (c < cols) {
                float v = rs * (g[i] - s1 - xh[i] * s2);
                if (add) v += to_f32(add[(long long)row * cols + c]);
                dx[(long long)row * cols + c] = from_f32<T>(v);
            }
        }
    }
    // reduce the 8 warps' dgamma/dbeta, 32 columns at a time
    for (int i = 0; i < kLnMaxPerLane; ++i) {
        if (i * 32 >= cols) break;
        sh[warp][lane] = dg[i];
        sh[warp][32 + lane] = db[i];
        __syncthreads();
        if (warp == 0) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) { a += sh[w][lane]; b += sh[w][32 + lane]; }
            const int c = lane + i * 32;
            if (c < cols) {
                partials[((size_t)blockIdx.x * 2 + 0) * cols + c] = a;
                partials[((size_t)blockIdx.x * 2 + 1) * cols + c] = b;
            }
        }
        __syncthreads();
    }
}

// out[w] (+)= sum_p partials[p][w]  (double accumulation, fixed order)
__global__ void reduce_partials_kernel(const float* __restrict__ partials, int nparts, int width, float* __restrict__ out0,
                                       float* __restrict__ out1, int split, int accumulate) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= width) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += (double)partials[(size_t)p * width + w];
    float* dst = (w < split) ? out0 + w : out1 + (w - split);
    *dst = accumulate ? *dst + (float)s : (float)s;
}

// ---------------------------------------------------------------- column sums: partial[cta][cols]
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long long ldx, float* __restrict__ partials, int rows, int cols) {
    // thread t owns columns t, t+256, ...; rows strided over CTAs: coalesced along columns
    for (int c = threadIdx.x; c < cols; c += 256) {
        float s = 0.f;
        for (int r = blockIdx.x; r < rows; r += gridDim.x) s += to_f32(x[(long long)r * ldx + c]);
        partials[(size_t)blockIdx.x * cols + c] = s;
    }
}

// ---------------------------------------------------------------- BatchNorm over channel-last data [rows][C]
// mode 0: sums of y and y^2;  mode 1 (backward): sums of dv and dv*xhat where dv = dz * act'(bn(y))
template <typename T, int MODE>
__global__ void __launch_bounds__(256) bn_reduce_kernel(const T* __restrict__ y, const T* __restrict__ dz, const float* __restrict__ mean,
                                                      const float* __restrict__ rstd, const float* __restrict__ scale,
                                                      const float* __restrict__ shift, int act, float* __restrict__ partials,
                                                      long long rows, int C) {
    __shared__ float sh[2][256];
    const int tid = threadIdx.x;
    const int cw = C < 256 ? C : 256;               // channels covered per pass
    const int rpi = 256 / cw;                       // rows per iteration (C < 256)
    const int c0 = tid % cw, rsub = tid / cw;
    for (int cb = 0; cb < C; cb += 256) {
        const int c = cb + c0;
        float a = 0.f, b = 0.f;
        if (c < C && rsub < rpi) {
            float mu = 0.f, rs = 0.f, sc = 0.f, shf = 0.f;
            if (MODE == 1) { mu = mean[c]; rs = rstd[c]; sc = scale[c]; shf = shift[c]; }
            for (long long r = (long long)blockIdx.x * rpi + rsub; r < rows; r += (long long)gridDim.x * rpi) {
                const float v = to_f32(y[r * C + c]);
                if (MODE == 0) { a += v; b += v * v; }
                else {
                    float d = to_f32(dz[r * C + c]);
                    const float u = v * sc + shf;
                    if (act == 1) d = u > 0.f ? d : 0.f;
                    else if (act == 2) { const float sg = 1.0f / (1.0f + __expf(-u)); d *= sg * (1.0f + u * (1.0f - sg)); }
                    a += d; b += d * (v - mu) * rs;
                }
            }
        }
        sh[0][tid] = a; sh[1][tid] = b;
        __syncthreads();
        if (tid < cw && cb + tid < C) {
            float sa = 0.f, sb = 0.f;
            for (int j = 0; j < rpi; ++j) { sa += sh[0][tid + j * cw]; sb += sh[1][tid + j * cw]; }
            partials[((size_t)blockIdx.x * 2 + 0) * C + cb + tid] = sa;
            partials[((size_t)blockIdx.x * 2 + 1) * C + cb + tid] = sb;
        }
        __syncthreads();
    }
}

// partials [nparts][2][C] -> batch statistics, affine scale/shift, running-stat update (momentum, unbiased variance)
__global__ void bn_finalize_kernel(const float* __restrict__ partials, int nparts, int C, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, long long* __restrict__ num_batches, float* __restrict__ mean,
                                   float* __restrict__ rstd, float* __restrict__ scale, float* __restrict__ shift, int training) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float mu, var;
    if (training) {
        double s = 0.0, q = 0.0;
        for (int p = 0; p < nparts; ++p) { s += (double)partials[((size_t)p * 2 + 0) * C + c]; q += (double)partials[((size_t)p * 2 + 1) * C + c]; }
        const double m = s / count;
        double v = q / count - m * m;
        if (v < 0.0) v = 0.0;
        mu = (float)m; var = (float)v;
        const double unbiased = count > 1.0 ? v * count / (count - 1.0) : v;
        running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mu;
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
        if (c == 0 && num_batches) *num_batches += 1;
    } else { mu = running_mean[c]; var = running_var[c]; }
    const float rs = rsqrtf(var + eps);
    mean[c] = mu; rstd[c] = rs;
    scale[c] = gamma[c] * rs;
    shift[c] = beta[c] - mu * gamma[c] * rs;
}

// backward finalize: partials -> (sum_dv, sum_dv_xhat); dgamma += sum_dv_xhat, dbeta += sum_dv
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partials, int nparts, int C, float* __restrict__ sums,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s = 0.0, q = 0.0;
    for (int p = 0; p < nparts; ++p) { s += (double)partials[((size_t)p * 2 + 0) * C + c]; q += (double)partials[((size_t)p * 2 + 1) * C + c]; }
    sums[c] = (float)s; sums[C + c] = (float)q;
    dgamma[c] += (float)q; dbeta[c] += (float)s;
}

// z = act(y * scale + shift)
template <typename T>
__global__ void bn_act_fwd_kernel(const T* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift, int act,
                                  T* __restrict__ z, long long total, int C) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        float u = to_f32(y[i]) * scale[c] + shift[c];
        if (act == 1) u = fmaxf(u, 0.f);
        else if (act == 2) u = u / (1.0f + __expf(-u));
        z[i] = from_f32<T>(u);
    }
}

// dy = gamma * rstd * (dv - sum_dv/R - xhat * sum_dv_xhat/R),  dv = dz * act'(bn(y))
template <typename T>
__global__ void bn_act_bwd_kernel(const T* __restrict__ dz, const T* __restrict__ y, const float* __restrict__ mean,
                                  const float* __restrict__ rstd, const float* __restrict__ scale, const float* __restrict__ shift,
                                  const float* __restrict__ sums, int act, T* __restrict__ dy, long long total, int C, float inv_rows) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const float v = to_f32(y[i]);
        const float u = v * scale[c] + shift[c];
        float d = to_f32(dz[i]);
        if (act == 1) d = u > 0.f ? d : 0.f;
        else if (act == 2) { const float sg = 1.0f / (1.0f + __expf(-u)); d *= sg * (1.0f + u * (1.0f - sg)); }
        const float xh = (v - mean[c]) * rstd[c];
        dy[i] = from_f32<T>(scale[c] * (d - sums[c] * inv_rows - xh * sums[C + c] * inv_rows));
    }
}

static int capped_grid(long long work_items, int per_cta, int cap_mult) {
    long long g = (work_items + per_cta - 1) / per_cta;
    const long long cap = (long long)sm_count() * cap_mult;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace sarssl

using namespace sarssl;

#define DISPATCH_T(dtype, ...)                                                         \
    do {                                                                               \
        if ((dtype) == SARSSL_F32) { using T = float; __VA_ARGS__; }                   \
        else if ((dtype) == SARSSL_BF16) { using T = __nv_bfloat16; __VA_ARGS__; }     \
        else { set_last_error("bad dtype %d", (int)(dtype)); return SARSSL_ERR_ARG; }  \
    } while (0)

extern "C" int sarssl_layernorm_fwd(const void* x, long long ldx, const float* gamma, const float* beta, void* out, long long ldo,
                                    float* mean, float* rstd, int rows, int cols, float eps, int dtype, cudaStream_t stream) {
    SARSSL_CHECK_ARG(x && gamma && beta && out && rows > 0 && cols > 0, "layernorm_fwd: bad arguments");
    SARSSL_CHECK_ARG(cols <= 32 * kLnMaxPerLane, "layernorm_fwd: cols=%d > %d", cols, 32 * kLnMaxPerLane);
    DISPATCH_T(dtype, (ln_fwd_kernel<T><<<(rows + 7) / 8, 256, 0, stream>>>(static_cast<const T*>(x), ldx, gamma, beta, static_cast<T*>(out), ldo,
                                                                        mean, rstd, rows, cols, eps)));
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" size_t sarssl_reduce_workspace_bytes(int cols) { return (size_t)sm_count() * 8 * 2 * (size_t)cols * sizeof(float) + 256; }

extern "C" int sarssl_layernorm_bwd(const void* dy, long long lddy, const void* x, long long ldx, const float* mean, const float* rstd,
                                    const float* gamma, const void* add, void* dx, float* dgamma, float* dbeta, int rows, int cols, int dtype,
                                    void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    SARSSL_CHECK_ARG(dy && x && mean && rstd && gamma && dx && dgamma && dbeta && workspace, "layernorm_bwd: null pointer");
    SARSSL_CHECK_ARG(cols <= 32 * kLnMaxPerLane, "layernorm_bwd: cols=%d too large", cols);
    const int grid = capped_grid(rows, 8 * 4, 4);
    if (workspace_bytes < (size_t)grid * 2 * cols * sizeof(float)) { set_last_error("layernorm_bwd: workspace too small"); return SARSSL_ERR_WORKSPACE; }
    float* partials = static_cast<float*>(workspace);
    DISPATCH_T(dtype, (ln_bwd_kernel<T><<<grid, 256, 0, stream>>>(static_cast<const T*>(dy), lddy, static_cast<const T*>(x), ldx, mean, rstd, gamma,
                                                               static_cast<const T*>(add), static_cast<T*>(dx), partials, rows, cols)));
    SARSSL_LAUNCH_CHECK();
    reduce_partials_kernel<<<(2 * cols + 255) / 256, 256, 0, stream>>>(partials, grid, 2 * cols, dgamma, dbeta, cols, 1);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_colsum(const void* x, long long ldx, float* out, int rows, int cols, int dtype, int accumulate, void* workspace,
                             size_t workspace_bytes, cudaStream_t stream) {
    SARSSL_CHECK_ARG(x && out && workspace && rows > 0 && cols > 0, "colsum: bad arguments");
    const int grid = capped_grid(rows, 64, 4);
    if (workspace_bytes < (size_t)grid * cols * sizeof(float)) { set_last_error("colsum: workspace too small"); return SARSSL_ERR_WORKSPACE; }
    float* partials = static_cast<float*>(workspace);
    DISPATCH_T(dtype, (colsum_kernel<T><<<grid, 256, 0, stream>>>(static_cast<const T*>(x), ldx, partials, rows, cols)));
    SARSSL_LAUNCH_CHECK();
    reduce_partials_kernel<<<(cols + 255) / 256, 256, 0, stream>>>(partials, grid, cols, out, out, cols, accumulate);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

// stats: 4*C floats = mean, rstd, scale, shift
extern "C" int sarssl_batchnorm_stats(const void* y, long long rows, int C, const float* gamma, const float* beta, float eps, float momentum,
                                      float* running_mean, float* running_var, long long* num_batches_tracked, float* stats, int training,
                                      int dtype, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    SARSSL_CHECK_ARG(y && gamma && beta && running_mean && running_var && stats && workspace && rows > 0 && C > 0, "batchnorm_stats: bad arguments");
    const int grid = capped_grid(rows, 512, 4);
    if (workspace_bytes < (size_t)grid * 2 * C * sizeof(float)) { set_last_error("batchnorm_stats: workspace too small"); return SARSSL_ERR_WORKSPACE; }
    float* partials = static_cast<float*>(workspace);
    if (training) {
        DISPATCH_T(dtype, (bn_reduce_kernel<T, 0><<<grid, 256, 0, stream>>>(static_cast<const T*>(y), nullptr, nullptr, nullptr, nullptr, nullptr, 0,
                                                                          partials, rows, C)));
        SARSSL_LAUNCH_CHECK();
    }
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(partials, grid, C, (double)rows, gamma, beta, eps, momentum, running_mean, running_var,
                                                           num_batches_tracked, stats, stats + C, stats + 2 * C, stats + 3 * C, training);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_batchnorm_act_fwd(const void* y, const float* stats, int act, void* z, long long rows, int C, int dtype, cudaStream_t stream) {
    SARSSL_CHECK_ARG(y && stats && z && rows > 0 && C > 0, "batchnorm_act_fwd: bad arguments");
    const long long total = rows * C;
    const int grid = capped_grid(total, 1024, 16);
    DISPATCH_T(dtype, (bn_act_fwd_kernel<T><<<grid, 256, 0, stream>>>(static_cast<const T*>(y), stats + 2 * C, stats + 3 * C, act, static_cast<T*>(z),
                                                                     total, C)));
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

// dz -> dy through act and batch-statistics BN; dgamma/dbeta are accumulated (+=)
extern "C" int sarssl_batchnorm_act_bwd(const void* dz, const void* y, const float* stats, int act, void* dy, float* dgamma, float* dbeta,
                                        long long rows, int C, int dtype, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    SARSSL_CHECK_ARG(dz && y && stats && dy && dgamma && dbeta && workspace, "batchnorm_act_bwd: null pointer");
    const int grid = capped_grid(rows, 512, 4);
    if (workspace_bytes < ((size_t)grid * 2 * C + 2 * C) * sizeof(float)) { set_last_error("batchnorm_act_bwd: workspace too small"); return SARSSL_ERR_WORKSPACE; }
    float* partials = static_cast<float*>(workspace);
    float* sums = partials + (size_t)grid * 2 * C;
    const float *mean = stats, *rstd = stats + C, *scale = stats + 2 * C, *shift = stats + 3 * C;
    DISPATCH_T(dtype, (bn_reduce_kernel<T, 1><<<grid, 256, 0, stream>>>(static_cast<const T*>(y), static_cast<const T*>(dz), mean, rstd, scale, shift,
                                                                      act, partials, rows, C)));
    SARSSL_LAUNCH_CHECK();
    bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(partials, grid, C, sums, dgamma, dbeta);
    SARSSL_LAUNCH_CHECK();
    const long long total = rows * C;
    const int g2 = capped_grid(total, 1024, 16);
    DISPATCH_T(dtype, (bn_act_bwd_kernel<T><<<g2, 256, 0, stream>>>(static_cast<const T*>(dz), static_cast<const T*>(y), mean, rstd, scale, shift, sums,
                                                                   act, static_cast<T*>(dy), total, C, 1.0f / (float)rows)));
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}
