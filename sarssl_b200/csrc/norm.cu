// Normalisation and reduction kernels of the Conformer / CNN-stem path (memory-bound, warp-shuffle reductions).
//   LayerNorm fwd/bwd            nn.LayerNorm in conformer/feed_forward.py:40, attention.py:139, convolution.py:137, Conformer.py:87
//   BatchNorm (batch statistics) nn.BatchNorm2d model.py:52-62, nn.BatchNorm1d convolution.py:141 - channel-last data [rows][C]
//   column sums                  bias / gain gradients
// All tensors are row-major [rows][cols] with fp32 or bf16 storage and fp32 math; every cross-CTA reduction goes
// through per-CTA partials that a second kernel adds in a fixed order (deterministic, final sums in double).
#include "common.cuh"
#include "vec.cuh"

namespace sarssl {

constexpr int kLnChunks = 4;           // cols <= 1024, cols % 8 == 0: lane owns elements [c*256 + lane*8, +8) of every 256-column chunk c

// ---------------------------------------------------------------- LayerNorm forward: one warp per row, R rows in flight per warp
// NCH = 256-column chunks per row (compile time); rows row0 + q * (gridDim.x * 8), q < R, are loaded before the first reduction.
template <typename T, int NCH, int R>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const T* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, T* __restrict__ out, long long ldo,
                                                   float* __restrict__ mean, float* __restrict__ rstd, int rows, int cols, float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * 8 + warp, rstride = gridDim.x * 8;
    float v[R][NCH][8];
#pragma unroll
    for (int q = 0; q < R; ++q) {
        const int row = row0 + q * rstride;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int col = c * 256 + lane * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) v[q][c][j] = 0.f;
            if (row < rows && col < cols) Vec8<T>::load(x + (long long)row * ldx + col, v[q][c]);
        }
    }
    float g[NCH][8], bt[NCH][8];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int col = c * 256 + lane * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) { g[c][j] = 0.f; bt[c][j] = 0.f; }
        if (col < cols) { load8f(gamma + col, g[c]); load8f(beta + col, bt[c]); }
    }
#pragma unroll
    for (int q = 0; q < R; ++q) {
        const int row = row0 + q * rstride;
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v[q][c][j];                 // (columns past `cols` hold zeros)
        const float mu = warp_sum(s) / cols;
        float qq = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            if (c * 256 + lane * 8 < cols) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { const float d = v[q][c][j] - mu; qq += d * d; }
            }
        }
        const float rs = rsqrtf(warp_sum(qq) / cols + eps);
        if (row >= rows) continue;
        if (lane == 0 && mean) { mean[row] = mu; rstd[row] = rs; }
        T* orow = out + (long long)row * ldo;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int col = c * 256 + lane * 8;
            if (col < cols) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = (v[q][c][j] - mu) * rs * g[c][j] + bt[c][j];
                Vec8<T>::store(orow + col, o);
            }
        }
    }
}

// ---------------------------------------------------------------- LayerNorm backward
// dx = add + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;  partial dgamma/dbeta per CTA.
// One warp per row; NCH = 256-column chunks per row (compile time, so a 256-wide model does not pay the registers of a 1024-wide
// one) and R rows in flight per warp: all loads of the R rows are issued before the first reduction (the kernel is latency-bound).
template <typename T, int NCH, int R>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const T* __restrict__ dy, long long lddy, const T* __restrict__ x, long long ldx,
                                                   const float* __restrict__ mean, const float* __restrict__ rstd,
                                                   const float* __restrict__ gamma, const T* __restrict__ add, T* __restrict__ dx,
                                                   float* __restrict__ partials, int rows, int cols) {
    __shared__ float sh[8][2][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float dg[NCH][8], db[NCH][8], gm[NCH][8];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int col = c * 256 + lane * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) { dg[c][j] = 0.f; db[c][j] = 0.f; gm[c][j] = 0.f; }
        if (col < cols) load8f(gamma + col, gm[c]);
    }
    const int rstride = gridDim.x * 8;
    for (int row0 = blockIdx.x * 8 + warp; row0 < rows; row0 += rstride * R) {
        float d[R][NCH][8], xv[R][NCH][8], av[R][NCH][8];
        float mu[R], rs[R];
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int row = row0 + q * rstride;
            const bool ok = row < rows;
            mu[q] = ok ? mean[row] : 0.f; rs[q] = ok ? rstd[row] : 0.f;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int col = c * 256 + lane * 8;
#pragma unroll
                for (int j = 0; j < 8; ++j) { d[q][c][j] = 0.f; xv[q][c][j] = 0.f; av[q][c][j] = 0.f; }
                if (ok && col < cols) {
                    Vec8<T>::load(dy + (long long)row * lddy + col, d[q][c]);
                    Vec8<T>::load(x + (long long)row * ldx + col, xv[q][c]);
                    if (add) Vec8<T>::load(add + (long long)row * cols + col, av[q][c]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int row = row0 + q * rstride;
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float xh = (xv[q][c][j] - mu[q]) * rs[q];
                    const float g = d[q][c][j] * gm[c][j];
                    dg[c][j] = fmaf(d[q][c][j], xh, dg[c][j]);
                    db[c][j] += d[q][c][j];
                    s1 += g;
                    s2 = fmaf(g, xh, s2);
                    xv[q][c][j] = xh; d[q][c][j] = g;
                }
            s1 = warp_sum(s1) / cols;
            s2 = warp_sum(s2) / cols;
            if (row < rows) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const int col = c * 256 + lane * 8;
                    if (col < cols) {
                        float o[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] = av[q][c][j] + rs[q] * (d[q][c][j] - s1 - xv[q][c][j] * s2);
                        Vec8<T>::store(dx + (long long)row * cols + col, o);
                    }
                }
            }
        }
    }
    // reduce the 8 warps' dgamma/dbeta, one 256-column chunk at a time
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        if (c * 256 >= cols) break;
#pragma unroll
        for (int j = 0; j < 8; ++j) { sh[warp][0][lane * 8 + j] = dg[c][j]; sh[warp][1][lane * 8 + j] = db[c][j]; }
        __syncthreads();
        {
            const int col = c * 256 + threadIdx.x;
            if (col < cols) {
                float a = 0.f, bsum = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) { a += sh[w][0][threadIdx.x]; bsum += sh[w][1][threadIdx.x]; }
                partials[((size_t)blockIdx.x * 2 + 0) * cols + col] = a;
                partials[((size_t)blockIdx.x * 2 + 1) * cols + col] = bsum;
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// out[w] (+)= sum_p partials[p][w]  (double accumulation, fixed order).  CTA = 32 columns x 8 partial slices: every thread walks
// a slice of the partials of its column (coalesced across the 32 columns), the 8 slices are combined through shared memory.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partials, int nparts, int width, float* __restrict__ out0,
                                                            float* __restrict__ out1, int split, int accumulate) {
    __shared__ double sh[8][33];
    const int cl = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int w = blockIdx.x * 32 + cl;
    double s = 0.0;
    if (w < width)
        for (int p = sl; p < nparts; p += 8) s += (double)partials[(size_t)p * width + w];
    sh[sl][cl] = s;
    __syncthreads();
    if (sl == 0 && w < width) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sh[k][cl];
        float* dst = (w < split) ? out0 + w : out1 + (w - split);
        *dst = accumulate ? *dst + (float)t : (float)t;
    }
}

// ---------------------------------------------------------------- column sums: partial[cta][cols]
// CTA = 256 threads = 32 column groups (8 columns each, one 16-byte load) x 8 row lanes; grid.x tiles the columns by 256,
// grid.y strides the rows.
// ATOMIC: the CTA's column sums are added straight into `partials` (= the fp32 output, [cols]) with red.global.add - no second
// kernel; the summation order then varies from run to run, like the split-K weight gradients these bias gradients sit next to.
template <typename T, bool ATOMIC>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long long ldx, float* __restrict__ partials, int rows, int cols) {
    __shared__ float sh[8][256];
    const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int col = blockIdx.x * 256 + cg * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (col < cols) {
        for (int r = blockIdx.y * 8 + rl; r < rows; r += gridDim.y * 8) {
            float v[8];
            Vec8<T>::load(x + (long long)r * ldx + col, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += v[j];
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sh[rl][cg * 8 + j] = acc[j];
    __syncthreads();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < cols) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += sh[w][threadIdx.x];
        if (ATOMIC) atomicAdd(partials + c, s);
        else partials[(size_t)blockIdx.y * cols + c] = s;
    }
}

// ---------------------------------------------------------------- BatchNorm over channel-last data [rows][C]
// mode 0: sums of y and y^2;  mode 1 (backward): sums of dv and dv*xhat where dv = dz * act'(bn(y)).
// Each thread owns one group of 8 consecutive elements of the flattened tensor per iteration (a 16-byte load); the thread's
// stride (gridDim.x * 256 * 8 elements) is a multiple of C or C a multiple of 8 groups, so its channel set is fixed:
//   C % 8 == 0 : channels cbase .. cbase+7 with cbase = (group * 8) % C  -> requires (256 * 8 * gridDim.x) % C == 0
//   C == 4     : channels 0..3 twice
template <typename T, int MODE>
__global__ void __launch_bounds__(256) bn_reduce_kernel(const T* __restrict__ y, const T* __restrict__ dz, const float* __restrict__ mean,
                                                      const float* __restrict__ rstd, const float* __restrict__ scale,
                                                      const float* __restrict__ shift, int act, float* __restrict__ partials,
                                                      long long total, int C) {
    __shared__ float sa[256][9], sb[256][9];      // per-thread partials, combined in a fixed order below (deterministic)
    const long long g0 = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long gs = (long long)gridDim.x * 256, ng = total / 8;
    const int cbase = (C == 4) ? 0 : (int)((g0 * 8) % C);
    // mode 1 accumulates sum(dv) and sum(dv * (y - mean)); rstd is applied once at the end (fewer live registers -> more loads in flight)
    float a[8], b[8], mu[8], sc[8], shf[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        a[j] = 0.f; b[j] = 0.f;
        const int c = (C == 4) ? (j & 3) : cbase + j;
        if (MODE == 1) { mu[j] = mean[c]; sc[j] = scale[c]; shf[j] = shift[c]; }
    }
    auto accumulate = [&](const float (&v)[8], const float (&d)[8]) {
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { a[j] += v[j]; b[j] += v[j] * v[j]; }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float u = v[j] * sc[j] + shf[j];
                float dv = d[j];
                if (act == 1) dv = u > 0.f ? dv : 0.f;
                else if (act == 2) { const float sg = 1.0f / (1.0f + __expf(-u)); dv *= sg * (1.0f + u * (1.0f - sg)); }
                a[j] += dv; b[j] = fmaf(dv, v[j] - mu[j], b[j]);
            }
        }
    };
    constexpr int U = 4;                          // independent 16-byte loads in flight per thread and operand
    long long g = g0;
    for (; g + (U - 1) * gs < ng; g += U * gs) {
        float v[U][8], d[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            Vec8<T>::load(y + (g + u * gs) * 8, v[u]);
            if (MODE == 1) Vec8<T>::load(dz + (g + u * gs) * 8, d[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) accumulate(v[u], d[u]);
    }
    for (; g < ng; g += gs) {
        float v[8], d[8];
        Vec8<T>::load(y + g * 8, v);
        if (MODE == 1) Vec8<T>::load(dz + g * 8, d);
        accumulate(v, d);
    }
    if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] *= rstd[(C == 4) ? (j & 3) : cbase + j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { sa[threadIdx.x][j] = a[j]; sb[threadIdx.x][j] = b[j]; }
    __syncthreads();
    // channel c is held by threads t with (t*8) % C == (c/8)*8, slot c%8 (C % 8 == 0, blockIdx.x*2048 % C == 0); C == 4: every thread, slots c and c+4
    for (int c = threadIdx.x; c < C; c += 256) {
        float ta = 0.f, tb = 0.f;
        if (C == 4) {
            for (int t = 0; t < 256; ++t) { ta += sa[t][c] + sa[t][c + 4]; tb += sb[t][c] + sb[t][c + 4]; }
        } else {
            const int per = C / 8;                  // threads t = c/8 + k*per
            for (int t = c / 8; t < 256; t += per) { ta += sa[t][c & 7]; tb += sb[t][c & 7]; }
        }
        partials[((size_t)blockIdx.x * 2 + 0) * C + c] = ta;
        partials[((size_t)blockIdx.x * 2 + 1) * C + c] = tb;
    }
}

// partials [nparts][2][C] -> batch statistics, affine scale/shift, running-stat update (momentum, unbiased variance).
// One warp per channel: lanes stride the partials, fixed-order double sums.
__global__ void __launch_bounds__(256) bn_finalize_kernel(const float* __restrict__ partials, int nparts, int C, double count, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, float momentum, float* __restrict__ running_mean,
                                                        float* __restrict__ running_var, long long* __restrict__ num_batches, float* __restrict__ mean,
                                                        float* __restrict__ rstd, float* __restrict__ scale, float* __restrict__ shift, int training) {
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= C) return;
    float mu, var;
    if (training) {
        double s = 0.0, q = 0.0;
        for (int p = lane; p < nparts; p += 32) { s += (double)partials[((size_t)p * 2 + 0) * C + c]; q += (double)partials[((size_t)p * 2 + 1) * C + c]; }
        s = warp_sum_d(s); q = warp_sum_d(q);
        const double m = s / count;
        double v = q / count - m * m;
        if (v < 0.0) v = 0.0;
        mu = (float)m; var = (float)v;
        if (lane == 0) {
            const double unbiased = count > 1.0 ? v * count / (count - 1.0) : v;
            running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mu;
            running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
            if (c == 0 && num_batches) *num_batches += 1;
        }
    } else { mu = running_mean[c]; var = running_var[c]; }
    if (lane == 0) {
        const float rs = rsqrtf(var + eps);
        mean[c] = mu; rstd[c] = rs;
        scale[c] = gamma[c] * rs;
        shift[c] = beta[c] - mu * gamma[c] * rs;
    }
}

// backward finalize: partials -> (sum_dv, sum_dv_xhat); dgamma += sum_dv_xhat, dbeta += sum_dv
__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const float* __restrict__ partials, int nparts, int C, float* __restrict__ sums,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= C) return;
    double s = 0.0, q = 0.0;
    for (int p = lane; p < nparts; p += 32) { s += (double)partials[((size_t)p * 2 + 0) * C + c]; q += (double)partials[((size_t)p * 2 + 1) * C + c]; }
    s = warp_sum_d(s); q = warp_sum_d(q);
    if (lane == 0) {
        sums[c] = (float)s; sums[C + c] = (float)q;
        dgamma[c] += (float)q; dbeta[c] += (float)s;
    }
}

// z = act(y * scale + shift).  Launched with bn_grid: every thread keeps one fixed group of 8 channels, so scale/shift live in registers.
template <typename T>
__global__ void __launch_bounds__(256) bn_act_fwd_kernel(const T* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift, int act,
                                                       T* __restrict__ z, long long total, int C) {
    const long long g0 = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long gs = (long long)gridDim.x * 256, ng = total / 8;
    const int cbase = (C == 4) ? 0 : (int)((g0 * 8) % C);
    float sc[8], shf[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { const int c = (C == 4) ? (j & 3) : cbase + j; sc[j] = scale[c]; shf[j] = shift[c]; }
    auto apply = [&](float (&v)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float u = v[j] * sc[j] + shf[j];
            if (act == 1) u = fmaxf(u, 0.f);
            else if (act == 2) u = __fdividef(u, 1.0f + __expf(-u));
            v[j] = u;
        }
    };
    constexpr int U = 4;
    long long g = g0;
    for (; g + (U - 1) * gs < ng; g += U * gs) {
        float v[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u) Vec8<T>::load(y + (g + u * gs) * 8, v[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) { apply(v[u]); Vec8<T>::store(z + (g + u * gs) * 8, v[u]); }
    }
    for (; g < ng; g += gs) {
        float v[8];
        Vec8<T>::load(y + g * 8, v);
        apply(v);
        Vec8<T>::store(z + g * 8, v);
    }
}

// dy = gamma * rstd * (dv - sum_dv/R - xhat * sum_dv_xhat/R),  dv = dz * act'(bn(y)); per channel this is dy = scale*dv - c0 - c1*y with
// c1 = scale * rstd * sum_dv_xhat / R and c0 = scale * sum_dv / R - c1 * mean (scale = gamma * rstd).  Same fixed-channel layout as above.
template <typename T>
__global__ void __launch_bounds__(256) bn_act_bwd_kernel(const T* __restrict__ dz, const T* __restrict__ y, const float* __restrict__ mean,
                                                       const float* __restrict__ rstd, const float* __restrict__ scale, const float* __restrict__ shift,
                                                       const float* __restrict__ sums, int act, T* __restrict__ dy, long long total, int C, float inv_rows) {
    const long long g0 = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long gs = (long long)gridDim.x * 256, ng = total / 8;
    const int cbase = (C == 4) ? 0 : (int)((g0 * 8) % C);
    float sc[8], shf[8], c0[8], c1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = (C == 4) ? (j & 3) : cbase + j;
        sc[j] = scale[c]; shf[j] = shift[c];
        c1[j] = sc[j] * rstd[c] * sums[C + c] * inv_rows;
        c0[j] = sc[j] * sums[c] * inv_rows - c1[j] * mean[c];
    }
    auto apply = [&](const float (&v)[8], float (&d)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float u = v[j] * sc[j] + shf[j];
            float dv = d[j];
            if (act == 1) dv = u > 0.f ? dv : 0.f;
            else if (act == 2) { const float sg = 1.0f / (1.0f + __expf(-u)); dv *= sg * (1.0f + u * (1.0f - sg)); }
            d[j] = fmaf(sc[j], dv, -fmaf(c1[j], v[j], c0[j]));
        }
    };
    constexpr int U = 4;
    long long g = g0;
    for (; g + (U - 1) * gs < ng; g += U * gs) {
        float v[U][8], d[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u) { Vec8<T>::load(y + (g + u * gs) * 8, v[u]); Vec8<T>::load(dz + (g + u * gs) * 8, d[u]); }
#pragma unroll
        for (int u = 0; u < U; ++u) { apply(v[u], d[u]); Vec8<T>::store(dy + (g + u * gs) * 8, d[u]); }
    }
    for (; g < ng; g += gs) {
        float v[8], d[8];
        Vec8<T>::load(y + g * 8, v);
        Vec8<T>::load(dz + g * 8, d);
        apply(v, d);
        Vec8<T>::store(dy + g * 8, d);
    }
}

// number of CTAs for the channel-last BN kernels: every thread must keep a fixed channel set, i.e. (grid * 2048) % C == 0
static int bn_grid(long long total, int C, int cap_ctas) {
    long long g = (total / 8 + 255) / 256;
    const long long cap = cap_ctas;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    if (C > 8 && (2048 % C) != 0) {              // C = 512, 768, ...: grid must be a multiple of C / gcd(C, 2048)
        long long a = C, b2 = 2048;
        while (b2) { long long t = a % b2; a = b2; b2 = t; }
        const long long unit = C / a;
        g = (g / unit) * unit;
        if (g < unit) g = unit;
    }
    return (int)g;
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
    SARSSL_CHECK_ARG(cols <= 256 * kLnChunks && cols % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0, "layernorm_fwd: cols=%d must be a multiple of 8, <= %d", cols,
                     256 * kLnChunks);
#define LN_FWD_LAUNCH(NCH, R)                                                                                                                  \
    DISPATCH_T(dtype, (ln_fwd_kernel<T, NCH, R><<<(rows + 8 * R - 1) / (8 * R), 256, 0, stream>>>(static_cast<const T*>(x), ldx, gamma, beta,    \
                                                                                              static_cast<T*>(out), ldo, mean, rstd, rows, cols, eps)))
    if (cols <= 256) LN_FWD_LAUNCH(1, 4);
    else if (cols <= 512) LN_FWD_LAUNCH(2, 2);
    else LN_FWD_LAUNCH(4, 1);
#undef LN_FWD_LAUNCH
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" size_t sarssl_reduce_workspace_bytes(int cols) { return (size_t)sm_count() * 8 * 2 * (size_t)cols * sizeof(float) + 256; }

extern "C" int sarssl_layernorm_bwd(const void* dy, long long lddy, const void* x, long long ldx, const float* mean, const float* rstd,
                                    const float* gamma, const void* add, void* dx, float* dgamma, float* dbeta, int rows, int cols, int dtype,
                                    void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    SARSSL_CHECK_ARG(dy && x && mean && rstd && gamma && dx && dgamma && dbeta && workspace, "layernorm_bwd: null pointer");
    SARSSL_CHECK_ARG(cols <= 256 * kLnChunks && cols % 8 == 0 && ldx % 8 == 0 && lddy % 8 == 0, "layernorm_bwd: cols=%d must be a multiple of 8, <= %d", cols,
                     256 * kLnChunks);
    float* partials = static_cast<float*>(workspace);
#define LN_BWD_LAUNCH(NCH, R)                                                                                                                    \
    do {                                                                                                                                         \
        int grid = capped_grid(rows, 8 * R, 64);                                                                                                 \
        const int cap = dtype == SARSSL_F32 ? resident_ctas(ln_bwd_kernel<float, NCH, R>, 256) : resident_ctas(ln_bwd_kernel<__nv_bfloat16, NCH, R>, 256); \
        if (grid > cap) grid = cap;                                                                                                              \
        if (workspace_bytes < (size_t)grid * 2 * cols * sizeof(float)) { set_last_error("layernorm_bwd: workspace too small"); return SARSSL_ERR_WORKSPACE; } \
        DISPATCH_T(dtype, (ln_bwd_kernel<T, NCH, R><<<grid, 256, 0, stream>>>(static_cast<const T*>(dy), lddy, static_cast<const T*>(x), ldx, mean, rstd, gamma, \
                                                                            static_cast<const T*>(add), static_cast<T*>(dx), partials, rows, cols)));    \
        nparts = grid;                                                                                                                           \
    } while (0)
    int nparts = 0;
    if (cols <= 256) LN_BWD_LAUNCH(1, 2);
    else if (cols <= 512) LN_BWD_LAUNCH(2, 1);
    else LN_BWD_LAUNCH(4, 1);
#undef LN_BWD_LAUNCH
    SARSSL_LAUNCH_CHECK();
    reduce_partials_kernel<<<(2 * cols + 31) / 32, 256, 0, stream>>>(partials, nparts, 2 * cols, dgamma, dbeta, cols, 1);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_colsum(const void* x, long long ldx, float* out, int rows, int cols, int dtype, int accumulate, void* workspace,
                             size_t workspace_bytes, cudaStream_t stream) {
    SARSSL_CHECK_ARG(x && out && workspace && rows > 0 && cols > 0, "colsum: bad arguments");
    SARSSL_CHECK_ARG(cols % 8 == 0 && ldx % 8 == 0, "colsum: cols and ldx must be multiples of 8");
    const int gx = (cols + 255) / 256;
    int grid = capped_grid(rows, 64, 64);
    {
        const int cap = dtype == SARSSL_F32 ? resident_ctas(colsum_kernel<float, false>, 256) : resident_ctas(colsum_kernel<__nv_bfloat16, false>, 256);
        if (grid > cap) grid = cap;
    }
    grid /= gx;
    if (grid < 1) grid = 1;
    if (accumulate && dtype == SARSSL_BF16) {       // bf16 training path: one launch, partial sums added with red.global.add (see colsum_kernel)
        colsum_kernel<__nv_bfloat16, true><<<dim3(gx, grid), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ldx, out, rows, cols);
        SARSSL_LAUNCH_CHECK();
        return SARSSL_OK;
    }
    if (workspace_bytes < (size_t)grid * cols * sizeof(float)) { set_last_error("colsum: workspace too small"); return SARSSL_ERR_WORKSPACE; }
    float* partials = static_cast<float*>(workspace);
    DISPATCH_T(dtype, (colsum_kernel<T, false><<<dim3(gx, grid), 256, 0, stream>>>(static_cast<const T*>(x), ldx, partials, rows, cols)));
    SARSSL_LAUNCH_CHECK();
    reduce_partials_kernel<<<(cols + 31) / 32, 256, 0, stream>>>(partials, grid, cols, out, out, cols, accumulate);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

// stats: 4*C floats = mean, rstd, scale, shift
extern "C" int sarssl_batchnorm_stats(const void* y, long long rows, int C, const float* gamma, const float* beta, float eps, float momentum,
                                      float* running_mean, float* running_var, long long* num_batches_tracked, float* stats, int training,
                                      int dtype, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    SARSSL_CHECK_ARG(y && gamma && beta && running_mean && running_var && stats && workspace && rows > 0 && C > 0, "batchnorm_stats: bad arguments");
    SARSSL_CHECK_ARG(C == 4 || C % 8 == 0, "batchnorm: C=%d must be 4 or a multiple of 8", C);
    SARSSL_CHECK_ARG((rows * C) % 8 == 0, "batchnorm: rows*C must be a multiple of 8");
    const int grid = bn_grid(rows * C, C, dtype == SARSSL_F32 ? resident_ctas(bn_reduce_kernel<float, 0>, 256) : resident_ctas(bn_reduce_kernel<__nv_bfloat16, 0>, 256));
    if (workspace_bytes < (size_t)grid * 2 * C * sizeof(float)) { set_last_error("batchnorm_stats: workspace too small"); return SARSSL_ERR_WORKSPACE; }
    float* partials = static_cast<float*>(workspace);
    if (training) {
        DISPATCH_T(dtype, (bn_reduce_kernel<T, 0><<<grid, 256, 0, stream>>>(static_cast<const T*>(y), nullptr, nullptr, nullptr, nullptr, nullptr, 0,
                                                                          partials, rows * C, C)));
        SARSSL_LAUNCH_CHECK();
    }
    bn_finalize_kernel<<<(C + 7) / 8, 256, 0, stream>>>(partials, grid, C, (double)rows, gamma, beta, eps, momentum, running_mean, running_var,
                                                           num_batches_tracked, stats, stats + C, stats + 2 * C, stats + 3 * C, training);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

// same as sarssl_batchnorm_stats, but the per-CTA partial sums [nparts][2][C] (sum, sum of squares) were produced by another kernel
// (the tensor-core conv epilogue)
extern "C" int sarssl_batchnorm_finalize(const float* partials, int nparts, long long rows, int C, const float* gamma, const float* beta, float eps,
                                         float momentum, float* running_mean, float* running_var, long long* num_batches_tracked, float* stats,
                                         cudaStream_t stream) {
    SARSSL_CHECK_ARG(partials && gamma && beta && running_mean && running_var && stats && nparts > 0 && rows > 0 && C > 0, "batchnorm_finalize: bad arguments");
    bn_finalize_kernel<<<(C + 7) / 8, 256, 0, stream>>>(partials, nparts, C, (double)rows, gamma, beta, eps, momentum, running_mean, running_var,
                                                       num_batches_tracked, stats, stats + C, stats + 2 * C, stats + 3 * C, 1);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_batchnorm_act_fwd(const void* y, const float* stats, int act, void* z, long long rows, int C, int dtype, cudaStream_t stream) {
    SARSSL_CHECK_ARG(y && stats && z && rows > 0 && C > 0, "batchnorm_act_fwd: bad arguments");
    SARSSL_CHECK_ARG((C == 4 || C % 8 == 0) && (rows * C) % 8 == 0, "batchnorm_act_fwd: C=%d must be 4 or a multiple of 8", C);
    const long long total = rows * C;
    const int grid = bn_grid(total, C, dtype == SARSSL_F32 ? resident_ctas(bn_act_fwd_kernel<float>, 256) : resident_ctas(bn_act_fwd_kernel<__nv_bfloat16>, 256));
    DISPATCH_T(dtype, (bn_act_fwd_kernel<T><<<grid, 256, 0, stream>>>(static_cast<const T*>(y), stats + 2 * C, stats + 3 * C, act, static_cast<T*>(z),
                                                                     total, C)));
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

// dz -> dy through act and batch-statistics BN; dgamma/dbeta are accumulated (+=)
extern "C" int sarssl_batchnorm_act_bwd(const void* dz, const void* y, const float* stats, int act, void* dy, float* dgamma, float* dbeta,
                                        long long rows, int C, int dtype, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    SARSSL_CHECK_ARG(dz && y && stats && dy && dgamma && dbeta && workspace, "batchnorm_act_bwd: null pointer");
    SARSSL_CHECK_ARG((C == 4 || C % 8 == 0) && (rows * C) % 8 == 0, "batchnorm_act_bwd: C=%d must be 4 or a multiple of 8", C);
    const int grid = bn_grid(rows * C, C, dtype == SARSSL_F32 ? resident_ctas(bn_reduce_kernel<float, 1>, 256) : resident_ctas(bn_reduce_kernel<__nv_bfloat16, 1>, 256));
    if (workspace_bytes < ((size_t)grid * 2 * C + 2 * C) * sizeof(float)) { set_last_error("batchnorm_act_bwd: workspace too small"); return SARSSL_ERR_WORKSPACE; }
    float* partials = static_cast<float*>(workspace);
    float* sums = partials + (size_t)grid * 2 * C;
    const float *mean = stats, *rstd = stats + C, *scale = stats + 2 * C, *shift = stats + 3 * C;
    DISPATCH_T(dtype, (bn_reduce_kernel<T, 1><<<grid, 256, 0, stream>>>(static_cast<const T*>(y), static_cast<const T*>(dz), mean, rstd, scale, shift,
                                                                      act, partials, rows * C, C)));
    SARSSL_LAUNCH_CHECK();
    bn_bwd_finalize_kernel<<<(C + 7) / 8, 256, 0, stream>>>(partials, grid, C, sums, dgamma, dbeta);
    SARSSL_LAUNCH_CHECK();
    const long long total = rows * C;
    const int g2 = bn_grid(total, C, dtype == SARSSL_F32 ? resident_ctas(bn_act_bwd_kernel<float>, 256) : resident_ctas(bn_act_bwd_kernel<__nv_bfloat16>, 256));
    DISPATCH_T(dtype, (bn_act_bwd_kernel<T><<<g2, 256, 0, stream>>>(static_cast<const T*>(dz), static_cast<const T*>(y), mean, rstd, scale, shift, sums,
                                                                   act, static_cast<T*>(dy), total, C, 1.0f / (float)rows)));
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}
