// Element-wise / row-wise kernels of the Conformer blocks (memory-bound): Swish/ReLU/GLU gradients, relative-position
// score assembly + softmax (+ its backward and the inverse shift), head-bias adds, strided adds, dtype casts and a
// small N-d permute used to pack weights for the GEMM-shaped kernels.
//   Swish / GLU               conformer/activation.py:19-42
//   relative shift + softmax  conformer/attention.py:87-97,105-113 (scale 1/sqrt(d_model), attention.py:57,91)
#include "common.cuh"
#include "rng.cuh"
#include "vec.cuh"

namespace sarssl {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// All element-wise kernels below move 8 elements (16 bytes of bf16) per thread per iteration; n / D / cols are multiples of 8.

// du = ds * dropmask(idx)/(1-p) * swish'(u)      (FFN: s = Dropout(Swish(u)), feed_forward.py:49-51)
template <typename T>
__global__ void __launch_bounds__(256) swish_bwd_kernel(const T* __restrict__ ds, const T* __restrict__ u, T* __restrict__ du, long long n, float drop_p,
                                                      unsigned long long seed, const unsigned long long* __restrict__ seed_dev) {
    if (seed_dev) seed += *seed_dev;
    const float ks = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
    const uint32_t thr = drop_threshold(drop_p);
    for (long long g = (long long)blockIdx.x * 256 + threadIdx.x; g * 8 < n; g += (long long)gridDim.x * 256) {
        float d[8], x[8];
        Vec8<T>::load(ds + g * 8, d);
        Vec8<T>::load(u + g * 8, x);
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            if (drop_p > 0.f) {
                const uint32_t kp = keep_pair(seed, (unsigned long long)(g * 4 + (j >> 1)), thr);
                d[j] = (kp & 1u) ? d[j] * ks : 0.f;
                d[j + 1] = (kp & 2u) ? d[j + 1] * ks : 0.f;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float sg = sigmoidf_(x[j]); d[j] *= sg * (1.0f + x[j] * (1.0f - sg)); }
        Vec8<T>::store(du + g * 8, d);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) relu_bwd_kernel(const T* __restrict__ dz, const T* __restrict__ z, T* __restrict__ dy, long long n) {
    for (long long g = (long long)blockIdx.x * 256 + threadIdx.x; g * 8 < n; g += (long long)gridDim.x * 256) {
        float d[8], x[8];
        Vec8<T>::load(dz + g * 8, d);
        Vec8<T>::load(z + g * 8, x);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = x[j] > 0.f ? d[j] : 0.f;
        Vec8<T>::store(dy + g * 8, d);
    }
}

// a[m][d] = g[m][d] * sigmoid(g[m][D + d])         (GLU over channels, convolution.py:139)
template <typename T>
__global__ void __launch_bounds__(256) glu_fwd_kernel(const T* __restrict__ g, T* __restrict__ a, long long rows, int D) {
    const long long n = rows * D;
    for (long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 8; i < n; i += (long long)gridDim.x * 256 * 8) {
        const long long m = i / D; const int d = (int)(i - m * D);
        float x[8], gt[8];
        Vec8<T>::load(g + m * 2 * D + d, x);
        Vec8<T>::load(g + m * 2 * D + D + d, gt);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] *= sigmoidf_(gt[j]);
        Vec8<T>::store(a + i, x);
    }
}
template <typename T>
__global__ void __launch_bounds__(256) glu_bwd_kernel(const T* __restrict__ da, const T* __restrict__ g, T* __restrict__ dg, long long rows, int D) {
    const long long n = rows * D;
    for (long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 8; i < n; i += (long long)gridDim.x * 256 * 8) {
        const long long m = i / D; const int d = (int)(i - m * D);
        float x[8], gt[8], dd[8], o1[8], o2[8];
        Vec8<T>::load(g + m * 2 * D + d, x);
        Vec8<T>::load(g + m * 2 * D + D + d, gt);
        Vec8<T>::load(da + i, dd);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float sg = sigmoidf_(gt[j]); o1[j] = dd[j] * sg; o2[j] = dd[j] * x[j] * sg * (1.0f - sg); }
        Vec8<T>::store(dg + m * 2 * D + d, o1);
        Vec8<T>::store(dg + m * 2 * D + D + d, o2);
    }
}

// qu = q + u_bias, qv = q + v_bias; q = first D columns of qkv rows (ld)        attention.py:87-88
template <typename T>
__global__ void __launch_bounds__(256) add_head_bias_kernel(const T* __restrict__ q, long long ld, const float* __restrict__ u, const float* __restrict__ v,
                                                          T* __restrict__ qu, T* __restrict__ qv, long long rows, int D) {
    const long long n = rows * D;
    for (long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 8; i < n; i += (long long)gridDim.x * 256 * 8) {
        const long long m = i / D; const int d = (int)(i - m * D);
        float x[8], ub[8], vb[8], o[8];
        Vec8<T>::load(q + m * ld + d, x);
        load8f(u + d, ub);
        load8f(v + d, vb);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = x[j] + ub[j];
        Vec8<T>::store(qu + i, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = x[j] + vb[j];
        Vec8<T>::store(qv + i, o);
    }
}

// out[m][c] (ldo) = a[m][c] (lda) + b[m][c] (ldb)
template <typename T>
__global__ void __launch_bounds__(256) add2_kernel(const T* __restrict__ a, long long lda, const T* __restrict__ b, long long ldb, T* __restrict__ out,
                                                 long long ldo, long long rows, int cols) {
    const long long n = rows * cols;
    for (long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 8; i < n; i += (long long)gridDim.x * 256 * 8) {
        const long long m = i / cols; const int c = (int)(i - m * cols);
        float x[8], y[8];
        Vec8<T>::load(a + m * lda + c, x);
        Vec8<T>::load(b + m * ldb + c, y);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] += y[j];
        Vec8<T>::store(out + m * ldo + c, x);
    }
}

// ---- relative-position scores: prob[b][h][i][:] = softmax_j( (content[b][h][i][j] + shift(pos)[i][j]) * scale )
// content [B][H][T][T], pos [H][B][T][T];  shift: j <= i -> pos[i][T-1-i+j];  j == i+1 -> 0;  j > i+1 -> pos[i+1][j-i-2]
// one warp per score row
template <typename T>
__global__ void __launch_bounds__(256) attn_softmax_fwd_kernel(const T* __restrict__ content, const T* __restrict__ pos, T* __restrict__ prob,
                                                             T* __restrict__ attn, int B, int H, int Tn, float scale, float drop_p,
                                                             unsigned long long seed, const unsigned long long* __restrict__ seed_dev) {
    if (seed_dev) seed += *seed_dev;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= (long long)B * H * Tn) return;
    const int i = (int)(row % Tn);
    const long long bh = row / Tn;
    const int h = (int)(bh % H), b = (int)(bh / H);
    const T* crow = content + row * Tn;
    const T* pbase = pos + ((long long)h * B + b) * Tn * Tn;
    float mx = -INFINITY;
    for (int j = lane; j < Tn; j += 32) {
        float ps = 0.f;
        if (j <= i) ps = to_f32(pbase[(long long)i * Tn + (Tn - 1 - i + j)]);
        else if (j > i + 1) ps = to_f32(pbase[(long long)(i + 1) * Tn + (j - i - 2)]);
        mx = fmaxf(mx, (to_f32(crow[j]) + ps) * scale);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < Tn; j += 32) {
        float ps = 0.f;
        if (j <= i) ps = to_f32(pbase[(long long)i * Tn + (Tn - 1 - i + j)]);
        else if (j > i + 1) ps = to_f32(pbase[(long long)(i + 1) * Tn + (j - i - 2)]);
        sum += __expf((to_f32(crow[j]) + ps) * scale - mx);
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    T* orow = prob + row * Tn;
    const float ks = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
    for (int j = lane; j < Tn; j += 32) {
        float ps = 0.f;
        if (j <= i) ps = to_f32(pbase[(long long)i * Tn + (Tn - 1 - i + j)]);
        else if (j > i + 1) ps = to_f32(pbase[(long long)(i + 1) * Tn + (j - i - 2)]);
        const float pv = __expf((to_f32(crow[j]) + ps) * scale - mx) * inv;
        orow[j] = from_f32<T>(pv);
        if (attn != nullptr)        // Dropout(attn) (attention.py:98) materialised for the tensor-core context GEMM
            attn[row * Tn + j] = from_f32<T>(keep_mask(seed, (unsigned long long)(row * Tn + j), drop_p) ? pv * ks : 0.f);
    }
}

// dscore = scale * P * (dP - sum_j dP*P),  dP = dattn * dropmask/(1-p);  in place over dattn
template <typename T>
__global__ void __launch_bounds__(256) attn_softmax_bwd_kernel(T* __restrict__ dattn, const T* __restrict__ prob, long long rows, int Tn,
                                                             float scale, float drop_p, unsigned long long seed, const unsigned long long* __restrict__ seed_dev) {
    if (seed_dev) seed += *seed_dev;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    T* drow = dattn + row * Tn;
    const T* prow = prob + row * Tn;
    const float ks = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
    float dot = 0.f;
    for (int j = lane; j < Tn; j += 32) {
        float d = to_f32(drow[j]);
        if (drop_p > 0.f) d = keep_mask(seed, (unsigned long long)(row * Tn + j), drop_p) ? d * ks : 0.f;
        dot += d * to_f32(prow[j]);
    }
    dot = warp_sum(dot);
    for (int j = lane; j < Tn; j += 32) {
        float d = to_f32(drow[j]);
        if (drop_p > 0.f) d = keep_mask(seed, (unsigned long long)(row * Tn + j), drop_p) ? d * ks : 0.f;
        drow[j] = from_f32<T>(scale * to_f32(prow[j]) * (d - dot));
    }
}

// inverse of the shift: dpos[h][b][r][k] = k >= T-1-r ? dscore[b][h][r][k-(T-1-r)] : (r >= 1 ? dscore[b][h][r-1][k+r+1] : 0)
template <typename T>
__global__ void attn_unshift_kernel(const T* __restrict__ dscore, T* __restrict__ dpos, int B, int H, int Tn) {
    const long long n = (long long)B * H * Tn * Tn;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(idx % Tn);
        long long r_ = idx / Tn;
        const int r = (int)(r_ % Tn); r_ /= Tn;
        const int b = (int)(r_ % B), h = (int)(r_ / B);
        const T* base = dscore + ((long long)b * H + h) * Tn * Tn;
        T v = from_f32<T>(0.f);
        if (k >= Tn - 1 - r) v = base[(long long)r * Tn + (k - (Tn - 1 - r))];
        else if (r >= 1) v = base[(long long)(r - 1) * Tn + (k + r + 1)];
        dpos[idx] = v;
    }
}

// ---- vector versions of the three score kernels for T % 8 == 0, T <= 1024 (one warp per row; lane owns the 8 consecutive columns
// of chunks lane, lane + 32, ... (NCH = ceil(T / 256) chunks), 16-byte global accesses).  The relative shift is a sliding window:
// shift(pos)[i][j] = S[T-1-i+j] with S = [pos row i | 0 | pos row i+1], and its inverse dpos[r][k] = S'[r+1+k] with
// S' = [dscore row r-1 | dscore row r]; the rows are staged in (dynamic) shared memory as fp32 with one pad word every 8 (lane stride
// 9 words: conflict-free for both the chunked stores and the windowed loads).
__device__ __forceinline__ int pad8(int i) { return i + (i >> 3); }
__device__ __forceinline__ float ex2_approx(float x) {          // 2^x, flush-to-zero (no denormal fix-up code around the MUFU)
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
constexpr int kScoreMaxT = 1024;
static inline int score_row_floats(int Tn) { return Tn + Tn / 8 + 8; }

template <typename T, int NCH>
__global__ void __launch_bounds__(256) attn_softmax_fwd_vec_kernel(const T* __restrict__ content, const T* __restrict__ pos, T* __restrict__ prob,
                                                                 T* __restrict__ attn, int B, int H, int Tn, float scale, float drop_p,
                                                                 unsigned long long seed, const unsigned long long* __restrict__ seed_dev) {
    if (seed_dev) seed += *seed_dev;
    extern __shared__ float score_smem[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rowf = Tn + Tn / 8 + 8;
    float* win0 = score_smem + (size_t)w * rowf;                  // this warp's staged (already shifted) positional row
    const long long row = (long long)blockIdx.x * 8 + w;
    if (row >= (long long)B * H * Tn) return;
    const int i = (int)(row % Tn);
    const long long bh = row / Tn;
    const int h = (int)(bh % H), b = (int)(bh / H);
    const T* pbase = pos + ((long long)h * B + b) * Tn * Tn;
    // The shift is applied while STAGING: source column s of pos row i lands at score column j = s - (T-1-i) (the j <= i part), source column s
    // of row i+1 at j = s + i + 2 (the j > i+1 part), column i+1 is zero.  The staged row is then read back by its owner lane at compile-time
    // offsets (9 c + e: conflict-free), with no per-element index arithmetic or selects.
    float v[NCH][8];
    const int sh0 = Tn - 1 - i, sh1 = i + 2;
    if (lane == 0 && i + 1 < Tn) win0[pad8(i + 1)] = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int c = lane + 32 * k;
        if (c * 8 < Tn) {
            float t[8];
            Vec8<T>::load(pbase + (long long)i * Tn + c * 8, t);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int j = c * 8 + e - sh0;
                if (j >= 0) win0[pad8(j)] = t[e];
            }
            if (i + 1 < Tn) {
                Vec8<T>::load(pbase + (long long)(i + 1) * Tn + c * 8, t);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int j = c * 8 + e + sh1;
                    if (j < Tn) win0[pad8(j)] = t[e];
                }
            }
            Vec8<T>::load(content + row * Tn + c * 8, v[k]);
        }
    }
    __syncwarp();
    const float scale2 = scale * 1.4426950408889634f;              // scores in units of log2: softmax through ex2 without a multiply per element
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int c = lane + 32 * k;
        if (c * 8 < Tn) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                v[k][e] = (v[k][e] + win0[c * 9 + e]) * scale2;
                mx = fmaxf(mx, v[k][e]);
            }
        }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        if ((lane + 32 * k) * 8 < Tn) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { v[k][e] = ex2_approx(v[k][e] - mx); sum += v[k][e]; }
        }
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    const float ks = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
    const uint32_t thr = drop_threshold(drop_p);
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int c = lane + 32 * k;
        if (c * 8 >= Tn) continue;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[k][e] *= inv;
        const long long off = row * Tn + c * 8;
        Vec8<T>::store(prob + off, v[k]);
        if (attn != nullptr) {            // Dropout(attn) (attention.py:98) materialised for the tensor-core context GEMM
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
                const uint32_t kp = keep_pair(seed, (unsigned long long)(off + e) >> 1, thr);
                v[k][e] = (kp & 1u) ? v[k][e] * ks : 0.f;
                v[k][e + 1] = (kp & 2u) ? v[k][e + 1] * ks : 0.f;
            }
            Vec8<T>::store(attn + off, v[k]);
        }
    }
}

template <typename T, int NCH>
__global__ void __launch_bounds__(256) attn_softmax_bwd_vec_kernel(T* __restrict__ dattn, const T* __restrict__ prob, long long rows, int Tn,
                                                                 float scale, float drop_p, unsigned long long seed, const unsigned long long* __restrict__ seed_dev) {
    if (seed_dev) seed += *seed_dev;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float d[NCH][8], pr[NCH][8];
    float dot = 0.f;
    const float ks = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
    const uint32_t thr = drop_threshold(drop_p);
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int c = lane + 32 * k;
        if (c * 8 >= Tn) continue;
        const long long off = row * Tn + c * 8;
        Vec8<T>::load(dattn + off, d[k]);
        Vec8<T>::load(prob + off, pr[k]);
    }
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int c = lane + 32 * k;
        if (c * 8 >= Tn) continue;
        const long long off = row * Tn + c * 8;
        if (drop_p > 0.f) {
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
                const uint32_t kp = keep_pair(seed, (unsigned long long)(off + e) >> 1, thr);
                d[k][e] = (kp & 1u) ? d[k][e] * ks : 0.f;
                d[k][e + 1] = (kp & 2u) ? d[k][e + 1] * ks : 0.f;
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) dot = fmaf(d[k][e], pr[k][e], dot);
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int c = lane + 32 * k;
        if (c * 8 >= Tn) continue;
#pragma unroll
        for (int e = 0; e < 8; ++e) d[k][e] = scale * pr[k][e] * (d[k][e] - dot);
        Vec8<T>::store(dattn + row * Tn + c * 8, d[k]);
    }
}

template <typename T, int NCH>
__global__ void __launch_bounds__(256) attn_unshift_vec_kernel(const T* __restrict__ dscore, T* __restrict__ dpos, int B, int H, int Tn) {
    extern __shared__ float score_smem[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* win = score_smem + (size_t)w * 2 * (Tn + Tn / 8 + 8);
    const long long orow = (long long)blockIdx.x * 8 + w;              // output row (h, b, r)
    if (orow >= (long long)B * H * Tn) return;
    const int r = (int)(orow % Tn);
    const long long hb = orow / Tn;
    const int b = (int)(hb % B), h = (int)(hb / B);
    const T* base = dscore + ((long long)b * H + h) * Tn * Tn;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int c = lane + 32 * k;
        if (c * 8 >= Tn) continue;
        float t[8];
        if (r >= 1) Vec8<T>::load(base + (long long)(r - 1) * Tn + c * 8, t);
        else {
#pragma unroll
            for (int e = 0; e < 8; ++e) t[e] = 0.f;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) win[pad8(c * 8 + e)] = t[e];
        Vec8<T>::load(base + (long long)r * Tn + c * 8, t);
#pragma unroll
        for (int e = 0; e < 8; ++e) win[pad8(Tn + c * 8 + e)] = t[e];
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int c = lane + 32 * k;
        if (c * 8 >= Tn) continue;
        float t[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] = win[pad8(r + 1 + c * 8 + e)];
        Vec8<T>::store(dpos + orow * Tn + c * 8, t);
    }
}

// pooled[b][d] = mean_t x[b][t][d] (fp32 out);  dx[b][t][d] = dpooled[b][d] / T        (torch.mean(embed, dim=1), model.py:705)
template <typename T>
__global__ void __launch_bounds__(256) mean_pool_fwd_kernel(const T* __restrict__ x, long long ldx, float* __restrict__ pooled, int Tn, int D) {
    const int b = blockIdx.y, d = blockIdx.x * 256 + threadIdx.x;
    if (d >= D) return;
    float s = 0.f;
    for (int t = 0; t < Tn; ++t) s += to_f32(x[((long long)b * Tn + t) * ldx + d]);
    pooled[(long long)b * D + d] = s / Tn;
}
template <typename T>
__global__ void __launch_bounds__(256) mean_pool_bwd_kernel(const float* __restrict__ dpooled, T* __restrict__ dx, long long ldx, int Tn, int D) {
    const int b = blockIdx.y, d = blockIdx.x * 256 + threadIdx.x;
    if (d >= D) return;
    const T v = from_f32<T>(dpooled[(long long)b * D + d] / Tn);
    for (int t = blockIdx.z; t < Tn; t += gridDim.z) dx[((long long)b * Tn + t) * ldx + d] = v;
}

template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ src, TD* __restrict__ dst, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = from_f32<TD>(to_f32(src[i]));
}

// dst (contiguous over dims d0..d3) [+]= src[i0*s0 + i1*s1 + i2*s2 + i3*s3]
template <typename TS, typename TD>
__global__ void permute4_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int d0, int d1, int d2, int d3, long long s0, long long s1,
                                long long s2, long long s3, int accumulate) {
    const long long n = (long long)d0 * d1 * d2 * d3;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int i3 = (int)(i % d3); long long r = i / d3;
        const int i2 = (int)(r % d2); r /= d2;
        const int i1 = (int)(r % d1); const int i0 = (int)(r / d1);
        const float v = to_f32(src[i0 * s0 + i1 * s1 + i2 * s2 + i3 * s3]);
        dst[i] = from_f32<TD>(accumulate ? to_f32(dst[i]) + v : v);
    }
}

__global__ void fill_f32_kernel(float* p, float v, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

static int ew_grid(long long n) {
    long long g = (n + 2047) / 2048;
    const long long cap = (long long)sm_count() * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace sarssl

using namespace sarssl;

#define DISPATCH_T(dtype, ...)                                                         \
    do {                                                                               \
        if ((dtype) == SARSSL_F32) { using T = float; __VA_ARGS__; }                   \
        else if ((dtype) == SARSSL_BF16) { using T = __nv_bfloat16; __VA_ARGS__; }     \
        else { set_last_error("bad dtype %d", (int)(dtype)); return SARSSL_ERR_ARG; }  \
    } while (0)

extern "C" int sarssl_swish_bwd(const void* ds, const void* u, void* du, long long n, float drop_p, unsigned long long seed,
                                const unsigned long long* seed_dev, int dtype,
                                cudaStream_t stream) {
    SARSSL_CHECK_ARG(ds && u && du && n > 0 && n % 8 == 0, "swish_bwd: bad arguments (n must be a multiple of 8)");
    DISPATCH_T(dtype, (swish_bwd_kernel<T><<<ew_grid(n), 256, 0, stream>>>(static_cast<const T*>(ds), static_cast<const T*>(u), static_cast<T*>(du), n,
                                                                          drop_p, seed, seed_dev)));
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_relu_bwd(const void* dz, const void* z, void* dy, long long n, int dtype, cudaStream_t stream) {
    SARSSL_CHECK_ARG(dz && z && dy && n > 0 && n % 8 == 0, "relu_bwd: bad arguments (n must be a multiple of 8)");
    DISPATCH_T(dtype, (relu_bwd_kernel<T><<<ew_grid(n), 256, 0, stream>>>(static_cast<const T*>(dz), static_cast<const T*>(z), static_cast<T*>(dy), n)));
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_glu_fwd(const void* g, void* a, long long rows, int D, int dtype, cudaStream_t stream) {
    SARSSL_CHECK_ARG(g && a && rows > 0 && D > 0 && D % 8 == 0, "glu_fwd: bad arguments (D must be a multiple of 8)");
    DISPATCH_T(dtype, (glu_fwd_kernel<T><<<ew_grid(rows * D), 256, 0, stream>>>(static_cast<const T*>(g), static_cast<T*>(a), rows, D)));
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_glu_bwd(const void* da, const void* g, void* dg, long long rows, int D, int dtype, cudaStream_t stream) {
    SARSSL_CHECK_ARG(da && g && dg && rows > 0 && D > 0 && D % 8 == 0, "glu_bwd: bad arguments (D must be a multiple of 8)");
    DISPATCH_T(dtype, (glu_bwd_kernel<T><<<ew_grid(rows * D), 256, 0, stream>>>(static_cast<const T*>(da), static_cast<const T*>(g), static_cast<T*>(dg),
                                                                               rows, D)));
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_add_head_bias(const void* q, long long ld, const float* u_bias, const float* v_bias, void* qu, void* qv, long long rows, int D,
                                    int dtype, cudaStream_t stream) {
    SARSSL_CHECK_ARG(q && u_bias && v_bias && qu && qv && rows > 0 && D > 0 && D % 8 == 0 && ld % 8 == 0, "add_head_bias: bad arguments (D, ld multiples of 8)");
    DISPATCH_T(dtype, (add_head_bias_kernel<T><<<ew_grid(rows * D), 256, 0, stream>>>(static_cast<const T*>(q), ld, u_bias, v_bias, static_cast<T*>(qu),
                                                                                     static_cast<T*>(qv), rows, D)));
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_add2(const void* a, long long lda, const void* b, long long ldb, void* out, long long ldo, long long rows, int cols, int dtype,
                           cudaStream_t stream) {
    SARSSL_CHECK_ARG(a && b && out && rows > 0 && cols > 0 && cols % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldo % 8 == 0, "add2: bad arguments (multiples of 8)");
    DISPATCH_T(dtype, (add2_kernel<T><<<ew_grid(rows * cols), 256, 0, stream>>>(static_cast<const T*>(a), lda, static_cast<const T*>(b), ldb,
                                                                               static_cast<T*>(out), ldo, rows, cols)));
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_attn_softmax_fwd(const void* content, const void* pos, void* prob, void* attn_dropped, int B, int H, int T_, float scale,
                                       float drop_p, unsigned long long seed, const unsigned long long* seed_dev, int dtype, cudaStream_t stream) {
    SARSSL_CHECK_ARG(content && pos && prob && B > 0 && H > 0 && T_ > 0, "attn_softmax_fwd: bad arguments");
    const long long rows = (long long)B * H * T_;
    if (T_ % 8 == 0 && T_ <= kScoreMaxT) {
        const size_t smem = (size_t)8 * score_row_floats(T_) * sizeof(float);
#define SCORE_FWD(NCH)                                                                                                                           \
        do {                                                                                                                                     \
            static bool set_f = false, set_h = false;                                                                                            \
            bool& set = dtype == SARSSL_F32 ? set_f : set_h;                                                                                     \
            if (!set) {                                                                                                                          \
                if (dtype == SARSSL_F32) SARSSL_CUDA(cudaFuncSetAttribute(attn_softmax_fwd_vec_kernel<float, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024)); \
                else SARSSL_CUDA(cudaFuncSetAttribute(attn_softmax_fwd_vec_kernel<__nv_bfloat16, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));      \
                set = true;                                                                                                                      \
            }                                                                                                                                    \
            DISPATCH_T(dtype, (attn_softmax_fwd_vec_kernel<T, NCH><<<(unsigned)((rows + 7) / 8), 256, smem, stream>>>(                           \
                                   static_cast<const T*>(content), static_cast<const T*>(pos), static_cast<T*>(prob), static_cast<T*>(attn_dropped), B, H, T_, \
                                   scale, drop_p, seed, seed_dev)));                                                                                       \
        } while (0)
        if (T_ <= 256) SCORE_FWD(1);
        else if (T_ <= 512) SCORE_FWD(2);
        else SCORE_FWD(4);
#undef SCORE_FWD
    } else
        DISPATCH_T(dtype, (attn_softmax_fwd_kernel<T><<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(static_cast<const T*>(content), static_cast<const T*>(pos),
                                                                                                    static_cast<T*>(prob), static_cast<T*>(attn_dropped), B, H, T_, scale,
                                                                                                    drop_p, seed, seed_dev)));
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_attn_softmax_bwd(void* dattn_inout, const void* prob, void* dpos, int B, int H, int T_, float scale, float drop_p,
                                       unsigned long long seed, const unsigned long long* seed_dev, int dtype, cudaStream_t stream) {
    SARSSL_CHECK_ARG(dattn_inout && prob && dpos && B > 0 && H > 0 && T_ > 0, "attn_softmax_bwd: bad arguments");
    const long long rows = (long long)B * H * T_;
    const bool vec = T_ % 8 == 0 && T_ <= kScoreMaxT;
    const size_t smem = (size_t)8 * 2 * score_row_floats(T_) * sizeof(float);
#define SCORE_BWD(NCH)                                                                                                                           \
    do {                                                                                                                                         \
        static bool set_f = false, set_h = false;                                                                                                \
        bool& set = dtype == SARSSL_F32 ? set_f : set_h;                                                                                         \
        if (!set) {                                                                                                                              \
            if (dtype == SARSSL_F32) SARSSL_CUDA(cudaFuncSetAttribute(attn_unshift_vec_kernel<float, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024)); \
            else SARSSL_CUDA(cudaFuncSetAttribute(attn_unshift_vec_kernel<__nv_bfloat16, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));          \
            set = true;                                                                                                                          \
        }                                                                                                                                        \
        DISPATCH_T(dtype, (attn_softmax_bwd_vec_kernel<T, NCH><<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(static_cast<T*>(dattn_inout),     \
                               static_cast<const T*>(prob), rows, T_, scale, drop_p, seed, seed_dev)));                                                    \
        SARSSL_LAUNCH_CHECK();                                                                                                                   \
        DISPATCH_T(dtype, (attn_unshift_vec_kernel<T, NCH><<<(unsigned)((rows + 7) / 8), 256, smem, stream>>>(static_cast<const T*>(dattn_inout), \
                               static_cast<T*>(dpos), B, H, T_)));                                                                               \
    } while (0)
    if (vec) {
        if (T_ <= 256) SCORE_BWD(1);
        else if (T_ <= 512) SCORE_BWD(2);
        else SCORE_BWD(4);
    } else {
        DISPATCH_T(dtype, (attn_softmax_bwd_kernel<T><<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(static_cast<T*>(dattn_inout), static_cast<const T*>(prob),
                                                                                                    rows, T_, scale, drop_p, seed, seed_dev)));
        SARSSL_LAUNCH_CHECK();
        DISPATCH_T(dtype, (attn_unshift_kernel<T><<<ew_grid(rows * T_), 256, 0, stream>>>(static_cast<const T*>(dattn_inout), static_cast<T*>(dpos), B, H, T_)));
    }
#undef SCORE_BWD
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_mean_pool_fwd(const void* x, long long ldx, float* pooled, int B, int T_, int D, int dtype, cudaStream_t stream) {
    SARSSL_CHECK_ARG(x && pooled && B > 0 && T_ > 0 && D > 0 && B <= 65535, "mean_pool_fwd: bad arguments");
    DISPATCH_T(dtype, (mean_pool_fwd_kernel<T><<<dim3((D + 255) / 256, B), 256, 0, stream>>>(static_cast<const T*>(x), ldx, pooled, T_, D)));
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_mean_pool_bwd(const float* dpooled, void* dx, long long ldx, int B, int T_, int D, int dtype, cudaStream_t stream) {
    SARSSL_CHECK_ARG(dpooled && dx && B > 0 && T_ > 0 && D > 0 && B <= 65535, "mean_pool_bwd: bad arguments");
    const int gz = T_ < 16 ? T_ : 16;
    DISPATCH_T(dtype, (mean_pool_bwd_kernel<T><<<dim3((D + 255) / 256, B, gz), 256, 0, stream>>>(dpooled, static_cast<T*>(dx), ldx, T_, D)));
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_cast(const void* src, int src_dtype, void* dst, int dst_dtype, long long n, cudaStream_t stream) {
    SARSSL_CHECK_ARG(src && dst && n > 0, "cast: bad arguments");
    const int g = ew_grid(n);
    if (src_dtype == SARSSL_F32 && dst_dtype == SARSSL_BF16) cast_kernel<float, __nv_bfloat16><<<g, 256, 0, stream>>>((const float*)src, (__nv_bfloat16*)dst, n);
    else if (src_dtype == SARSSL_BF16 && dst_dtype == SARSSL_F32) cast_kernel<__nv_bfloat16, float><<<g, 256, 0, stream>>>((const __nv_bfloat16*)src, (float*)dst, n);
    else if (src_dtype == SARSSL_F32 && dst_dtype == SARSSL_F32) cast_kernel<float, float><<<g, 256, 0, stream>>>((const float*)src, (float*)dst, n);
    else cast_kernel<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, stream>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, n);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_permute4(const void* src, int src_dtype, void* dst, int dst_dtype, const int* dims4_host, const long long* src_strides4_host,
                               int accumulate, cudaStream_t stream) {
    SARSSL_CHECK_ARG(src && dst && dims4_host && src_strides4_host, "permute4: null pointer");
    const int* d = dims4_host; const long long* s = src_strides4_host;
    const long long n = (long long)d[0] * d[1] * d[2] * d[3];
    SARSSL_CHECK_ARG(n > 0, "permute4: empty");
    const int g = ew_grid(n);
    if (src_dtype == SARSSL_F32 && dst_dtype == SARSSL_BF16)
        permute4_kernel<float, __nv_bfloat16><<<g, 256, 0, stream>>>((const float*)src, (__nv_bfloat16*)dst, d[0], d[1], d[2], d[3], s[0], s[1], s[2], s[3], accumulate);
    else if (src_dtype == SARSSL_BF16 && dst_dtype == SARSSL_F32)
        permute4_kernel<__nv_bfloat16, float><<<g, 256, 0, stream>>>((const __nv_bfloat16*)src, (float*)dst, d[0], d[1], d[2], d[3], s[0], s[1], s[2], s[3], accumulate);
    else if (src_dtype == SARSSL_F32 && dst_dtype == SARSSL_F32)
        permute4_kernel<float, float><<<g, 256, 0, stream>>>((const float*)src, (float*)dst, d[0], d[1], d[2], d[3], s[0], s[1], s[2], s[3], accumulate);
    else
        permute4_kernel<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, stream>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, d[0], d[1], d[2], d[3], s[0], s[1], s[2], s[3], accumulate);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_fill_f32(float* p, float value, long long n, cudaStream_t stream) {
    SARSSL_CHECK_ARG(p && n > 0, "fill: bad arguments");
    fill_f32_kernel<<<ew_grid(n), 256, 0, stream>>>(p, value, n);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

namespace sarssl {
// dst = alpha * src * dropout_mask(offset)/(1-p): re-applies a forward dropout mask to the incoming gradient
template <typename T>
__global__ void __launch_bounds__(256) scale_dropout_kernel(const T* __restrict__ src, T* __restrict__ dst, long long n, float alpha, float drop_p,
                                                          unsigned long long seed, const unsigned long long* __restrict__ seed_dev) {
    if (seed_dev) seed += *seed_dev;
    const float ks = drop_p > 0.f ? alpha / (1.0f - drop_p) : alpha;
    const uint32_t thr = drop_threshold(drop_p);
    for (long long g = (long long)blockIdx.x * 256 + threadIdx.x; g * 8 < n; g += (long long)gridDim.x * 256) {
        float v[8];
        Vec8<T>::load(src + g * 8, v);
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            const uint32_t kp = drop_p > 0.f ? keep_pair(seed, (unsigned long long)(g * 4 + (j >> 1)), thr) : 3u;
            v[j] = (kp & 1u) ? v[j] * ks : 0.f;
            v[j + 1] = (kp & 2u) ? v[j + 1] * ks : 0.f;
        }
        Vec8<T>::store(dst + g * 8, v);
    }
}
}  // namespace sarssl

extern "C" int sarssl_scale_dropout(const void* src, void* dst, long long n, float alpha, float drop_p, unsigned long long seed,
                                    const unsigned long long* seed_dev, int dtype,
                                    cudaStream_t stream) {
    SARSSL_CHECK_ARG(src && dst && n > 0 && n % 8 == 0, "scale_dropout: bad arguments (n must be a multiple of 8)");
    DISPATCH_T(dtype, (sarssl::scale_dropout_kernel<T><<<sarssl::ew_grid(n), 256, 0, stream>>>(static_cast<const T*>(src), static_cast<T*>(dst), n, alpha,
                                                                                             drop_p, seed, seed_dev)));
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}
