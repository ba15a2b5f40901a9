// Strided-batched GEMM on CUDA cores with fp32 accumulation and a fused epilogue.
//
// Role in the design: (1) the exact-fp32 arithmetic mode the north star's 1e-4 forward-activation gate is checked in,
// (2) the small / oddly shaped contractions (attention score and context products, per-head strides) and
// (3) the numerical cross-check for the tcgen05 kernels in gemm_tc.cu, which carry the big bf16 GEMMs.
//
//   C[z][m][n] = resid[z][m][n] + beta * drop( act( alpha * sum_k A[z][m][k] * B[z][n][k] + bias[n] ) )
//
// A element (m, k) lives at A + z1*sAb1 + z2*sAb2 + m*sAm + k*sAk (same for B with n), so row-major, transposed and
// per-head interleaved operands need no copies.  Optional A-side dropout re-creates a forward dropout mask on the fly
// (counter-based RNG keyed by the element offset), which is how the backward pass applies masks it never stored.
#include "common.cuh"
#include "rng.cuh"

namespace sarssl {

constexpr int BM = 64, BN = 64, BK = 16, TPB = 256;

struct GemmP {
    const void* A; const void* B; void* C; void* pre; const void* resid; const float* bias;
    long long sAm, sAk, sAb1, sAb2, sBn, sBk, sBb1, sBb2, ldc, sCb1, sCb2, ldr;
    int M, N, K, nb2;
    float alpha, beta;
    int act, accumulate, has_resid;
    float drop_p; unsigned long long drop_seed;        // epilogue dropout (index = C element offset)
    float a_drop_p; unsigned long long a_drop_seed;    // A-operand dropout mask (index = A element offset)
    const unsigned long long* seed_dev;                // nullable: device word added to both seeds
};

template <typename T> __device__ __forceinline__ float ldf(const T* p) { return to_f32(*p); }

template <typename TA, typename TC, bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(TPB) gemm_simt_kernel(GemmP p) {
    if (p.seed_dev) { const unsigned long long so = *p.seed_dev; p.drop_seed += so; p.a_drop_seed += so; }
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int z = blockIdx.z, z1 = z / p.nb2, z2 = z - z1 * p.nb2;
    const TA* A = static_cast<const TA*>(p.A) + z1 * p.sAb1 + z2 * p.sAb2;
    const TA* B = static_cast<const TA*>(p.B) + z1 * p.sBb1 + z2 * p.sBb2;
    const long long a_base_off = z1 * p.sAb1 + z2 * p.sAb2;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const bool a_drop = p.a_drop_p > 0.f;
    const float a_keep_scale = a_drop ? 1.0f / (1.0f - p.a_drop_p) : 1.0f;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < p.K; k0 += BK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + i * TPB;
            {
                const int kk = A_KMAJOR ? (idx & (BK - 1)) : (idx >> 6);
                const int mm = A_KMAJOR ? (idx >> 4) : (idx & (BM - 1));
                const int m = m0 + mm, k = k0 + kk;
                float v = 0.f;
                if (m < p.M && k < p.K) {
                    const long long off = (long long)m * p.sAm + (long long)k * p.sAk;
                    v = ldf(A + off);
                    if (a_drop) v = keep_mask(p.a_drop_seed, (unsigned long long)(a_base_off + off), p.a_drop_p) ? v * a_keep_scale : 0.f;
                }
                As[kk][mm] = v;
            }
            {
                const int kk = B_KMAJOR ? (idx & (BK - 1)) : (idx >> 6);
                const int nn = B_KMAJOR ? (idx >> 4) : (idx & (BN - 1));
                const int n = n0 + nn, k = k0 + kk;
                float v = 0.f;
                if (n < p.N && k < p.K) v = ldf(B + (long long)n * p.sBn + (long long)k * p.sBk);
                Bs[kk][nn] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    // epilogue
    TC* C = static_cast<TC*>(p.C) + z1 * p.sCb1 + z2 * p.sCb2;
    TC* pre = p.pre ? static_cast<TC*>(p.pre) + z1 * p.sCb1 + z2 * p.sCb2 : nullptr;
    const TC* resid = p.has_resid ? static_cast<const TC*>(p.resid) : nullptr;
    const bool drop = p.drop_p > 0.f;
    const float keep_scale = drop ? 1.0f / (1.0f - p.drop_p) : 1.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            const long long off = (long long)m * p.ldc + n;
            float v = acc[i][j] * p.alpha + (p.bias ? p.bias[n] : 0.f);
            if (pre) pre[off] = from_f32<TC>(v);
            if (p.act == 1) v = fmaxf(v, 0.f);
            else if (p.act == 2) v = v / (1.0f + __expf(-v));
            if (drop) v = keep_mask(p.drop_seed, (unsigned long long)(z1 * p.sCb1 + z2 * p.sCb2 + off), p.drop_p) ? v * keep_scale : 0.f;
            if (resid) v = to_f32(resid[z1 * p.sCb1 + z2 * p.sCb2 + (long long)m * p.ldr + n]) + p.beta * v;
            else v *= p.beta;
            if (p.accumulate) v += to_f32(C[off]);
            C[off] = from_f32<TC>(v);
        }
    }
}

template <typename TA, typename TC>
static void launch(const GemmP& p, int nbatch, cudaStream_t s) {
    dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, nbatch);
    const bool ak = p.sAk == 1, bk = p.sBk == 1;
    if (ak && bk) gemm_simt_kernel<TA, TC, true, true><<<grid, TPB, 0, s>>>(p);
    else if (ak && !bk) gemm_simt_kernel<TA, TC, true, false><<<grid, TPB, 0, s>>>(p);
    else if (!ak && bk) gemm_simt_kernel<TA, TC, false, true><<<grid, TPB, 0, s>>>(p);
    else gemm_simt_kernel<TA, TC, false, false><<<grid, TPB, 0, s>>>(p);
}

}  // namespace sarssl

using namespace sarssl;

extern "C" int sarssl_gemm(const sarssl_gemm_args* a, cudaStream_t stream) {
    SARSSL_CHECK_ARG(a && a->A && a->B && a->C, "gemm: null pointer");
    SARSSL_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0 && a->nb1 > 0 && a->nb2 > 0, "gemm: bad dims M=%d N=%d K=%d nb=%dx%d", a->M, a->N, a->K,
                     a->nb1, a->nb2);
    SARSSL_CHECK_ARG((long long)a->nb1 * a->nb2 <= 65535, "gemm: batch %d x %d exceeds the grid limit", a->nb1, a->nb2);
    SARSSL_CHECK_ARG(a->ab_dtype == SARSSL_F32 || a->ab_dtype == SARSSL_BF16, "gemm: bad ab_dtype");
    SARSSL_CHECK_ARG(a->c_dtype == SARSSL_F32 || a->c_dtype == SARSSL_BF16, "gemm: bad c_dtype");
    SARSSL_CHECK_ARG(a->act >= 0 && a->act <= 2, "gemm: bad activation %d", a->act);
    GemmP p;
    p.A = a->A; p.B = a->B; p.C = a->C; p.pre = a->pre_out; p.resid = a->resid; p.bias = a->bias;
    p.sAm = a->sAm; p.sAk = a->sAk; p.sAb1 = a->sAb1; p.sAb2 = a->sAb2;
    p.sBn = a->sBn; p.sBk = a->sBk; p.sBb1 = a->sBb1; p.sBb2 = a->sBb2;
    p.ldc = a->ldc; p.sCb1 = a->sCb1; p.sCb2 = a->sCb2; p.ldr = a->ldr ? a->ldr : a->ldc;
    p.M = a->M; p.N = a->N; p.K = a->K; p.nb2 = a->nb2;
    p.alpha = a->alpha; p.beta = a->beta; p.act = a->act; p.accumulate = a->accumulate; p.has_resid = a->resid != nullptr;
    p.drop_p = a->drop_p; p.drop_seed = a->drop_seed; p.a_drop_p = a->a_drop_p; p.a_drop_seed = a->a_drop_seed; p.seed_dev = a->seed_dev;
    const int nbatch = a->nb1 * a->nb2;
    if (a->ab_dtype == SARSSL_F32 && a->c_dtype == SARSSL_F32) launch<float, float>(p, nbatch, stream);
    else if (a->ab_dtype == SARSSL_BF16 && a->c_dtype == SARSSL_BF16) launch<__nv_bfloat16, __nv_bfloat16>(p, nbatch, stream);
    else if (a->ab_dtype == SARSSL_BF16 && a->c_dtype == SARSSL_F32) launch<__nv_bfloat16, float>(p, nbatch, stream);
    else launch<float, __nv_bfloat16>(p, nbatch, stream);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}
