// Fused multi-tensor Adam (torch.optim.Adam, betas (0.9, 0.999), eps 1e-8, weight_decay 0, as created at learner.py:83)
// over one flat fp32 parameter arena: p, g, m, v are views of four equally laid out flat buffers, so one launch updates all
// 17.5 M parameters; optionally scales the gradient (1/world after the all-reduce, 1/accum for micro-batching), refreshes
// the bf16 compute copy and clears the gradient for the next step in the same pass.
#include "common.cuh"

namespace sarssl {

__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            __nv_bfloat16* __restrict__ p_bf16, long long n, float step_size, float b1, float b2, float omb1, float omb2, float eps,
                            float bc2_sqrt, float grad_scale, int zero_grad, const float* __restrict__ hyper) {
    if (hyper) { step_size = hyper[0]; bc2_sqrt = hyper[1]; }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * grad_scale;
        const float mi = b1 * m[i] + omb1 * gi;
        const float vi = b2 * v[i] + omb2 * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        const float pi = p[i] - step_size * (mi / denom);
        p[i] = pi;
        if (p_bf16) p_bf16[i] = __float2bfloat16_rn(pi);
        if (zero_grad) g[i] = 0.f;
    }
}

}  // namespace sarssl

extern "C" int sarssl_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* param_bf16, long long n, int step, double lr,
                                double beta1, double beta2, double eps, float grad_scale, int zero_grad, const float* hyper_dev, cudaStream_t stream) {
    SARSSL_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "adam_step: bad arguments");
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    long long g = (n + 1023) / 1024;
    const long long cap = (long long)sarssl::sm_count() * 16;
    if (g > cap) g = cap;
    sarssl::adam_kernel<<<(unsigned)g, 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, static_cast<__nv_bfloat16*>(param_bf16), n, (float)(lr / bc1),
                                                        (float)beta1, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps,
                                                        (float)sqrt(bc2), grad_scale, zero_grad, hyper_dev);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_adam_hyper_host(float* hyper2_host, int step, double lr, double beta1, double beta2) {
    SARSSL_CHECK_ARG(hyper2_host && step >= 1, "adam_hyper_host: bad arguments");
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    hyper2_host[0] = (float)(lr / bc1);
    hyper2_host[1] = (float)sqrt(bc2);
    return SARSSL_OK;
}
