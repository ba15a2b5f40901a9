// Issue-rate check: packed fma.rn.f32x2 (FFMA2) against scalar FFMA on sm_100a.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_scalar(float* out, int iters) {
    float a[8], b = threadIdx.x * 1e-3f + 1.0f, c = 0.5f;
    for (int i = 0; i < 8; ++i) a[i] = i + threadIdx.x;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, int iters) {
    unsigned long long a[4], b, c;
    float2 bb = make_float2(threadIdx.x * 1e-3f + 1.0f, threadIdx.x * 1e-3f + 1.0f), cc = make_float2(0.5f, 0.5f);
    b = *reinterpret_cast<unsigned long long*>(&bb); c = *reinterpret_cast<unsigned long long*>(&cc);
    for (int i = 0; i < 4; ++i) { float2 t = make_float2(2 * i + threadIdx.x, 2 * i + 1 + threadIdx.x); a[i] = *reinterpret_cast<unsigned long long*>(&t); }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 4; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));
    float s = 0; for (int i = 0; i < 4; ++i) { float2 t = *reinterpret_cast<float2*>(&a[i]); s += t.x + t.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int v = 0; v < 2; ++v) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (v == 0) k_scalar<<<148 * 8, 256>>>(out, iters); else k_packed<<<148 * 8, 256>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double fma = 148.0 * 8 * 256 * 8.0 * iters;
        printf("%s: %.3f ms  %.1f TFLOP/s fp32 (%.2e FMA lanes)\n", v == 0 ? "scalar FFMA " : "packed FFMA2", ms, 2 * fma / ms / 1e9, fma);
    }
    return 0;
}
