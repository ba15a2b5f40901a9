// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16 operands from 128B-swizzled shared memory, M = 128 per CTA, K = 16) as a function of N,
// for one CTA and for a CTA pair (cta_group::2), issued back to back the way the conv kernels issue them.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/umma_rate scripts/micro/umma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo) { return ((saddr >> 4) & 0x3FFF) | (((lbo >> 4) & 0x3FFF) << 16); }
__device__ __forceinline__ constexpr uint32_t desc_hi(uint32_t sbo) { return ((sbo >> 4) & 0x3FFF) | (1u << 14) | (2u << 29); }
template <int G>
__device__ __forceinline__ void umma(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
    if (G == 1)
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %2};\n\tsetp.ne.b32 p, %5, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(hi), "r"(b_lo), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %2};\n\tsetp.ne.b32 p, %5, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(hi), "r"(b_lo), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// pattern: every k-step issues op(N1) into columns [0, N1) and, if N2 > 0, op(N2) into columns [256, 256 + N2); the A slice rotates over `nslots` boxes
template <int G>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N1, int N2, int iters, int same_a, int d2off, long long* out) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    uint32_t rank = 0;
    if (G == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = threadIdx.x; i < 192 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3C003C00u, 0x3C003C00u, 0, 0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) {
        if (G == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (G == 2) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (warp == 1 && rank == 0 && (threadIdx.x & 31) == 0) {
        const uint32_t mfield = (G == 2 ? 256u : 128u) >> 4;
        const uint32_t id0 = (1u << 4) | (1u << 7) | (1u << 10) | (mfield << 24);
        const uint32_t id1 = id0 | ((uint32_t)(N1 >> 3) << 17), id2 = id0 | ((uint32_t)(N2 >> 3) << 17);
        const uint32_t hi = desc_hi(1024);
        const uint32_t a_lo = desc_lo(smem_u32(smem), 16), b_lo = desc_lo(smem_u32(smem + 104 * 1024), 16);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t a = a_lo + (same_a ? 0u : (uint32_t)(it % 6) * (17408u >> 4));
#pragma unroll
            for (int dw = 0; dw < 3; ++dw)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    umma<G>(tmem, a + dw * 8 + k * 2, b_lo + dw * 256 + k * 2, hi, id1, 1u);
                    if (N2 > 0) umma<G>(tmem + d2off, a + dw * 8 + k * 2, b_lo + dw * 256 + 2048 + k * 2, hi, id2, 1u);
                }
        }
        if (G == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
        while (!try_wait(&bar, 0)) {}
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (G == 2) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    if (warp == 1) {
        if (G == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

int main() {
    long long* out;
    cudaMalloc(&out, 148 * sizeof(long long));
    const int smem = 193 * 1024 + 1024;
    cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 400;
    struct Cfg { int g, n1, n2, same, d2; } cfgs[] = {
        {1, 128, 64, 0, 256}, {1, 128, 64, 0, 128}, {1, 128, 64, 0, 192}, {1, 128, 64, 0, 384}, {1, 128, 128, 0, 128}, {1, 128, 128, 0, 256}, {1, 64, 64, 0, 64}, {1, 64, 64, 0, 128},
        {1, 64, 64, 0, 256}, {1, 128, 0, 0, 0}, {1, 64, 0, 0, 0},
    };
    for (auto c : cfgs) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(out, 0, 148 * sizeof(long long));
            if (c.g == 1) rate_kernel<1><<<148, 128, smem>>>(c.n1, c.n2, iters, c.same, c.d2, out);
            else {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                cudaLaunchKernelEx(&cfg, rate_kernel<2>, c.n1, c.n2, iters, c.same, c.d2, out);
            }
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("cfg g=%d n1=%d n2=%d: %s\n", c.g, c.n1, c.n2, cudaGetErrorString(e)); return 1; }
        }
        long long h[148];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0; int cnt = 0; double sum = 0;
        for (int i = 0; i < 148; ++i) if (h[i] > 0) { if (h[i] > mx) mx = h[i]; sum += h[i]; ++cnt; }
        const double ksteps = iters * 12.0;
        const double cols = c.n1 + c.n2;
        printf("cta_group::%d  N=%3d%s  second accumulator at column %3d: %7.1f clk per k-step (avg over %d issuers, max %7.1f); ideal math %5.1f clk (%.0f%% of peak)\n", c.g, c.n1,
               c.n2 ? (c.n2 == 64 ? "+ 64" : c.n2 == 128 ? "+128" : "+256") : "    ", c.d2, sum / cnt / ksteps, cnt, mx / ksteps, cols / 2, 100.0 * (cols / 2) / (sum / cnt / ksteps));
    }
    return 0;
}
