// Hardware probe (diagnostic entry, used by tests/test_tc_probe_gpu.py): does a UMMA shared-memory descriptor whose start
// address is offset by whole 128-byte rows inside a 1024-byte swizzle atom read the rows TMA wrote there?
// The implicit-GEMM 3x3 convolution (conv_tc.cu) relies on it: one TMA box of W+2 pixels per image row serves the three
// horizontal taps by sliding the descriptor start by one pixel row (128 B) instead of re-loading shifted boxes.
//   D[128][64] = A[row_off .. row_off+128)[0..64) * B[64][64]^T       (mode 0: A K-major,  tile [136 rows][64 k])
//   D[128][64] = sum_k A[row_off + k][0..128) (x) B[k][0..64)          (mode 1: A, B MN-major, tiles [72 k rows][64 mn])
#include "common.cuh"
#include <cuda.h>

namespace sarssl {
__device__ __forceinline__ uint64_t probe_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) |
           ((uint64_t)(base_off & 7) << 49) | (2ull << 61);
}

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* D, int mode,
                                                  int row_off, int use_base_offset) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sA = smem;                    // mode 0: 136 x 128 B = 17408 B;  mode 1: two boxes of 72 x 128 B (9216 B each, padded to 10240)
    unsigned char* sB = smem + 20480;            // 64 x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 20480 + 8192);
    uint64_t* done = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); mbar_fence_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        if (mode == 0) {
            mbar_expect_tx(bar, 136 * 128 + 64 * 128);
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(sA)), "l"(&tmA), "r"(0), "r"(0), "r"(smem_u32(bar)) : "memory");
        } else {
            mbar_expect_tx(bar, 2 * 72 * 128 + 64 * 128);
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(sA)), "l"(&tmA), "r"(0), "r"(0), "r"(smem_u32(bar)) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(sA + 10240)), "l"(&tmA), "r"(64), "r"(0), "r"(smem_u32(bar)) : "memory");
        }
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(sB)), "l"(&tmB), "r"(0), "r"(0), "r"(smem_u32(bar)) : "memory");
        mbar_wait(bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = smem_u32(sA) + row_off * 128, b0 = smem_u32(sB);
        const uint32_t bo = use_base_offset ? (uint32_t)row_off : 0u;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((mode ? 1u : 0u) << 15) | ((mode ? 1u : 0u) << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
        for (int k = 0; k < 4; ++k) {
            uint64_t da, db;
            if (mode == 0) { da = probe_desc(a0 + k * 32, 16, 1024, bo); db = probe_desc(b0 + k * 32, 16, 1024, 0); }
            else { da = probe_desc(a0 + k * 2048, 10240, 1024, bo); db = probe_desc(b0 + k * 2048, 8192, 1024, 0); }
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(k) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(done)) : "memory");
    }
    mbar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 64; c0 += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) D[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
}  // namespace sarssl

using namespace sarssl;

// A: mode 0 -> bf16 [136][64]; mode 1 -> bf16 [72][128].  B: bf16 [64][64] (mode 0: [n][k], mode 1: [k][n]).  D: f32 [128][64].
extern "C" int sarssl_probe_umma_row_offset(const void* A, const void* B, float* D, int mode, int row_off, int use_base_offset, cudaStream_t stream) {
    SARSSL_CHECK_ARG(A && B && D && (mode == 0 || mode == 1) && row_off >= 0 && row_off < 8, "probe: bad arguments");
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        set_last_error("probe: cuTensorMapEncodeTiled unavailable");
        return SARSSL_ERR_UNSUPPORTED;
    }
    EncFn enc = reinterpret_cast<EncFn>(f);
    CUtensorMap ma, mb;
    cuuint32_t es[2] = {1, 1};
    cuuint64_t da[2], sa[1]; cuuint32_t ba[2];
    if (mode == 0) { da[0] = 64; da[1] = 136; sa[0] = 128; ba[0] = 64; ba[1] = 136; }
    else { da[0] = 128; da[1] = 72; sa[0] = 256; ba[0] = 64; ba[1] = 72; }
    cuuint64_t dbm[2] = {64, 64}, sb[1] = {128}; cuuint32_t bb[2] = {64, 64};
    if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(A), da, sa, ba, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(B), dbm, sb, bb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        set_last_error("probe: tensor map encode failed");
        return SARSSL_ERR_ARG;
    }
    SARSSL_CUDA(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    probe_kernel<<<1, 128, 32768, stream>>>(ma, mb, D, mode, row_off, use_base_offset);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}
