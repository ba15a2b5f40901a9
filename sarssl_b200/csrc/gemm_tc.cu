// bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA) for sm_100a.
//
//   C[m][n] = resid[m][n] + beta * drop( act( alpha * sum_k A[m][k] * B[n][k] + bias[n] ) )      (same epilogue as gemm_simt.cu)
//
// Carries the dense contractions of the path: FFN / projection / pointwise-conv / decoder / patch-embedding GEMMs and their
// data gradients (K-major operands: activations [M][K] and weights [N][K], or pre-transposed weights for dgrad), and the
// weight gradients dW = dY^T X where the contraction runs over the token rows (MN-major operands, same kernel).
//
// One CTA computes a 128 x BN tile: warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + MMA issuer (one
// elected lane issues UMMA 128xBNx16 instructions), warps 2-5 = epilogue (each owns the 32 TMEM lanes its warp-id % 4
// selects).  A kStages-deep ring of 128B-swizzled shared-memory tiles is handed from TMA to MMA through full/empty
// mbarriers; tcgen05.commit releases a stage when the MMAs reading it retire and finally signals the epilogue.
// Two CTAs fit per SM (smem and TMEM), so one tile's epilogue overlaps the other's main loop.
#include "common.cuh"
#include "rng.cuh"
#include <cuda.h>

namespace sarssl {

constexpr int TBM = 128, TBK = 64, kStages = 3, kTcThreads = 192;

struct TcEpi {
    void* C; void* pre; const void* resid; const float* bias;
    long long ldc, ldr;
    int M, N, K;
    float alpha, beta;
    int act, accumulate, c_is_bf16;
    float drop_p; unsigned long long drop_seed;
};

// ---- PTX wrappers ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version = 1 [46,48), layout type [61,64) (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           (1ull << 46) | (2ull << 61);
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): D = f32 [4,6) = 1, A/B = bf16 [7,10) / [10,13) = 1, a_major bit 15, b_major bit 16
// (0 = K-major, 1 = MN-major), N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int BN>
struct TcSmem {
    static constexpr int kABytes = TBM * TBK * 2, kBBytes = BN * TBK * 2;
    static constexpr int kBytes = kStages * (kABytes + kBBytes) + 1024 /*align slack*/ + 256 /*barriers*/;
};

// A_MN / B_MN: operand is MN-major (rows of the global matrix run along the contraction dimension), used by weight gradients.
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kTcThreads) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcEpi p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kA = TcSmem<BN>::kABytes, kB = TcSmem<BN>::kBBytes;
    unsigned char* sA = smem;
    unsigned char* sB = smem + kStages * kA;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * (kA + kB));
    uint64_t* empty = full + kStages;
    uint64_t* accum_full = empty + kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * BN;
    const int num_kb = (p.K + TBK - 1) / TBK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(accum_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, BN);          // BN fp32 columns x 128 lanes
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], kA + kB);
                if (!A_MN) tma_load_2d(sA + s * kA, &tmA, kb * TBK, m0, &full[s]);
                else { tma_load_2d(sA + s * kA, &tmA, m0, kb * TBK, &full[s]); tma_load_2d(sA + s * kA + kA / 2, &tmA, m0 + 64, kb * TBK, &full[s]); }
                if (!B_MN) tma_load_2d(sB + s * kB, &tmB, kb * TBK, n0, &full[s]);
                else {
#pragma unroll
                    for (int j = 0; j < BN / 64; ++j) tma_load_2d(sB + s * kB + j * 8192, &tmB, n0 + 64 * j, kb * TBK, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(TBM, BN, A_MN, B_MN);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(sA + s * kA), b_addr = smem_u32(sB + s * kB);
#pragma unroll
                for (int k = 0; k < TBK / 16; ++k) {
                    // K-major: 8-row groups 1024 B apart, advance 32 B per 16-element k step inside the 128 B swizzle atom.
                    // MN-major: 64-element MN groups 8192 B apart (LBO), 8-k-row groups 1024 B apart (SBO), advance 2048 B per k step.
                    const uint64_t da = A_MN ? smem_desc(a_addr + k * 2048, 8192, 1024) : smem_desc(a_addr + k * 32, 16, 1024);
                    const uint64_t db = B_MN ? smem_desc(b_addr + k * 2048, 8192, 1024) : smem_desc(b_addr + k * 32, 16, 1024);
                    umma_bf16(tmem_base, da, db, idesc, (kb | k) != 0);
                }
                umma_commit(&empty[s]);                 // stage reusable once these MMAs have read it
            }
            umma_commit(accum_full);                    // accumulator complete
        }
    } else {
        const int q = warp & 3;                         // TMEM lane quarter this warp may touch
        mbar_wait(accum_full, 0);
        tc_fence_after();
        const int m = m0 + q * 32 + lane;
        const bool drop = p.drop_p > 0.f;
        const float keep_scale = drop ? 1.0f / (1.0f - p.drop_p) : 1.0f;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, r);
            if (m < p.M) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = n0 + c0 + j;
                    if (n >= p.N) break;
                    const long long off = (long long)m * p.ldc + n;
                    float v = __uint_as_float(r[j]) * p.alpha + (p.bias ? p.bias[n] : 0.f);
                    if (p.pre) {
                        if (p.c_is_bf16) static_cast<__nv_bfloat16*>(p.pre)[off] = __float2bfloat16_rn(v);
                        else static_cast<float*>(p.pre)[off] = v;
                    }
                    if (p.act == 1) v = fmaxf(v, 0.f);
                    else if (p.act == 2) v = v / (1.0f + __expf(-v));
                    if (drop) v = keep_mask(p.drop_seed, (unsigned long long)off, p.drop_p) ? v * keep_scale : 0.f;
                    if (p.resid) {
                        const long long roff = (long long)m * p.ldr + n;
                        const float rv = p.c_is_bf16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(p.resid)[roff]) : static_cast<const float*>(p.resid)[roff];
                        v = rv + p.beta * v;
                    } else v *= p.beta;
                    if (p.c_is_bf16) static_cast<__nv_bfloat16*>(p.C)[off] = __float2bfloat16_rn(v);
                    else {
                        float* cp = static_cast<float*>(p.C) + off;
                        *cp = p.accumulate ? *cp + v : v;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, BN);
}

// ---- host: tensor maps ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
    }
    return fn;
}

// 2-D bf16 row-major matrix [rows][cols] with row pitch ld (elements); box = box_cols x box_rows, 128B swizzle
static int make_map_2d(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_cols, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_last_error("cuTensorMapEncodeTiled unavailable"); return SARSSL_ERR_UNSUPPORTED; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled failed (%d): rows=%lld cols=%lld ld=%lld box=%dx%d base=%p", (int)r, rows, cols, ld, box_cols, box_rows, base); return SARSSL_ERR_ARG; }
    return SARSSL_OK;
}

template <int BN, bool A_MN, bool B_MN>
static int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, const TcEpi& e, cudaStream_t stream) {
    static bool configured = false;
    const int smem = TcSmem<BN>::kBytes;
    if (!configured) {
        SARSSL_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    dim3 grid((e.N + BN - 1) / BN, (e.M + TBM - 1) / TBM);
    gemm_tc_kernel<BN, A_MN, B_MN><<<grid, kTcThreads, smem, stream>>>(ma, mb, e);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

}  // namespace sarssl

using namespace sarssl;

// Same argument block as sarssl_gemm; requirements of the tensor-core path (otherwise SARSSL_ERR_UNSUPPORTED, and the caller
// uses sarssl_gemm): bf16 operands, no batching, no A-side dropout, each operand either K-major (sXk == 1) or MN-major
// (sAm == 1 / sBn == 1) - any mix -, leading dimensions multiples of 8 elements, 16-byte aligned bases.
extern "C" int sarssl_gemm_tc(const sarssl_gemm_args* a, cudaStream_t stream) {
    SARSSL_CHECK_ARG(a && a->A && a->B && a->C, "gemm_tc: null pointer");
    const bool a_k = a->sAk == 1, a_mn = a->sAm == 1 && !a_k, b_k = a->sBk == 1, b_mn = a->sBn == 1 && !b_k;
    if (a->ab_dtype != SARSSL_BF16 || a->nb1 != 1 || a->nb2 != 1 || a->a_drop_p > 0.f || !(a_k || a_mn) || !(b_k || b_mn)) {
        set_last_error("gemm_tc: unsupported configuration (needs bf16, unbatched, unit stride along K or along M/N for each operand)");
        return SARSSL_ERR_UNSUPPORTED;
    }
    const long long lda = a_k ? a->sAm : a->sAk, ldb = b_k ? a->sBn : a->sBk;
    if ((lda % 8) || (ldb % 8) || !aligned16(a->A) || !aligned16(a->B) || a->M < 1 || a->N < 1 || a->K < 1) {
        set_last_error("gemm_tc: operands must be 16-byte aligned with leading dimensions multiple of 8");
        return SARSSL_ERR_UNSUPPORTED;
    }
    TcEpi e;
    e.C = a->C; e.pre = a->pre_out; e.resid = a->resid; e.bias = a->bias; e.ldc = a->ldc; e.ldr = a->ldr ? a->ldr : a->ldc;
    e.M = a->M; e.N = a->N; e.K = a->K; e.alpha = a->alpha; e.beta = a->beta; e.act = a->act; e.accumulate = a->accumulate;
    e.c_is_bf16 = a->c_dtype == SARSSL_BF16; e.drop_p = a->drop_p; e.drop_seed = a->drop_seed;
    if (e.accumulate && e.c_is_bf16) { set_last_error("gemm_tc: accumulate needs an fp32 C"); return SARSSL_ERR_UNSUPPORTED; }
    const bool bn64 = a->N <= 64;
    const int BN = bn64 ? 64 : 128;
    CUtensorMap ma, mb;
    int rc;
    // K-major operand: global [M or N rows][K cols], one box of 64 k x (128 | BN) rows.  MN-major: global [K rows][M or N cols], 64 x 64 boxes.
    if ((rc = a_k ? make_map_2d(&ma, a->A, a->M, a->K, lda, TBK, TBM) : make_map_2d(&ma, a->A, a->K, a->M, lda, 64, TBK))) return rc;
    if ((rc = b_k ? make_map_2d(&mb, a->B, a->N, a->K, ldb, TBK, BN) : make_map_2d(&mb, a->B, a->K, a->N, ldb, 64, TBK))) return rc;
    if (a_k && b_k) return bn64 ? launch_tc<64, false, false>(ma, mb, e, stream) : launch_tc<128, false, false>(ma, mb, e, stream);
    if (a_k && b_mn) return bn64 ? launch_tc<64, false, true>(ma, mb, e, stream) : launch_tc<128, false, true>(ma, mb, e, stream);
    if (a_mn && b_k) return bn64 ? launch_tc<64, true, false>(ma, mb, e, stream) : launch_tc<128, true, false>(ma, mb, e, stream);
    return bn64 ? launch_tc<64, true, true>(ma, mb, e, stream) : launch_tc<128, true, true>(ma, mb, e, stream);
}
