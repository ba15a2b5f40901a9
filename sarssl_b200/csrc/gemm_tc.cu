// bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA) for sm_100a.
//
//   C[m][n] = resid[m][n] + beta * drop( act( alpha * sum_k A[m][k] * B[n][k] + bias[n] ) )      (same epilogue as gemm_simt.cu)
//
// Carries the dense contractions of the path: FFN / projection / pointwise-conv / decoder / patch-embedding GEMMs and their
// data gradients (K-major operands: activations [M][K] and weights [N][K], or pre-transposed weights for dgrad), and the
// weight gradients dW = dY^T X where the contraction runs over the token rows (MN-major operands, same kernel).
//
// Persistent: one CTA per SM walks the 128 x BN output tiles (n fastest, so concurrently running CTAs share A rows through L2).
// warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + MMA issuer (warp-uniform loop, one elected lane issues UMMA
// 128xBNx16 instructions), warps 2-9 = epilogue (each owns the 32 TMEM lanes its warp-id % 4 selects and every other 32-column chunk).
// A ring of 128B-swizzled shared-memory stages is handed from TMA to MMA through full/empty mbarriers; tcgen05.commit releases a
// stage when the MMAs reading it retire and signals the epilogue when a tile's accumulator is complete.  The accumulator is double
// buffered in TMEM, so the epilogue of tile i (TMEM -> registers -> bias / activation / dropout / residual -> bf16 -> swizzled smem
// stage -> TMA store) overlaps the main loop of tile i+1; short-K GEMMs are no longer dominated by per-CTA set-up.
#include "common.cuh"
#include "rng.cuh"
#include <cuda.h>

namespace sarssl {

constexpr int TBM = 128, TBK = 64, kEpiWarps = 16, kTcThreads = 64 + 32 * kEpiWarps;       // 2 control warps + 16 epilogue warps (one 32-column chunk each at BN = 128; round 2, after the local-memory fix: FFN-1 -4 %, residual epilogue -9 % against 8)

struct TcEpi {
    void* C; void* pre; const void* resid; const float* bias;
    long long ldc, ldr, sCb1, sCb2;
    int M, N, K, nb2, splitk, kb_per_split, tiles_m, tiles_n;
    long long ntiles;
    int a_z1, a_z2, b_z1, b_z2;          // 1 if the operand really advances along that batch axis (0: broadcast, coordinate stays 0)
    float alpha, beta;
    int act, accumulate, c_is_bf16, vec_ok, tma_store;
    float drop_p; unsigned long long drop_seed;
    const unsigned long long* seed_dev;  // nullable: device word added to drop_seed
};

// ---- PTX wrappers ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(smem_dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
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
// same instruction with the two descriptors given as (lo, hi) words: the issuing thread keeps `hi` and a per-stage base `lo` in
// registers and only adds a small constant per instruction (descriptor construction was the issue-rate bottleneck).
__device__ __forceinline__ void umma_bf16_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
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

template <int BN, int NS, int EPI = 0>
struct TcSmem {
    static constexpr int kABytes = TBM * TBK * 2, kBBytes = BN * TBK * 2;
    // bf16 output tile, 128B-swizzled boxes of 64 columns (C and the pre-activation).  BN = 256: the fp32 weight-gradient variant (EPI 0) stages nothing
    // (partial sums go from registers to red.global); the bf16 variants stage ONE half tile (two boxes) and store the tile in two rounds - the room
    // pays for the fourth 48 KB stage.
    static constexpr int kStageC = BN == 256 ? 2 * (TBM * 128) : (BN / 64) * (TBM * 128);
    static constexpr int kNStage = BN == 256 ? (EPI != 0 ? 1 : 0) : 2;
    static constexpr int kBytes = NS * (kABytes + kBBytes) + kNStage * kStageC + 1024 /*align slack*/ + 256 /*barriers*/;
};


// A_MN / B_MN: operand is MN-major (rows of the global matrix run along the contraction dimension), used by weight gradients.
// NS = pipeline stages: 3 for long K; 2 for K <= 512 (4-8 k-blocks), which lets three CTAs share an SM so that more epilogue warps
// are in flight - those GEMMs are epilogue / store bound.
template <int BN, bool A_MN, bool B_MN, int NS, int EPI = 0>
__global__ void __launch_bounds__(kTcThreads) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                           const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmPre, TcEpi p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kA = TcSmem<BN, NS, EPI>::kABytes, kB = TcSmem<BN, NS, EPI>::kBBytes, kSC = TcSmem<BN, NS, EPI>::kStageC;
    constexpr bool WIDE_ST = BN == 256 && EPI != 0;      // bf16 output of a 128 x 256 tile: staged and stored in two halves of 128 columns
    constexpr int kStages = NS;
    unsigned char* sA = smem;
    unsigned char* sB = smem + kStages * kA;
    unsigned char* stC = smem + kStages * (kA + kB);              // staged C tile
    unsigned char* stP = stC + kSC;                               // staged pre-activation tile
    uint64_t* full = reinterpret_cast<uint64_t*>(stC + TcSmem<BN, NS, EPI>::kNStage * kSC);
    uint64_t* empty = full + kStages;
    uint64_t* tfull = empty + kStages;                            // [2] accumulator ready
    uint64_t* tempty = tfull + 2;                                 // [2] accumulator drained (every epilogue warp arrives)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb_total = (p.K + TBK - 1) / TBK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], kEpiWarps); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);      // two accumulators of BN fp32 columns x 128 lanes
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile id -> (batch/split slice zz, m tile, n tile); n fastest
    auto decode = [&](long long t, int& zz, int& m0, int& n0) {
        n0 = (int)(t % p.tiles_n) * BN; t /= p.tiles_n;
        m0 = (int)(t % p.tiles_m) * TBM;
        zz = (int)(t / p.tiles_m);
    };

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (long long t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
                int zz, m0, n0;
                decode(t, zz, m0, n0);
                const int zs = zz % p.splitk, zb = zz / p.splitk;
                const int z1 = zb / p.nb2, z2 = zb - z1 * p.nb2;
                const int az1 = z1 * p.a_z1, az2 = z2 * p.a_z2, bz1 = z1 * p.b_z1, bz2 = z2 * p.b_z2;
                const int kb_begin = zs * p.kb_per_split, kb_end = min(num_kb_total, kb_begin + p.kb_per_split);
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], kA + kB);
                    if (!A_MN) tma_load_4d(sA + s * kA, &tmA, kb * TBK, m0, az2, az1, &full[s]);
                    else { tma_load_4d(sA + s * kA, &tmA, m0, kb * TBK, az2, az1, &full[s]); tma_load_4d(sA + s * kA + kA / 2, &tmA, m0 + 64, kb * TBK, az2, az1, &full[s]); }
                    if (!B_MN) tma_load_4d(sB + s * kB, &tmB, kb * TBK, n0, bz2, bz1, &full[s]);
                    else {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j) tma_load_4d(sB + s * kB + j * 8192, &tmB, n0 + 64 * j, kb * TBK, bz2, bz1, &full[s]);
                    }
                    if (++s == kStages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        {   // the whole warp runs the loop (uniform control flow keeps descriptors in uniform registers); one elected lane issues
            // BN = 256: two N = 128 instructions per k step into the two halves of the accumulator.  Instructions that accumulate into the columns
            // their predecessor wrote wait for it (scripts/micro/umma_rate.cu: one N = 256 accumulator 171 clk per step, two alternating N = 128
            // halves 128 clk); the point of the wide tile is the A operand, which is fetched from L2 once per 256 output columns instead of twice.
            constexpr uint32_t idesc = make_idesc(TBM, BN == 256 ? 128 : BN, A_MN, B_MN);
            // K-major: 8-row groups 1024 B apart (SBO), advance 32 B per 16-element k step inside the 128 B swizzle atom.
            // MN-major: 64-element MN groups 8192 B apart (LBO), 8-k-row groups 1024 B apart (SBO), advance 2048 B per k step.
            constexpr uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
            constexpr uint32_t a_step = A_MN ? (2048u >> 4) : (32u >> 4), b_step = B_MN ? (2048u >> 4) : (32u >> 4);
            const uint32_t a_lo0 = ((smem_u32(sA) >> 4) & 0x3FFF) | ((A_MN ? (8192u >> 4) : 1u) << 16);
            const uint32_t b_lo0 = ((smem_u32(sB) >> 4) & 0x3FFF) | ((B_MN ? (8192u >> 4) : 1u) << 16);
            int s = 0;
            uint32_t ph = 0, it = 0;
            for (long long t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++it) {
                int zz, m0, n0;
                decode(t, zz, m0, n0);
                const int zs = zz % p.splitk;
                const int kb_begin = zs * p.kb_per_split, kb_end = min(num_kb_total, kb_begin + p.kb_per_split);
                const int num_kb = kb_end - kb_begin;
                const uint32_t acc = it & 1u;
                mbar_wait(&tempty[acc], ((it >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * BN;
                for (int i = 0; i < num_kb; ++i) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a_lo = a_lo0 + (uint32_t)s * (kA >> 4), b_lo = b_lo0 + (uint32_t)s * (kB >> 4);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < TBK / 16; ++k) {
                            umma_bf16_w(d, a_lo + k * a_step, hi, b_lo + k * b_step, hi, idesc, (uint32_t)((i | k) != 0));
                            if (BN == 256)      // columns 128-255: B rows (K-major) or the two 64-column groups (MN-major) 16 KB further
                                umma_bf16_w(d + 128, a_lo + k * a_step, hi, b_lo + (16384u >> 4) + k * b_step, hi, idesc, (uint32_t)((i | k) != 0));
                        }
                        umma_commit(&empty[s]);             // stage reusable once these MMAs have read it
                        if (i == num_kb - 1) umma_commit(&tfull[acc]);     // accumulator complete
                    }
                    __syncwarp();
                    if (++s == kStages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else {
        // EPI != 0: the epilogue variant is fixed at compile time (bit 6 = fast marker, bit 0 bias, bits 1-2 activation, bit 3 second
        // (pre-activation) output, bit 4 dropout, bit 5 residual; bf16 C through the TMA store, alpha = 1, no split-K): the flag tests
        // below fold away and each instantiation is straight-line code.  EPI == 0 keeps every test at run time.
        constexpr bool FAST = EPI != 0;
        const bool f_tma = BN == 256 ? WIDE_ST : (FAST ? true : (p.tma_store != 0));
        const bool f_vec = FAST ? true : (p.vec_ok != 0);
        const bool f_bf16 = FAST ? true : (p.c_is_bf16 != 0);
        const bool f_bias = FAST ? ((EPI & 1) != 0) : (p.bias != nullptr);
        const int f_act = FAST ? ((EPI >> 1) & 3) : p.act;
        const bool f_pre = FAST ? ((EPI & 8) != 0) : (p.pre != nullptr);
        const bool f_resid = FAST ? ((EPI & 32) != 0) : (p.resid != nullptr);
        const float f_alpha = FAST ? 1.0f : p.alpha;
        const int q = warp & 3;                         // TMEM lane quarter this warp may touch
        const int cpart = (warp - 2) >> 2;              // warps sharing a lane quarter interleave the tile's 32-column chunks
        const bool drop = FAST ? ((EPI & 16) != 0) : (p.drop_p > 0.f);
        const unsigned long long drop_seed = p.drop_seed + ((drop && p.seed_dev) ? *p.seed_dev : 0ull);
        const uint32_t thr = drop_threshold(p.drop_p);
        const float keep_scale = drop ? 1.0f / (1.0f - p.drop_p) : 1.0f;
        const bool atomic = FAST ? false : (p.splitk > 1);
        uint32_t it = 0;
        for (long long t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++it) {
            int zz, m0, n0;
            decode(t, zz, m0, n0);
            const int zb = zz / p.splitk;
            const int z1 = zb / p.nb2, z2 = zb - z1 * p.nb2;
            const uint32_t acc = it & 1u;
            if (f_tma) {                          // the previous tile's TMA store must have finished reading the staging buffers
                if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
            }
            mbar_wait(&tfull[acc], (it >> 1) & 1u);
            tc_fence_after();
            const int m = m0 + q * 32 + lane;
            const long long zoff = (long long)z1 * p.sCb1 + (long long)z2 * p.sCb2;
#pragma unroll 1
            for (int c0 = cpart * 32; c0 < BN; c0 += 32 * (kEpiWarps / 4)) {
                if (WIDE_ST && c0 >= 128) {           // second half: the first half's store must have left the staging boxes
                    if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
                }
                const int sb = WIDE_ST ? ((c0 & 127) >> 6) : (c0 >> 6);      // staging box of this chunk
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + c0, r);
                const int nb = n0 + c0;
                if (!f_tma && (m >= p.M || nb >= p.N)) continue;
                if (f_tma && nb >= p.N) continue;              // (N % 32 == 0 on this path: whole chunks only)
                const long long off = zoff + (long long)m * p.ldc + nb;
                const long long roff = zoff + (long long)m * p.ldr + nb;
                if (f_vec && (FAST || nb + 32 <= p.N)) {
                    // ---------------- vector path: 32 consecutive columns of one row
                    float v[32];
                    const bool row_ok = m < p.M;                     // rows past M only exist to fill the staged tile; TMA clips them
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 bz = f_bias ? __ldg(reinterpret_cast<const float4*>(p.bias + nb + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        v[j] = __uint_as_float(r[j]) * f_alpha + bz.x; v[j + 1] = __uint_as_float(r[j + 1]) * f_alpha + bz.y;
                        v[j + 2] = __uint_as_float(r[j + 2]) * f_alpha + bz.z; v[j + 3] = __uint_as_float(r[j + 3]) * f_alpha + bz.w;
                    }
                    if (f_pre) {
                        if (f_bf16) {
                            unsigned char* sbox = stP + sb * (TBM * 128) + (q * 32 + lane) * 128;
                            uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.pre) + off);
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                __nv_bfloat162 t0 = __floats2bfloat162_rn(v[j], v[j + 1]), t1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                                __nv_bfloat162 t2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), t3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                                const uint4 pk = make_uint4(*reinterpret_cast<uint32_t*>(&t0), *reinterpret_cast<uint32_t*>(&t1), *reinterpret_cast<uint32_t*>(&t2),
                                                            *reinterpret_cast<uint32_t*>(&t3));
                                if (f_tma) *reinterpret_cast<uint4*>(sbox + (((((c0 & 63) + j) >> 3) ^ ((q * 32 + lane) & 7)) << 4)) = pk;
                                else dst[j >> 3] = pk;
                            }
                        } else {
                            float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.pre) + off);
#pragma unroll
                            for (int j = 0; j < 32; j += 4) dst[j >> 2] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        }
                    }
                    if (f_act == 1) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                    } else if (f_act == 2) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __fdividef(v[j], 1.0f + __expf(-v[j]));
                    }
                    if (drop) {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {              // off is even here (ldc % 8 == 0, nb % 32 == 0)
                            const uint32_t kp = keep_pair(drop_seed, (unsigned long long)(off + j) >> 1, thr);
                            v[j] = (kp & 1u) ? v[j] * keep_scale : 0.f;
                            v[j + 1] = (kp & 2u) ? v[j + 1] * keep_scale : 0.f;
                        }
                    }
                    if (f_resid && row_ok) {
                        if (f_bf16) {
                            const uint4* src = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.resid) + roff);
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                const uint4 u = src[j >> 3];
                                const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x)), f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
                                const float2 f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.z)), f3 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.w));
                                v[j] = f0.x + p.beta * v[j]; v[j + 1] = f0.y + p.beta * v[j + 1]; v[j + 2] = f1.x + p.beta * v[j + 2]; v[j + 3] = f1.y + p.beta * v[j + 3];
                                v[j + 4] = f2.x + p.beta * v[j + 4]; v[j + 5] = f2.y + p.beta * v[j + 5]; v[j + 6] = f3.x + p.beta * v[j + 6]; v[j + 7] = f3.y + p.beta * v[j + 7];
                            }
                        } else {
                            const float4* src = reinterpret_cast<const float4*>(static_cast<const float*>(p.resid) + roff);
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 u = src[j >> 2];
                                v[j] = u.x + p.beta * v[j]; v[j + 1] = u.y + p.beta * v[j + 1]; v[j + 2] = u.z + p.beta * v[j + 2]; v[j + 3] = u.w + p.beta * v[j + 3];
                            }
                        }
                    } else if (!f_resid && p.beta != 1.0f) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= p.beta;
                    }
                    if (f_bf16) {
                        unsigned char* sbox = stC + sb * (TBM * 128) + (q * 32 + lane) * 128;
                        uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.C) + off);
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            __nv_bfloat162 t0 = __floats2bfloat162_rn(v[j], v[j + 1]), t1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                            __nv_bfloat162 t2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), t3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                            const uint4 pk = make_uint4(*reinterpret_cast<uint32_t*>(&t0), *reinterpret_cast<uint32_t*>(&t1), *reinterpret_cast<uint32_t*>(&t2),
                                                        *reinterpret_cast<uint32_t*>(&t3));
                            if (f_tma) *reinterpret_cast<uint4*>(sbox + (((((c0 & 63) + j) >> 3) ^ ((q * 32 + lane) & 7)) << 4)) = pk;
                            else dst[j >> 3] = pk;
                        }
                    } else {
                        float* cp = static_cast<float*>(p.C) + off;
                        if (atomic) {                               // split-K partial sums: 16-byte vector reductions (4x fewer L2 atomic operations)
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
                        } else if (p.accumulate) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                float4 u = *reinterpret_cast<float4*>(cp + j);
                                u.x += v[j]; u.y += v[j + 1]; u.z += v[j + 2]; u.w += v[j + 3];
                                *reinterpret_cast<float4*>(cp + j) = u;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        }
                    }
                    if (WIDE_ST) {                    // this half (128 columns = two boxes) is staged: store it; the accumulator is free after the second
                        if (c0 >= 128) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tempty[acc]);
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
                        if (warp == 2 && lane == 0) {
#pragma unroll
                            for (int j = 0; j < 2; ++j)
                                asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(&tmC),
                                             "r"(n0 + (c0 & 128) + 64 * j), "r"(m0), "r"(z2), "r"(z1), "r"(smem_u32(stC + j * (TBM * 128))) : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    }
                    continue;
                }
                // ---------------- scalar path (ragged N or unaligned C); fully unrolled: a run-time index would push r[] into local memory
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = nb + j;
                    if (n >= p.N) continue;
                    float v = __uint_as_float(r[j]) * f_alpha + (f_bias ? p.bias[n] : 0.f);
                    if (f_pre) {
                        if (f_bf16) static_cast<__nv_bfloat16*>(p.pre)[off + j] = __float2bfloat16_rn(v);
                        else static_cast<float*>(p.pre)[off + j] = v;
                    }
                    if (f_act == 1) v = fmaxf(v, 0.f);
                    else if (f_act == 2) v = __fdividef(v, 1.0f + __expf(-v));
                    if (drop) v = keep_mask(drop_seed, (unsigned long long)(off + j), p.drop_p) ? v * keep_scale : 0.f;
                    if (f_resid) {
                        const float rv = f_bf16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(p.resid)[roff + j]) : static_cast<const float*>(p.resid)[roff + j];
                        v = rv + p.beta * v;
                    } else v *= p.beta;
                    if (f_bf16) static_cast<__nv_bfloat16*>(p.C)[off + j] = __float2bfloat16_rn(v);
                    else {
                        float* cp = static_cast<float*>(p.C) + off + j;
                        if (atomic) atomicAdd(cp, v);
                        else *cp = p.accumulate ? *cp + v : v;
                    }
                }
            }
            if (WIDE_ST) continue;                    // released and stored half by half above
            // accumulator drained: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (f_tma) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
                if (warp == 2 && lane == 0) {
#pragma unroll
                    for (int j = 0; j < BN / 64; ++j) {
                        if (n0 + 64 * j < p.N) {
                            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(&tmC), "r"(n0 + 64 * j), "r"(m0),
                                         "r"(z2), "r"(z1), "r"(smem_u32(stC + j * (TBM * 128))) : "memory");
                            if (f_pre)
                                asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(&tmPre), "r"(n0 + 64 * j),
                                             "r"(m0), "r"(z2), "r"(z1), "r"(smem_u32(stP + j * (TBM * 128))) : "memory");
                        }
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (f_tma && warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 2 * BN);
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

// bf16 matrix [rows][cols] with row pitch ld (elements), replicated over two batch axes (element strides sb1, sb2; a zero
// stride means "broadcast": that axis is collapsed to size 1).  box = box_cols x box_rows x 1 x 1, 128B swizzle.
static int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_cols, int box_rows, int nb1, long long sb1,
                    int nb2, long long sb2) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_last_error("cuTensorMapEncodeTiled unavailable"); return SARSSL_ERR_UNSUPPORTED; }
    const bool use1 = nb1 > 1 && sb1 != 0, use2 = nb2 > 1 && sb2 != 0;
    cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)(use2 ? nb2 : 1), (cuuint64_t)(use1 ? nb1 : 1)};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)(use2 ? sb2 * 2 : ld * 2 * rows), (cuuint64_t)(use1 ? sb1 * 2 : ld * 2 * rows)};
    cuuint32_t box[4] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed (%d): rows=%lld cols=%lld ld=%lld box=%dx%d nb=%dx%d sb=%lld,%lld base=%p", (int)r, rows, cols, ld, box_cols,
                       box_rows, nb1, nb2, sb1, sb2, base);
        return SARSSL_ERR_UNSUPPORTED;
    }
    return SARSSL_OK;
}

template <int BN, bool A_MN, bool B_MN, int NS, int EPI = 0>
static int launch_tc_ns(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mp, const TcEpi& e, cudaStream_t stream) {
    static bool configured = false;
    const int smem = TcSmem<BN, NS, EPI>::kBytes;
    if (!configured) {
        SARSSL_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, A_MN, B_MN, NS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    const int grid = (int)(e.ntiles < sm_count() ? e.ntiles : sm_count());
    gemm_tc_kernel<BN, A_MN, B_MN, NS, EPI><<<grid, kTcThreads, smem, stream>>>(ma, mb, mc, mp, e);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

template <int BN, bool A_MN, bool B_MN>
static int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mp, const TcEpi& e, cudaStream_t stream) {
    return launch_tc_ns<BN, A_MN, B_MN, 4>(ma, mb, mc, mp, e, stream);
}

// SARSSL_GEMM_WIDE=0 keeps the 128 x 128 tiles everywhere (A/B measurements)
static bool wide_tiles_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("SARSSL_GEMM_WIDE"); on = e ? (atoi(e) != 0) : 1; }
    return on != 0;
}

// SARSSL_GEMM_WIDE_MIN_K: shortest K for which the bf16 GEMMs use the wide tile (A/B measurements; default 256: in-step 43.15 ms against 43.32 ms at 512)
static int wide_min_k() {
    static int k = -1;
    if (k < 0) { const char* e = getenv("SARSSL_GEMM_WIDE_MIN_K"); k = e ? atoi(e) : 256; }
    return k;
}

}  // namespace sarssl

using namespace sarssl;

// Same argument block as sarssl_gemm.  Requirements of the tensor-core path (otherwise SARSSL_ERR_UNSUPPORTED and the caller
// uses sarssl_gemm): bf16 operands, no A-side dropout, each operand K-major (unit stride along K) or MN-major (unit stride along
// M / N), leading dimensions and batch strides multiples of 8 elements, 16-byte aligned bases.
extern "C" int sarssl_gemm_tc(const sarssl_gemm_args* a, cudaStream_t stream) {
    SARSSL_CHECK_ARG(a && a->A && a->B && a->C, "gemm_tc: null pointer");
    const bool a_k = a->sAk == 1, a_mn = a->sAm == 1 && !a_k, b_k = a->sBk == 1, b_mn = a->sBn == 1 && !b_k;
    if (a->ab_dtype != SARSSL_BF16 || a->a_drop_p > 0.f || !(a_k || a_mn) || !(b_k || b_mn) || a->nb1 < 1 || a->nb2 < 1) {
        set_last_error("gemm_tc: unsupported configuration (needs bf16, no A-side dropout, unit stride along K or along M/N for each operand)");
        return SARSSL_ERR_UNSUPPORTED;
    }
    const long long lda = a_k ? a->sAm : a->sAk, ldb = b_k ? a->sBn : a->sBk;
    if ((lda % 8) || (ldb % 8) || (a->sAb1 % 8) || (a->sAb2 % 8) || (a->sBb1 % 8) || (a->sBb2 % 8) || !aligned16(a->A) || !aligned16(a->B) || a->M < 1 ||
        a->N < 1 || a->K < 1 || (long long)a->nb1 * a->nb2 > 16384) {
        set_last_error("gemm_tc: operands must be 16-byte aligned with leading dimensions / batch strides multiple of 8");
        return SARSSL_ERR_UNSUPPORTED;
    }
    const int nbatch = a->nb1 * a->nb2;
    TcEpi e;
    e.C = a->C; e.pre = a->pre_out; e.resid = a->resid; e.bias = a->bias; e.ldc = a->ldc; e.ldr = a->ldr ? a->ldr : a->ldc;
    e.sCb1 = a->sCb1; e.sCb2 = a->sCb2; e.nb2 = a->nb2;
    e.M = a->M; e.N = a->N; e.K = a->K; e.alpha = a->alpha; e.beta = a->beta; e.act = a->act; e.accumulate = a->accumulate;
    e.c_is_bf16 = a->c_dtype == SARSSL_BF16; e.drop_p = a->drop_p; e.drop_seed = a->drop_seed; e.seed_dev = a->seed_dev;
    e.a_z1 = (a->nb1 > 1 && a->sAb1 != 0); e.a_z2 = (a->nb2 > 1 && a->sAb2 != 0); e.b_z1 = (a->nb1 > 1 && a->sBb1 != 0); e.b_z2 = (a->nb2 > 1 && a->sBb2 != 0);
    if (e.accumulate && e.c_is_bf16) { set_last_error("gemm_tc: accumulate needs an fp32 C"); return SARSSL_ERR_UNSUPPORTED; }
    const size_t esz = e.c_is_bf16 ? 2 : 4;
    e.vec_ok = (a->ldc % 8 == 0) && (e.ldr % 8 == 0) && (a->sCb1 % 8 == 0) && (a->sCb2 % 8 == 0) && aligned16(a->C) && (!a->pre_out || aligned16(a->pre_out)) &&
               (!a->resid || aligned16(a->resid)) && (!a->bias || aligned16(a->bias));
    (void)esz;
    const bool bn64 = a->N <= 64;
    // 128 x 256 tiles for the split-K weight gradients (both operands MN-major, fp32 partial sums): the tall operand is read from L2 once per 256
    // output columns; these GEMMs move 32 KB of operands per 128 x 128 x 64 block and are bound by that traffic.  Measured at K = 32768: 2048 x 512
    // 768 -> 865 TFLOP/s, 1024 x 3072 856 -> 1078, 512 x 1024 659 -> 825; 1024 x 256 alone loses at that K (twice the split-K slices: 603 -> 499) but
    // inside the step (K = 65536) wide tiles for every shape from 1024 x 256 up gave -0.43 ms per step against -0.18 ms when those were excluded.
    const bool bn256 = a_mn && b_mn && a->N % 256 == 0 && (long long)a->M * a->N >= 256LL * 1024 && a->accumulate && a->c_dtype != SARSSL_BF16 && !a->pre_out && !a->resid && !a->bias && a->act == 0 &&
                       a->drop_p == 0.f && wide_tiles_enabled();
    // ... and for the bf16 GEMMs of the layer types below when N is a multiple of 256 and K is long enough for the operand traffic to matter
    // (forward linears without a second output, data gradients): the tile is staged and stored in two halves
    const int fast_flags = 64 | (a->bias ? 1 : 0) | ((a->act & 3) << 1) | (a->pre_out ? 8 : 0) | (a->drop_p > 0.f ? 16 : 0) | (a->resid ? 32 : 0);
    const bool fast_shape = e.c_is_bf16 && e.vec_ok && a->alpha == 1.0f && !a->accumulate && (a->resid || a->beta == 1.0f);
    const bool wide_fast = fast_shape && a->N % 256 == 0 && a->K >= wide_min_k() && wide_tiles_enabled() &&
                           ((a_k && b_k && (fast_flags == 64 || fast_flags == (64 | 1) || fast_flags == (64 | 1 | 2) || fast_flags == (64 | 1 | 16 | 32) ||
                                            fast_flags == (64 | 1 | 32))) ||
                            (a_k && b_mn && fast_flags == 64));
    const int BN = bn64 ? 64 : ((bn256 || wide_fast) ? 256 : 128);
    // split-K: weight gradients (fp32 accumulate, plain epilogue) have few output tiles and a very long K
    const int num_kb = (a->K + TBK - 1) / TBK;
    int splitk = 1;
    if (e.accumulate && !e.c_is_bf16 && !a->pre_out && !a->resid && !a->bias && a->act == 0 && a->drop_p == 0.f) {
        const long long tiles = (long long)((a->N + BN - 1) / BN) * ((a->M + TBM - 1) / TBM) * nbatch;
        const long long want = ((tiles <= 16 ? 3LL : 4LL) * sm_count() + tiles - 1) / tiles;     // 3-4 work items per SM balance best (measured)
        splitk = (int)(want < 1 ? 1 : want);
        if (splitk > num_kb / 4) splitk = num_kb / 4 > 0 ? num_kb / 4 : 1;       // at least 4 k-blocks per slice
    }
    e.kb_per_split = (num_kb + splitk - 1) / splitk;
    e.splitk = (num_kb + e.kb_per_split - 1) / e.kb_per_split;
    e.tiles_m = (a->M + TBM - 1) / TBM;
    e.tiles_n = (a->N + BN - 1) / BN;
    e.ntiles = (long long)e.tiles_m * e.tiles_n * nbatch * e.splitk;
    CUtensorMap ma, mb, mc, mp;
    int rc;
    // bf16 outputs leave through a swizzled shared-memory stage + TMA store (full 128-byte lines instead of 16-byte pieces per thread);
    e.tma_store = e.c_is_bf16 && e.vec_ok && (a->N % 32 == 0);
    memset(&mc, 0, sizeof(mc));
    memset(&mp, 0, sizeof(mp));
    if (e.tma_store) {
        if ((rc = make_map(&mc, a->C, a->M, a->N, a->ldc, 64, TBM, a->nb1, a->sCb1, a->nb2, a->sCb2))) return rc;
        if (a->pre_out && (rc = make_map(&mp, a->pre_out, a->M, a->N, a->ldc, 64, TBM, a->nb1, a->sCb1, a->nb2, a->sCb2))) return rc;
    }
    // K-major operand: global [M or N rows][K cols], one box of 64 k x (128 | BN) rows.  MN-major: global [K rows][M or N cols], 64 x 64 boxes.
    if ((rc = a_k ? make_map(&ma, a->A, a->M, a->K, lda, TBK, TBM, a->nb1, a->sAb1, a->nb2, a->sAb2)
                  : make_map(&ma, a->A, a->K, a->M, lda, 64, TBK, a->nb1, a->sAb1, a->nb2, a->sAb2))) return rc;
    if ((rc = b_k ? make_map(&mb, a->B, a->N, a->K, ldb, TBK, BN, a->nb1, a->sBb1, a->nb2, a->sBb2)
                  : make_map(&mb, a->B, a->K, a->N, ldb, 64, TBK, a->nb1, a->sBb1, a->nb2, a->sBb2))) return rc;
    // compile-time epilogue variants for the layer types of the model (everything else runs the generic epilogue)
    if (wide_fast) {
        if (a_k && b_mn) return launch_tc_ns<256, false, true, 4, 64>(ma, mb, mc, mp, e, stream);
        switch (fast_flags) {
            case 64: return launch_tc_ns<256, false, false, 4, 64>(ma, mb, mc, mp, e, stream);
            case 64 | 1: return launch_tc_ns<256, false, false, 4, 64 | 1>(ma, mb, mc, mp, e, stream);
            case 64 | 1 | 2: return launch_tc_ns<256, false, false, 4, 64 | 1 | 2>(ma, mb, mc, mp, e, stream);
            case 64 | 1 | 16 | 32: return launch_tc_ns<256, false, false, 4, 64 | 1 | 16 | 32>(ma, mb, mc, mp, e, stream);
            default: return launch_tc_ns<256, false, false, 4, 64 | 1 | 32>(ma, mb, mc, mp, e, stream);
        }
    }
    if (!bn64 && e.tma_store && e.splitk == 1 && a->alpha == 1.0f && !a->accumulate) {
        const int flags = 64 | (a->bias ? 1 : 0) | ((a->act & 3) << 1) | (a->pre_out ? 8 : 0) | (a->drop_p > 0.f ? 16 : 0) | (a->resid ? 32 : 0);
        const bool beta_ok = a->resid || a->beta == 1.0f;
        if (beta_ok && a_k && b_k) {
            switch (flags) {
                case 64: return launch_tc_ns<128, false, false, 4, 64>(ma, mb, mc, mp, e, stream);                       // plain
                case 64 | 1: return launch_tc_ns<128, false, false, 4, 64 | 1>(ma, mb, mc, mp, e, stream);               // bias
                case 64 | 1 | 2: return launch_tc_ns<128, false, false, 4, 64 | 1 | 2>(ma, mb, mc, mp, e, stream);       // bias + ReLU
                case 64 | 1 | 4 | 8 | 16: return launch_tc_ns<128, false, false, 4, 64 | 1 | 4 | 8 | 16>(ma, mb, mc, mp, e, stream);   // bias + Swish + pre + dropout (FFN 1)
                case 64 | 1 | 4 | 8: return launch_tc_ns<128, false, false, 4, 64 | 1 | 4 | 8>(ma, mb, mc, mp, e, stream);             // the same in eval mode
                case 64 | 1 | 16 | 32: return launch_tc_ns<128, false, false, 4, 64 | 1 | 16 | 32>(ma, mb, mc, mp, e, stream);         // bias + dropout + residual
                case 64 | 1 | 32: return launch_tc_ns<128, false, false, 4, 64 | 1 | 32>(ma, mb, mc, mp, e, stream);                   // the same in eval mode
                default: break;
            }
        }
        if (beta_ok && a_k && b_mn && flags == 64) return launch_tc_ns<128, false, true, 4, 64>(ma, mb, mc, mp, e, stream);    // data gradients
    }
    if (a_k && b_k) return bn64 ? launch_tc<64, false, false>(ma, mb, mc, mp, e, stream) : launch_tc<128, false, false>(ma, mb, mc, mp, e, stream);
    if (a_k && b_mn) return bn64 ? launch_tc<64, false, true>(ma, mb, mc, mp, e, stream) : launch_tc<128, false, true>(ma, mb, mc, mp, e, stream);
    if (a_mn && b_k) return bn64 ? launch_tc<64, true, false>(ma, mb, mc, mp, e, stream) : launch_tc<128, true, false>(ma, mb, mc, mp, e, stream);
    if (bn256) return launch_tc<256, true, true>(ma, mb, mc, mp, e, stream);
    return bn64 ? launch_tc<64, true, true>(ma, mb, mc, mp, e, stream) : launch_tc<128, true, true>(ma, mb, mc, mp, e, stream);
}
