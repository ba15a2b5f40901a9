// 3x3 convolution 64 -> 64 (padding 1) on the tcgen05 tensor cores as an implicit GEMM - the CNN patch-embedding stem's
// dominant cost (69 % of the model FLOPs; reference: nn.Conv2d(64, 64, 3, padding=1) at model.py:54,57 through cuDNN).
//
// Images are channel-last bf16 [B][H][W][64]; one pixel is one 128-byte row, exactly one row of a 128B-swizzled UMMA operand.
//
// Forward / data gradient  (conv3x3_tc_kernel, persistent, 1 CTA per SM):
//   out[b,h,w,:] = sum_{dh,dw} in[b,h+dh,w+dw,:] . Wp[:, tap(dh,dw), :]        (data gradient = same kernel, mirrored weights)
//   * output tile = 128 consecutive pixels of TWO vertically adjacent image rows, accumulator 128 x 128 fp32 in TMEM (double
//     buffered): columns 0-63 = row h, 64-127 = row h+1.  Input rows h and h+1 feed both output rows, so their UMMAs run with
//     N = 128 against two weight taps that sit side by side in shared memory ([W(dh) | W(dh-1)]); rows h-1 and h+2 feed one output
//     row each (N = 64).  With N = 64 alone every instruction re-reads its 128 x 16 A slice for half the math and the kernel is bound
//     by shared-memory operand reads (measured: the same FLOPs issued as N = 128 run 18 % faster);
//   * A operand: ONE TMA box per image row ([w0-1, w0+135) x 64 ch, out-of-bounds pixels zero-filled = the padding) serves the
//     three horizontal taps by sliding the UMMA descriptor start by one pixel row (128 B) - swizzling is address based, see
//     tests/test_tc_probe_gpu.py; a rolling ring of row boxes lets vertically consecutive tiles re-use two of their three rows,
//     so shared memory is filled with ~1.06 x the input instead of 9 x (which would saturate the L2 -> SM path);
//   * B operand: all 9 taps of the packed weights (72 KB) stay resident in shared memory for the life of the CTA;
//   * warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue (TMEM -> registers -> bf16 -> 128-byte row stores).
//
// Weight gradient  (conv3x3_wgrad_tc_kernel, persistent, split over pixels):
//   dWp[o, tap, ci] = sum_pixels dy[p, o] * in[p + tap, ci]; the contraction runs over pixels, so both operands are MN-major;
//   the whole result lives in TMEM.  Output rows are processed in pairs: the dy boxes of rows h and h+1 sit side by side in shared memory,
//   so the input rows that feed both (h, h+1) run as N = 128 instructions ([dh | dh-1]); rows h-1 and h+2 feed one tap row each (N = 64).
//   Taps dw = -1 / 0 fill the 128 lanes of an instruction through the pixel shift (the second 64-channel group starts one pixel later).
//   Tap dw = +1 has no horizontal partner, so it is stacked VERTICALLY instead: input rows h and h+1 occupy neighbouring ring slots (the
//   ring starts at an odd slot, so such a pair never wraps) and form one M = 128 operand whose two groups are a slot apart; against
//   [dy(h) | dy(h+1)] one instruction yields (dh 0 | dh -1) in the lower and (dh +1 | dh 0) in the upper lanes.  Only rows h-1 / h+2 of that
//   tap still issue half-empty instructions: 640 instead of 768 column-units per k-step.  A tap that receives contributions in two TMEM
//   regions gets two slots in the per-CTA partial sums; the fixed-order reduction kernel adds them.
#include "common.cuh"
#include <cuda.h>

namespace sarssl {

// ---- shared PTX helpers (same encodings as gemm_tc.cu) ----------------------------------------------------------------------
__device__ __forceinline__ void ctma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)), "l"(map),
                 "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ctma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t cdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
// descriptor words: lo = start>>4 | (LBO>>4)<<16, hi = SBO>>4 | version 1 (bit 46) | SWIZZLE_128B (bits 61-63).  The MMA-issuing thread
// keeps `hi` and a per-operand base `lo` in registers and only adds a small constant per instruction.
__device__ __forceinline__ uint32_t cdesc_lo(uint32_t saddr, uint32_t lbo) { return ((saddr >> 4) & 0x3FFF) | (((lbo >> 4) & 0x3FFF) << 16); }
__device__ __forceinline__ constexpr uint32_t cdesc_hi(uint32_t sbo) { return ((sbo >> 4) & 0x3FFF) | (1u << 14) | (2u << 29); }
__device__ __forceinline__ void cumma2(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void cumma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db),
                 "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void ccommit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one lane of a converged warp (elect.sync); the rest of the MMA warp runs the same uniform loop so descriptor arithmetic stays in
// uniform registers instead of being R2UR-broadcast before every instruction
__device__ __forceinline__ bool celect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void cfence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cfence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ctmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- operand transform: BatchNorm + ReLU applied to a TMA-landed row box in place -------------------------------------------------
// The stem's second 3x3 convolution consumes z = relu(scale * y + shift) (model.py:52-56).  Round 1 materialised z with a separate pass
// (read y, write z: 0.66 ms per encoder at 256 clips) because TMA feeds the UMMA directly.  Here four extra warps sit between the
// TMA's `full` barrier and the `ready` barrier the MMA warp waits on: they rewrite the box (pixels x 64 channels bf16, 128-byte swizzled
// rows) in shared memory - smem -> registers -> fma + max -> bf16 -> smem, same arithmetic and rounding as bn_act_fwd_kernel - so z never
// exists in HBM.  Zero padding must stay zero: whole out-of-image rows are skipped, out-of-image pixels of a row are masked.
// Thread t of the 128 owns the logical 16-byte chunk t & 7 (= 8 fixed channels: their scale / shift live in registers) of pixels
// (t >> 3) + 16 i; the physical chunk is (t & 7) ^ (pixel & 7).
struct XfConst {
    float sc[8], sh[8];
    __device__ __forceinline__ void load(const float* scale, const float* shift, int t) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j] = scale[(t & 7) * 8 + j]; sh[j] = shift[(t & 7) * 8 + j]; }
    }
};
// NITER = ceil(npix / 16) chunks per thread: all loads first, then the arithmetic, then the stores (one shared-memory round trip of latency
// per box instead of NITER: the transform sits on the TMA -> MMA path of a ring that cannot be made deeper in the forward kernel).
template <int NITER>
__device__ __forceinline__ void xf_box(unsigned char* box, int npix, int w_first, int W, int t, const XfConst& c) {
    const int ch = t & 7, r0 = t >> 3;
    uint4 u[NITER];
    bool ok[NITER];
#pragma unroll
    for (int i = 0; i < NITER; ++i) {
        const int r = r0 + 16 * i, w = w_first + r;
        ok[i] = r < npix && w >= 0 && w < W;
        if (ok[i]) u[i] = *reinterpret_cast<const uint4*>(box + r * 128 + ((ch ^ (r & 7)) << 4));
    }
#pragma unroll
    for (int i = 0; i < NITER; ++i) {
        uint32_t* pw = reinterpret_cast<uint32_t*>(&u[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float lo = __uint_as_float(pw[j] << 16), hi = __uint_as_float(pw[j] & 0xFFFF0000u);
            const float a = fmaxf(fmaf(lo, c.sc[2 * j], c.sh[2 * j]), 0.f), b = fmaxf(fmaf(hi, c.sc[2 * j + 1], c.sh[2 * j + 1]), 0.f);
            const __nv_bfloat162 o = __floats2bfloat162_rn(a, b);
            pw[j] = *reinterpret_cast<const uint32_t*>(&o);
        }
    }
#pragma unroll
    for (int i = 0; i < NITER; ++i) {
        const int r = r0 + 16 * i;
        if (ok[i]) *reinterpret_cast<uint4*>(box + r * 128 + ((ch ^ (r & 7)) << 4)) = u[i];
    }
}
__device__ __forceinline__ void xf_publish(uint64_t* bar, int lane) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy writes -> visible to the tensor core's async-proxy reads
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

// ============================================================================================================================
// forward / data gradient
// ============================================================================================================================
constexpr int CW = 128;                  // output pixels per tile
constexpr int CBOXR = 136;               // rows per input box: CW + 2 halo pixels, rounded up to 8
constexpr int CBOXB = CBOXR * 128;       // 17408 B (multiple of 1024)
constexpr int CSLOTS = 6;                // ring of row boxes
constexpr int CSEG = 32;                 // image rows per work unit (2 halo row loads amortised over 32 tiles)
constexpr int CWB = 9 * 8192;            // resident weights
constexpr int COUTB = CW * 128;           // 16 KB output staging tile (x2)
constexpr int kConvSmem = CSLOTS * CBOXB + CWB + 2 * COUTB + 1024 + 256;
constexpr int kConvThreads = 192;          // weight-gradient kernel: 2 control + 4 epilogue warps
constexpr int kConvFwdThreads = 320;       // forward kernel: 2 control + 8 epilogue warps (the epilogue paces the tile rate)
constexpr int kConvXfThreads = 448;        // + 4 operand-transform warps (fused BatchNorm + ReLU of the input)

struct ConvP {
    const float* in_scale;               // XF: per-channel scale / shift of the input's BatchNorm (the input is y, the operand relu(scale * y + shift))
    const float* in_shift;
    __nv_bfloat16* out;
    float* bn_partials;                  // nullable: [gridDim.x][2][64] per-CTA sums of out and out^2 (BatchNorm batch statistics of the conv output)
    int B, H, W, tiles_w, nseg_h;
    long long nunits;
};

template <bool XF>
__global__ void __launch_bounds__(XF ? kConvXfThreads : kConvFwdThreads, 1) conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmIn,
                                                                                      const __grid_constant__ CUtensorMap tmW,
                                                                                      const __grid_constant__ CUtensorMap tmOut, ConvP p) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    unsigned char* ring = smem;
    unsigned char* wsm = smem + CSLOTS * CBOXB;
    unsigned char* osm = smem + CSLOTS * CBOXB + CWB;            // 2 x 16 KB, 1024-aligned
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + CSLOTS * CBOXB + CWB + 2 * COUTB);
    uint64_t* empty = full + CSLOTS;
    uint64_t* wbar = empty + CSLOTS;
    uint64_t* tfull = wbar + 1;          // [2]
    uint64_t* tempty = tfull + 2;        // [2]
    uint64_t* ready = tempty + 2;        // [CSLOTS] XF: box transformed (one arrival per transform warp)
    uint32_t* slot_tmem = reinterpret_cast<uint32_t*>(ready + CSLOTS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmIn) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOut) : "memory");
        for (int s = 0; s < CSLOTS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&ready[s], 4); }
        mbar_init(wbar, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_tmem)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    cfence_before();
    __syncthreads();
    cfence_after();
    const uint32_t tmem = *slot_tmem;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(wbar, CWB);
            // global tap t = (dh+1)*3 + (dw+1) -> slot (dw+1)*3 + (1-dh): for a fixed dw the taps lie in the order dh = +1, 0, -1, so
            // [W(dh) | W(dh-1)] is one contiguous 128-row B operand
            for (int t = 0; t < 9; ++t) ctma_load_2d(wsm + ((t % 3) * 3 + (2 - t / 3)) * 8192, &tmW, t * 64, 0, wbar);
            int s = 0;
            uint32_t ephase = 0;                                // bit s = phase of empty[s] to wait for next (starts "already free")
            for (long long u = blockIdx.x; u < p.nunits; u += gridDim.x) {
                const int hs = (int)(u % p.nseg_h);
                const long long r = u / p.nseg_h;
                const int wt = (int)(r % p.tiles_w), b = (int)(r / p.tiles_w);
                const int h0 = hs * CSEG, h1 = min(p.H, h0 + CSEG), w0 = wt * CW;
                const int npair = (h1 - h0 + 1) >> 1;
                for (int row = h0 - 1; row <= h0 + 2 * npair; ++row) {       // rows past the image are zero-filled by TMA
                    mbar_wait(&empty[s], ((ephase >> s) & 1u) ^ 1u);
                    ephase ^= 1u << s;
                    mbar_expect_tx(&full[s], CBOXB);
                    ctma_load_4d(ring + s * CBOXB, &tmIn, 0, w0 - 1, row, b, &full[s]);
                    if (++s == CSLOTS) s = 0;
                }
            }
        }
    } else if (warp == 1) {
        {
            constexpr uint32_t idesc64 = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
            constexpr uint32_t idesc128 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
            constexpr uint32_t hi = cdesc_hi(1024);
            mbar_wait(wbar, 0);
            const uint32_t w_lo = cdesc_lo(smem_u32(wsm), 16), ring_lo = cdesc_lo(smem_u32(ring), 16);
            int slot0 = 0;                                      // ring slot of the oldest row box (row h - 1) of the current pair
            uint32_t fphase = 0;                                // bit s = phase of full[s] expected next
            uint32_t it = 0;
            for (long long u = blockIdx.x; u < p.nunits; u += gridDim.x) {
                const int hs = (int)(u % p.nseg_h);
                const int h0 = hs * CSEG, h1 = min(p.H, h0 + CSEG);
                const int npair = (h1 - h0 + 1) >> 1;
                for (int t = 0; t < npair; ++t, ++it) {
                    // rows h-1 .. h+2 = slots slot0 .. slot0+3; each box's full barrier is waited exactly once, in order
                    int sl[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) { sl[j] = slot0 + j; if (sl[j] >= CSLOTS) sl[j] -= CSLOTS; }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {               // (unrolled: a run-time index would push sl[] into local memory)
                        if (j < 2 && t != 0) continue;
                        mbar_wait(XF ? &ready[sl[j]] : &full[sl[j]], (fphase >> sl[j]) & 1u);
                        fphase ^= 1u << sl[j];
                    }
                    const uint32_t acc = it & 1u;
                    mbar_wait(&tempty[acc], ((it >> 1) & 1u) ^ 1u);
                    cfence_after();
                    const uint32_t d = tmem + acc * 128;
                    const uint32_t a0 = ring_lo + (uint32_t)sl[0] * (CBOXB >> 4), a1 = ring_lo + (uint32_t)sl[1] * (CBOXB >> 4),
                                   a2 = ring_lo + (uint32_t)sl[2] * (CBOXB >> 4), a3 = ring_lo + (uint32_t)sl[3] * (CBOXB >> 4);
                    const bool last = t == npair - 1;
                    if (celect_one()) {
                        // weight slot (dw, dh) = dw*3 + (1 - dh) (dw, dh as 0..2 / -1..1).  A tcgen05.mma that accumulates into the columns its predecessor
                        // wrote waits ~43 clk for it (scripts/micro/umma_rate.cu), so the two N = 64 streams - input row h-1 into columns 0-63, input row
                        // h+2 into columns 64-127 - are issued ALTERNATELY (independent accumulators: 48 clk each instead of 92); their first instructions
                        // initialise the 128 columns.  Rows h-1 and h die with this pair: their slots are released before row h+1's instructions.
#pragma unroll
                        for (int dw = 0; dw < 3; ++dw)
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                cumma2(d, a0 + dw * 8 + k * 2, hi, w_lo + (dw * 3 + 2) * 512 + k * 2, hi, idesc64, (uint32_t)((dw | k) != 0));       // row h-1: dh = -1
                                cumma2(d + 64, a3 + dw * 8 + k * 2, hi, w_lo + (dw * 3 + 0) * 512 + k * 2, hi, idesc64, (uint32_t)((dw | k) != 0));  // row h+2: dh = +1
                            }
#pragma unroll
                        for (int dw = 0; dw < 3; ++dw)
#pragma unroll
                            for (int k = 0; k < 4; ++k)         // input row h:   dh = 0 for output row h | dh = -1 for output row h+1
                                cumma2(d, a1 + dw * 8 + k * 2, hi, w_lo + (dw * 3 + 1) * 512 + k * 2, hi, idesc128, 1u);
                        ccommit(&empty[sl[0]]);
                        ccommit(&empty[sl[1]]);
#pragma unroll
                        for (int dw = 0; dw < 3; ++dw)
#pragma unroll
                            for (int k = 0; k < 4; ++k)         // input row h+1: dh = +1 for output row h | dh = 0 for output row h+1
                                cumma2(d, a2 + dw * 8 + k * 2, hi, w_lo + (dw * 3 + 0) * 512 + k * 2, hi, idesc128, 1u);
                        if (last) { ccommit(&empty[sl[2]]); ccommit(&empty[sl[3]]); }
                        ccommit(&tfull[acc]);
                    }
                    __syncwarp();
                    slot0 += last ? 4 : 2; if (slot0 >= CSLOTS) slot0 -= CSLOTS;
                }
            }
        }
    } else if (XF && warp >= 10) {
        // operand-transform warps: walk the producer's box sequence, rewrite each landed row box, hand it to the MMA warp
        const int t = threadIdx.x - 320;
        XfConst xc;
        xc.load(p.in_scale, p.in_shift, t);
        int s = 0;
        uint32_t fph = 0;
        for (long long u = blockIdx.x; u < p.nunits; u += gridDim.x) {
            const int hs = (int)(u % p.nseg_h);
            const long long r = u / p.nseg_h;
            const int wt = (int)(r % p.tiles_w);
            const int h0 = hs * CSEG, h1 = min(p.H, h0 + CSEG), w0 = wt * CW;
            const int npair = (h1 - h0 + 1) >> 1;
            for (int row = h0 - 1; row <= h0 + 2 * npair; ++row) {
                mbar_wait(&full[s], (fph >> s) & 1u);
                fph ^= 1u << s;
                if (row >= 0 && row < p.H) xf_box<(CBOXR + 15) / 16>(ring + s * CBOXB, CBOXR, w0 - 1, p.W, t, xc);
                xf_publish(&ready[s], lane);
                if (++s == CSLOTS) s = 0;
            }
        }
    } else {
        const int q = warp & 3;
        const int chalf = (warp - 2) >> 2;                      // two warps share a lane quarter: channels [32*chalf, 32*chalf + 32)
        const int row = q * 32 + lane;                          // pixel row of the tile = TMEM lane
        // fused BatchNorm statistics: thread (channel sc, row quarter sh) adds its 32 staged bf16 values of every output row
        const int et = (warp - 2) * 32 + lane, sc = et & 63, sh = et >> 6;
        float bn_s = 0.f, bn_q = 0.f;
        uint32_t it = 0;
        for (long long u = blockIdx.x; u < p.nunits; u += gridDim.x) {
            const int hs = (int)(u % p.nseg_h);
            const long long r = u / p.nseg_h;
            const int wt = (int)(r % p.tiles_w), b = (int)(r / p.tiles_w);
            const int h0 = hs * CSEG, h1 = min(p.H, h0 + CSEG), w0 = wt * CW;
            for (int h = h0; h < h1; h += 2, ++it) {
                const uint32_t acc = it & 1u;
                const int nrow = min(2, h1 - h);                // the last pair of an odd segment has one real output row
                // both staging buffers were last read by the previous pair's TMA stores (issued a whole tile's MMA time ago)
                if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("bar.sync 1, 256;" ::: "memory");
                mbar_wait(&tfull[acc], (it >> 1) & 1u);
                cfence_after();
#pragma unroll
                for (int rp = 0; rp < 2; ++rp) {
                    unsigned char* stage = osm + rp * COUTB;
                    const int c0 = chalf * 32;
                    uint32_t v[32];
                    ctmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + acc * 128 + rp * 64 + c0, v);
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        uint4 pk;
                        __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
                        __nv_bfloat162 t1 = __floats2bfloat162_rn(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                        __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
                        __nv_bfloat162 t3 = __floats2bfloat162_rn(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
                        pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
                        pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
                        const int chunk = (c0 + j) >> 3;                       // 16-byte chunk of the 128-byte pixel row
                        *reinterpret_cast<uint4*>(stage + row * 128 + ((chunk ^ (row & 7)) << 4)) = pk;     // 128B swizzle, as the TMA store expects
                    }
                }
                cfence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);       // TMEM buffer free for the MMA warp
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (warp == 2 && lane == 0) {
                    for (int rp = 0; rp < nrow; ++rp)
                        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(&tmOut), "r"(0), "r"(w0), "r"(h + rp),
                                     "r"(b), "r"(smem_u32(osm + rp * COUTB)) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (p.bn_partials != nullptr) {
                    const int rmax = min(32, p.W - w0 - sh * 32);                 // rows beyond the image edge hold garbage
                    for (int rp = 0; rp < nrow; ++rp) {
                        const unsigned char* colp = osm + rp * COUTB + ((sc & 7) << 1);
                        for (int r = 0; r < rmax; ++r) {
                            const int rr = sh * 32 + r;
                            const float xv = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(colp + rr * 128 + (((sc >> 3) ^ (rr & 7)) << 4)));
                            bn_s += xv; bn_q += xv * xv;
                        }
                    }
                }
            }
        }
        if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (p.bn_partials != nullptr) {
            // combine the four row quarters through the (now idle) staging buffer, fixed order
            asm volatile("bar.sync 1, 256;" ::: "memory");
            float* red = reinterpret_cast<float*>(osm);
            red[et * 2] = bn_s; red[et * 2 + 1] = bn_q;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (sh == 0) {
                p.bn_partials[((size_t)blockIdx.x * 2 + 0) * 64 + sc] = (red[sc * 2] + red[(64 + sc) * 2]) + (red[(128 + sc) * 2] + red[(192 + sc) * 2]);
                p.bn_partials[((size_t)blockIdx.x * 2 + 1) * 64 + sc] = (red[sc * 2 + 1] + red[(64 + sc) * 2 + 1]) + (red[(128 + sc) * 2 + 1] + red[(192 + sc) * 2 + 1]);
            }
        }
    }
    cfence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

// ============================================================================================================================
// weight gradient
// ============================================================================================================================
constexpr int GK = 64;                   // pixels per k-block
constexpr int GBOXR = 72;                // input rows per box: GK + 2 halo, rounded up to 8
constexpr int GBOXB = 10240;             // 72 * 128 = 9216, padded to a multiple of 1024
constexpr int GSLOTS = 12;               // ring of input row boxes (even; deep enough to hide TMA latency + the operand transform: a pair lasts ~1.3 us)
constexpr int GDY = 8;                   // ring of dy boxes (8 KB each)
constexpr int kWgradSmem = GSLOTS * GBOXB + GDY * 8192 + 1024 + 512;          // data + alignment slack + 53 mbarriers and the TMEM slot
constexpr int GSEG = 32;

constexpr int GTAPS = 12;                // partial-sum slots per CTA: taps 0..8 + a second slot for the three dw = +1 taps (9 + dh + 1)
struct WgradP {
    const float* in_scale;               // XF: the `in` operand is relu(scale * in + shift), applied to the landed boxes by the (otherwise idle) epilogue warps
    const float* in_shift;
    float* partials;                     // [gridDim.x][GTAPS][64 ci][64 o]
    int B, H, W, tiles_w, nseg_h;
    long long nunits;
};

// TMEM: 7 regions of 64 columns (see the MMA warp and the epilogue for the tap each lane half of a region holds)
template <bool XF>
__global__ void __launch_bounds__(kConvThreads) conv3x3_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmDy,
                                                                      WgradP p) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    unsigned char* ring = smem;
    unsigned char* dyr = smem + GSLOTS * GBOXB;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + GSLOTS * GBOXB + GDY * 8192);
    uint64_t* empty = full + GSLOTS;
    uint64_t* dfull = empty + GSLOTS;
    uint64_t* dempty = dfull + GDY;
    uint64_t* done = dempty + GDY;
    uint64_t* ready = done + 1;          // [GSLOTS] XF: box transformed
    uint32_t* slot_tmem = reinterpret_cast<uint32_t*>(ready + GSLOTS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmIn) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDy) : "memory");
        for (int s = 0; s < GSLOTS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&ready[s], 4); }
        for (int s = 0; s < GDY; ++s) { mbar_init(&dfull[s], 1); mbar_init(&dempty[s], 1); }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    cfence_before();
    __syncthreads();
    cfence_after();
    const uint32_t tmem = *slot_tmem;
    if (warp >= 2) {                                            // all accumulators start at zero: every UMMA accumulates, whatever pair shape comes first
        const uint32_t zero[1] = {0u};
        for (int c = 0; c < 448; c += 32) {
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
                ::"r"(tmem + ((uint32_t)((warp & 3) * 32) << 16) + c), "r"(zero[0]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    cfence_before();
    __syncthreads();
    cfence_after();

    if (warp == 0) {
        if (lane == 0) {
            int s = 1, sd = 0;                                  // odd start: rows (h, h+1) of a pair always sit in slots (even, even + 1)
            uint32_t ephase = 0, dephase = 0;
            for (long long u = blockIdx.x; u < p.nunits; u += gridDim.x) {
                const int hs = (int)(u % p.nseg_h);
                const long long r = u / p.nseg_h;
                const int wt = (int)(r % p.tiles_w), b = (int)(r / p.tiles_w);
                const int h0 = hs * GSEG, h1 = min(p.H, h0 + GSEG), w0 = wt * GK;
                const int npair = (h1 - h0 + 1) >> 1;
                // input rows h0-1 .. h0+2*npair (rows outside the image are zero-filled); the two dy rows of a pair follow its last input row.
                // A trailing half pair still loads its (unused) second dy row and fourth input row so that the slot accounting is uniform.
                for (int row = h0 - 1; row <= h0 + 2 * npair; ++row) {
                    mbar_wait(&empty[s], ((ephase >> s) & 1u) ^ 1u);
                    ephase ^= 1u << s;
                    mbar_expect_tx(&full[s], GBOXR * 128);
                    ctma_load_4d(ring + s * GBOXB, &tmIn, 0, w0 - 1, row, b, &full[s]);
                    if (++s == GSLOTS) s = 0;
                    const int rel = row - h0;                   // after input row h+2 of the pair (h, h+1): dy rows h, h+1
                    if (rel >= 2 && (rel & 1) == 0) {
#pragma unroll 1
                        for (int rp = 0; rp < 2; ++rp) {
                            mbar_wait(&dempty[sd], ((dephase >> sd) & 1u) ^ 1u);
                            dephase ^= 1u << sd;
                            mbar_expect_tx(&dfull[sd], 8192);
                            ctma_load_4d(dyr + sd * 8192, &tmDy, 0, w0, row - 2 + rp, b, &dfull[sd]);
                            if (++sd == GDY) sd = 0;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        {
            // A (input taps) MN-major M = 128: two 64-channel groups LBO (one pixel row) apart; B (dy) MN-major, N = 64 or two rows N = 128
            constexpr uint32_t idesc64 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
            constexpr uint32_t idesc128 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
            constexpr uint32_t hi = cdesc_hi(1024);
            const uint32_t ring_lo = cdesc_lo(smem_u32(ring), 128), dy_lo0 = cdesc_lo(smem_u32(dyr), 8192);
            const uint32_t stk_lo = cdesc_lo(smem_u32(ring), GBOXB);      // vertically stacked operand: the second 64-channel group is the next ring slot
            int slot0 = 1, sd = 0;
            uint32_t fphase = 0, dphase = 0;
            for (long long u = blockIdx.x; u < p.nunits; u += gridDim.x) {
                const int hs = (int)(u % p.nseg_h);
                const int h0 = hs * GSEG, h1 = min(p.H, h0 + GSEG);
                const int npair = (h1 - h0 + 1) >> 1;
                for (int t = 0; t < npair; ++t) {
                    int sl[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) { sl[j] = slot0 + j; if (sl[j] >= GSLOTS) sl[j] -= GSLOTS; }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {               // (unrolled: a run-time index would push sl[] into local memory)
                        if (j < 2 && t != 0) continue;
                        mbar_wait(XF ? &ready[sl[j]] : &full[sl[j]], (fphase >> sl[j]) & 1u);
                        fphase ^= 1u << sl[j];
                    }
#pragma unroll
                    for (int rp = 0; rp < 2; ++rp) {            // dy rows h, h+1 = slots sd, sd+1 (sd is even: never wraps inside a pair)
                        mbar_wait(&dfull[sd + rp], (dphase >> (sd + rp)) & 1u);
                        dphase ^= 1u << (sd + rp);
                    }
                    cfence_after();
                    const uint32_t dy_lo = dy_lo0 + (uint32_t)sd * (8192 >> 4);
                    const uint32_t a0 = ring_lo + (uint32_t)sl[0] * (GBOXB >> 4), a1 = ring_lo + (uint32_t)sl[1] * (GBOXB >> 4),
                                   a2 = ring_lo + (uint32_t)sl[2] * (GBOXB >> 4), a3 = ring_lo + (uint32_t)sl[3] * (GBOXB >> 4);
                    const uint32_t a12 = stk_lo + (uint32_t)sl[1] * (GBOXB >> 4) + 16;        // [in(h) | in(h+1)] at dw = +1 (sl[1] is even: sl[2] = sl[1] + 1)
                    const bool last = t == npair - 1;
                    const bool full_pair = h0 + 2 * t + 1 < h1;  // a trailing half pair has only output row h
                    // TMEM columns.  Taps dw = -1 / 0 (lanes 0-63 / 64-127): [dh=+1 | dh=0 | dh=-1] at 0 / 64 / 128, so an N = 128 instruction whose B is
                    // [dy(h) | dy(h+1)] lands on (dh, dh-1).  Tap dw = +1: X at 192 (lanes 0-63 = in(h): [dh=0 | dh=-1], lanes 64-127 = in(h+1):
                    // [dh=+1 | dh=0]), Y1 at 320 (in(h-1) x dy(h): dh=-1), Y2 at 384 (in(h+2) x dy(h+1): dh=+1).  Everything accumulates (zeroed above).
                    if (celect_one()) {
                        if (full_pair) {
                            // rows h-1 and h die with this pair: everything that reads them first, then their slots are released half a pair early
#pragma unroll
                            for (int k = 0; k < 4; ++k)         // input row h-1: dh = -1 for dy(h)
                                cumma2(tmem + 128, a0 + k * 128, hi, dy_lo + k * 128, hi, idesc64, 1u);
#pragma unroll
                            for (int k = 0; k < 4; ++k)         // dw = +1, row h-1 (upper lanes unused)
                                cumma2(tmem + 320, a0 + 16 + k * 128, hi, dy_lo + k * 128, hi, idesc64, 1u);
#pragma unroll
                            for (int k = 0; k < 4; ++k)         // input row h:   dh = 0 for dy(h), dh = -1 for dy(h+1)
                                cumma2(tmem + 64, a1 + k * 128, hi, dy_lo + k * 128, hi, idesc128, 1u);
#pragma unroll
                            for (int k = 0; k < 4; ++k)         // dw = +1, rows h | h+1 stacked against both dy rows
                                cumma2(tmem + 192, a12 + k * 128, hi, dy_lo + k * 128, hi, idesc128, 1u);
                            ccommit(&empty[sl[0]]);
                            ccommit(&empty[sl[1]]);
#pragma unroll
                            for (int k = 0; k < 4; ++k)         // input row h+1: dh = +1 for dy(h), dh = 0 for dy(h+1)
                                cumma2(tmem, a2 + k * 128, hi, dy_lo + k * 128, hi, idesc128, 1u);
#pragma unroll
                            for (int k = 0; k < 4; ++k)         // input row h+2: dh = +1 for dy(h+1)
                                cumma2(tmem, a3 + k * 128, hi, dy_lo + 512 + k * 128, hi, idesc64, 1u);
#pragma unroll
                            for (int k = 0; k < 4; ++k)         // dw = +1, row h+2 (upper lanes unused)
                                cumma2(tmem + 384, a3 + 16 + k * 128, hi, dy_lo + 512 + k * 128, hi, idesc64, 1u);
                        } else {
#pragma unroll
                            for (int dh = 0; dh < 3; ++dh) {                      // single output row: input rows h-1, h, h+1 against dy(h)
                                const uint32_t a = dh == 0 ? a0 : (dh == 1 ? a1 : a2);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    cumma2(tmem + (2 - dh) * 64, a + k * 128, hi, dy_lo + k * 128, hi, idesc64, 1u);
                            }
#pragma unroll
                            for (int k = 0; k < 4; ++k)         // dw = +1: rows h | h+1 stacked against dy(h) -> (dh = 0 | dh = +1) in X's first column block
                                cumma2(tmem + 192, a12 + k * 128, hi, dy_lo + k * 128, hi, idesc64, 1u);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                cumma2(tmem + 320, a0 + 16 + k * 128, hi, dy_lo + k * 128, hi, idesc64, 1u);
                        }
                        if (!full_pair) { ccommit(&empty[sl[0]]); ccommit(&empty[sl[1]]); }
                        if (last) { ccommit(&empty[sl[2]]); ccommit(&empty[sl[3]]); }
                        ccommit(&dempty[sd]);
                        ccommit(&dempty[sd + 1]);
                    }
                    __syncwarp();
                    slot0 += last ? 4 : 2; if (slot0 >= GSLOTS) slot0 -= GSLOTS;
                    sd += 2; if (sd >= GDY) sd = 0;
                }
            }
            if (celect_one()) ccommit(done);
            __syncwarp();
        }
    }
    // every CTA writes its partial (zeros if it had no work: the accumulators were zeroed) so the reduction can read a fixed number of partials
    const bool has_work = (long long)blockIdx.x < p.nunits;
    if (warp >= 2) {
        const int q = warp & 3;
        float* dst = p.partials + (size_t)blockIdx.x * GTAPS * 64 * 64;
        if (XF) {                                               // main loop: these warps are the operand transform (they only drain TMEM at the very end)
            const int t = threadIdx.x - 64;
            XfConst xc;
            xc.load(p.in_scale, p.in_shift, t);
            int s = 1;
            uint32_t fph = 0;
            for (long long u = blockIdx.x; u < p.nunits; u += gridDim.x) {
                const int hs = (int)(u % p.nseg_h);
                const long long r = u / p.nseg_h;
                const int wt = (int)(r % p.tiles_w);
                const int h0 = hs * GSEG, h1 = min(p.H, h0 + GSEG), w0 = wt * GK;
                const int npair = (h1 - h0 + 1) >> 1;
                for (int row = h0 - 1; row <= h0 + 2 * npair; ++row) {
                    mbar_wait(&full[s], (fph >> s) & 1u);
                    fph ^= 1u << s;
                    if (row >= 0 && row < p.H) xf_box<(GBOXR + 15) / 16>(ring + s * GBOXB, GBOXR, w0 - 1, p.W, t, xc);
                    xf_publish(&ready[s], lane);
                    if (++s == GSLOTS) s = 0;
                }
            }
        }
        if (has_work) {
            mbar_wait(done, 0);
            cfence_after();
        }
        const int l = q * 32 + lane, up = l >> 6, ci = l & 63;
        // region r at TMEM columns r*64 -> partial slot of this lane (-1: nothing useful in this lane)
        //   0..2: taps dw = -1 (lanes 0-63) / dw = 0 (lanes 64-127), dh + 1 = 2 - r
        //   3: X first block  (in(h) x dy(h): dh 0 | in(h+1) x dy(h): dh +1)        4: X second block (in(h) x dy(h+1): dh -1 | in(h+1) x dy(h+1): dh 0, 2nd slot)
        //   5: Y1 (in(h-1) x dy(h): dh -1, 2nd slot)                                 6: Y2 (in(h+2) x dy(h+1): dh +1, 2nd slot)
        for (int r = 0; r < 7; ++r) {
            int slot;
            if (r < 3) slot = (2 - r) * 3 + up;
            else if (r == 3) slot = up ? 8 : 5;
            else if (r == 4) slot = up ? 10 : 2;
            else if (r == 5) slot = up ? -1 : 9;
            else slot = up ? -1 : 11;
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 32) {
                uint32_t v[32];
                ctmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + r * 64 + c0, v);
                if (slot >= 0) {
                    float* o = dst + ((size_t)slot * 64 + ci) * 64 + c0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(o + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                }
            }
        }
    }
    cfence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// partials [nparts][GTAPS][ci][o] -> dWp[o][tap][ci] (+)=; slot 9 + dh + 1 holds the second contribution to tap (dh, dw = +1)
__global__ void conv_wgrad_reduce_kernel(const float* __restrict__ partials, int nparts, float* __restrict__ out, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // index into [tap][ci][o]
    if (i >= 576 * 64) return;
    const int o = i & 63, tc = i >> 6, tap = tc >> 6, ci = tc & 63;
    const int second = (tap % 3 == 2) ? (((9 + tap / 3) * 64 + ci) * 64 + o) : -1;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) {
        const float* base = partials + (size_t)p * GTAPS * 64 * 64;
        s += (double)base[i];
        if (second >= 0) s += (double)base[second];
    }
    float* d = out + (size_t)o * 576 + tc;
    *d = accumulate ? *d + (float)s : (float)s;
}

typedef CUresult (*CEncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                           CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CEncFn cenc() {
    static CEncFn fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<CEncFn>(f);
    }
    return fn;
}

// channel-last image [B][H][W][64] bf16, box = 64 ch x box_w pixels of one row
static int image_map(CUtensorMap* m, const void* base, int B, int H, int W, int box_w) {
    CEncFn fn = cenc();
    if (!fn) { set_last_error("cuTensorMapEncodeTiled unavailable"); return SARSSL_ERR_UNSUPPORTED; }
    cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
    cuuint32_t box[4] = {64, (cuuint32_t)box_w, 1, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_last_error("conv tensor map failed (%d) B=%d H=%d W=%d box_w=%d", (int)r, B, H, W, box_w); return SARSSL_ERR_ARG; }
    return SARSSL_OK;
}

}  // namespace sarssl

using namespace sarssl;

extern "C" int sarssl_conv3x3_tc_grid(int B, int H, int W) {
    const long long nunits = (long long)B * ((W + CW - 1) / CW) * ((H + CSEG - 1) / CSEG);
    return (int)(nunits < sm_count() ? nunits : sm_count());
}

extern "C" int sarssl_conv3x3_tc(const void* in, const void* weight_packed, void* out, float* bn_partials, const float* in_scale, const float* in_shift,
                                 int B, int H, int W, cudaStream_t stream) {
    SARSSL_CHECK_ARG(in && weight_packed && out && B > 0 && H > 0 && W > 0, "conv3x3_tc: bad arguments");
    SARSSL_CHECK_ARG(aligned16(in) && aligned16(weight_packed) && aligned16(out), "conv3x3_tc: buffers must be 16-byte aligned");
    CUtensorMap mi, mw, mo;
    int rc;
    if ((rc = image_map(&mi, in, B, H, W, CBOXR))) return rc;
    if ((rc = image_map(&mo, out, B, H, W, CW))) return rc;
    {
        CEncFn fn = cenc();
        cuuint64_t dims[2] = {576, 64}, strides[1] = {576 * 2};
        cuuint32_t box[2] = {64, 64}, es[2] = {1, 1};
        if (fn(&mw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(weight_packed), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            set_last_error("conv3x3_tc: weight tensor map failed");
            return SARSSL_ERR_ARG;
        }
    }
    SARSSL_CHECK_ARG((in_scale == nullptr) == (in_shift == nullptr), "conv3x3_tc: in_scale and in_shift come together");
    static bool configured = false;
    if (!configured) {
        SARSSL_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvSmem));
        SARSSL_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvSmem));
        configured = true;
    }
    ConvP p;
    p.in_scale = in_scale; p.in_shift = in_shift;
    p.out = static_cast<__nv_bfloat16*>(out); p.bn_partials = bn_partials; p.B = B; p.H = H; p.W = W;
    p.tiles_w = (W + CW - 1) / CW; p.nseg_h = (H + CSEG - 1) / CSEG;
    p.nunits = (long long)B * p.tiles_w * p.nseg_h;
    const int grid = sarssl_conv3x3_tc_grid(B, H, W);
    if (in_scale) conv3x3_tc_kernel<true><<<grid, kConvXfThreads, kConvSmem, stream>>>(mi, mw, mo, p);
    else conv3x3_tc_kernel<false><<<grid, kConvFwdThreads, kConvSmem, stream>>>(mi, mw, mo, p);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" size_t sarssl_conv3x3_wgrad_tc_workspace_bytes(void) { return (size_t)sm_count() * GTAPS * 64 * 64 * sizeof(float); }

extern "C" int sarssl_conv3x3_wgrad_tc(const void* dy, const void* in, const float* in_scale, const float* in_shift, float* dweight_packed, int accumulate,
                                       int B, int H, int W, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    SARSSL_CHECK_ARG(dy && in && dweight_packed && workspace && B > 0 && H > 0 && W > 0, "conv3x3_wgrad_tc: bad arguments");
    if (workspace_bytes < sarssl_conv3x3_wgrad_tc_workspace_bytes()) { set_last_error("conv3x3_wgrad_tc: workspace too small"); return SARSSL_ERR_WORKSPACE; }
    CUtensorMap mi, md;
    int rc;
    if ((rc = image_map(&mi, in, B, H, W, GBOXR))) return rc;
    if ((rc = image_map(&md, dy, B, H, W, GK))) return rc;
    SARSSL_CHECK_ARG((in_scale == nullptr) == (in_shift == nullptr), "conv3x3_wgrad_tc: in_scale and in_shift come together");
    static bool configured = false;
    if (!configured) {
        SARSSL_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgradSmem));
        SARSSL_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgradSmem));
        configured = true;
    }
    WgradP p;
    p.in_scale = in_scale; p.in_shift = in_shift;
    p.partials = static_cast<float*>(workspace); p.B = B; p.H = H; p.W = W;
    p.tiles_w = (W + GK - 1) / GK; p.nseg_h = (H + GSEG - 1) / GSEG;
    p.nunits = (long long)B * p.tiles_w * p.nseg_h;
    const int grid = (int)(p.nunits < sm_count() ? p.nunits : sm_count());
    if (in_scale) conv3x3_wgrad_tc_kernel<true><<<grid, kConvThreads, kWgradSmem, stream>>>(mi, md, p);
    else conv3x3_wgrad_tc_kernel<false><<<grid, kConvThreads, kWgradSmem, stream>>>(mi, md, p);
    SARSSL_LAUNCH_CHECK();
    conv_wgrad_reduce_kernel<<<(576 * 64 + 255) / 256, 256, 0, stream>>>(p.partials, grid, dweight_packed, accumulate);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}
