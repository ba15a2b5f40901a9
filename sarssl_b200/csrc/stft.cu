// Multi-channel STFT front-end for sm_100a.
//
// Replaces (reference, paths under /root/reference/code):
//   STFT.forward                      common/utils_module.py:49-72    -> sarssl_stft_spectrum
//   STFTLearner.data_preprocess       learner.py:525-572 (+ AddChToBatch utils_module.py:124-148)
//                                                                      -> sarssl_stft_frontend
// Both are HBM-bound (about 6 MFLOP per 1.5 MB clip).  Design:
//   * frames are 512 samples every 256; the two channels of a pair are interleaved in memory exactly like
//     one complex sample, so a frame block is staged into shared memory with ONE 1-D bulk TMA copy
//     (cp.async.bulk + mbarrier) and ONE complex 512-point FFT (fft512.cuh, 64 lanes) gives both spectra;
//   * output is written frame-major ("patch layout" [clip][frame][bin][re/im][mic], 4 KB contiguous per
//     frame) - the layout PatchSplit would produce (utils_module.py:196-205) and the one the encoder stem and
//     the loss kernel consume - with 16-byte streaming stores;
//   * the per-clip normalisation (divide by mean |X_ch0| over all 257 x nt bins) needs a clip-wide reduction
//     before anything can be written.  The fused kernel is persistent: CTAs pull (clip, frame-block) items
//     from an atomic queue, keep their un-scaled spectra in shared memory, publish a partial magnitude sum,
//     rendezvous on a per-clip arrival counter in global memory, then scale and write.  HBM traffic is the
//     algorithmic minimum (read the waveform once, write the spectra once).  Items are handed out in order,
//     so all peers of a clip are resident whenever grid >= items-per-clip (checked on the host; otherwise the
//     generic three-kernel path below is used).
//   * generic path (any channel count, any length): spectrum kernel -> clip scale kernel -> pair/normalise.
#include "common.cuh"
#include "fft512.cuh"
#include "fft512w.cuh"
#include "../../include/sarssl_b200.h"
#include <atomic>

namespace sarssl {

constexpr int kFPI = 8;            // frames per work item
constexpr int kGroups = 4;         // 64-lane FFT groups per CTA
constexpr int kThreads = kGroups * kFftLanes;
constexpr int kHop = 256;
constexpr int kBins = 257;

__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(kFftLanes) : "memory"); }

// One frame: windowed load from the staged samples, 3 radix-8 passes, split.  On return v[] holds Z and the
// group's scratch holds Z for split_bin(); the caller must group_bar() again before reusing the scratch.
__device__ __forceinline__ void fft_frame(const float2* frame, const FftLane& c, float* sre, float* sim, int l, int g,
                                          float2 (&v)[8]) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        float2 s = frame[l + 64 * r];
        v[r] = make_float2(s.x * c.win[r], s.y * c.win[r]);
    }
    fft_pass1(v, c, sre, sim, l);
    group_bar(g);
    fft_pass2_load(v, sre, sim, l);
    group_bar(g);
    fft_pass2_store(v, c, sre, sim, l);
    group_bar(g);
    fft_pass3(v, sre, sim, l);
    group_bar(g);
    split_store(v, sre, sim, l);
    group_bar(g);
}

struct FusedSmem {
    float2 in[(kFPI + 1) * kHop];                 // 18 KB   staged samples (ch0, ch1)
    float4 out[kFPI * kHop];                      // 32 KB   un-scaled bins 1..256, (re0, re1, im0, im1)
    float scratch[kGroups][kFftScratchFloats];    // 18 KB
    float red[32];
    uint64_t bar;
    int item;
    float scale;
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// counters[0] = work queue head, counters[1] = error flag, counters[2 + b] = arrivals for clip b
__global__ void __launch_bounds__(kThreads) stft_frontend_fused_kernel(const float* __restrict__ sig, float4* __restrict__ out,
                                                                     float* partials, unsigned* counters, int nb,
                                                                     long long nsample, int nt, int ipc, float eps) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FusedSmem& sm = *reinterpret_cast<FusedSmem*>(smem_raw);
    const int tid = threadIdx.x, g = tid >> 6, l = tid & 63;
    FftLane lane;
    lane.init(l);
    float* sre = sm.scratch[g];
    float* sim = sre + kFftPlane;
    if (tid == 0) {
        mbar_init(&sm.bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t phase = 0;
    const int nitems = nb * ipc;
    for (;;) {
        if (tid == 0) sm.item = (int)atomicAdd(&counters[0], 1u);
        __syncthreads();
        const int item = sm.item;
        if (item >= nitems) break;
        const int b = item / ipc, fb = item - b * ipc;
        const int f0 = fb * kFPI;
        const int nfr = min(kFPI, nt - f0);
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)(nfr + 1) * kHop * sizeof(float2);
            mbar_expect_tx(&sm.bar, bytes);
            tma_bulk_g2s(sm.in, sig + ((size_t)b * nsample + (size_t)f0 * kHop) * 2, bytes, &sm.bar);
        }
        mbar_wait(&sm.bar, phase);
        phase ^= 1u;

        float part = 0.f;
        for (int fr = g; fr < nfr; fr += kGroups) {
            float2 v[8];
            fft_frame(sm.in + fr * kHop, lane, sre, sim, l, g, v);
            float4* orow = sm.out + fr * kHop;
#pragma unroll
            for (int k3 = 0; k3 < 4; ++k3) {
                const float4 o = split_bin(v, sre, sim, l, k3);
                part += sqrtf(o.x * o.x + o.z * o.z);
                const int k = l + 64 * k3;
                if (k >= 1) orow[k - 1] = o;
            }
            if (l == 0) {
                const float4 o = split_bin(v, sre, sim, 0, 4);
                part += sqrtf(o.x * o.x + o.z * o.z);
                orow[255] = o;
            }
            group_bar(g);
        }
        const float total = block_sum(part, sm.red);
        if (tid == 0) {
            partials[(size_t)b * ipc + fb] = total;
            __threadfence();
            atomicAdd(&counters[2 + b], 1u);
            unsigned spins = 0;
            while (ld_acquire_u32(&counters[2 + b]) < (unsigned)ipc) {
                __nanosleep(64);
                if (++spins > (1u << 24)) {          // ~ seconds: never expected; report instead of hanging
                    atomicExch(&counters[1], 1u);
                    break;
                }
            }
        }
        __syncthreads();
        if (tid < 32) {                              // fixed-order (deterministic) sum of the clip's partials
            float s = 0.f;
            for (int i = tid; i < ipc; i += 32) s += __ldcg(&partials[(size_t)b * ipc + i]);
            s = warp_sum(s);
            if (tid == 0) sm.scale = 1.0f / (s / (float)((long long)kBins * nt) + eps);
        }
        __syncthreads();
        const float scale = sm.scale;
        float4* dst = out + ((size_t)b * nt + f0) * kHop;
        for (int i = tid; i < nfr * kHop; i += kThreads) {
            float4 o = sm.out[i];
            o.x *= scale; o.y *= scale; o.z *= scale; o.w *= scale;
            st_stream_f4(dst + i, o);
        }
        __syncthreads();
    }
}

// ---------------- fused kernel, independent warps (default) ---------------------------------------------------------------------
// Every warp of the persistent grid is its own worker with its own transform (fft512w.cuh: 16 points per lane, one shared-memory
// transpose, no block-level barrier after start-up): frame f = warp_id + r * warps of the clip-major frame list in round r.
// The per-clip normalisation needs the clip's nt partial sums of |X_ch0| before anything can be written.  A warp that waited for its
// clip right after publishing would meet the clip's other 256 warps after every frame and the SM would alternate between a compute
// phase and a store phase (measured: 1.3 ms).  Instead the wait is software-pipelined kLag rounds deep: a warp transforms frame r,
// publishes its partial sum, parks the frame's un-scaled bins in TENSOR MEMORY (tcgen05.st: the SM's 256 KB of TMEM are otherwise idle
// in this kernel; 8 warps x 4 frames x 4 KB = 128 KB per CTA, two CTAs per SM) and only then finishes frame r - kLag (tcgen05.ld, scale,
// streaming store).  With one round of slack (bins parked in shared memory: all that fits there) the clip's slowest warp still paced
// all 257 of them (0.454 ms, 85 % of the polls had to wait); three rounds absorb the jitter.
// The rendezvous carries no fence and no atomic: a partial sum is published as ONE 64-bit store {sum, launch epoch}; a clip is complete
// when all nt of its words carry this launch's epoch (L2-coherent loads; the words are the only data exchanged, so nothing else has to
// be ordered).  (__threadfence + atomicAdd + ld.acquire cost 48 % of the stall samples of the first version: profiles/r02_stft_frontend.txt.)
// One 4 KB input buffer per warp: the next frame's bulk-TMA copy is issued as soon as this frame's samples are in registers.
// Deadlock-free when every warp is resident and warps >= nt - 1: a warp publishes round r before it waits for an earlier round, and a
// clip that holds a round r - 1 frame ends before frame (r + 1) * warps.
constexpr int kW2Warps = 8;
constexpr int kLag = 3;                           // rounds between publishing a frame and storing it; kLag + 1 TMEM slots per warp
struct Warp2Smem {
    float2 in[kW2Warps][kFftN];                   // 32 KB   per-warp staged samples (ch0, ch1)
    float2 tb[kW2Warps][kWTransFloat2];           // 33 KB   per-warp transpose buffer, then the frame's Z in bin order
    float2 win2[kFftN];                           // 4 KB    periodic Hann window as (w, w) pairs (one packed multiply per sample pair)
    uint64_t bar[kW2Warps];
    uint32_t tmem_slot;
};

__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ unsigned long long ld_cg_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// 32 lanes x 32 columns (one frame's 256 bins x float4, lane l holding bins 1 + l + 32 r) <-> tensor memory
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float4 (&o)[8]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "f"(o[0].x), "f"(o[0].y), "f"(o[0].z), "f"(o[0].w), "f"(o[1].x), "f"(o[1].y), "f"(o[1].z), "f"(o[1].w), "f"(o[2].x), "f"(o[2].y),
          "f"(o[2].z), "f"(o[2].w), "f"(o[3].x), "f"(o[3].y), "f"(o[3].z), "f"(o[3].w), "f"(o[4].x), "f"(o[4].y), "f"(o[4].z), "f"(o[4].w), "f"(o[5].x),
          "f"(o[5].y), "f"(o[5].z), "f"(o[5].w), "f"(o[6].x), "f"(o[6].y), "f"(o[6].z), "f"(o[6].w), "f"(o[7].x), "f"(o[7].y), "f"(o[7].z), "f"(o[7].w)
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float4 (&o)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=f"(o[0].x), "=f"(o[0].y), "=f"(o[0].z), "=f"(o[0].w), "=f"(o[1].x), "=f"(o[1].y), "=f"(o[1].z), "=f"(o[1].w), "=f"(o[2].x), "=f"(o[2].y),
          "=f"(o[2].z), "=f"(o[2].w), "=f"(o[3].x), "=f"(o[3].y), "=f"(o[3].z), "=f"(o[3].w), "=f"(o[4].x), "=f"(o[4].y), "=f"(o[4].z), "=f"(o[4].w),
          "=f"(o[5].x), "=f"(o[5].y), "=f"(o[5].z), "=f"(o[5].w), "=f"(o[6].x), "=f"(o[6].y), "=f"(o[6].z), "=f"(o[6].w), "=f"(o[7].x), "=f"(o[7].y),
          "=f"(o[7].z), "=f"(o[7].w)
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kW2Warps * 32, 2) stft_frontend_warp2_kernel(const float* __restrict__ sig, float4* __restrict__ out,
                                                                              unsigned long long* partials, unsigned* counters, unsigned epoch, int nb,
                                                                              long long nsample, int nt, float eps) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Warp2Smem& sm = *reinterpret_cast<Warp2Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    WarpFftLane lc;
    lc.init(lane);
    for (int i = tid; i < kFftN; i += kW2Warps * 32) {
        const float w = 0.5f - 0.5f * cospif((float)i / 256.0f);
        sm.win2[i] = make_float2(w, w);
    }
    if (lane == 0) {
        mbar_init(&sm.bar[warp], 1);
        mbar_fence_init();
    }
    if (warp == 0) {                                                 // half of the SM's tensor memory (two CTAs per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // a warp reaches the 32 TMEM lanes of its quarter (warp % 4); the two warps of a quarter split the CTA's 256 columns
    const uint32_t tm = sm.tmem_slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 128;
    const unsigned total = (unsigned)nb * (unsigned)nt, unt = (unsigned)nt;
    const unsigned gw = blockIdx.x * kW2Warps + warp, G = gridDim.x * kW2Warps;
    const unsigned db = G / unt, dt = G - db * unt;                  // frame f + G = (b + db, t + dt) with one carry
    const float inv_bins = 1.0f / (float)((long long)kBins * nt);
    const float2 sgn = lane < 16 ? make_float2(1.f, 1.f) : make_float2(-1.f, -1.f);
    float2* in = sm.in[warp];
    float2* tb = sm.tb[warp];
    auto issue = [&](unsigned b, unsigned t) {
        if (lane == 0) {
            mbar_expect_tx(&sm.bar[warp], kFftN * sizeof(float2));
            tma_bulk_g2s(in, sig + ((size_t)b * nsample + (size_t)t * kHop) * 2, kFftN * sizeof(float2), &sm.bar[warp]);
        }
    };
    // scale and store the frame parked in TMEM slot `slot`, once its clip is complete
    auto finish = [&](unsigned b, unsigned t, int slot) {
        const unsigned long long* pp = partials + (size_t)b * unt;
        float s;
        for (unsigned spins = 0;; ++spins) {
            s = 0.f;
            bool ok = true;
            for (unsigned i = lane; i < unt; i += 32) {              // fixed order: deterministic
                const unsigned long long w = ld_cg_u64(pp + i);
                ok = ok && (unsigned)(w >> 32) == epoch;
                s += __uint_as_float((unsigned)w);
            }
            if (__all_sync(0xffffffffu, ok)) break;
            __nanosleep(64);
            if (spins > (1u << 22)) { if (lane == 0) atomicExch(&counters[1], 1u); break; }      // never expected; report instead of hanging
        }
        s = warp_sum(s);
        // the parked bins are 2 X and the partial sums 2 sum|X_ch0| (the split's factor 1/2 is folded in here): out = 2 X * scale
        const float scale = 0.5f / (0.5f * s * inv_bins + eps);
        const float2 sc2 = make_float2(scale, scale);
        float4 o[8];
        tmem_ld32(tm + (uint32_t)slot * 32, o);
        float4* dst = out + ((size_t)b * unt + t) * kHop + lane;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float2 lo = cscale2(make_float2(o[r].x, o[r].y), sc2), hi = cscale2(make_float2(o[r].z, o[r].w), sc2);
            st_stream_f4(dst + 32 * r, make_float4(lo.x, lo.y, hi.x, hi.y));
        }
    };
    uint32_t ph = 0;
    unsigned b = gw / unt, t = gw - b * unt;
    unsigned hb = b, ht = t;                                         // oldest parked frame (frames of one warp advance by (db, dt) per round)
    int nheld = 0, wslot = 0, rslot = 0;
    if (gw < total) issue(b, t);
    for (unsigned f = gw; f < total; f += G) {
        mbar_wait(&sm.bar[warp], ph);
        ph ^= 1u;
        float2 v[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) v[n1] = cscale2(in[32 * n1 + lane], sm.win2[32 * n1 + lane]);
        __syncwarp();                                                // every lane has its samples: the buffer can take the next frame
        unsigned nb_ = b + db, nt_ = t + dt;
        if (nt_ >= unt) { nt_ -= unt; ++nb_; }
        if (f + G < total) issue(nb_, nt_);
        wfft_stage1(v, lc, tb, lane);
        __syncwarp();
        wfft_stage2(v, tb, lane);
        {
            float2 o[16];
#pragma unroll
            for (int s = 0; s < 16; ++s) o[s] = make_float2(__shfl_xor_sync(0xffffffffu, v[s].x, 16), __shfl_xor_sync(0xffffffffu, v[s].y, 16));
            __syncwarp();                                            // every lane's transpose reads are done before the buffer is re-written
            wfft_store_z(v, o, sgn, tb, lane);
        }
        __syncwarp();
        float4 o[8];
        float part = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int k = 1 + lane + 32 * r;
            o[r] = wfft_split2(tb[wz_pos(k)], tb[wz_pos(512 - k)]);
            part += sqrt_approx(o[r].x * o[r].x + o[r].z * o[r].z);
        }
        {   // bin 0 (DC) is dropped from the patches but counts in the mean magnitude: X_ch0[0] = Re Z[0]
            const float z0 = tb[0].x;
            if (lane == 0) part += 2.f * fabsf(z0);
        }
        __syncwarp();                                                // the Z buffer is read: the next frame's transpose may overwrite it
        part = warp_sum(part);
        if (lane == 0) __stcg(&partials[(size_t)b * unt + t], ((unsigned long long)epoch << 32) | __float_as_uint(part));
        tmem_st32(tm + (uint32_t)wslot * 32, o);
        wslot = (wslot + 1) & kLag;
        if (++nheld > kLag) {
            finish(hb, ht, rslot);
            rslot = (rslot + 1) & kLag;
            --nheld;
            hb += db; ht += dt;
            if (ht >= unt) { ht -= unt; ++hb; }
        }
        b = nb_; t = nt_;
    }
    while (nheld > 0) {
        finish(hb, ht, rslot);
        rslot = (rslot + 1) & kLag;
        --nheld;
        hb += db; ht += dt;
        if (ht >= unt) { ht -= unt; ++hb; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sm.tmem_slot), "r"(256) : "memory");
}

// ---------------- generic path ----------------
// grid.x = nb * ipc * npair.  Pair p covers channels (2p, 2p+1) (second one zero if absent).
// spec: complex64 [nb][nt][257][nch] (interleaved re, im);  partials: |X_ch0| sums per (clip, frame block).
__global__ void __launch_bounds__(kThreads) stft_spectrum_kernel(const float* __restrict__ sig, float2* __restrict__ spec,
                                                               float* __restrict__ partials, int nb, long long nsample, int nch,
                                                               int nt, int ipc, int npair) {
    __shared__ __align__(16) float2 in_s[(kFPI + 1) * kHop];
    __shared__ float scratch[kGroups][kFftScratchFloats];
    __shared__ float red[32];
    const int tid = threadIdx.x, g = tid >> 6, l = tid & 63;
    int item = blockIdx.x;
    const int p = item % npair; item /= npair;
    const int fb = item % ipc;
    const int b = item / ipc;
    const int f0 = fb * kFPI, nfr = min(kFPI, nt - f0);
    const int c0 = 2 * p, c1 = 2 * p + 1;
    const float* src = sig + ((size_t)b * nsample + (size_t)f0 * kHop) * nch;
    for (int i = tid; i < (nfr + 1) * kHop; i += kThreads)
        in_s[i] = make_float2(src[(size_t)i * nch + c0], c1 < nch ? src[(size_t)i * nch + c1] : 0.f);
    __syncthreads();
    FftLane lane;
    lane.init(l);
    float* sre = scratch[g];
    float* sim = sre + kFftPlane;
    float part = 0.f;
    for (int fr = g; fr < nfr; fr += kGroups) {
        float2 v[8];
        fft_frame(in_s + fr * kHop, lane, sre, sim, l, g, v);
        float2* orow = spec + ((size_t)b * nt + f0 + fr) * kBins * nch;
#pragma unroll
        for (int k3 = 0; k3 < 5; ++k3) {
            if (k3 == 4 && l != 0) break;
            const float4 o = split_bin(v, sre, sim, l, k3);
            const int k = l + 64 * k3;
            orow[(size_t)k * nch + c0] = make_float2(o.x, o.z);
            if (c1 < nch) orow[(size_t)k * nch + c1] = make_float2(o.y, o.w);
            if (p == 0) part += sqrtf(o.x * o.x + o.z * o.z);
        }
        group_bar(g);
    }
    if (partials != nullptr && p == 0) {
        const float total = block_sum(part, red);
        if (tid == 0) partials[(size_t)b * ipc + fb] = total;
    }
}

__global__ void clip_scale_kernel(const float* __restrict__ partials, float* __restrict__ scale, int nb, int ipc, int nt, float eps) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= nb) return;
    float s = 0.f;
    for (int i = lane; i < ipc; i += 32) s += partials[(size_t)b * ipc + i];
    s = warp_sum(s);
    if (lane == 0) scale[b] = 1.0f / (s / (float)((long long)kBins * nt) + eps);
}

// spec [nb][nt][257][nch] -> patches [(b*(nch-1) + j-1)][nt][256] float4 (re0, re_j, im0, im_j), bins 1..256, scaled
__global__ void pair_normalize_kernel(const float2* __restrict__ spec, const float* __restrict__ scale, float4* __restrict__ out,
                                      int nb, int nt, int nch) {
    const size_t total = (size_t)nb * (nch - 1) * nt * 256;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int f = (int)(i & 255);
        size_t r = i >> 8;
        const int t = (int)(r % nt); r /= nt;
        const int j = (int)(r % (nch - 1)) + 1;
        const int b = (int)(r / (nch - 1));
        const float s = scale[b];
        const float2* row = spec + (((size_t)b * nt + t) * kBins + (f + 1)) * nch;
        const float2 a = row[0], c = row[j];
        st_stream_f4(out + i, make_float4(a.x * s, c.x * s, a.y * s, c.y * s));
    }
}

// generalised pairing for the off-hot-path variants of data_preprocess: ch_mode 'MM' (all channel pairs a < b in AddChToBatch's order,
// utils_module.py:136-143) and fre_used_ratio 0.5 (bins 0..127, learner.py:516-517).  spec [nb][nt][257][nch] -> [(b, pair)][nt][nbins] float4
__global__ void pair_normalize_ex_kernel(const float2* __restrict__ spec, const float* __restrict__ scale, float4* __restrict__ out, int nb, int nt,
                                         int nch, int all_pairs, int first_bin, int nbins) {
    const int npair = all_pairs ? nch * (nch - 1) / 2 : nch - 1;
    const size_t total = (size_t)nb * npair * nt * nbins;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int f = (int)(i % nbins);
        size_t r = i / nbins;
        const int t = (int)(r % nt); r /= nt;
        int pr = (int)(r % npair);
        const int b = (int)(r / npair);
        int ca = 0, cb = pr + 1;
        if (all_pairs) {                                             // pairs (0,1) (0,2) .. (0,nch-1) (1,2) ..
            ca = 0;
            while (pr >= nch - 1 - ca) { pr -= nch - 1 - ca; ++ca; }
            cb = ca + 1 + pr;
        }
        const float s = scale[b];
        const float2* row = spec + (((size_t)b * nt + t) * kBins + (first_bin + f)) * nch;
        const float2 a = row[ca], c = row[cb];
        st_stream_f4(out + i, make_float4(a.x * s, c.x * s, a.y * s, c.y * s));
    }
}

static int items_per_clip(int nt) { return (nt + kFPI - 1) / kFPI; }

// Clears the work-queue head and the per-clip arrival counters; counters[1], the rendezvous-timeout flag, is STICKY: it is only ever
// set by the kernels, so a caller can check it once per epoch (sarssl_stft_frontend_error_flag) instead of once per launch.
static cudaError_t reset_counters(unsigned* counters, int nb, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(counters, 0, sizeof(unsigned), stream);
    if (e != cudaSuccess) return e;
    return cudaMemsetAsync(counters + 2, 0, 248 + (size_t)nb * sizeof(unsigned), stream);
}

}  // namespace sarssl

using namespace sarssl;

extern "C" int sarssl_stft_num_frames(long long nsample, int win_len, int hop) {
    if (nsample < win_len || hop <= 0) return 0;
    return (int)((nsample - win_len) / hop + 1);
}

extern "C" size_t sarssl_stft_workspace_bytes(int nb, long long nsample, int nch, int generic) {
    const int nt = sarssl_stft_num_frames(nsample, 512, 256);
    const int ipc = items_per_clip(nt > 0 ? nt : 1);
    size_t bytes = 256;                                            // counters: head, error, pad
    bytes += ((size_t)nb * sizeof(unsigned) + 255) / 256 * 256;    // per-clip arrivals
    bytes += ((size_t)nb * (nt > 0 ? nt : 1) * sizeof(unsigned long long) + 255) / 256 * 256; // partial sums (one {sum, epoch} word per frame)
    bytes += ((size_t)nb * sizeof(unsigned long long) + 255) / 256 * 256;       // per-clip scale (generic path: float; independent-warp kernel: {scale, epoch})
    if (nch != 2 || generic) bytes += (size_t)nb * (nt > 0 ? nt : 1) * kBins * nch * sizeof(float2);   // spectrum temp
    return bytes;
}

static int check_stft_args(const void* sig, const void* out, int nb, long long nsample, int nch, int win_len, int hop, int nfft) {
    SARSSL_CHECK_ARG(sig && out, "stft: null pointer");
    SARSSL_CHECK_ARG(nb > 0 && nch > 0, "stft: nb=%d nch=%d must be positive", nb, nch);
    if (win_len != 512 || nfft != 512 || hop != 256) {
        set_last_error("stft: only win_len = nfft = 512, hop = 256 is implemented (run_pretrain.py:67-72); got %d/%d/%d", win_len, nfft, hop);
        return SARSSL_ERR_UNSUPPORTED;
    }
    SARSSL_CHECK_ARG(nsample >= win_len, "stft: nsample=%lld shorter than one frame", nsample);
    return SARSSL_OK;
}

extern "C" int sarssl_stft_spectrum(const float* sig, float* spec, int nb, long long nsample, int nch, int win_len, int hop,
                                    int nfft, cudaStream_t stream) {
    int rc = check_stft_args(sig, spec, nb, nsample, nch, win_len, hop, nfft);
    if (rc) return rc;
    const int nt = sarssl_stft_num_frames(nsample, win_len, hop), ipc = items_per_clip(nt), npair = (nch + 1) / 2;
    stft_spectrum_kernel<<<nb * ipc * npair, kThreads, 0, stream>>>(sig, reinterpret_cast<float2*>(spec), nullptr, nb, nsample, nch, nt, ipc, npair);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_stft_frontend(const float* sig, float* patches, int nb, long long nsample, int nch, int win_len, int hop,
                                    int nfft, float eps, int force_generic, void* workspace, size_t workspace_bytes,
                                    cudaStream_t stream) {
    int rc = check_stft_args(sig, patches, nb, nsample, nch, win_len, hop, nfft);
    if (rc) return rc;
    SARSSL_CHECK_ARG(nch >= 2, "stft_frontend: needs at least 2 microphones (got %d)", nch);
    SARSSL_CHECK_ARG(workspace != nullptr, "stft_frontend: null workspace");
    if (workspace_bytes < sarssl_stft_workspace_bytes(nb, nsample, nch, 0)) {
        set_last_error("stft_frontend: workspace %zu < required %zu", workspace_bytes, sarssl_stft_workspace_bytes(nb, nsample, nch, 0));
        return SARSSL_ERR_WORKSPACE;
    }
    SARSSL_CHECK_ARG(aligned16(sig) && aligned16(patches) && aligned16(workspace), "stft_frontend: buffers must be 16-byte aligned");
    const int nt = sarssl_stft_num_frames(nsample, win_len, hop), ipc = items_per_clip(nt);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    unsigned* counters = reinterpret_cast<unsigned*>(ws);
    size_t off = 256 + ((size_t)nb * sizeof(unsigned) + 255) / 256 * 256;
    float* partials = reinterpret_cast<float*>(ws + off);
    off += ((size_t)nb * nt * sizeof(unsigned long long) + 255) / 256 * 256;
    float* scale = reinterpret_cast<float*>(ws + off);
    off += ((size_t)nb * sizeof(unsigned long long) + 255) / 256 * 256;

    // force_generic: 0 = the fastest applicable kernel (independent-warp kernel, else the 64-lane-group kernel, else the generic path),
    // 1 = generic three-kernel path, 4 = fused kernel with 64-lane FFT groups and a CTA-level clip rendezvous (round 1's default; kept as
    // the A/B partner of scripts/stft_variants.py), 5 = independent-warp kernel
    const bool pair_ok = nch == 2 && (nsample % 2 == 0);            // the bulk-TMA staging needs 16-byte aligned clip starts
    if (pair_ok && (force_generic == 0 || force_generic == 5) && (long long)nb * nt < (1ll << 31)) {
        static int max_ctas_v5 = -1;
        if (max_ctas_v5 < 0) {
            SARSSL_CUDA(cudaFuncSetAttribute(stft_frontend_warp2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Warp2Smem)));
            // Residency from the kernel's own resource use: cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for this kernel (measured on
            // the B200s of this pool, CUDA 12.9; ncu's launch__occupancy_limit_{registers,shared_mem} both say 2 and two CTAs per SM do run: 0.39 ms
            // against 0.54 ms with one).  If fewer CTAs than assumed were ever resident the rendezvous would time out and raise the sticky error flag.
            cudaFuncAttributes fa;
            SARSSL_CUDA(cudaFuncGetAttributes(&fa, stft_frontend_warp2_kernel));
            int dev = 0, regs_sm = 0, smem_sm = 0, thr_sm = 0;
            SARSSL_CUDA(cudaGetDevice(&dev));
            SARSSL_CUDA(cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev));
            SARSSL_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
            SARSSL_CUDA(cudaDeviceGetAttribute(&thr_sm, cudaDevAttrMaxThreadsPerMultiProcessor, dev));
            const int regs_cta = ((fa.numRegs + 7) / 8 * 8) * kW2Warps * 32;
            const int smem_cta = (int)sizeof(Warp2Smem) + (int)fa.sharedSizeBytes + 1024;          // + the driver's 1 KB per CTA
            int n = regs_sm / regs_cta;
            if (smem_sm / smem_cta < n) n = smem_sm / smem_cta;
            if (thr_sm / (kW2Warps * 32) < n) n = thr_sm / (kW2Warps * 32);
            max_ctas_v5 = n < 2 ? n : 2;                              // each CTA takes half of the SM's tensor memory
        }
        const long long resident5 = (long long)max_ctas_v5 * sm_count(), frames = (long long)nb * nt, ctas = (frames + kW2Warps - 1) / kW2Warps;
        if (max_ctas_v5 > 0 && (ctas <= resident5 || resident5 * kW2Warps >= nt)) {
            static std::atomic<unsigned> g_epoch{0};
            unsigned epoch = ++g_epoch;
            if (epoch == 0) epoch = ++g_epoch;                        // 0 is what a fresh (zeroed) workspace holds
            stft_frontend_warp2_kernel<<<(int)(ctas < resident5 ? ctas : resident5), kW2Warps * 32, sizeof(Warp2Smem), stream>>>(
                sig, reinterpret_cast<float4*>(patches), reinterpret_cast<unsigned long long*>(partials), counters, epoch, nb, nsample, nt, eps);
            SARSSL_LAUNCH_CHECK();
            return SARSSL_OK;
        }
    }
    static int max_ctas_v1 = -1;
    if (max_ctas_v1 < 0) {
        SARSSL_CUDA(cudaFuncSetAttribute(stft_frontend_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem)));
        int n = 0;
        SARSSL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, stft_frontend_fused_kernel, kThreads, sizeof(FusedSmem)));
        max_ctas_v1 = n;
    }
    const long long resident = (long long)max_ctas_v1 * sm_count();
    const bool fused_ok = force_generic != 1 && pair_ok && max_ctas_v1 > 0 && resident >= ipc;
    SARSSL_CUDA(reset_counters(counters, nb, stream));
    if (fused_ok) {
        const long long items = (long long)nb * ipc;
        const int grid = (int)(items < resident ? items : resident);
        stft_frontend_fused_kernel<<<grid, kThreads, sizeof(FusedSmem), stream>>>(sig, reinterpret_cast<float4*>(patches), partials, counters, nb, nsample, nt, ipc, eps);
        SARSSL_LAUNCH_CHECK();
        return SARSSL_OK;
    }
    // generic: spectrum (+ partial sums) -> per-clip scale -> pair + normalise
    float2* spec = reinterpret_cast<float2*>(ws + off);
    if (workspace_bytes < sarssl_stft_workspace_bytes(nb, nsample, nch, 1)) {
        set_last_error("stft_frontend(generic path): workspace %zu < required %zu (query with generic=1)", workspace_bytes,
                       sarssl_stft_workspace_bytes(nb, nsample, nch, 1));
        return SARSSL_ERR_WORKSPACE;
    }
    const int npair = (nch + 1) / 2;
    stft_spectrum_kernel<<<nb * ipc * npair, kThreads, 0, stream>>>(sig, spec, partials, nb, nsample, nch, nt, ipc, npair);
    SARSSL_LAUNCH_CHECK();
    clip_scale_kernel<<<(nb + 3) / 4, 128, 0, stream>>>(partials, scale, nb, ipc, nt, eps);
    SARSSL_LAUNCH_CHECK();
    const size_t total = (size_t)nb * (nch - 1) * nt * 256;
    const int blocks = (int)((total + 255) / 256 < (size_t)sm_count() * 16 ? (total + 255) / 256 : (size_t)sm_count() * 16);
    pair_normalize_kernel<<<blocks, 256, 0, stream>>>(spec, scale, reinterpret_cast<float4*>(patches), nb, nt, nch);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_stft_frontend_ex(const float* sig, float* patches, int nb, long long nsample, int nch, int win_len, int hop, int nfft, float eps,
                                       int all_pairs, int first_bin, int nbins, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    int rc = check_stft_args(sig, patches, nb, nsample, nch, win_len, hop, nfft);
    if (rc) return rc;
    SARSSL_CHECK_ARG(nch >= 2 && workspace != nullptr, "stft_frontend_ex: needs at least 2 microphones and a workspace");
    SARSSL_CHECK_ARG(first_bin >= 0 && nbins > 0 && first_bin + nbins <= kBins, "stft_frontend_ex: bins [%d, %d) outside [0, 257)", first_bin, first_bin + nbins);
    if (workspace_bytes < sarssl_stft_workspace_bytes(nb, nsample, nch, 1)) {
        set_last_error("stft_frontend_ex: workspace %zu < required %zu (query with generic=1)", workspace_bytes, sarssl_stft_workspace_bytes(nb, nsample, nch, 1));
        return SARSSL_ERR_WORKSPACE;
    }
    SARSSL_CHECK_ARG(aligned16(sig) && aligned16(patches) && aligned16(workspace), "stft_frontend_ex: buffers must be 16-byte aligned");
    const int nt = sarssl_stft_num_frames(nsample, win_len, hop), ipc = items_per_clip(nt), npair_ch = (nch + 1) / 2;
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    size_t off = 256 + ((size_t)nb * sizeof(unsigned) + 255) / 256 * 256;
    float* partials = reinterpret_cast<float*>(ws + off);
    off += ((size_t)nb * nt * sizeof(unsigned long long) + 255) / 256 * 256;
    float* scale = reinterpret_cast<float*>(ws + off);
    off += ((size_t)nb * sizeof(unsigned long long) + 255) / 256 * 256;
    float2* spec = reinterpret_cast<float2*>(ws + off);
    stft_spectrum_kernel<<<nb * ipc * npair_ch, kThreads, 0, stream>>>(sig, spec, partials, nb, nsample, nch, nt, ipc, npair_ch);
    SARSSL_LAUNCH_CHECK();
    clip_scale_kernel<<<(nb + 3) / 4, 128, 0, stream>>>(partials, scale, nb, ipc, nt, eps);
    SARSSL_LAUNCH_CHECK();
    const int npair = all_pairs ? nch * (nch - 1) / 2 : nch - 1;
    const size_t total = (size_t)nb * npair * nt * nbins;
    const int blocks = (int)((total + 255) / 256 < (size_t)sm_count() * 16 ? (total + 255) / 256 : (size_t)sm_count() * 16);
    pair_normalize_ex_kernel<<<blocks, 256, 0, stream>>>(spec, scale, reinterpret_cast<float4*>(patches), nb, nt, nch, all_pairs, first_bin, nbins);
    SARSSL_LAUNCH_CHECK();
    return SARSSL_OK;
}

extern "C" int sarssl_stft_frontend_error_flag(const void* workspace, int* flag_host, cudaStream_t stream) {
    SARSSL_CHECK_ARG(workspace && flag_host, "null pointer");
    unsigned v = 0;
    SARSSL_CUDA(cudaMemcpyAsync(&v, static_cast<const unsigned*>(workspace) + 1, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
    SARSSL_CUDA(cudaStreamSynchronize(stream));
    *flag_host = (int)v;
    return SARSSL_OK;
}
