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

// ---------------- pipelined fused kernel -------------------------------------------------------------------------------------
// Same work split as stft_frontend_fused_kernel (item = kPF consecutive frames of one clip, one frame per 64-lane group), but the
// three latencies that kernel exposes are overlapped inside the CTA:
//   * the bulk-TMA load of item n+1 is in flight while item n is transformed            (double-buffered input);
//   * the per-clip rendezvous and the scaled store of item n-1 happen after item n has been transformed and published, so the
//     other CTAs working on that clip had a whole transform's time to arrive          (double-buffered un-scaled output).
// Items are assigned statically, item(n) = blockIdx.x + n * gridDim.x, with gridDim.x >= items per clip: a CTA owns at most one
// item of any clip, items n-1 and n+1 lie in different clips, and a CTA only ever waits for a clip after publishing everything it
// has transformed - with all CTAs resident the rendezvous cannot dead-lock (induction over the clip index).
constexpr int kPF = 4;
struct PipeSmem {
    float2 in[2][(kPF + 1) * kHop];               // 2 x 10 KB  staged samples (ch0, ch1)
    float4 out[2][kPF * kHop];                    // 2 x 16 KB  un-scaled bins 1..256, (re0, re1, im0, im1)
    float scratch[kGroups][kFftScratchFloats];    // 18 KB
    float red[32];
    uint64_t bar[2];
    float scale;
};

__global__ void __launch_bounds__(kThreads) stft_frontend_pipe_kernel(const float* __restrict__ sig, float4* __restrict__ out,
                                                                    float* partials, unsigned* counters, int nb,
                                                                    long long nsample, int nt, int ipc, float eps) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PipeSmem& sm = *reinterpret_cast<PipeSmem*>(smem_raw);
    const int tid = threadIdx.x, g = tid >> 6, l = tid & 63;
    FftLane lane;
    lane.init(l);
    float* sre = sm.scratch[g];
    float* sim = sre + kFftPlane;
    const long long nitems = (long long)nb * ipc;
    auto issue_load = [&](long long item, int buf) {          // thread 0 only
        const int b = (int)(item / ipc), fb = (int)(item - (long long)b * ipc);
        const int f0 = fb * kPF, nfr = min(kPF, nt - f0);
        const uint32_t bytes = (uint32_t)(nfr + 1) * kHop * sizeof(float2);
        mbar_expect_tx(&sm.bar[buf], bytes);
        tma_bulk_g2s(sm.in[buf], sig + ((size_t)b * nsample + (size_t)f0 * kHop) * 2, bytes, &sm.bar[buf]);
    };
    if (tid == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        mbar_fence_init();
        if (blockIdx.x < nitems) issue_load(blockIdx.x, 0);
    }
    __syncthreads();
    uint32_t phases = 0;                                      // bit buf = parity to wait for on bar[buf]
    int pb = -1, pfb = 0, pnfr = 0;                            // the published, not yet stored item (clip, item in clip, frames)
    for (int n = 0;; ++n) {
        const long long item = (long long)blockIdx.x + (long long)n * gridDim.x;
        const int buf = n & 1;
        const bool valid = item < nitems;
        int b = 0, fb = 0, nfr = 0;
        if (valid) {
            if (tid == 0 && item + gridDim.x < nitems) issue_load(item + gridDim.x, buf ^ 1);      // in[buf ^ 1] was last read two barriers ago
            b = (int)(item / ipc); fb = (int)(item - (long long)b * ipc);
            nfr = min(kPF, nt - fb * kPF);
            mbar_wait(&sm.bar[buf], (phases >> buf) & 1u);
            phases ^= 1u << buf;
            float part = 0.f;
            if (g < nfr) {
                float2 v[8];
                fft_frame(sm.in[buf] + g * kHop, lane, sre, sim, l, g, v);
                float4* orow = sm.out[buf] + g * kHop;
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
            }
            const float total = block_sum(part, sm.red);
            if (tid == 0) {
                partials[(size_t)b * ipc + fb] = total;
                __threadfence();
                atomicAdd(&counters[2 + b], 1u);
            }
        }
        if (pb >= 0) {                                        // finish the previous item: rendezvous, scale, store
            if (tid == 0) {
                unsigned spins = 0;
                while (ld_acquire_u32(&counters[2 + pb]) < (unsigned)ipc) {
                    __nanosleep(32);
                    if (++spins > (1u << 24)) { atomicExch(&counters[1], 1u); break; }      // never expected; report instead of hanging
                }
            }
            __syncthreads();
            if (tid < 32) {                                   // fixed-order (deterministic) sum of the clip's partials
                float s = 0.f;
                for (int i = tid; i < ipc; i += 32) s += __ldcg(&partials[(size_t)pb * ipc + i]);
                s = warp_sum(s);
                if (tid == 0) sm.scale = 1.0f / (s / (float)((long long)kBins * nt) + eps);
            }
            __syncthreads();
            const float scale = sm.scale;
            float4* dst = out + ((size_t)pb * nt + (size_t)pfb * kPF) * kHop;
            const float4* src = sm.out[buf ^ 1];
            for (int i = tid; i < pnfr * kHop; i += kThreads) {
                float4 o = src[i];
                o.x *= scale; o.y *= scale; o.z *= scale; o.w *= scale;
                st_stream_f4(dst + i, o);
            }
        }
        __syncthreads();
        if (!valid) break;
        pb = b; pfb = fb; pnfr = nfr;
    }
}

// ---------------- second-generation fused kernel: independent warps, one transform each (fft512w.cuh) ----------------
// Every warp of the persistent grid is its own worker: frame f = warp_global_id + round * total_warps of the (clip-major) frame list.
// Per frame: bulk-TMA its 512 samples (double buffered: the next frame's samples are in flight while this one is transformed),
// FFT in registers + one smem transpose, stage the un-scaled bins in the transpose buffer, publish |X_ch0| partial sum and arrive on
// the clip's counter, spin (lane 0) until the clip's nt frames have arrived, add the nt partials in a fixed order, scale and write the
// 4 KB frame row.  No block-level barrier exists after start-up, so the SM's warps drift into different phases and overlap load /
// compute / store.  Deadlock-free: frames are assigned round-robin in list order, a warp publishes before it waits, and it only waits
// for frames of the same or an earlier round; all warps are resident (grid sized by the occupancy query).
constexpr int kWarpsPerCta = 8;
struct WarpWorkerSmem {
    float2 in[kWarpsPerCta][2][kFftN];            // 64 KB   per-warp double-buffered samples
    float2 tb[kWarpsPerCta][kWTransFloat2];       // 33 KB   per-warp transpose buffer / staged output
    float win[kFftN];                             // 2 KB
    uint64_t bar[kWarpsPerCta][2];
};

__global__ void __launch_bounds__(kWarpsPerCta * 32, 2) stft_frontend_warp_kernel(const float* __restrict__ sig, float4* __restrict__ out, float* partials,
                                                                                 unsigned* counters, int nb, long long nsample, int nt, float eps) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    WarpWorkerSmem& sm = *reinterpret_cast<WarpWorkerSmem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    WarpFftLane lc;
    lc.init(lane);
    for (int i = tid; i < kFftN; i += kWarpsPerCta * 32) sm.win[i] = 0.5f - 0.5f * cospif((float)i / 256.0f);
    if (lane == 0) {
        mbar_init(&sm.bar[warp][0], 1);
        mbar_init(&sm.bar[warp][1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const long long total = (long long)nb * nt;
    const long long gw = (long long)blockIdx.x * kWarpsPerCta + warp, G = (long long)gridDim.x * kWarpsPerCta;
    const int k1 = lane & 15, h = lane >> 4;
    const int kbase = h == 0 ? k1 : (((16 - k1) & 15) + 128);       // bin of out[j] is kbase + 16*j
    float2* tb = sm.tb[warp];
    float4* stage = reinterpret_cast<float4*>(tb);
    uint32_t ph0 = 0, ph1 = 0;
    int buf = 0;
    auto issue = [&](long long f, int bsel) {
        if (lane == 0) {
            const long long b = f / nt, t = f - b * nt;
            mbar_expect_tx(&sm.bar[warp][bsel], kFftN * sizeof(float2));
            tma_bulk_g2s(sm.in[warp][bsel], sig + ((size_t)b * nsample + (size_t)t * kHop) * 2, kFftN * sizeof(float2), &sm.bar[warp][bsel]);
        }
    };
    if (gw < total) issue(gw, 0);
    for (long long f = gw; f < total; f += G) {
        const long long b = f / nt;
        const int t = (int)(f - b * nt);
        if (f + G < total) issue(f + G, buf ^ 1);                    // prefetch the next frame of this warp
        mbar_wait(&sm.bar[warp][buf], buf ? ph1 : ph0);
        if (buf) ph1 ^= 1u; else ph0 ^= 1u;
        const float2* frame = sm.in[warp][buf];
        float2 v[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            const float2 sx = frame[32 * n1 + lane];
            const float w = sm.win[32 * n1 + lane];
            v[n1] = make_float2(sx.x * w, sx.y * w);
        }
        wfft_stage1(v, lc, tb, lane);
        __syncwarp();
        wfft_stage2(v, tb, lane);
        wfft_combine(v, lane);
        float4 o[8], nyq;
        wfft_split_all(v, lane, o, nyq);
        float part = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = kbase + 16 * j;
            part += sqrtf(o[j].x * o[j].x + o[j].z * o[j].z);
            if (k >= 1) stage[k - 1] = o[j];
        }
        if (lane == 16) {
            part += sqrtf(nyq.x * nyq.x + nyq.z * nyq.z);
            stage[255] = nyq;
        }
        part = warp_sum(part);
        if (lane == 0) {
            partials[(size_t)b * nt + t] = part;
            __threadfence();
            atomicAdd(&counters[2 + b], 1u);
            unsigned spins = 0;
            while (ld_acquire_u32(&counters[2 + b]) < (unsigned)nt) {
                __nanosleep(32);
                if (++spins > (1u << 24)) { atomicExch(&counters[1], 1u); break; }
            }
        }
        __syncwarp();
        float s = 0.f;
        for (int i = lane; i < nt; i += 32) s += __ldcg(&partials[(size_t)b * nt + i]);
        s = warp_sum(s);
        const float scale = 1.0f / (s / (float)((long long)kBins * nt) + eps);
        float4* dst = out + ((size_t)b * nt + t) * kHop;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float4 ov = stage[lane + 32 * r];
            ov.x *= scale; ov.y *= scale; ov.z *= scale; ov.w *= scale;
            st_stream_f4(dst + lane + 32 * r, ov);
        }
        __syncwarp();
        buf ^= 1;
    }
}

// ---------------- third-generation fused kernel: independent warps, rendezvous one frame behind ---------------------------------
// The warp-worker kernel above stalls because every warp waits for its clip's rendezvous right after publishing: the 257 warps of a clip
// meet after every frame and the SM alternates between a compute phase and a store phase.  Here the wait is software-pipelined: a warp
// transforms frame r (round-robin list order, f = warp_id + r * warps), publishes its |X_ch0| partial sum, and only THEN finishes frame
// r - 1 (whose un-scaled bins wait in the second transpose buffer): by that time the other warps have had a whole transform's time to
// publish their frames of that clip, so the spin almost never spins, nothing runs in lock step, and loads, butterflies and the scaled
// stores of different warps overlap freely.  The input needs one 4 KB buffer per warp: the next frame's bulk-TMA copy is issued as soon
// as this frame's samples are in registers.
// Deadlock-free when every warp is resident and warps >= nt - 1: a warp publishes round r before it waits for round r - 1, and a clip that
// holds a round r - 1 frame ends before frame (r + 1) * warps.
constexpr int kW2Warps = 8;
struct Warp2Smem {
    float2 in[kW2Warps][kFftN];                   // 32 KB   per-warp staged samples (ch0, ch1)
    float2 tb[kW2Warps][2][kWTransFloat2];        // 66 KB   per-warp transpose buffer / staged un-scaled output, double buffered
    float win[kFftN];                             // 2 KB
    uint64_t bar[kW2Warps];
};

__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// The rendezvous carries no fence and no atomic: a frame's |X_ch0| partial sum is published as ONE 64-bit store {sum, launch epoch}, and a
// clip is complete when all nt of its words carry this launch's epoch (read with L2-coherent loads; the words are the only data exchanged,
// so nothing else has to be ordered).  __threadfence + atomicAdd + ld.acquire cost 48 % of this kernel's stall samples before.
__device__ __forceinline__ unsigned long long ld_cg_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kW2Warps * 32, 2) stft_frontend_warp2_kernel(const float* __restrict__ sig, float4* __restrict__ out,
                                                                              unsigned long long* partials, unsigned* counters, unsigned epoch, int nb,
                                                                              long long nsample, int nt, float eps) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Warp2Smem& sm = *reinterpret_cast<Warp2Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    WarpFftLane lc;
    lc.init(lane);
    for (int i = tid; i < kFftN; i += kW2Warps * 32) sm.win[i] = 0.5f - 0.5f * cospif((float)i / 256.0f);
    if (lane == 0) {
        mbar_init(&sm.bar[warp], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const long long total = (long long)nb * nt;
    const long long gw = (long long)blockIdx.x * kW2Warps + warp, G = (long long)gridDim.x * kW2Warps;
    const int k1 = lane & 15, h = lane >> 4;
    const int kbase = h == 0 ? k1 : (((16 - k1) & 15) + 128);       // bin of out[j] is kbase + 16*j
    const float inv_bins = 1.0f / (float)((long long)kBins * nt);
    float2* in = sm.in[warp];
    auto issue = [&](long long f) {
        if (lane == 0) {
            const long long b = f / nt, t = f - b * nt;
            mbar_expect_tx(&sm.bar[warp], kFftN * sizeof(float2));
            tma_bulk_g2s(in, sig + ((size_t)b * nsample + (size_t)t * kHop) * 2, kFftN * sizeof(float2), &sm.bar[warp]);
        }
    };
    // scale and store the frame staged in `stage` once its clip is complete
    auto finish = [&](long long f, const float4* stage) {
        const long long b = f / nt;
        const int t = (int)(f - b * nt);
        const unsigned long long* pp = partials + (size_t)b * nt;
        float s;
        for (unsigned spins = 0;; ++spins) {
            s = 0.f;
            bool ok = true;
            for (int i = lane; i < nt; i += 32) {                    // fixed order: deterministic
                const unsigned long long w = ld_cg_u64(pp + i);
                ok = ok && (unsigned)(w >> 32) == epoch;
                s += __uint_as_float((unsigned)w);
            }
            if (__all_sync(0xffffffffu, ok)) break;
            __nanosleep(64);
            if (spins > (1u << 22)) { if (lane == 0) atomicExch(&counters[1], 1u); break; }      // never expected; report instead of hanging
        }
        s = warp_sum(s);
        const float scale = 1.0f / (s * inv_bins + eps);
        float4* dst = out + ((size_t)b * nt + t) * kHop;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float4 ov = stage[lane + 32 * r];
            ov.x *= scale; ov.y *= scale; ov.z *= scale; ov.w *= scale;
            st_stream_f4(dst + lane + 32 * r, ov);
        }
    };
    uint32_t ph = 0;
    int cur = 0;
    long long pf = -1;                                               // published, not yet stored
    if (gw < total) issue(gw);
    for (long long f = gw; f < total; f += G) {
        const long long b = f / nt;
        const int t = (int)(f - b * nt);
        mbar_wait(&sm.bar[warp], ph);
        ph ^= 1u;
        float2 v[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            const float2 sx = in[32 * n1 + lane];
            const float w = sm.win[32 * n1 + lane];
            v[n1] = make_float2(sx.x * w, sx.y * w);
        }
        __syncwarp();                                                // every lane has its samples: the buffer can take the next frame
        if (f + G < total) issue(f + G);
        float2* tb = sm.tb[warp][cur];
        wfft_stage1(v, lc, tb, lane);
        __syncwarp();
        wfft_stage2(v, tb, lane);
        wfft_combine(v, lane);
        float4 o[8], nyq;
        wfft_split_all(v, lane, o, nyq);
        float4* stage = reinterpret_cast<float4*>(tb);
        float part = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = kbase + 16 * j;
            part += sqrt_approx(o[j].x * o[j].x + o[j].z * o[j].z);
            if (k >= 1) stage[k - 1] = o[j];
        }
        if (lane == 16) {
            part += sqrt_approx(nyq.x * nyq.x + nyq.z * nyq.z);
            stage[255] = nyq;
        }
        part = warp_sum(part);
        if (lane == 0) __stcg(&partials[(size_t)b * nt + t], ((unsigned long long)epoch << 32) | __float_as_uint(part));
        __syncwarp();                                                // staged bins visible to the whole warp
        if (pf >= 0) finish(pf, reinterpret_cast<const float4*>(sm.tb[warp][cur ^ 1]));
        pf = f;
        cur ^= 1;
    }
    if (pf >= 0) finish(pf, reinterpret_cast<const float4*>(sm.tb[warp][cur ^ 1]));
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
    bytes += ((size_t)nb * sizeof(float) + 255) / 256 * 256;       // per-clip scale (generic path)
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
    off += ((size_t)nb * sizeof(float) + 255) / 256 * 256;

    // force_generic: 0 / 4 = fused kernel with 64-lane FFT groups (default, fastest measured), 1 = generic three-kernel path,
    // 3 = the fused kernel with load / rendezvous / store pipelined inside the CTA (same speed: the kernel is bound by the FFT's
    // group barriers and shared-memory round trips, not by the exposed latencies),
    // 2 = experimental warp-worker kernel (one warp per transform, no block barriers; measured slower: the per-clip rendezvous
    // keeps its independent warps in lock step - profiles/r01_stft_variants.txt)
    if (force_generic == 3) {      // pipelined kernel (its own frames-per-item granularity); measured equal to the default, profiles/r01_stft_variants.txt
        static int max_ctas_v3 = -1;
        if (max_ctas_v3 < 0) {
            SARSSL_CUDA(cudaFuncSetAttribute(stft_frontend_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PipeSmem)));
            int n = 0;
            SARSSL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, stft_frontend_pipe_kernel, kThreads, sizeof(PipeSmem)));
            max_ctas_v3 = n;
        }
        const int ipc3 = (nt + kPF - 1) / kPF;
        const long long resident3 = (long long)max_ctas_v3 * sm_count(), items3 = (long long)nb * ipc3;
        if (nch == 2 && (nsample % 2 == 0) && resident3 >= ipc3) {
            SARSSL_CUDA(reset_counters(counters, nb, stream));
            const int grid = (int)(items3 < resident3 ? items3 : resident3);
            stft_frontend_pipe_kernel<<<grid, kThreads, sizeof(PipeSmem), stream>>>(sig, reinterpret_cast<float4*>(patches), partials, counters, nb, nsample, nt,
                                                                                   ipc3, eps);
            SARSSL_LAUNCH_CHECK();
            return SARSSL_OK;
        }
    }
    if (force_generic == 5) {      // independent warps, rendezvous one frame behind
        static int max_ctas_v5 = -1;
        if (max_ctas_v5 < 0) {
            SARSSL_CUDA(cudaFuncSetAttribute(stft_frontend_warp2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Warp2Smem)));
            int n = 0;
            SARSSL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, stft_frontend_warp2_kernel, kW2Warps * 32, sizeof(Warp2Smem)));
            max_ctas_v5 = n;
        }
        const long long resident5 = (long long)max_ctas_v5 * sm_count(), frames = (long long)nb * nt, ctas = (frames + kW2Warps - 1) / kW2Warps;
        if (nch == 2 && (nsample % 2 == 0) && max_ctas_v5 > 0 && (ctas <= resident5 || resident5 * kW2Warps >= nt)) {
            static std::atomic<unsigned> g_epoch{0};
            unsigned epoch = ++g_epoch;
            if (epoch == 0) epoch = ++g_epoch;                        // 0 is what a fresh (zeroed) workspace holds
            stft_frontend_warp2_kernel<<<(int)(ctas < resident5 ? ctas : resident5), kW2Warps * 32, sizeof(Warp2Smem), stream>>>(
                sig, reinterpret_cast<float4*>(patches), reinterpret_cast<unsigned long long*>(partials), counters, epoch, nb, nsample, nt, eps);
            SARSSL_LAUNCH_CHECK();
            return SARSSL_OK;
        }
    }
    static int max_ctas_v1 = -1, max_ctas_v2 = -1;
    const bool v1 = force_generic != 2;
    const size_t smem = v1 ? sizeof(FusedSmem) : sizeof(WarpWorkerSmem);
    if (max_ctas_v1 < 0) {
        SARSSL_CUDA(cudaFuncSetAttribute(stft_frontend_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem)));
        SARSSL_CUDA(cudaFuncSetAttribute(stft_frontend_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WarpWorkerSmem)));
        int n = 0;
        SARSSL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, stft_frontend_fused_kernel, kThreads, sizeof(FusedSmem)));
        max_ctas_v1 = n;
        SARSSL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, stft_frontend_warp_kernel, kWarpsPerCta * 32, sizeof(WarpWorkerSmem)));
        max_ctas_v2 = n;
    }
    const int max_ctas_per_sm = v1 ? max_ctas_v1 : max_ctas_v2;
    const long long resident = (long long)max_ctas_per_sm * sm_count();
    // the bulk-TMA staging needs 16-byte aligned clip starts: nsample * 2 floats * 4 B -> nsample even
    // (the warp-worker kernel needs every frame of a clip's round resident: resident warps >= nt)
    const bool fused_ok = force_generic != 1 && nch == 2 && (nsample % 2 == 0) && max_ctas_per_sm > 0 &&
                          (v1 ? resident >= ipc : resident * kWarpsPerCta >= nt);
    SARSSL_CUDA(reset_counters(counters, nb, stream));
    if (fused_ok) {
        const long long items = (long long)nb * ipc;
        const int grid = (int)(items < resident ? items : resident);
        if (v1)
            stft_frontend_fused_kernel<<<grid, kThreads, smem, stream>>>(sig, reinterpret_cast<float4*>(patches), partials, counters, nb, nsample, nt, ipc, eps);
        else {
            const long long frames = (long long)nb * nt, ctas = (frames + kWarpsPerCta - 1) / kWarpsPerCta;
            stft_frontend_warp_kernel<<<(int)(ctas < resident ? ctas : resident), kWarpsPerCta * 32, smem, stream>>>(sig, reinterpret_cast<float4*>(patches), partials,
                                                                                                                  counters, nb, nsample, nt, eps);
        }
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

extern "C" int sarssl_stft_frontend_error_flag(const void* workspace, int* flag_host, cudaStream_t stream) {
    SARSSL_CHECK_ARG(workspace && flag_host, "null pointer");
    unsigned v = 0;
    SARSSL_CUDA(cudaMemcpyAsync(&v, static_cast<const unsigned*>(workspace) + 1, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
    SARSSL_CUDA(cudaStreamSynchronize(stream));
    *flag_host = (int)v;
    return SARSSL_OK;
}
