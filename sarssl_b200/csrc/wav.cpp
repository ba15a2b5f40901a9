// Host-side WAV decoder for the input pipeline in front of the path (reference: soundfile.read inside
// FixMicSigDataset.__getitem__, code/dataset.py:142-178).  RIFF/WAVE, PCM 8/16/24/32-bit, IEEE float 32/64-bit, plain or
// WAVE_FORMAT_EXTENSIBLE headers, any chunk order; samples are returned as float32 in [-1, 1) scaled like libsndfile
// (int16 / 2^15, int24 / 2^23, int32 / 2^31, uint8 -> (v - 128) / 2^7), interleaved [nsample][nch] - the layout the front-end
// kernel takes.  Plain C++ (no CUDA): called through the same C ABI from worker threads (ctypes releases the GIL).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sarssl_b200.h"
#include "common.cuh"

namespace {

struct WavHeader {
    int format = 0;          // 1 PCM, 3 IEEE float
    int nch = 0, fs = 0, bits = 0, block_align = 0;
    long long data_off = -1, data_bytes = 0;
};

uint32_t rd32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

// returns 0 or an error code (message in last_error)
int parse_header(FILE* f, const char* path, WavHeader& h) {
    unsigned char b[40];
    if (fread(b, 1, 12, f) != 12 || memcmp(b, "RIFF", 4) != 0 || memcmp(b + 8, "WAVE", 4) != 0) {
        sarssl::set_last_error("wav: %s is not a RIFF/WAVE file", path);
        return SARSSL_ERR_ARG;
    }
    bool have_fmt = false;
    for (;;) {
        if (fread(b, 1, 8, f) != 8) break;
        const uint32_t sz = rd32(b + 4);
        const long long body = ftell(f);
        if (memcmp(b, "fmt ", 4) == 0) {
            const size_t n = sz < 40 ? sz : 40;
            if (n < 16 || fread(b, 1, n, f) != n) { sarssl::set_last_error("wav: %s has a truncated fmt chunk", path); return SARSSL_ERR_ARG; }
            h.format = rd16(b); h.nch = rd16(b + 2); h.fs = (int)rd32(b + 4); h.block_align = rd16(b + 12); h.bits = rd16(b + 14);
            if (h.format == 0xFFFE && n >= 26) h.format = rd16(b + 24);           // WAVE_FORMAT_EXTENSIBLE: first two bytes of the sub-format GUID
            have_fmt = true;
        } else if (memcmp(b, "data", 4) == 0) {
            h.data_off = body;
            h.data_bytes = sz;
            if (have_fmt) break;
        }
        if (fseek(f, (long)(body + sz + (sz & 1)), SEEK_SET) != 0) break;
    }
    if (!have_fmt || h.data_off < 0) { sarssl::set_last_error("wav: %s has no fmt / data chunk", path); return SARSSL_ERR_ARG; }
    const bool pcm = h.format == 1 && (h.bits == 8 || h.bits == 16 || h.bits == 24 || h.bits == 32);
    const bool flt = h.format == 3 && (h.bits == 32 || h.bits == 64);
    if (!(pcm || flt) || h.nch < 1) {
        sarssl::set_last_error("wav: %s: format tag %d with %d bits is not supported (PCM 8/16/24/32, float 32/64)", path, h.format, h.bits);
        return SARSSL_ERR_UNSUPPORTED;
    }
    if (h.block_align != h.nch * h.bits / 8) h.block_align = h.nch * h.bits / 8;
    return SARSSL_OK;
}

}  // namespace

extern "C" int sarssl_wav_info(const char* path, int* fs, int* nch, long long* nsample) {
    SARSSL_CHECK_ARG(path && fs && nch && nsample, "wav_info: null pointer");
    FILE* f = fopen(path, "rb");
    if (!f) { sarssl::set_last_error("wav: cannot open %s", path); return SARSSL_ERR_ARG; }
    WavHeader h;
    const int rc = parse_header(f, path, h);
    fclose(f);
    if (rc) return rc;
    *fs = h.fs; *nch = h.nch; *nsample = h.data_bytes / h.block_align;
    return SARSSL_OK;
}

// Decodes frames [first, first + count) of all channels into out[count][nch] (float32).  Frames past the end of the file are
// zero-filled; *nread receives the number of real frames.
// expect_nch / expect_fs > 0 and exact != 0 turn a channel-count, sample-rate or length mismatch into an error (the batch reader's checks)
static int wav_read_impl(const char* path, long long first, long long count, float* out, long long* nread, int expect_nch, int expect_fs, int exact) {
    FILE* f = fopen(path, "rb");
    if (!f) { sarssl::set_last_error("wav: cannot open %s", path); return SARSSL_ERR_ARG; }
    WavHeader h;
    int rc = parse_header(f, path, h);
    if (rc) { fclose(f); return rc; }
    if ((expect_nch > 0 && h.nch != expect_nch) || (expect_fs > 0 && h.fs != expect_fs) || (exact && h.data_bytes / h.block_align != first + count)) {
        fclose(f);
        sarssl::set_last_error("wav: %s has %d channels at %d Hz, %lld frames; the batch expects %d channels at %d Hz%s", path, h.nch, h.fs,
                               (long long)(h.data_bytes / h.block_align), expect_nch, expect_fs, exact ? " and an exact length" : "");
        return SARSSL_ERR_ARG;
    }
    const long long total = h.data_bytes / h.block_align;
    long long n = total - first;
    if (n < 0) n = 0;
    if (n > count) n = count;
    std::vector<unsigned char> buf((size_t)n * h.block_align);
    if (n > 0) {
        if (fseek(f, (long)(h.data_off + first * h.block_align), SEEK_SET) != 0 || fread(buf.data(), 1, buf.size(), f) != buf.size()) {
            fclose(f);
            sarssl::set_last_error("wav: %s is shorter than its header says", path);
            return SARSSL_ERR_ARG;
        }
    }
    fclose(f);
    const long long nv = n * h.nch;
    const unsigned char* p = buf.data();
    if (h.format == 3 && h.bits == 32) memcpy(out, p, (size_t)nv * 4);
    else if (h.format == 3) { for (long long i = 0; i < nv; ++i) { double d; memcpy(&d, p + 8 * i, 8); out[i] = (float)d; } }
    else if (h.bits == 16) {                                     // the common case: little-endian host, a loop the compiler vectorises
        const int16_t* q = reinterpret_cast<const int16_t*>(p);  // (std::vector storage is suitably aligned)
        for (long long i = 0; i < nv; ++i) out[i] = (float)q[i] * (1.0f / 32768.0f);
    }
    else if (h.bits == 24) {
        for (long long i = 0; i < nv; ++i) {
            const int32_t v = (int32_t)((uint32_t)p[3 * i] << 8 | (uint32_t)p[3 * i + 1] << 16 | (uint32_t)p[3 * i + 2] << 24) >> 8;
            out[i] = (float)v * (1.0f / 8388608.0f);
        }
    } else if (h.bits == 32) { for (long long i = 0; i < nv; ++i) out[i] = (float)((double)(int32_t)rd32(p + 4 * i) * (1.0 / 2147483648.0)); }
    else { for (long long i = 0; i < nv; ++i) out[i] = (float)((int)p[i] - 128) * (1.0f / 128.0f); }
    for (long long i = nv; i < count * h.nch; ++i) out[i] = 0.f;
    if (nread) *nread = n;
    return SARSSL_OK;
}

extern "C" int sarssl_wav_read_f32(const char* path, long long first, long long count, float* out, long long* nread) {
    SARSSL_CHECK_ARG(path && out && first >= 0 && count >= 0, "wav_read_f32: bad arguments");
    return wav_read_impl(path, first, count, out, nread, 0, 0, 0);
}

// One call decodes a whole batch: file i -> out[i][count][nch], `nthreads` host threads pulling files from a shared counter (no Python,
// no GIL per clip).  Every file must have `nch` channels at `fs` Hz (fs <= 0: not checked) and, with exact != 0, exactly first + count frames;
// otherwise frames past the end are zero-filled.  On failure the first failing file's message is in sarssl_last_error() and *bad_index names it.
extern "C" int sarssl_wav_read_batch_f32(const char* const* paths, int nfiles, long long first, long long count, int nch, int fs, int exact,
                                         float* out, int nthreads, int* bad_index) {
    SARSSL_CHECK_ARG(paths && out && nfiles > 0 && first >= 0 && count >= 0 && nch > 0, "wav_read_batch_f32: bad arguments");
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nfiles) nthreads = nfiles;
    std::atomic<int> next(0), bad(-1), bad_rc(SARSSL_OK);
    std::string bad_msg;
    auto work = [&]() {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= nfiles || bad.load() >= 0) return;
            const int rc = wav_read_impl(paths[i], first, count, out + (size_t)i * (size_t)count * nch, nullptr, nch, fs, exact);
            if (rc != SARSSL_OK) {
                int none = -1;
                if (bad.compare_exchange_strong(none, i)) { bad_rc = rc; bad_msg = sarssl_last_error(); }     // (the message is thread-local)
                return;
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    if (bad_index) *bad_index = bad.load();
    if (bad.load() >= 0) { sarssl::set_last_error("%s", bad_msg.c_str()); return bad_rc.load(); }
    return SARSSL_OK;
}
