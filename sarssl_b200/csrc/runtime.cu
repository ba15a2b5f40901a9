// Library-wide plumbing: version, thread-local error text, cached device properties.
#include "common.cuh"
#include "../../include/sarssl_b200.h"
#include <stdarg.h>

namespace sarssl {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int resident_ctas_impl(const void* kernel, int threads, size_t smem) {
    struct Entry { const void* k; int threads; size_t smem; int n; };
    static Entry cache[256];
    static int used = 0;
    for (int i = 0; i < used; ++i)
        if (cache[i].k == kernel && cache[i].threads == threads && cache[i].smem == smem) return cache[i].n;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const int n = per_sm * sm_count();
    if (used < 256) cache[used++] = {kernel, threads, smem, n};
    return n;
}

}  // namespace sarssl

extern "C" int sarssl_version(void) { return 100; }   // 0.1.0
extern "C" const char* sarssl_last_error(void) { return sarssl::g_err; }
