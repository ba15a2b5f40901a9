// Data-parallel gradient exchange: one NCCL communicator per process (one process per GPU), sum all-reduce of contiguous
// buckets of the flat fp32 gradient arena over NVLink 5 / NVSwitch.  Replaces nn.DataParallel's per-step weight broadcast +
// gradient reduce-to-GPU-0 (learner.py:25-31).  The buckets are enqueued on a side stream as backward
// finishes each parameter group, so the exchange overlaps the rest of backward (sarssl_b200/parallel.py).
#include "common.cuh"
#include <nccl.h>
#include <cstdlib>

namespace sarssl {
static ncclComm_t g_comm = nullptr;
static int g_world = 1, g_rank = 0;
}  // namespace sarssl

using namespace sarssl;

#define SARSSL_NCCL(call)                                                                            \
    do {                                                                                             \
        ncclResult_t r__ = (call);                                                                   \
        if (r__ != ncclSuccess) {                                                                    \
            set_last_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, ncclGetErrorString(r__));    \
            return SARSSL_ERR_NCCL;                                                                  \
        }                                                                                            \
    } while (0)

extern "C" int sarssl_comm_unique_id_bytes(void) { return (int)sizeof(ncclUniqueId); }

extern "C" int sarssl_comm_get_unique_id(void* id_host) {
    SARSSL_CHECK_ARG(id_host, "comm_get_unique_id: null pointer");
    ncclUniqueId id;
    SARSSL_NCCL(ncclGetUniqueId(&id));
    memcpy(id_host, &id, sizeof(id));
    return SARSSL_OK;
}

extern "C" int sarssl_comm_init(int rank, int world, const void* id_host) {
    SARSSL_CHECK_ARG(id_host && world >= 1 && rank >= 0 && rank < world, "comm_init: bad arguments rank=%d world=%d", rank, world);
    if (g_comm) { set_last_error("comm_init: communicator already initialised"); return SARSSL_ERR_ARG; }
    ncclUniqueId id;
    memcpy(&id, id_host, sizeof(id));
    // The all-reduce runs on a side stream under backward, whose persistent kernels hold every SM with one large CTA each, so NCCL's CTAs compete
    // with them.  Capping NCCL's CTAs (SARSSL_NCCL_MAX_CTAS=4) was measured at 8 GPUs: 44.90 ms/step against 44.56 with NCCL's default - no gain,
    // so the default stays NCCL's own choice (0); the knob remains for other topologies.
    ncclConfig_t config = NCCL_CONFIG_INITIALIZER;
    int max_ctas = 0;
    if (const char* e = getenv("SARSSL_NCCL_MAX_CTAS")) max_ctas = atoi(e);
    if (max_ctas > 0) { config.minCTAs = 1; config.maxCTAs = max_ctas; }
    SARSSL_NCCL(ncclCommInitRankConfig(&g_comm, world, id, rank, &config));
    g_world = world; g_rank = rank;
    return SARSSL_OK;
}

extern "C" int sarssl_comm_world_size(void) { return g_comm ? g_world : 1; }

// in-place sum over ranks of n fp32 values starting at buf, enqueued on `stream`
extern "C" int sarssl_allreduce_sum_f32(float* buf, size_t n, cudaStream_t stream) {
    SARSSL_CHECK_ARG(buf && n > 0, "allreduce: bad arguments");
    if (!g_comm) { set_last_error("allreduce: communicator not initialised"); return SARSSL_ERR_NCCL; }
    SARSSL_NCCL(ncclAllReduce(buf, buf, n, ncclFloat, ncclSum, g_comm, stream));
    return SARSSL_OK;
}

extern "C" int sarssl_comm_destroy(void) {
    if (g_comm) {
        ncclCommDestroy(g_comm);
        g_comm = nullptr;
        g_world = 1;
    }
    return SARSSL_OK;
}
