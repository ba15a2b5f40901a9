"""Data-parallel gradient exchange (replaces nn.DataParallel of learner.py:25-31).

One process per GPU.  Parameters are replicated and receive identical fused-Adam updates on every rank; the only collective
is a sum all-reduce of the flat fp32 gradient arena, cut into a few contiguous buckets (decoder, spectral encoder, spatial
encoder blocks, spatial stem + first block) that are enqueued on a side CUDA stream as soon as backward has finished the
bucket's parameters, so the exchange overlaps the rest of backward.  BatchNorm statistics stay rank-local, like the
per-replica statistics of the reference's DataParallel.  The 1/world factor is folded into the Adam kernel's grad_scale.

Backends: "nccl" = the library's own communicator (sarssl_comm_*, NCCL over NVLink/NVSwitch); "torch" = torch.distributed
all_reduce on the same buckets (used by the CPU/gloo tests of the bucketing logic)."""
import ctypes as C

import torch

from ._lib import check, lib


def bucket_ranges(store):
    """[(name, offset, numel)] in the order backward completes them: decoder, spec encoder, spat blocks 2..1, spat block 0 + stem."""
    def span(prefixes):
        ks = [k for k in store.order if any(k.startswith(p) for p in prefixes)]
        if not ks:                                # a model without any head / decoder (SARSSL(downstream_head='')): an empty bucket
            return store.total, 0
        lo = min(store.offsets[k][0] for k in ks)
        hi = max(store.offsets[k][0] + (store.offsets[k][1] + 3) // 4 * 4 for k in ks)
        return lo, hi - lo
    out = [("decoder",) + span(["decoder.", "mlp_head.", "joint_head.", "head_mch.", "spec_spat_decoder.", "spec_decoder.", "spat_decoder."]), ("spec_encoder",) + span(["spec_encoder."]),
           ("spat_blocks_1_2",) + span(["spat_encoder.embed.layers.1.", "spat_encoder.embed.layers.2."]),
           ("spat_stem_block_0",) + span(["spat_encoder.patch_embed.", "spat_encoder.embed.layers.0."])]
    covered = sum(n for _, _, n in out)
    assert covered == store.total, (covered, store.total)
    return out


class GradientSync:
    @staticmethod
    def create(model, backend=None):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return None
        if backend is None:
            backend = "nccl" if model.store.flat.is_cuda else "torch"
        return GradientSync(model, backend)

    def __init__(self, model, backend):
        import torch.distributed as dist
        self.model, self.backend = model, backend
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.buckets = bucket_ranges(model.store)
        self.pending = {name for name, _, _ in self.buckets}
        self.stream = None
        self.defer = False                  # True while accumulating micro-batches: backward must not start the exchange yet
        if backend == "nccl" and lib().sarssl_comm_world_size() == self.world:
            # the library's communicator is process-global (one process per GPU): a second model / learner in the same process shares it
            self.stream = torch.cuda.Stream(device=model.store.flat.device)
        elif backend == "nccl":
            L = lib()
            nbytes = L.sarssl_comm_unique_id_bytes()
            buf = (C.c_ubyte * nbytes)()
            if self.rank == 0:
                check(L.sarssl_comm_get_unique_id(buf), "sarssl_comm_get_unique_id")
            obj = [bytes(buf)]
            dist.broadcast_object_list(obj, src=0)
            idbuf = (C.c_ubyte * nbytes).from_buffer_copy(obj[0])
            check(L.sarssl_comm_init(self.rank, self.world, idbuf), "sarssl_comm_init")
            self.stream = torch.cuda.Stream(device=model.store.flat.device)
        model.grad_sync = self
        model.dp = (self.rank, self.world)

    # called by the engine as backward finishes each parameter group
    def bucket_ready(self, name):
        if self.defer or name not in self.pending:      # gradient accumulation: only the group's last backward announces buckets
            return
        self.pending.discard(name)
        _, off, n = next(b for b in self.buckets if b[0] == name)
        g = self.model.store.grad
        if self.backend == "nccl":
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(g.device))
            self.stream.wait_event(ev)
            check(lib().sarssl_allreduce_sum_f32(C.c_void_p(g.data_ptr() + 4 * off), n, C.c_void_p(self.stream.cuda_stream)), "sarssl_allreduce_sum_f32")
        else:
            import torch.distributed as dist
            dist.all_reduce(g[off:off + n])

    def all_reduce(self):
        """Finish the exchange (reduce whatever backward did not announce), make the compute stream wait for it, and return the
        1/world scale the optimizer applies."""
        self.defer = False
        for name, _, _ in self.buckets:
            self.bucket_ready(name)
        if self.backend == "nccl":
            torch.cuda.current_stream(self.model.store.grad.device).wait_stream(self.stream)
        self.pending = {name for name, _, _ in self.buckets}
        return 1.0 / self.world

    def close(self):
        if self.backend == "nccl":
            lib().sarssl_comm_destroy()
