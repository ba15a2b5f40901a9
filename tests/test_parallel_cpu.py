"""CPU, world_size 2 over gloo: the host-side logic of the data-parallel path - gradient buckets cover the flat arena exactly,
the bucketed all-reduce sums across ranks, the 1/world scale, and the per-rank slices of the global mask stream."""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sarssl_b200.model import SARSSL
        from sarssl_b200.parallel import GradientSync, bucket_ranges
        m = SARSSL(sig_shape=(256, 16, 2, 2), device="cpu")
        m.device = torch.device("cpu")
        m.patch_mask.device = "cpu"
        st = m.store
        # buckets tile the arena without gaps or overlaps
        cover = np.zeros(st.total, dtype=np.int32)
        for _, off, n in bucket_ranges(st):
            cover[off:off + n] += 1
        assert (cover == 1).all()
        st.grad.copy_(torch.arange(st.total, dtype=torch.float32) % 1000 * (rank + 1))
        sync = GradientSync.create(m, backend="torch")
        assert sync is not None and m.dp == (rank, world)
        sync.bucket_ready("decoder")                    # announced early by backward; the rest is flushed by all_reduce()
        scale = sync.all_reduce()
        want = torch.arange(st.total, dtype=torch.float32) % 1000 * sum(r + 1 for r in range(world))
        assert scale == 1.0 / world and torch.equal(st.grad, want)
        # a second step re-arms every bucket
        st.grad.fill_(float(rank + 1))
        sync.all_reduce()
        assert torch.equal(st.grad, torch.full((st.total,), float(sum(r + 1 for r in range(world)))))
        # masks: each rank keeps its rows of the global stream
        random.seed(400000005)
        pidx, cidx, flag = m.patch_mask.draw(3, 16, 2, None, dp=m.dp)
        random.seed(400000005)
        ref = [(random.sample(range(16), 8), random.randint(0, 1)) for _ in range(3 * world)]
        mine = ref[rank * 3:(rank + 1) * 3]
        assert pidx.tolist() == [r[0] for r in mine] and cidx.tolist() == [r[1] for r in mine]
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_and_mask_slices_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}
