"""GPU, 2 devices (skipped on single-GPU boxes): the data-parallel step over the library's own NCCL communicator -
bucketed gradient all-reduce overlapped with backward, identical fused-Adam updates on every rank."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import random
        from oracle import sarssl_oracle as O
        from sarssl_b200.learner import STFTLearner
        from sarssl_b200.model import SARSSL
        from sarssl_b200.optim import FusedAdam
        nb, nt = 2, 16
        sig = O.synthetic_waveforms(nb * world, (nt + 1) * 256, 2, seed=3)[rank * nb:(rank + 1) * nb]     # this rank's shard of the global batch
        m = SARSSL(sig_shape=(256, nt, 2, 2), device=dev)
        m.load_state_dict(O.synthetic_state_dict(7))
        m.to(dev)
        m.set_dropout(0.0)
        m.train()
        L = STFTLearner(m, 512, 0.5, 512, 1, 16000)
        L.device = dev
        L.mul_gpu()
        assert m.dp == (rank, world) and L.grad_sync.backend == "nccl"
        x, = L.data_preprocess(sig.to(dev))
        random.seed(400000001)                      # same seed on every rank: global mask stream, sliced per rank
        loss, diff, vis = m(x)
        # local gradients first (no exchange), for the reference sum
        m.grad_sync, keep = None, m.grad_sync
        loss.backward()
        local = m.store.grad.clone()
        m.grad_sync = keep
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        want = sum(gathered)
        # now the real path: backward announces buckets, all_reduce() flushes and returns 1/world
        m.store.grad.zero_()
        x, = L.data_preprocess(sig.to(dev))
        random.seed(400000001)
        loss2, _, _ = m(x)
        loss2.backward()
        scale = keep.all_reduce()
        torch.cuda.synchronize()
        assert scale == 1.0 / world
        err = float((m.store.grad - want).norm() / want.norm())
        assert err < 1e-5, err
        opt = FusedAdam(m, lr=1e-3)
        opt.step(1e-3, grad_scale=scale)
        flat = [torch.empty_like(m.store.flat) for _ in range(world)]
        dist.all_gather(flat, m.store.flat)
        assert all(torch.equal(flat[0], f) for f in flat)                 # replicas stay bit-identical
        masks = [torch.empty_like(vis.mask_patch_idx) for _ in range(world)]
        dist.all_gather(masks, vis.mask_patch_idx)
        random.seed(400000001)
        ref = [random.sample(range(nt), nt // 2) + [random.randint(0, 1)] for _ in range(nb * world)]
        got = torch.cat(masks).cpu().tolist()
        assert got == [r[:-1] for r in ref]
        # a second model / learner in the same process shares the library's process-global communicator (bench.py's sub-runs do this)
        m2 = SARSSL(sig_shape=(256, nt, 2, 2), device=dev)
        m2.load_state_dict(O.synthetic_state_dict(7))
        m2.to(dev)
        m2.set_dropout(0.0)
        m2.train()
        L2 = STFTLearner(m2, 512, 0.5, 512, 1, 16000)
        L2.device = dev
        L2.mul_gpu()
        x, = L2.data_preprocess(sig.to(dev))
        random.seed(400000001)
        l3, _, _ = m2(x)
        l3.backward()
        L2.grad_sync.all_reduce()
        torch.cuda.synchronize()
        err2 = float((m2.store.grad - want).norm() / want.norm())
        assert err2 < 1e-5, err2
        keep.close()
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def _graph_worker(rank, world, port, out):
    """The data-parallel step replayed as a CUDA graph (the NCCL all-reduce is captured on its side stream): replicas stay bit-identical and
    the trained weights equal the eager data-parallel epoch's.  (Collectives of torch's communicator and of the library's own are never left
    in flight together: two NCCL communicators running concurrently on one GPU may dead-lock, hence the synchronisations.)"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from oracle import sarssl_oracle as O
        from sarssl_b200 import ops
        from sarssl_b200.learner import STFTLearner
        from sarssl_b200.model import SARSSL
        nb, nt = 2, 16
        finals = {}
        for mode in (False, None):
            m3 = SARSSL(sig_shape=(256, nt, 2, 2), device=dev)
            m3.load_state_dict(O.synthetic_state_dict(7))
            m3.to(dev)
            m3.set_dropout(0.1)
            m3.set_compute_dtype(torch.bfloat16)
            m3.rng_state = ops.mt_seed(99)
            m3.train()
            L3 = STFTLearner(m3, 512, 0.5, 512, 1, 16000)
            L3.device = dev
            L3.mul_gpu()
            data = [[O.synthetic_waveforms(nb * world, (nt + 1) * 256, 2, seed=70 + i)[rank * nb:(rank + 1) * nb]] for i in range(4)]
            launches0 = m3._engine().k.launches
            L3.pretrain_epoch(data, lr=1e-3, epoch=1, use_graph=mode)
            eager_launches = m3.engine.k.launches - launches0
            torch.cuda.synchronize()
            flat = [torch.empty_like(m3.store.flat) for _ in range(world)]
            dist.all_gather(flat, m3.store.flat)
            torch.cuda.synchronize()
            assert all(torch.equal(flat[0], f) for f in flat), mode
            finals[mode] = (m3.store.flat.clone(), eager_launches)
        assert finals[None][1] <= 0.6 * finals[False][1]                   # one eager step + one recording pass of launches; steps 2-4 were replays
        gerr = float((finals[None][0] - finals[False][0]).norm() / finals[False][0].norm())
        assert gerr < 2e-3, gerr
        out[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_data_parallel_graph_replay_two_gpus():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_graph_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}


@pytest.mark.timeout(300)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_data_parallel_step_two_gpus():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}
