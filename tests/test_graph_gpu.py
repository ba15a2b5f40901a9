"""GPU: the pre-training step replayed as one CUDA graph (sarssl_b200/graph.py) is interchangeable with the eager step: same masks (host
MT19937 stream), same dropout stream (per-step part of the seed read from device memory), same Adam arithmetic (bias corrections from
device scalars)."""
import random

import pytest
import torch

from oracle import sarssl_oracle as O
from sarssl_b200 import ops
from sarssl_b200.learner import STFTLearner
from sarssl_b200.model import SARSSL

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def build(nt, dtype, dropout, seed):
    m = SARSSL(sig_shape=(256, nt, 2, 2), device=DEV)
    m.load_state_dict(O.synthetic_state_dict(7))
    m.to(DEV)
    m.set_compute_dtype(dtype)
    m.set_dropout(dropout)
    m.rng_state = ops.mt_seed(seed)
    m.train()
    L = STFTLearner(m, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=1, fs=16000, task=None, ch_mode="M")
    L.device = DEV
    return m, L


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_graph_replay_equals_eager_epoch_with_dropout(dtype):
    nb, nt, steps = 4, 16, 6
    batches = [[O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=50 + i)] for i in range(steps)]
    res = {}
    for mode in (False, None):
        m, L = build(nt, dtype, 0.1, 1234)
        l0 = L.pretrain_epoch(batches, lr=1e-3, epoch=1, use_graph=mode)[:2]
        launches_mid = m.engine.k.launches
        l1 = L.pretrain_epoch(batches, lr=5e-4, epoch=2, use_graph=mode)[:2]          # a new epoch: fresh Adam state, another learning rate
        res[mode] = (l0, l1, m.store.flat.detach().clone(), m.state_dict()["spat_encoder.patch_embed.4.running_mean"].clone(),
                     m.engine.k.launches - launches_mid, m.engine.step_seed)
    (e0, e1, pe, rme, ne, se), (g0, g1, pg, rmg, ng, sg) = res[False], res[None]
    assert ng == 0 and ne > 1000                                   # the second epoch ran without a single eager launch
    assert se == sg == 2 * steps
    tol = 1e-5 if dtype == torch.float32 else 2e-3                 # (split-K partial sums are added in arrival order: last-bit differences run to run)
    for a, b in zip(e0 + e1, g0 + g1):
        assert abs(a - b) <= tol * abs(a), (e0, e1, g0, g1)
    assert float((pe - pg).norm() / pe.norm()) < tol
    assert torch.allclose(rme, rmg, rtol=1e-3 if dtype == torch.bfloat16 else 1e-5, atol=1e-6)


def test_graph_step_draws_the_reference_mask_stream():
    """Masks of replayed steps come from the same CPython `random` stream as the reference's PatchMask (global stream, model.rng_state None)."""
    nb, nt = 3, 16
    m, L = build(nt, torch.float32, 0.0, 1)
    m.rng_state = None
    batches = [[O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=3)]] * 3
    random.seed(77)
    _, _, vis = L.pretrain_epoch(batches, lr=1e-4, epoch=1)
    random.seed(77)
    want = [O.draw_masks(nb, nt, nt // 2, 2) for _ in range(3)][-1]
    assert torch.equal(torch.as_tensor(vis.mask_patch_idx).cpu(), want[0])
    assert torch.equal(torch.as_tensor(vis.mask_ch_idx).cpu().reshape(-1), want[1].reshape(-1))
    assert vis["mask"].shape == (nb, 256, nt, 2)


def test_finetune_graph_replay_equals_eager_epoch():
    """The fine-tuning step (train_epoch) replayed as a CUDA graph against the eager epoch: same losses, same final weights."""
    nb, nt, steps = 4, 16, 5
    data = [(O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=60 + i), {"TDOA": (torch.arange(nb, dtype=torch.float32) - nb / 2) * 2e-5 * (i + 1)})
            for i in range(steps)]
    res = {}
    for mode in (False, None):
        m = SARSSL(sig_shape=(256, nt, 2, 2), pretrain=False, device=DEV)
        m.load_state_dict(O.synthetic_state_dict(9, pretrain=False, dembed_ds=768))
        m.to(DEV)
        m.set_compute_dtype(torch.bfloat16)
        m.set_dropout(0.1)
        m.train()
        L = STFTLearner(m, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=1, fs=16000, task="TDOA", ch_mode="M")
        L.device = DEV
        l0 = L.train_epoch(data, lr=2e-5, use_graph=mode)
        mid = m.engine.k.launches
        l1, mae = L.train_epoch(data, lr=1e-5, return_metric=True, use_graph=mode)
        res[mode] = (l0, l1, float(mae), m.store.flat.detach().clone(), m.engine.k.launches - mid)
    (e0, e1, em, pe, ne), (g0, g1, gm, pg, ng) = res[False], res[None]
    assert ng == 0 and ne > 1000
    for a, b in ((e0, g0), (e1, g1), (em, gm)):
        assert abs(a - b) <= 5e-3 * abs(a) + 1e-9, res
    assert float((pe - pg).norm() / pe.norm()) < 1e-4
