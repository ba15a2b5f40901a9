"""GPU parity of the frozen-encoder continuation of pre-training (reference model.py:603-666, gen_loss_spec
:749-774, run_pretrain.py:315-390): SARSSL(pretrain=False, pretrain_frozen_encoder=True) against the fixture the real reference produced
and against the oracle, and a pretrain_epoch in which only `spec_spat_decoder` moves."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import sarssl_oracle as O
from sarssl_b200.learner import STFTLearner
from sarssl_b200.model import SARSSL

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def build(nt, sd_seed, dtype=torch.float32, freeze=True):
    m = SARSSL(sig_shape=(256, nt, 2, 2), pretrain=False, pretrain_frozen_encoder=True, device=DEV)
    m.load_state_dict(O.synthetic_state_dict(sd_seed, pretrain=False, frozen=True))
    m.to(DEV)
    m.set_dropout(0.0)
    m.set_compute_dtype(dtype)
    m.train()
    if freeze:                                   # run_pretrain.py:364-371
        for k, p in m.named_parameters():
            if "encoder" in k:
                p.requires_grad = False
    L = STFTLearner(m, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=1, fs=16000, task=None, ch_mode="M")
    L.device = DEV
    return m, L


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def test_frozen_branch_matches_reference_fixture_fp32():
    g = np.load(os.path.join(GOLDEN, "frozen_nt16_b3.npz"))
    nb, nt = int(g["nb"]), int(g["nt"])
    m, L = build(nt, int(g["sd_seed"]))
    assert list(m.state_dict().keys()) == [str(k) for k in g["keys"]]
    x, = L.data_preprocess(O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=int(g["sig_seed"])).to(DEV))
    random.seed(int(g["mask_seed"]))
    loss, zero, vis = m(x)
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 1e-4 * float(g["loss"]) and float(zero) == 0.0
    assert rel(vis["pred"].cpu(), torch.from_numpy(g["pred"])) < 1e-4
    for k, p in m.named_parameters():
        if "grad_norm/" + k in g.files:
            gr = p.grad.detach().cpu().reshape(-1)
            assert abs(float(gr.norm()) - float(g["grad_norm/" + k])) <= 2e-3 * float(g["grad_norm/" + k]), k
            samp = gr.numpy()[O.fixture_sample_idx(k, gr.numel(), 512)]
            assert np.abs(samp - g["grad_rand/" + k]).max() <= 2e-3 * np.abs(g["grad_rand/" + k]).max() + 1e-8, k
        else:                                     # encoders: the reference has p.grad None; here the arena stays zero
            assert float(p.grad.abs().max()) == 0.0, k
    for k in g.files:
        if k.startswith("bn/"):                   # train-mode BatchNorm still updates its running statistics (the encoders are not in eval())
            assert np.allclose(m.state_dict()[k[3:]].cpu().numpy(), g[k], rtol=1e-4, atol=1e-6), k


def test_unfrozen_encoders_get_gradients_matching_oracle():
    """The same branch with trainable encoders (nothing in model.py:603-666 requires them frozen): every gradient against oracle autograd,
    which exercises input mode 4 (un-masked channel of the masked frames only) through the stem's backward kernels."""
    nb, nt = 2, 16
    m, L = build(nt, 5, freeze=False)
    sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=3)
    x, = L.data_preprocess(sig.to(DEV))
    random.seed(77)
    loss, _, vis = m(x)
    loss.backward()
    sd = O.synthetic_state_dict(5, pretrain=False, frozen=True)
    names = [k for k, v in sd.items() if v.is_floating_point() and "running" not in k and not k.endswith(".pe") and not k.startswith(("spec_decoder", "spat_decoder"))]
    for k in names:
        sd[k].requires_grad_(True)
    random.seed(77)
    pidx, cidx = O.draw_masks(nb, nt, nt // 2, 2)
    lo, _, _ = O.frozen_encoder_forward(O.preprocess(sig), sd, pidx, cidx, training=True)
    lo.backward()
    assert abs(float(loss) - float(lo)) <= 1e-4 * float(lo)
    gmax = max(float(sd[k].grad.norm()) for k in names)

    def err(k):        # analytically-zero gradients (key_proj bias: softmax is shift invariant) hold rounding noise only, hence the floor
        a, b = m.store.p(k).grad.cpu().double(), sd[k].grad.double()
        return float((a - b).norm() / (b.norm() + 1e-4 * gmax))
    worst = max((err(k), k) for k in names)
    assert worst[0] < 1e-2, worst
    for k in ("spec_decoder.proj.0.weight", "spat_decoder.proj.2.bias"):           # built but unused by forward (model.py:638-651 are commented out)
        assert float(m.store.p(k).grad.abs().max()) == 0.0


def test_frozen_pretrain_epoch_moves_only_the_decoder_bf16():
    nb, nt = 4, 64
    m, L = build(nt, 9, dtype=torch.bfloat16)
    m.set_dropout(0.1)
    sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=4)
    before = {k: p.detach().clone() for k, p in m.named_parameters()}
    random.seed(1)
    l0, z0, _ = L.pretrain_epoch([[sig]], lr=1e-3, epoch=1)
    l1, z1, vis = L.pretrain_epoch([[sig]] * 6, lr=1e-3, epoch=2)
    assert np.isfinite(l0) and l1 < l0 and z0 == 0.0 == z1
    assert vis["pred"].shape == (nb, 256, nt, 2, 2) and vis["mask"].shape == (nb, 256, nt, 2)
    for k, p in m.named_parameters():
        moved = not torch.equal(p.detach(), before[k])
        assert moved == k.startswith("spec_spat_decoder."), k
    lt = L.pretest_epoch([[sig]], return_diff=False)
    assert np.isfinite(lt)
