"""GPU parity of the downstream fine-tuning branch (reference model.py:667-719 + learner.py:170-269,620-653):
SARSSL(pretrain=False) forward / backward and the STFTLearner train / test step against the oracle and the reference fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import sarssl_oracle as O
from sarssl_b200.learner import STFTLearner
from sarssl_b200.model import SARSSL

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DDS = {"spec_spat": 768, "spec": 512, "spat": 256}


def build(nt, embed, sd_seed, dtype=torch.float32):
    m = SARSSL(sig_shape=(256, nt, 2, 2), pretrain=False, downstream_embed=embed, device=DEV)
    m.load_state_dict(O.synthetic_state_dict(sd_seed, pretrain=False, dembed_ds=DDS[embed]))
    m.to(DEV)
    m.set_dropout(0.0)
    m.set_compute_dtype(dtype)
    m.train()
    L = STFTLearner(m, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=1, fs=16000, task="TDOA", ch_mode="M")
    L.device = DEV
    return m, L


@pytest.mark.parametrize("name,embed", [("downstream_nt16_b4", "spec_spat"), ("downstream_spat_nt64_b2", "spat")])
def test_downstream_step_matches_reference_fixture_fp32(name, embed):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    nb, nt = int(g["nb"]), int(g["nt"])
    sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=int(g["sig_seed"]))
    m, L = build(nt, embed, int(g["sd_seed"]))
    x, tar = L.data_preprocess(sig.to(DEV), {"TDOA": torch.from_numpy(g["labels"])})
    assert np.allclose(tar.cpu().numpy(), g["tar"])
    pred, emb = m(x)
    loss = L.loss(pred_batch=pred, gt_batch=tar)
    loss.backward()
    assert pred.shape == (nb, 1) and emb.shape == (nb, DDS[embed])
    assert np.allclose(pred.detach().cpu().numpy(), g["pred"], rtol=2e-4, atol=2e-4)
    assert np.allclose(emb.cpu().numpy(), g["embed"], rtol=2e-4, atol=2e-4)
    assert abs(float(loss) - float(g["loss"])) <= 1e-4 * float(g["loss"])
    assert abs(float(L.evaluate(pred_batch=pred, gt_batch=tar)) - float(g["mae"])) <= 1e-4 * float(g["mae"])
    gmax = max(float(g[k]) for k in g.files if k.startswith("grad_norm/"))
    for k, p in m.named_parameters():
        ref = float(g["grad_norm/" + k])
        assert abs(float(p.grad.norm()) - ref) <= 1e-2 * ref + 1e-5 * gmax, k
        samp = p.grad.detach().cpu().reshape(-1)[torch.from_numpy(g["grad_idx/" + k])].numpy()
        assert np.abs(samp - g["grad_val/" + k]).max() <= 1e-2 * np.abs(g["grad_val/" + k]).max() + 1e-4 * gmax, k
    if embed == "spat":        # the spectral encoder is run but feeds nothing: zero gradient, like autograd in the reference
        assert float(m.store.p("spec_encoder.patch_embed.3.weight").grad.abs().max()) == 0.0


def test_downstream_train_and_test_epoch_bf16():
    nb, nt = 8, 64
    sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=12)
    labels = (torch.arange(nb, dtype=torch.float32) - nb / 2) * 2e-5
    m, L = build(nt, "spec_spat", 9, dtype=torch.bfloat16)
    m.set_dropout(0.1)
    data = [(sig, {"TDOA": labels})] * 4
    first = L.train_epoch(data[:1], lr=2e-5)          # (Adam moves every weight by ~lr per step: keep it small for the synthetic weights)
    last, mae = L.train_epoch(data * 3, lr=2e-5, return_metric=True)
    assert np.isfinite(first) and np.isfinite(last) and last < first
    tl, tm, vis = L.test_epoch(data[:2], return_metric=True, return_vis=True)
    assert np.isfinite(tl) and vis["embed"].shape == (2 * nb, 768) and vis["label"].shape == (2 * nb, 1)
