"""CPU, build container only (skipped where /root/reference is absent, e.g. the GPU box): the oracle restatement against the
REAL reference run live - front-end, masks, forward loss, every gradient, BN running statistics."""
import random

import pytest
import torch

from oracle import ref_shim
from oracle import sarssl_oracle as O

pytestmark = pytest.mark.skipif(ref_shim.reference_root() is None, reason="reference checkout not present")


def test_oracle_matches_live_reference():
    rm, rl, rops, ru = ref_shim.load_reference()
    nb, nt = 2, 8
    sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=21)
    net = rm.SARSSL(sig_shape=(256, nt, 2, 2), pretrain=True, device="cpu")
    net.load_state_dict(O.synthetic_state_dict(3))
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    L = rl.STFTLearner(net, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=1, fs=16000, task=None, ch_mode="M")
    L.cpu()
    x_ref, = L.data_preprocess(sig)
    x = O.preprocess(sig)
    assert (x - x_ref).abs().max() <= 1e-5 * x_ref.abs().max()
    net.train()
    random.seed(77)
    loss_ref, diff_ref, vis_ref = net(x_ref)
    loss_ref.backward()
    sd = O.synthetic_state_dict(3)
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k and not k.endswith(".pe"):
            v.requires_grad_(True)
    random.seed(77)
    pidx, cidx = O.draw_masks(nb, nt, nt // 2, 2)
    loss, diff, vis = O.pretrain_forward(x, sd, pidx, cidx, training=True)
    loss.backward()
    assert abs(float(loss) - float(loss_ref)) <= 1e-5 * float(loss_ref) and abs(float(diff) - float(diff_ref)) <= 1e-5 * float(diff_ref)
    assert torch.equal(vis["mask"], vis_ref["mask"])
    gmax = max(float(p.grad.norm()) for p in net.parameters())
    for k, p in net.named_parameters():
        assert float((sd[k].grad - p.grad).norm()) <= 1e-3 * float(p.grad.norm()) + 1e-6 * gmax, k
    for k, v in net.state_dict().items():
        if "running_" in k:
            assert torch.allclose(v, sd[k].detach(), rtol=1e-5, atol=1e-7), k


@pytest.mark.parametrize("ch_mode,ratio,nch", [("MM", 1, 4), ("M", 0.5, 3), ("MM", 0.5, 3)])
def test_preprocess_variants_match_live_reference(ch_mode, ratio, nch):
    """data_preprocess with ch_mode 'MM' (all microphone pairs) and fre_used_ratio 0.5 (bins 0..127): oracle vs the real STFTLearner."""
    rm, rl, _, _ = ref_shim.load_reference()
    sig = O.synthetic_waveforms(2, 9 * 256, nch, seed=31)
    net = rm.SARSSL(sig_shape=(256, 8, 2, 2), pretrain=True, device="cpu")
    L = rl.STFTLearner(net, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=ratio, fs=16000, task=None, ch_mode=ch_mode)
    L.cpu()
    want, = L.data_preprocess(sig)
    got = O.preprocess(sig, ch_mode=ch_mode, fre_used_ratio=ratio)
    assert got.shape == want.shape and torch.allclose(got, want, rtol=1e-5, atol=1e-6)
