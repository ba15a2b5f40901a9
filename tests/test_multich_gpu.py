"""GPU parity of the two remaining model classes of code/model.py: SARSSL_MultiCH (model.py:793-821) and
MCConformer (model.py:824-912) against fixtures produced by the real reference and against oracle autograd."""
import os

import numpy as np
import pytest
import torch

from oracle import sarssl_oracle as O
from sarssl_b200.model import MCConformer, SARSSL, SARSSL_MultiCH
from sarssl_b200.optim import FusedAdam

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def test_multich_matches_reference_fixture_fp32():
    g = np.load(os.path.join(GOLDEN, "multich_nt16_b2x3.npz"))
    nb, nt, P, factor = int(g["nb"]), int(g["nt"]), int(g["nmic_pair"]), int(g["factor"])
    m = SARSSL_MultiCH(sig_shape=(256, nt, 2, 2), nmic_pair=P, task="TDOA", device=DEV)
    assert list(m.state_dict().keys()) == [str(k) for k in g["keys"]]                   # model_sch.* / head_mch.* like the reference
    m.load_state_dict(O.synthetic_state_dict(int(g["sd_seed"]), pretrain=False, head="", prefix="model_sch.", nmic_pair=P, factor=factor))
    m.to(DEV)
    m.set_dropout(0.0)
    m.train()
    x = O.preprocess(O.synthetic_waveforms(nb * P, (nt + 1) * 256, 2, seed=int(g["sig_seed"]))).to(DEV)
    pred, emb = m(x)
    assert pred.shape == (nb, factor) and emb.shape == (nb, P * 256)
    assert np.allclose(pred.detach().cpu().numpy(), g["pred"], rtol=2e-4, atol=2e-4) and np.allclose(emb.cpu().numpy(), g["embed"], rtol=2e-4, atol=2e-4)
    loss = torch.nn.functional.mse_loss(pred, torch.from_numpy(g["tar"]).to(DEV))
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 1e-4 * float(g["loss"])
    gmax = max(float(g[k]) for k in g.files if k.startswith("grad_norm/"))
    for k, p in m.named_parameters():
        gr = p.grad.detach().cpu().reshape(-1)
        ref = float(g["grad_norm/" + k])
        assert abs(float(gr.norm()) - ref) <= 1e-2 * ref + 1e-5 * gmax, k
        samp = gr.numpy()[O.fixture_sample_idx(k, gr.numel(), 256)]
        assert np.abs(samp - g["grad_rand/" + k]).max() <= 1e-2 * np.abs(g["grad_rand/" + k]).max() + 1e-4 * gmax, k
    before = m.store.p("head_mch.1.weight").detach().clone()
    FusedAdam(m, lr=1e-3).step(1e-3)                                                   # the head lives in the same arena as the encoders
    assert not torch.equal(before, m.store.p("head_mch.1.weight").detach())


def test_headless_downstream_model_returns_the_embedding():
    """SARSSL(pretrain=False, downstream_head='') - the inner model of SARSSL_MultiCH (model.py:797): pred is the time-mean embedding."""
    nt = 16
    m = SARSSL(sig_shape=(256, nt, 2, 2), pretrain=False, downstream_head="", downstream_embed="spat", device=DEV)
    sd = O.synthetic_state_dict(9, pretrain=False, head="")
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    m.to(DEV)
    m.set_dropout(0.0)
    m.train()
    x = O.preprocess(O.synthetic_waveforms(2, (nt + 1) * 256, 2, seed=4)).to(DEV)
    pred, emb = m(x)
    assert pred.shape == (2, 256) and torch.equal(pred, emb)
    _, want = O.downstream_forward(x.cpu(), {**sd, "mlp_head.0.weight": torch.ones(256), "mlp_head.0.bias": torch.zeros(256),
                                             "mlp_head.1.weight": torch.zeros(1, 256), "mlp_head.1.bias": torch.zeros(1)}, embed_use="spat")
    assert rel(emb.cpu(), want) < 1e-4
    (pred * torch.linspace(-1, 1, 256, device=DEV)).sum().backward()
    assert float(m.store.p("spat_encoder.patch_embed.3.weight").grad.abs().max()) > 0 and float(m.store.p("spec_encoder.patch_embed.3.weight").grad.abs().max()) == 0.0


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
def test_mcconformer_matches_reference_fixture(dtype, tol):
    g = np.load(os.path.join(GOLDEN, "mcconformer_nt16_b2.npz"))
    nb, nt = int(g["nb"]), int(g["nt"])
    m = MCConformer(sig_shape=[256, nt, 2, 2], patch_shape=(256, 1), device=DEV)
    assert list(m.state_dict().keys()) == [str(k) for k in g["keys"]]
    m.load_state_dict(O.synthetic_state_dict(int(g["sd_seed"])))
    m.to(DEV)
    m.set_compute_dtype(dtype)
    m.eval()
    x = O.preprocess(O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=int(g["sig_seed"]))).to(DEV)
    with torch.no_grad():
        y = m(x)
    assert y.shape == g["data_pred"].shape and rel(y.float().cpu(), torch.from_numpy(g["data_pred"])) < tol


def test_mcconformer_backward_matches_oracle_autograd():
    nb, nt = 2, 16
    m = MCConformer(sig_shape=[256, nt, 2, 2], device=DEV)
    sd = O.synthetic_state_dict(21)
    m.load_state_dict(sd)
    m.to(DEV)
    m.set_dropout(0.0)
    m.train()
    x = O.preprocess(O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=6))
    tar = torch.randn(nb, 256, nt, 2, 2, generator=torch.Generator().manual_seed(1))
    y = m(x.to(DEV))
    torch.nn.functional.mse_loss(y, tar.to(DEV)).backward()
    names = [k for k, v in sd.items() if v.is_floating_point() and "running" not in k and not k.endswith(".pe")]
    for k in names:
        sd[k].requires_grad_(True)
    yo = O.mcconformer_forward(x, sd, training=True)
    torch.nn.functional.mse_loss(yo, tar).backward()
    assert rel(y.detach().cpu(), yo.detach()) < 1e-4
    gmax = max(float(sd[k].grad.norm()) for k in names)
    worst = max((float((m.store.p(k).grad.cpu().double() - sd[k].grad.double()).norm() / (sd[k].grad.double().norm() + 1e-4 * gmax)), k) for k in names)
    assert worst[0] < 1e-2, worst


def test_joint_head_dlabel3_matches_reference_and_oracle_autograd():
    """SARSSL(pretrain=False, downstream_dlabel=3): joint_head = LayerNorm, Linear, ReLU, Linear (model.py:501-507,709-710).  Forward against the
    real reference's output (fixture), gradients against oracle autograd."""
    g = np.load(os.path.join(GOLDEN, "joint_head_nt16_b2.npz"))
    nb, nt, dl = int(g["nb"]), int(g["nt"]), int(g["dlabel"])
    sd = O.synthetic_state_dict(int(g["sd_seed"]), pretrain=False, dembed_ds=768, dlabel=dl)
    m = SARSSL(sig_shape=(256, nt, 2, 2), pretrain=False, downstream_dlabel=dl, device=DEV)
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    m.to(DEV)
    m.set_dropout(0.0)
    m.train()
    x = O.preprocess(O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=int(g["sig_seed"])))
    pred, emb = m(x.to(DEV))
    assert pred.shape == (nb, dl) and np.allclose(pred.detach().cpu().numpy(), g["pred"], rtol=2e-4, atol=2e-4)
    assert np.allclose(emb.cpu().numpy(), g["embed"], rtol=2e-4, atol=2e-4)
    tar = torch.linspace(-1, 1, nb * dl).reshape(nb, dl)
    torch.nn.functional.mse_loss(pred, tar.to(DEV)).backward()
    names = [k for k, v in sd.items() if v.is_floating_point() and "running" not in k and not k.endswith(".pe")]
    for k in names:
        sd[k].requires_grad_(True)
    po, _ = O.downstream_forward(x, sd, "spec_spat", training=True)
    torch.nn.functional.mse_loss(po, tar).backward()
    gmax = max(float(sd[k].grad.norm()) for k in names)
    worst = max((float((m.store.p(k).grad.cpu().double() - sd[k].grad.double()).norm() / (sd[k].grad.double().norm() + 1e-4 * gmax)), k) for k in names)
    assert worst[0] < 1e-2, worst
