"""CPU: the oracle restatement (oracle/sarssl_oracle.py) against fixtures produced by the real reference."""
import math
import os
import random

import numpy as np
import pytest
import torch

from oracle import sarssl_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def test_stft_is_windowed_dft_known_answer():
    # explicit O(N^2) DFT in float64 on a tiny signal: the definition, independent of torch.fft
    sig = O.synthetic_waveforms(1, 3 * 256, 2, seed=3)
    S = O.stft(sig)
    n = torch.arange(512, dtype=torch.float64)
    w = 0.5 - 0.5 * torch.cos(2 * math.pi * n / 512)
    k = torch.arange(257, dtype=torch.float64)[:, None]
    E = torch.exp(-2j * math.pi * k * n[None, :] / 512)
    for t in range(2):
        for ch in range(2):
            fr = sig[0, t * 256:t * 256 + 512, ch].double() * w
            X = (E * fr[None, :]).sum(1)
            assert (S[0, :, t, ch].to(torch.complex128) - X).abs().max() < 1e-5


def test_stft_preprocess_match_reference_fixture():
    g = load("tiny_nt16_b3")
    sig = O.synthetic_waveforms(int(g["nb"]), (int(g["nt"]) + 1) * 256, 2, seed=int(g["sig_seed"]))
    S = O.stft(sig)
    ref = torch.complex(torch.from_numpy(g["stft_re"]), torch.from_numpy(g["stft_im"]))
    assert (S - ref).abs().max() <= 1e-4 * ref.abs().max()
    x = O.preprocess(sig)
    xr = torch.from_numpy(g["x"])
    assert x.shape == xr.shape
    assert (x - xr).norm() <= 1e-5 * xr.norm()


def test_istft_roundtrip_rectangular():
    sig = O.synthetic_waveforms(2, 9 * 256, 2, seed=9)
    # reference STFT analysis is Hann, synthesis rectangular: istft(stft(x)) == x * hann-OLA / count
    y = O.istft(O.stft(sig))
    assert y.shape == (2, 9 * 256, 2)
    w = O.hann_periodic(512)
    # interior samples are covered by two frames whose Hann windows sum to 1 -> y = x / 2
    assert (y[:, 256:8 * 256] - 0.5 * sig[:, 256:8 * 256]).abs().max() < 1e-6


def test_mask_stream_matches_cpython_fixture():
    g = load("mask_streams")
    for key in g.files:
        if not key.startswith("p/"):
            continue
        parts = key.split("/")
        seed, nb, npatch = int(parts[1]), int(parts[2]), int(parts[3])
        nmic = 5 if key.endswith("nmic5") else 2
        random.seed(seed)
        p, c = O.draw_masks(nb, npatch, npatch // 2, nmic)
        assert np.array_equal(p.numpy(), g[key])
        assert np.array_equal(c.numpy(), g["c/" + key[2:]])


def test_relative_shift_is_the_pad_view_trick():
    T = 7
    pos = torch.randn(2, 3, T, T)
    z = torch.cat([pos.new_zeros(2, 3, T, 1), pos], dim=-1).view(2, 3, T + 1, T)[:, :, 1:].reshape(2, 3, T, T)
    assert torch.equal(O.relative_shift(pos), z)


def test_lr_schedule_fixture():
    g = load("lr_schedule")
    mine = np.array([O.cosine_lr(e) for e in range(1, 31)])
    assert np.allclose(mine, g["lr"], rtol=1e-6, atol=1e-12)


def _run_oracle(g):
    nb, nt = int(g["nb"]), int(g["nt"])
    sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=int(g["sig_seed"]))
    x = O.preprocess(sig)
    sd = O.synthetic_state_dict(int(g["sd_seed"]))
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k and not k.endswith(".pe"):
            v.requires_grad_(True)
    random.seed(int(g["mask_seed"]))
    pidx, cidx = O.draw_masks(nb, nt, nt // 2, 2)
    taps = {}
    loss, diff, vis = O.pretrain_forward(x, sd, pidx, cidx, training=True, taps=taps)
    loss.backward()
    return x, sd, pidx, cidx, loss, diff, vis, taps


@pytest.mark.parametrize("name", ["tiny_nt16_b3", "full_nt256_b2"])
def test_forward_backward_match_reference_fixture(name):
    g = load(name)
    x, sd, pidx, cidx, loss, diff, vis, taps = _run_oracle(g)
    assert np.array_equal(pidx.numpy(), g["mask_patch_idx"]) and np.array_equal(cidx.numpy(), g["mask_ch_idx"])
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert abs(float(diff) - float(g["diff"])) <= 1e-5 * abs(float(g["diff"]))
    if "pred" in g.files:
        pr = torch.from_numpy(g["pred"])
        assert (vis["pred"] - pr).norm() <= 1e-4 * pr.norm()
    else:
        got = vis["pred"].reshape(-1)[torch.from_numpy(g["pred_sample_idx"])]
        assert np.abs(got.numpy() - g["pred_sample"]).max() <= 1e-4 * float(g["pred_norm"]) / math.sqrt(got.numel())
    # gradients: norm-wise per tensor, with an absolute floor for analytically-zero grads (key_proj bias)
    gmax = max(float(g[k]) for k in g.files if k.startswith("grad_norm/"))
    for k in g.files:
        if not k.startswith("grad_norm/"):
            continue
        key = k[len("grad_norm/"):]
        mine = sd[key].grad.reshape(-1)
        ref_norm = float(g[k])
        assert abs(float(mine.norm()) - ref_norm) <= 1e-3 * ref_norm + 1e-7 * gmax, key
        samp = mine[torch.from_numpy(g["grad_idx/" + key])].numpy()
        assert np.abs(samp - g["grad_val/" + key]).max() <= 5e-3 * np.abs(g["grad_val/" + key]).max() + 1e-6 * gmax, key
    for k in g.files:
        if k.startswith("bn/"):
            assert np.allclose(sd[k[3:]].detach().numpy(), g[k], rtol=1e-4, atol=1e-6), k


def test_eval_mode_matches_reference_fixture():
    g = load("tiny_nt16_b3")
    x, sd, *_ = _run_oracle(g)          # one training step updates the running statistics first, as in the fixture
    sd = {k: v.detach() for k, v in sd.items()}
    random.seed(int(g["mask_seed"]) + 1)
    pidx, cidx = O.draw_masks(int(g["nb"]), int(g["nt"]), int(g["nt"]) // 2, 2)
    with torch.no_grad():
        le, de, _ = O.pretrain_forward(x, sd, pidx, cidx, training=False)
    assert abs(float(le) - float(g["eval_loss"])) <= 1e-5 * float(g["eval_loss"])
    assert abs(float(de) - float(g["eval_diff"])) <= 1e-5 * float(g["eval_diff"])


@pytest.mark.parametrize("name,embed", [("downstream_nt16_b4", "spec_spat"), ("downstream_spat_nt64_b2", "spat")])
def test_downstream_branch_matches_reference_fixture(name, embed):
    """Downstream fine-tuning branch: oracle vs the reference's SARSSL(pretrain=False) + TDOA MSE step."""
    g = load(name)
    nb, nt = int(g["nb"]), int(g["nt"])
    dds = {"spec_spat": 768, "spec": 512, "spat": 256}[embed]
    sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=int(g["sig_seed"]))
    sd = O.synthetic_state_dict(int(g["sd_seed"]), pretrain=False, dembed_ds=dds)
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k and not k.endswith(".pe"):
            v.requires_grad_(True)
    pred, emb = O.downstream_forward(O.preprocess(sig), sd, embed, training=True)
    tar = torch.from_numpy(g["tar"])
    assert np.allclose(tar.numpy(), g["labels"][:, None] * 16000)
    loss = torch.nn.functional.mse_loss(pred, tar)
    loss.backward()
    assert np.allclose(pred.detach().numpy(), g["pred"], rtol=1e-4, atol=1e-5) and np.allclose(emb.detach().numpy(), g["embed"], rtol=1e-4, atol=1e-5)
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * float(g["loss"])
    gmax = max(float(g[k]) for k in g.files if k.startswith("grad_norm/"))
    for k in g.files:
        if k.startswith("grad_norm/"):
            key = k[len("grad_norm/"):]
            gr = sd[key].grad
            mine = float(gr.norm()) if gr is not None else 0.0
            assert abs(mine - float(g[k])) <= 1e-3 * float(g[k]) + 1e-6 * gmax, key


def test_frozen_encoder_branch_matches_reference_fixture():
    """oracle.frozen_encoder_forward vs the reference's SARSSL(pretrain_frozen_encoder=True) forward/backward
    (model.py:603-666, gen_loss_spec :749-774) - loss, the prediction and the gradients of the only trainable module (spec_spat_decoder)."""
    import random
    g = np.load(os.path.join(GOLDEN, "frozen_nt16_b3.npz"))
    nb, nt = int(g["nb"]), int(g["nt"])
    sd = O.synthetic_state_dict(int(g["sd_seed"]), pretrain=False, frozen=True)
    assert list(sd.keys()) == [str(k) for k in g["keys"]]                       # the reference's state_dict keys, in its order
    names = [k for k in sd if k.startswith("spec_spat_decoder.")]
    for k in names:
        sd[k].requires_grad_(True)
    x = O.preprocess(O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=int(g["sig_seed"])))
    random.seed(int(g["mask_seed"]))
    pidx, cidx = O.draw_masks(nb, nt, nt // 2, 2)
    loss, zero, vis = O.frozen_encoder_forward(x, sd, pidx, cidx, training=True)
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * float(g["loss"]) and float(zero) == 0.0 == float(g["zero"])
    assert np.allclose(vis["pred"].numpy(), g["pred"], rtol=1e-4, atol=1e-5)
    assert sorted(k[10:] for k in g.files if k.startswith("grad_norm/")) == sorted(names)      # the encoders received no gradient in the reference
    for k in names:
        gr = sd[k].grad.reshape(-1)
        assert abs(float(gr.norm()) - float(g["grad_norm/" + k])) <= 1e-4 * float(g["grad_norm/" + k]), k
        samp = gr.numpy()[O.fixture_sample_idx(k, gr.numel(), 512)]
        assert np.abs(samp - g["grad_rand/" + k]).max() <= 1e-4 * np.abs(g["grad_rand/" + k]).max() + 1e-9, k
    for k in g.files:
        if k.startswith("bn/"):
            assert np.allclose(sd[k[3:]].numpy(), g[k], rtol=1e-4, atol=1e-6), k


def test_multich_matches_reference_fixture():
    """oracle.multich_forward vs the reference's SARSSL_MultiCH (model.py:793-821): prediction, embedding, gradients of an MSE step."""
    g = np.load(os.path.join(GOLDEN, "multich_nt16_b2x3.npz"))
    nb, nt, P, factor = int(g["nb"]), int(g["nt"]), int(g["nmic_pair"]), int(g["factor"])
    sd = O.synthetic_state_dict(int(g["sd_seed"]), pretrain=False, head="", prefix="model_sch.", nmic_pair=P, factor=factor)
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    names = [k for k, v in sd.items() if v.is_floating_point() and "running" not in k and not k.endswith(".pe")]
    for k in names:
        sd[k].requires_grad_(True)
    x = O.preprocess(O.synthetic_waveforms(nb * P, (nt + 1) * 256, 2, seed=int(g["sig_seed"])))
    pred, emb = O.multich_forward(x, sd, P, training=True)
    assert np.allclose(pred.detach().numpy(), g["pred"], rtol=1e-4, atol=1e-5) and np.allclose(emb.detach().numpy(), g["embed"], rtol=1e-4, atol=1e-5)
    loss = torch.nn.functional.mse_loss(pred, torch.from_numpy(g["tar"]))
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * float(g["loss"])
    gmax = max(float(g[k]) for k in g.files if k.startswith("grad_norm/"))
    for k in names:
        gr = (sd[k].grad if sd[k].grad is not None else torch.zeros_like(sd[k])).reshape(-1)
        assert abs(float(gr.norm()) - float(g["grad_norm/" + k])) <= 2e-3 * float(g["grad_norm/" + k]) + 1e-6 * gmax, k


def test_mcconformer_matches_reference_fixture():
    """oracle.mcconformer_forward vs the reference's MCConformer (model.py:824-912), eval mode."""
    g = np.load(os.path.join(GOLDEN, "mcconformer_nt16_b2.npz"))
    nb, nt = int(g["nb"]), int(g["nt"])
    sd = O.synthetic_state_dict(int(g["sd_seed"]))
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    x = O.preprocess(O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=int(g["sig_seed"])))
    with torch.no_grad():
        y = O.mcconformer_forward(x, sd, training=False)
    assert y.shape == g["data_pred"].shape and np.allclose(y.numpy(), g["data_pred"], rtol=1e-4, atol=1e-5)
