"""TEST INFRASTRUCTURE ONLY - generate tests/golden/*.npz by running the REAL reference (build container only).

    python oracle/make_golden.py

The reference (/root/reference/code, imported through oracle/ref_shim.py) is fed seeded synthetic waveforms
(oracle.synthetic_waveforms) and a seeded synthetic state_dict (oracle.synthetic_state_dict); dropout is set to
p=0 (its Philox stream cannot be matched) while the model stays in train() mode so BatchNorm uses batch
statistics.  Masks come from python `random` seeded right before the forward, exactly like run_pretrain.py:249.
The fixtures hold the reference's own outputs; tests compare the oracle restatement and the CUDA path with them.
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim, sarssl_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def sample_idx(n, k=16):
    return np.unique(np.linspace(0, n - 1, num=min(k, n)).astype(np.int64))


def run_case(name, nb, nt, sig_seed, sd_seed, mask_seed, keep_full, grad_samples=0):
    rm, rl, rops, ru = ref_shim.load_reference()
    nsample = (nt + 1) * 256
    sig = O.synthetic_waveforms(nb, nsample, 2, seed=sig_seed)
    net = rm.SARSSL(sig_shape=(256, nt, 2, 2), pretrain=True, device="cpu")
    net.load_state_dict(O.synthetic_state_dict(sd_seed))
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    L = rl.STFTLearner(net, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=1, fs=16000, task=None,
                       ch_mode="M")
    L.cpu()
    out = {"nb": nb, "nt": nt, "sig_seed": sig_seed, "sd_seed": sd_seed, "mask_seed": mask_seed}
    S = L.stft(sig)
    x, = L.data_preprocess(sig)
    if keep_full:
        out["stft_re"], out["stft_im"] = S.real.numpy(), S.imag.numpy()
        out["x"] = x.numpy()
    else:
        out["x_sample_idx"] = sample_idx(x.numel(), 4096)
        out["x_sample"] = x.reshape(-1).numpy()[out["x_sample_idx"]]
    out["x_norm"] = float(x.norm())

    # activations at module boundaries (hooks on the pieces; EmbedEncoder calls .forward directly, so hook below it)
    taps = {}

    def hook(key):
        def f(mod, inp, res):
            taps[key] = res.detach()
        return f

    for enc in ("spec_encoder", "spat_encoder"):
        e = getattr(net, enc)
        for i in (2, 5, 8, 11):
            e.patch_embed[i].register_forward_hook(hook(f"{enc}.patch_embed.{i}"))
        e.patch_embed.register_forward_hook(hook(f"{enc}.patch_embed"))
        for l, blk in enumerate(e.embed.layers):
            for j in range(5):
                blk.sequential[j].register_forward_hook(hook(f"{enc}.embed.layers.{l}.sequential.{j}"))
    net.train()
    random.seed(mask_seed)
    loss, diff, vis = net(x)
    loss.backward()
    random.seed(mask_seed)
    pidx, cidx = O.draw_masks(nb, nt, nt // 2, 2)
    out["mask_patch_idx"], out["mask_ch_idx"] = pidx.numpy(), cidx.numpy()
    assert torch.equal(vis["mask"], O.dense_masks(pidx, cidx, nt, 256)[0].permute(0, 2, 1, 3))
    out["loss"], out["diff"] = float(loss), float(diff)
    if keep_full:
        out["pred"] = vis["pred"].numpy()
    else:
        out["pred_sample_idx"] = sample_idx(vis["pred"].numel(), 4096)
        out["pred_sample"] = vis["pred"].reshape(-1).numpy()[out["pred_sample_idx"]]
    out["pred_norm"] = float(vis["pred"].norm())
    for k, v in taps.items():
        if k.endswith("patch_embed"):      # (nb, D, 1, nt) -> tokens (nb, nt, D)
            v = v[:, :, 0].transpose(1, 2)
        elif ".patch_embed." in k:         # (nb, C, nf, nt) -> our image layout (nb, nt, nf, C)
            v = v.permute(0, 3, 2, 1)
        v = v.contiguous()
        idx = sample_idx(v.numel(), 512)
        out["tap_idx/" + k] = idx
        out["tap_val/" + k] = v.reshape(-1).numpy()[idx]
        out["tap_norm/" + k] = float(v.norm())
    for k, p in net.named_parameters():
        g = p.grad.reshape(-1)
        idx = sample_idx(g.numel(), 64)
        out["grad_idx/" + k] = idx
        out["grad_val/" + k] = g.numpy()[idx]
        out["grad_norm/" + k] = float(g.norm())
        if grad_samples:                   # a large pseudo-random sample for norm-wise gradient gates (indices are re-derived, not stored)
            out["grad_rand/" + k] = g.numpy()[O.fixture_sample_idx(k, g.numel(), grad_samples)].astype(np.float32)
    if grad_samples:
        out["grad_samples"] = grad_samples
    for k, v in net.state_dict().items():
        if "running_" in k:
            out["bn/" + k] = v.numpy().copy()
    # eval-mode forward (pretest_epoch path): running statistics, masks still random
    net.eval()
    random.seed(mask_seed + 1)
    with torch.no_grad():
        le, de, _ = net(x)
    out["eval_loss"], out["eval_diff"] = float(le), float(de)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", out["loss"], "diff", out["diff"], "eval", out["eval_loss"])


def run_downstream_case(name, nb, nt, embed, sig_seed, sd_seed):
    """Downstream fine-tuning branch: reference SARSSL(pretrain=False) + STFTLearner(task='TDOA') MSE step."""
    rm, rl, rops, ru = ref_shim.load_reference()
    dds = {"spec_spat": 768, "spec": 512, "spat": 256}[embed]
    sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=sig_seed)
    labels = (torch.arange(nb, dtype=torch.float32) - nb / 2) * 1e-4                   # TDOA in seconds
    net = rm.SARSSL(sig_shape=(256, nt, 2, 2), pretrain=False, device="cpu", downstream_embed=embed)
    net.load_state_dict(O.synthetic_state_dict(sd_seed, pretrain=False, dembed_ds=dds))
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    L = rl.STFTLearner(net, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=1, fs=16000, task="TDOA", ch_mode="M")
    L.cpu()
    net.train()
    x, tar = L.data_preprocess(sig, {"TDOA": labels})
    pred, emb = net(x)
    loss = L.loss(pred_batch=pred, gt_batch=tar)
    loss.backward()
    out = {"nb": nb, "nt": nt, "sig_seed": sig_seed, "sd_seed": sd_seed, "labels": labels.numpy(), "tar": tar.numpy(), "pred": pred.detach().numpy(),
           "embed": emb.detach().numpy(), "loss": float(loss), "mae": float(L.evaluate(pred_batch=pred, gt_batch=tar))}
    for k, p in net.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        g = g.reshape(-1)
        idx = sample_idx(g.numel(), 64)
        out["grad_idx/" + k] = idx
        out["grad_val/" + k] = g.numpy()[idx]
        out["grad_norm/" + k] = float(g.norm())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", out["loss"], "pred", out["pred"].ravel()[:3])


def run_frozen_case(name, nb, nt, sig_seed, sd_seed, mask_seed):
    """reference SARSSL(pretrain=False, pretrain_frozen_encoder=True) with the encoders' requires_grad cleared as
    run_pretrain.py:364-371 does ('encoder' in key); one forward + backward."""
    rm, rl, rops, ru = ref_shim.load_reference()
    sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=sig_seed)
    net = rm.SARSSL(sig_shape=(256, nt, 2, 2), pretrain=False, device="cpu", pretrain_frozen_encoder=True)
    net.load_state_dict(O.synthetic_state_dict(sd_seed, pretrain=False, frozen=True))
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    for k, p in net.named_parameters():
        if "encoder" in k:
            p.requires_grad = False
    L = rl.STFTLearner(net, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=1, fs=16000, task=None, ch_mode="M")
    L.cpu()
    net.train()
    x, = L.data_preprocess(sig)
    random.seed(mask_seed)
    loss, zero, vis = net(x)
    loss.backward()
    out = {"nb": nb, "nt": nt, "sig_seed": sig_seed, "sd_seed": sd_seed, "mask_seed": mask_seed, "loss": float(loss), "zero": float(zero),
           "pred": vis["pred"].numpy(), "keys": np.array(list(net.state_dict().keys()))}
    for k, p in net.named_parameters():
        if p.grad is None:
            continue
        g = p.grad.reshape(-1)
        out["grad_rand/" + k] = g.numpy()[O.fixture_sample_idx(k, g.numel(), 512)].astype(np.float32)
        out["grad_norm/" + k] = float(g.norm())
    for k, v in net.state_dict().items():
        if "running_" in k:
            out["bn/" + k] = v.numpy().copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", out["loss"], "grads for", sorted(k[10:] for k in out if k.startswith("grad_norm/")))


def run_multich_case(name, nb, nt, nmic_pair, task, sig_seed, sd_seed):
    """SARSSL_MultiCH (model.py:793-821): forward + an MSE step on the head output; fixtures hold pred, embed and gradient samples."""
    rm, rl, rops, ru = ref_shim.load_reference()
    factor = nmic_pair if task == "TDOA" else 1
    net = rm.SARSSL_MultiCH(sig_shape=(256, nt, 2, 2), nmic_pair=nmic_pair, task=task, device="cpu")
    net.load_state_dict(O.synthetic_state_dict(sd_seed, pretrain=False, head="", prefix="model_sch.", nmic_pair=nmic_pair, factor=factor))
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    net.train()
    sig = O.synthetic_waveforms(nb * nmic_pair, (nt + 1) * 256, 2, seed=sig_seed)
    x = O.preprocess(sig)
    pred, emb = net(x)
    tar = torch.linspace(-1.0, 1.0, nb * factor).reshape(nb, factor)
    loss = torch.nn.functional.mse_loss(pred, tar)
    loss.backward()
    out = {"nb": nb, "nt": nt, "nmic_pair": nmic_pair, "factor": factor, "sig_seed": sig_seed, "sd_seed": sd_seed, "pred": pred.detach().numpy(),
           "embed": emb.detach().numpy(), "tar": tar.numpy(), "loss": float(loss), "keys": np.array(list(net.state_dict().keys()))}
    for k, p in net.named_parameters():
        g = (p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1)
        out["grad_rand/" + k] = g.numpy()[O.fixture_sample_idx(k, g.numel(), 256)].astype(np.float32)
        out["grad_norm/" + k] = float(g.norm())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", out["loss"], "pred", out["pred"].ravel()[:4])


def run_mcconformer_case(name, nb, nt, sig_seed, sd_seed):
    """MCConformer (model.py:824-912) in eval mode: data_pred of the un-masked input."""
    rm, rl, rops, ru = ref_shim.load_reference()
    net = rm.MCConformer(sig_shape=[256, nt, 2, 2], patch_shape=(256, 1), spec_model=["cnn", "conformer"], spat_model=["cnn", "conformer"],
                         dembed={"spec": 512, "spat": 256})
    net.load_state_dict(O.synthetic_state_dict(sd_seed))
    net.eval()
    x = O.preprocess(O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=sig_seed))
    with torch.no_grad():
        y = net(x)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), nb=nb, nt=nt, sig_seed=sig_seed, sd_seed=sd_seed, data_pred=y.numpy(),
                        keys=np.array(list(net.state_dict().keys())))
    print(name, "data_pred", tuple(y.shape), float(y.norm()))


def mask_streams():
    """Known-answer vectors for the mask RNG: python `random` (CPython MT19937) under the reference's seeds."""
    out = {}
    for seed, nb, npatch in ((400000002, 8, 256), (100000001, 4, 256), (1, 5, 16), (12345678901234, 3, 999),
                             (7, 2, 1024), (9, 3, 5000)):
        random.seed(seed)
        p, c = O.draw_masks(nb, npatch, npatch // 2, 2)
        out[f"p/{seed}/{nb}/{npatch}"] = p.numpy()
        out[f"c/{seed}/{nb}/{npatch}"] = c.numpy()
    random.seed(5)
    p, c = O.draw_masks(4, 64, 32, 5)      # nmic = 5 exercises a different _randbelow width
    out["p/5/4/64/nmic5"], out["c/5/4/64/nmic5"] = p.numpy(), c.numpy()
    np.savez_compressed(os.path.join(OUT, "mask_streams.npz"), **out)
    print("mask_streams ok")


def lr_table():
    rm, rl, rops, ru = ref_shim.load_reference()
    f = ru.create_learning_rate_schedule(total_steps=30, base=1e-3, decay_type="cosine", warmup_steps=1, linear_end=1e-6)
    np.savez(os.path.join(OUT, "lr_schedule.npz"), lr=np.array([float(f(e)) for e in range(1, 31)], dtype=np.float64))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    mask_streams()
    lr_table()
    run_case("tiny_nt16_b3", nb=3, nt=16, sig_seed=5, sd_seed=7, mask_seed=11, keep_full=True)
    run_case("full_nt256_b2", nb=2, nt=256, sig_seed=6, sd_seed=7, mask_seed=400000001, keep_full=False)
    run_case("full_nt256_b8", nb=8, nt=256, sig_seed=16, sd_seed=7, mask_seed=400000003, keep_full=False, grad_samples=2048)
    run_downstream_case("downstream_nt16_b4", nb=4, nt=16, embed="spec_spat", sig_seed=8, sd_seed=9)
    run_downstream_case("downstream_spat_nt64_b2", nb=2, nt=64, embed="spat", sig_seed=10, sd_seed=9)
    run_frozen_case("frozen_nt16_b3", nb=3, nt=16, sig_seed=12, sd_seed=13, mask_seed=21)
    run_multich_case("multich_nt16_b2x3", nb=2, nt=16, nmic_pair=3, task="TDOA", sig_seed=14, sd_seed=15)
    run_mcconformer_case("mcconformer_nt16_b2", nb=2, nt=16, sig_seed=18, sd_seed=19)
