"""GPU parity of the fused model (SARSSL.forward / backward through the C ABI) against the CPU oracle and the fixtures the
real reference produced.  fp32 mode: forward within 1e-4 relative; gradients norm-wise per tensor."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import sarssl_oracle as O
from sarssl_b200 import ops
from sarssl_b200.learner import STFTLearner
from sarssl_b200.model import SARSSL

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def build(nt, sd_seed=7, dtype=torch.float32):
    m = SARSSL(sig_shape=(256, nt, 2, 2), device=DEV)
    m.load_state_dict(O.synthetic_state_dict(sd_seed))
    m.to(DEV)
    m.set_dropout(0.0)
    m.set_compute_dtype(dtype)
    m.train()
    return m


def run_ours(m, sig, mask_seed):
    L = STFTLearner(m, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=1, fs=16000, task=None, ch_mode="M")
    x, = L.data_preprocess(sig.to(DEV))
    random.seed(mask_seed)
    loss, diff, vis = m(x)
    loss.backward()
    return loss, diff, vis


def run_oracle(sig, nt, sd_seed, mask_seed):
    x = O.preprocess(sig)
    sd = O.synthetic_state_dict(sd_seed)
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k and not k.endswith(".pe"):
            v.requires_grad_(True)
    random.seed(mask_seed)
    pidx, cidx = O.draw_masks(sig.shape[0], nt, nt // 2, 2)
    taps = {}
    loss, diff, vis = O.pretrain_forward(x, sd, pidx, cidx, training=True, taps=taps)
    loss.backward()
    return loss, diff, vis, sd, taps


def check_grads(m, ref_grad, tol, floor=1e-6):
    """Norm-wise relative error per tensor; `floor` (relative to the largest gradient norm) absorbs gradients that are
    analytically zero (key_proj bias: softmax is shift invariant) and only hold rounding noise."""
    gmax = max(float(g.norm()) for g in ref_grad.values())
    worst = ("", 0.0)
    for k, p in m.named_parameters():
        g, r = p.grad.detach().cpu().double(), ref_grad[k].double()
        err = float((g - r).norm()) / (float(r.norm()) + floor * gmax)
        if err > worst[1]:
            worst = (k, err)
    assert worst[1] < tol, f"worst gradient mismatch {worst}"


@pytest.mark.parametrize("nb,nt", [(3, 16), (2, 7)])
def test_forward_backward_fp32_vs_oracle(nb, nt):
    sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=5)
    m = build(nt)
    loss, diff, vis = run_ours(m, sig, 11)
    rl, rd, rvis, sd, taps = run_oracle(sig, nt, 7, 11)
    assert rel(vis["pred"].cpu(), rvis["pred"]) < 1e-4
    assert torch.equal(vis["mask"].cpu(), rvis["mask"])
    assert rel(vis["tar"].cpu(), rvis["tar"]) < 1e-5
    assert abs(float(loss) - float(rl)) < 1e-4 * float(rl) and abs(float(diff) - float(rd)) < 1e-4 * float(rd)
    check_grads(m, {k: sd[k].grad for k, _ in m.named_parameters()}, 1e-2)   # BN bias grads are cancellation-heavy sums (the CPU oracle itself moves by ~2e-3 between runs)
    for k, v in m.state_dict().items():                       # BatchNorm running statistics after one step
        if "running_" in k:
            assert torch.allclose(v.cpu(), sd[k].detach(), rtol=1e-4, atol=1e-6), k
        if "num_batches" in k:
            assert int(v) == 1


def test_matches_reference_fixture_tiny():
    g = np.load(os.path.join(GOLDEN, "tiny_nt16_b3.npz"))
    nb, nt = int(g["nb"]), int(g["nt"])
    sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=int(g["sig_seed"]))
    m = build(nt, int(g["sd_seed"]))
    loss, diff, vis = run_ours(m, sig, int(g["mask_seed"]))
    assert np.array_equal(vis.mask_patch_idx.cpu().numpy(), g["mask_patch_idx"])
    assert np.array_equal(vis.mask_ch_idx.cpu().numpy(), g["mask_ch_idx"][:, 0])
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * float(g["loss"])
    assert abs(float(diff) - float(g["diff"])) < 1e-4 * float(g["diff"])
    assert rel(vis["pred"].cpu(), torch.from_numpy(g["pred"])) < 1e-4
    gmax = max(float(g[k]) for k in g.files if k.startswith("grad_norm/"))
    for k, p in m.named_parameters():
        mine = p.grad.detach().cpu().reshape(-1)
        assert abs(float(mine.norm()) - float(g["grad_norm/" + k])) <= 2e-3 * float(g["grad_norm/" + k]) + 1e-6 * gmax, k
        samp = mine[torch.from_numpy(g["grad_idx/" + k])].numpy()
        assert np.abs(samp - g["grad_val/" + k]).max() <= 5e-3 * np.abs(g["grad_val/" + k]).max() + 5e-5 * gmax, k     # floor: a ReLU flipping at |u| ~ 1e-7 moves one term
    # eval-mode forward with the updated running statistics (pretest_epoch path)
    m.eval()
    L = STFTLearner(m, 512, 0.5, 512, 1, 16000)
    x, = L.data_preprocess(sig.to(DEV))
    random.seed(int(g["mask_seed"]) + 1)
    with torch.no_grad():
        le, de, _ = m(x)
    assert abs(float(le) - float(g["eval_loss"])) < 1e-4 * float(g["eval_loss"])


def test_matches_reference_fixture_full_size():
    g = np.load(os.path.join(GOLDEN, "full_nt256_b2.npz"))
    nb, nt = int(g["nb"]), int(g["nt"])
    sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=int(g["sig_seed"]))
    m = build(nt, int(g["sd_seed"]))
    loss, diff, vis = run_ours(m, sig, int(g["mask_seed"]))
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * float(g["loss"])
    got = vis["pred"].contiguous().reshape(-1)[torch.from_numpy(g["pred_sample_idx"]).to(DEV)].cpu().numpy()
    assert np.abs(got - g["pred_sample"]).max() <= 1e-3 * np.abs(g["pred_sample"]).max()
    gmax = max(float(g[k]) for k in g.files if k.startswith("grad_norm/"))
    for k, p in m.named_parameters():
        assert abs(float(p.grad.norm()) - float(g["grad_norm/" + k])) <= 5e-3 * float(g["grad_norm/" + k]) + 1e-6 * gmax, k


def _autocast_reference_grads(sig, nt, sd_seed, mask_seed):
    """Calibration: the reference's own mixed-precision route - the oracle's functional model on CUDA under
    torch.autocast(bfloat16) (what learner.py:100 does with --use-amp, in bf16)."""
    x = O.preprocess(sig).to(DEV)
    sd = {k: v.to(DEV) for k, v in O.synthetic_state_dict(sd_seed).items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k and not k.endswith(".pe"):
            v.requires_grad_(True)
    random.seed(mask_seed)
    pidx, cidx = O.draw_masks(sig.shape[0], nt, nt // 2, 2)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss, _, _ = O.pretrain_forward(x, sd, pidx.to(DEV), cidx.to(DEV), training=True)
    loss.backward()
    return {k: v.grad.float().cpu() for k, v in sd.items() if v.requires_grad}


def test_bf16_mode_loss_and_gradients():
    """bf16 activation storage (fp32 accumulation / statistics / master weights), tcgen05 GEMMs.  Gates: loss and prediction
    within 2e-2 of the fp32 oracle (north star).  Gradients: at this 3-clip x 16-frame size the CNN-stem gradients are
    chaotic in the bf16 rounding noise (two bf16 runs that differ only in accumulation order sit 6 % apart, torch
    autocast(bf16) itself is 4 % median / 14 % worst from fp32 - profiles/r01_bf16_calibration.txt), so the gate is relative to
    what torch autocast achieves: median <= 1.25 x, every tensor <= 2.5 x + 3e-2, and an absolute ceiling of 0.25."""
    nb, nt = 3, 16
    sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=5)
    m = build(nt, dtype=torch.bfloat16)
    loss, diff, vis = run_ours(m, sig, 11)
    assert m.engine.k.tc_launches > 50                                   # the tensor-core GEMM path really ran
    rl, rd, rvis, sd, taps = run_oracle(sig, nt, 7, 11)
    assert abs(float(loss) - float(rl)) < 2e-2 * float(rl)
    assert rel(vis["pred"].float().cpu(), rvis["pred"]) < 2e-2
    ac = _autocast_reference_grads(sig, nt, 7, 11)
    gmax = max(float(sd[k].grad.norm()) for k, _ in m.named_parameters())
    errs, autos = [], []
    for k, p in m.named_parameters():
        r = sd[k].grad.double()
        den = float(r.norm()) + 1e-3 * gmax
        mine = float((p.grad.detach().cpu().double() - r).norm()) / den
        auto = float((ac[k].double() - r).norm()) / den
        assert mine <= 2.5 * auto + 3e-2 and mine < 0.25, (k, mine, auto)
        errs.append(mine)
        autos.append(auto)
    errs.sort()
    autos.sort()
    assert errs[len(errs) // 2] <= 1.25 * autos[len(autos) // 2] + 5e-3, (errs[len(errs) // 2], autos[len(autos) // 2])


@pytest.mark.parametrize("nb,nt", [(2, 12), (1, 249)])
def test_bf16_ragged_frame_counts(nb, nt):
    """Frame counts that are not multiples of 8 / 128 (4.000 s clips give nt = 249): the tensor-core kernels clip
    their tiles, the attention GEMMs whose leading dimension is not 16-byte aligned take the CUDA-core kernel."""
    sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=4)
    m = build(nt, dtype=torch.bfloat16)
    loss, diff, vis = run_ours(m, sig, 13)
    rl, rd, rvis, sd, taps = run_oracle(sig, nt, 7, 13)
    assert abs(float(loss) - float(rl)) < 2e-2 * float(rl) and abs(float(diff) - float(rd)) < 1e-4 * float(rd)
    assert rel(vis["pred"].float().cpu(), rvis["pred"]) < 3e-2
    gdec = m.store.p("decoder.proj.2.weight").grad.cpu()
    assert rel(gdec, sd["decoder.proj.2.weight"].grad) < 5e-2


def test_dropout_statistics_and_train_step():
    """Dropout on: the loss stays finite, differs between steps (new masks), and a few fused-Adam steps reduce it."""
    nb, nt = 4, 16
    sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=8)
    m = build(nt)
    m.set_dropout(0.1)
    L = STFTLearner(m, 512, 0.5, 512, 1, 16000)
    L.device = DEV
    random.seed(3)
    first, _, _ = L.pretrain_epoch([[sig]] * 2, lr=1e-3, epoch=1)
    last, _, _ = L.pretrain_epoch([[sig]] * 6, lr=1e-3, epoch=2)
    assert np.isfinite(first) and np.isfinite(last) and last < first
    le, de, _ = L.pretest_epoch([[sig]])
    assert np.isfinite(le) and de > 0


def test_pretrain_evaluate_matches_oracle_math():
    """Evaluation tail (learner.py:574-618): iSTFT with zero DC, max normalisation, reconstruction errors."""
    nb, nt = 3, 16
    sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=5)
    m = build(nt)
    L = STFTLearner(m, 512, 0.5, 512, 1, 16000)
    L.device = DEV
    random.seed(11)
    le, de, vis, res = L.pretest_epoch([[sig]], return_diff=True, return_eval=True)
    pred, gt, mask = vis["pred"].cpu(), vis["tar"].cpu(), vis["mask"].cpu()           # (nb,nf,nt,2,2), (nb,nf,nt,2)

    def to_sig(t):                                                                    # the reference's recipe, on the CPU oracle
        z = torch.view_as_complex(t.permute(0, 1, 2, 4, 3).contiguous())              # (nb, nf, nt, nch)
        z = torch.cat((torch.zeros_like(z[:, 0:1]), z), dim=1)
        s = O.istft(z)
        return s / s.max()

    assert rel(res["sig_pred"].cpu(), to_sig(pred)) < 1e-4 and rel(res["sig_tar"].cpu(), to_sig(gt)) < 1e-4
    md = mask[:, :, :, None, :].expand(-1, -1, -1, 2, -1)
    d = (pred - gt) ** 2
    assert abs(float(res["mse"]) - float(d.mean())) < 1e-4 * float(d.mean())
    assert abs(float(res["mse_mask"]) - float((d * (1 - md)).sum() / (1 - md).sum())) < 1e-4 * float(res["mse_mask"])
    assert abs(float(res["mse_mask_ch"]) - float((d * (1 - md)).sum(4).mean())) < 1e-4 * float(res["mse_mask_ch"])
    assert res["pesq"].shape == (nb, 2)


@pytest.mark.parametrize("nt", [1024, 512, 264])
def test_long_clip_config5(nt):
    """BASELINE.json configs[4] shape (16.4 s clips, nt = 1024) and the other chunk counts of the score kernels (2 and 4 chunks of 256
    columns per lane, a ragged last chunk).  fp32 forward vs the CPU oracle (1e-4), then a bf16 forward + backward: loss within 2e-2
    of fp32, finite gradients."""
    nb = 1
    sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=21)
    m = build(nt)
    L = STFTLearner(m, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=1, fs=16000, task=None, ch_mode="M")
    x, = L.data_preprocess(sig.to(DEV))
    random.seed(3)
    with torch.no_grad():
        loss32, diff32, vis = m(x)
    sd = O.synthetic_state_dict(7)
    random.seed(3)
    pidx, cidx = O.draw_masks(nb, nt, nt // 2, 2)
    with torch.no_grad():
        rl, rd, rvis = O.pretrain_forward(O.preprocess(sig), sd, pidx, cidx, training=True)
    assert rel(vis["pred"].cpu(), rvis["pred"]) < 1e-4
    assert abs(float(loss32) - float(rl)) < 1e-4 * float(rl) and abs(float(diff32) - float(rd)) < 1e-4 * float(rd)
    mb = build(nt, dtype=torch.bfloat16)
    random.seed(3)
    loss16, _, _ = mb(x)
    loss16.backward()
    assert abs(float(loss16) - float(rl)) < 2e-2 * float(rl)
    for k, p in mb.named_parameters():
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()), k


def test_half_spectrum_model_nf128_matches_oracle():
    """fre_used_ratio 0.5 (learner.py:516-517): 128-bin spectrograms, SARSSL(sig_shape=(128, nt, 2, 2)) - loss and gradients against oracle autograd."""
    import random
    from sarssl_b200.learner import STFTLearner
    nb, nt, nf = 2, 16, 128
    sd = O.synthetic_state_dict(11, nf=nf)
    m = SARSSL(sig_shape=(nf, nt, 2, 2), patch_shape=(nf, 1), device=DEV)
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    m.to(DEV)
    m.set_dropout(0.0)
    m.train()
    L = STFTLearner(m, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=0.5, fs=16000, task=None, ch_mode="M")
    L.device = DEV
    sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=8)
    x, = L.data_preprocess(sig)
    assert x.shape == (nb, 2, nf, nt, 2)
    random.seed(3)
    loss, diff, vis = m(x)
    loss.backward()
    names = [k for k, v in sd.items() if v.is_floating_point() and "running" not in k and not k.endswith(".pe")]
    for k in names:
        sd[k].requires_grad_(True)
    random.seed(3)
    pidx, cidx = O.draw_masks(nb, nt, nt // 2, 2)
    lo, do_, _ = O.pretrain_forward(O.preprocess(sig, fre_used_ratio=0.5), sd, pidx, cidx, training=True)
    lo.backward()
    assert abs(float(loss) - float(lo)) <= 1e-4 * float(lo) and abs(float(diff) - float(do_)) <= 1e-4 * float(do_)
    gmax = max(float(sd[k].grad.norm()) for k in names)
    worst = max((float((m.store.p(k).grad.cpu().double() - sd[k].grad.double()).norm() / (sd[k].grad.double().norm() + 1e-4 * gmax)), k) for k in names)
    assert worst[0] < 1e-2, worst
