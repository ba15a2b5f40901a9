"""GPU parity: STFT / data_preprocess / iSTFT / masked-loss kernels (through the C ABI) against the CPU oracle and the
reference-generated fixtures.  Tolerances: fp32 forward 1e-4 relative (BASELINE.json north_star)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import sarssl_oracle as O
from sarssl_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def to_ref_layout(patches):            # (nb, nt, nf, 2, 2) [t,f,r,m] -> reference (nb, 2[m], nf, nt, 2[r])
    return patches.permute(0, 4, 2, 1, 3)


@pytest.mark.parametrize("nb,nsample,nch", [(1, 512, 2), (2, 4352, 2), (3, 2 * 256 + 511, 2), (2, 9 * 256 + 1, 1), (2, 2560, 3),
                                            (1, 5000, 5), (5, 65792, 2)])
def test_stft_spectrum_matches_oracle(nb, nsample, nch):
    sig = O.synthetic_waveforms(nb, nsample, nch, seed=nb + nch)
    want = O.stft(sig)                                         # (nb, nf, nt, nch)
    got = ops.stft_spectrum(sig.to(DEV)).permute(0, 2, 1, 3).cpu()
    assert got.shape == want.shape
    assert (got - want).abs().max() <= 1e-4 * want.abs().max()
    assert rel(torch.view_as_real(got), torch.view_as_real(want)) < 1e-5


def test_stft_matches_reference_fixture():
    g = np.load(os.path.join(GOLDEN, "tiny_nt16_b3.npz"))
    sig = O.synthetic_waveforms(int(g["nb"]), (int(g["nt"]) + 1) * 256, 2, seed=int(g["sig_seed"]))
    ref = torch.complex(torch.from_numpy(g["stft_re"]), torch.from_numpy(g["stft_im"]))
    got = ops.stft_spectrum(sig.to(DEV)).permute(0, 2, 1, 3).cpu()
    assert (got - ref).abs().max() <= 1e-4 * ref.abs().max()
    x = to_ref_layout(ops.stft_frontend(sig.to(DEV))).cpu()
    xr = torch.from_numpy(g["x"])
    assert x.shape == xr.shape and rel(x, xr) < 1e-5
    assert (x - xr).abs().max() <= 1e-4 * xr.abs().max()


@pytest.mark.parametrize("nb,nsample,nch,generic", [(1, 512, 2, 0), (3, 4352, 2, 0), (3, 4352, 2, 1), (2, 65792, 2, 0), (2, 65792, 2, 1), (4, 64000, 2, 0),
                                                    (2, 2 * 256 + 511, 2, 0), (2, 4352, 3, 0), (1, 3000, 4, 0), (1, 262400, 2, 0), (3, 10 * 256, 2, 0),
                                                    (3, 4352, 2, 4), (2, 65792, 2, 4), (4, 64000, 2, 4), (1, 262400, 2, 4), (1, 512, 2, 4), (5, 7 * 256 + 512, 2, 4),
                                                    (1, 512, 2, 5), (3, 4352, 2, 5), (2, 65792, 2, 5), (4, 64000, 2, 5), (1, 262400, 2, 5), (40, 65792, 2, 5),
                                                    (5, 7 * 256 + 512, 2, 5), (2, 2 * 256 + 511, 2, 5)])
def test_frontend_matches_oracle(nb, nsample, nch, generic):
    sig = O.synthetic_waveforms(nb, nsample, nch, seed=17 + nsample % 97)
    want = O.preprocess(sig)                                   # (nb*(nch-1), 2, 256, nt, 2)
    got = to_ref_layout(ops.stft_frontend(sig.to(DEV), force_generic=generic)).cpu()
    ops.stft_frontend_check(DEV)
    assert got.shape == want.shape
    assert rel(got, want) < 1e-5
    assert (got - want).abs().max() <= 1e-4 * want.abs().max()


def test_frontend_full_batch_properties():
    """config 2 size (batch 1024): spot-check clips against the oracle + size-independent properties."""
    nb, nsample = 1024, 65792
    g = torch.Generator(device=DEV).manual_seed(3)
    sig = 0.1 * torch.randn(nb, nsample, 2, device=DEV, generator=g)
    out = ops.stft_frontend(sig)
    ops.stft_frontend_check(DEV)
    assert out.shape == (nb, 256, 256, 2, 2) and bool(torch.isfinite(out).all())
    for b in (0, 511, 1023):
        want = O.preprocess(sig[b:b + 1].cpu())
        assert rel(to_ref_layout(out[b:b + 1]).cpu(), want) < 1e-5
    # scale invariance (the front-end divides by the mean magnitude): f(3x) == f(x) up to eps = 1e-6
    out3 = ops.stft_frontend(3.0 * sig)
    assert rel(out3, out) < 1e-4
    # determinism: the clip reduction is order-fixed
    assert torch.equal(ops.stft_frontend(sig), out)
    # clip independence: permuting the batch permutes the output
    perm = torch.randperm(nb, device=DEV)
    assert torch.equal(ops.stft_frontend(sig[perm].contiguous()), out[perm])


@pytest.mark.parametrize("nb,nt,nch", [(2, 16, 2), (1, 1, 2), (2, 9, 3), (1, 256, 2)])
def test_istft_matches_oracle(nb, nt, nch):
    sig = O.synthetic_waveforms(nb, (nt + 1) * 256, nch, seed=nt)
    S = O.stft(sig)                                            # reference layout (nb, nf, nt, nch)
    want = O.istft(S)
    got_ref_layout = ops.istft(S.to(DEV)).cpu()                # strided read of the reference layout
    assert got_ref_layout.shape == want.shape
    assert (got_ref_layout - want).abs().max() <= 1e-5 * max(1.0, float(want.abs().max()))
    ours = ops.stft_spectrum(sig.to(DEV)).permute(0, 2, 1, 3)  # frame-major storage, reference indexing
    got = ops.istft(ours).cpu()
    assert (got - want).abs().max() <= 1e-5 * max(1.0, float(want.abs().max()))


def _loss_case(nb, nt, nf, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    vec = torch.randn(nb, nt, nf, 2, 2, generator=g)
    pred = torch.randn(nb, nt, nf * 4, generator=g)
    state = ops.mt_seed(seed)
    pidx, cidx, flag = ops.draw_masks(state, nb, nt, nt // 2, 2)
    pred_q = pred.to(dtype).float()
    p = pred_q.clone().requires_grad_(True)
    loss, diff = O.masked_loss(p.view(nb, nt, nf, 2, 2), vec, torch.from_numpy(pidx), torch.from_numpy(cidx)[:, None])
    loss.backward()
    out2, dpred = ops.masked_loss(pred.to(DEV).to(dtype), vec.to(DEV), torch.from_numpy(flag).to(DEV),
                                  torch.from_numpy(cidx).to(torch.int32).to(DEV), nt // 2)
    return loss, diff, p.grad, out2.cpu(), dpred.float().cpu()


@pytest.mark.parametrize("nb,nt,nf", [(2, 16, 256), (3, 7, 256), (1, 2, 64), (8, 256, 256)])
def test_masked_loss_fp32(nb, nt, nf):
    if nt // 2 == 0:
        pytest.skip("no masked frame")
    loss, diff, grad, out2, dpred = _loss_case(nb, nt, nf, 5 + nt, torch.float32)
    assert abs(float(out2[0]) - float(loss)) <= 1e-5 * float(loss)
    assert abs(float(out2[1]) - float(diff)) <= 1e-5 * float(diff)
    assert rel(dpred, grad) < 1e-5
    assert int((dpred != 0).sum()) == int((grad != 0).sum())       # exact sparsity pattern: masked frame x masked channel


def test_masked_loss_bf16():
    loss, diff, grad, out2, dpred = _loss_case(4, 32, 256, 9, torch.bfloat16)
    assert abs(float(out2[0]) - float(loss)) <= 1e-4 * float(loss)
    assert rel(dpred, grad) < 1e-2                                   # bf16 rounding of the stored gradient


@pytest.mark.parametrize("ch_mode,ratio,nch,nb,nsample", [("MM", 1, 4, 2, 9 * 256), ("M", 0.5, 3, 2, 9 * 256), ("MM", 0.5, 3, 3, 4352), ("MM", 1, 2, 2, 4352),
                                                           ("M", 0.5, 2, 2, 65792)])
def test_learner_preprocess_variants_match_oracle(ch_mode, ratio, nch, nb, nsample):
    """STFTLearner.data_preprocess off the pre-training configuration: ch_mode 'MM' (all microphone pairs, utils_module.py:136-143) and
    fre_used_ratio 0.5 (bins 0..127, learner.py:516-517); the oracle's versions are checked against the real reference in
    tests/test_oracle_vs_reference.py."""
    from sarssl_b200.learner import STFTLearner
    from sarssl_b200.model import SARSSL
    m = SARSSL(sig_shape=(256, 16, 2, 2), device=DEV)
    L = STFTLearner(m, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=ratio, fs=16000, task=None, ch_mode=ch_mode)
    L.device = DEV
    sig = O.synthetic_waveforms(nb, nsample, nch, seed=41)
    got, = L.data_preprocess(sig)
    want = O.preprocess(sig, ch_mode=ch_mode, fre_used_ratio=ratio)
    assert got.shape == want.shape
    assert rel(got.cpu(), want) < 1e-5 and (got.cpu() - want).abs().max() <= 1e-4 * want.abs().max()
