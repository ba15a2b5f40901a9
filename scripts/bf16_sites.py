"""TEST INFRASTRUCTURE (CPU): which bf16-rounded tensor of the CNN stem produces the gradient error of the bf16 mode?
The oracle's stem is re-run with bf16 rounding injected at ONE kind of site at a time (forward value and / or the gradient flowing back
through it) and the stem weight gradients are compared with the un-rounded run.

    python scripts/bf16_sites.py [nb] [nt]
"""
import os
import random
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sarssl_oracle as O  # noqa: E402


class Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fwd, bwd):
        ctx.bwd = bwd
        return x.bfloat16().float() if fwd else x

    @staticmethod
    def backward(ctx, g):
        return (g.bfloat16().float() if ctx.bwd else g), None, None


SITES = {}          # name -> (fwd, bwd)


def site(x, name):
    f, b = SITES.get(name, (False, False))
    return Round.apply(x, f, b) if (f or b) else x


def cnn_stem(img, sd, pre, bn, taps=None):
    y = img
    for i, pad in ((0, 0), (3, 1), (6, 1), (9, 0)):
        y = F.conv2d(y, sd[f"{pre}.{i}.weight"], None, padding=pad)
        y = site(y, f"y{i}")                   # conv output = BatchNorm input (stored bf16; its gradient dy is the wgrad / dgrad operand)
        y = O._bn(y, sd, f"{pre}.{i + 1}", bn)
        y = F.relu(y)
        y = site(y, f"z{i}")                   # BatchNorm + ReLU output (stored bf16; its gradient dz is the dgrad conv's output)
    w = sd[f"{pre}.12.weight"]
    y = F.conv2d(y, w, None, stride=(w.shape[2], 1))
    return y[:, :, 0].transpose(1, 2)


O.cnn_stem = cnn_stem
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 256
torch.set_num_threads(8)
sig = O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=16)
x = O.preprocess(sig)
keys = [f"spec_encoder.patch_embed.{i}.weight" for i in (0, 1, 3, 4, 6, 7, 9, 10, 12)] + [f"spat_encoder.patch_embed.{i}.weight" for i in (3, 6)]


def run(sites):
    SITES.clear()
    SITES.update(sites)
    sd = O.synthetic_state_dict(7)
    for k in keys:
        sd[k].requires_grad_(True)
    random.seed(400000003)
    pidx, cidx = O.draw_masks(nb, nt, nt // 2, 2)
    loss, _, _ = O.pretrain_forward(x, sd, pidx, cidx, training=True)
    loss.backward()
    return {k: sd[k].grad.clone() for k in keys}


ref = run({})
cases = {"all fwd y": {f"y{i}": (True, False) for i in (0, 3, 6, 9)}, "all fwd z": {f"z{i}": (True, False) for i in (0, 3, 6, 9)},
         "all bwd dy": {f"y{i}": (False, True) for i in (0, 3, 6, 9)}, "all bwd dz": {f"z{i}": (False, True) for i in (0, 3, 6, 9)},
         "bwd dz6 only": {"z6": (False, True)}, "bwd dz3 only": {"z3": (False, True)}, "bwd dz0 only": {"z0": (False, True)},
         "bwd dy6 only": {"y6": (False, True)}, "bwd dy3 only": {"y3": (False, True)}, "bwd dy9 only": {"y9": (False, True)},
         "everything": {f"{a}{i}": (True, True) for a in "yz" for i in (0, 3, 6, 9)}}
print(f"nb {nb} nt {nt}: norm-wise relative error of the stem weight gradients per rounding site")
print(" " * 16 + "  ".join(k.replace("_encoder.patch_embed", "").replace(".weight", "")[:8].rjust(8) for k in keys))
for name, sites in cases.items():
    got = run(sites)
    print(name.ljust(16) + "  ".join(f"{float((got[k] - ref[k]).norm() / ref[k].norm()):8.1e}" for k in keys), flush=True)
