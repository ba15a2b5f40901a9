"""TEST INFRASTRUCTURE (CPU): second site experiment - where do the >2e-2 errors of the conformer conv-module gradients come from?
Rounding is injected (a) only in the CNN stem, (b) only at chosen tensors of the conformer conv module."""
import os, random, sys
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


SITES = {}


def site(x, name):
    f, b = SITES.get(name, (False, False))
    return Round.apply(x, f, b) if (f or b) else x


def cnn_stem(img, sd, pre, bn, taps=None):
    y = img
    for i, pad in ((0, 0), (3, 1), (6, 1), (9, 0)):
        y = F.conv2d(y, sd[f"{pre}.{i}.weight"], None, padding=pad)
        y = site(y, "stem_y")
        y = F.relu(O._bn(y, sd, f"{pre}.{i + 1}", bn))
        y = site(y, "stem_z")
    w = sd[f"{pre}.12.weight"]
    y = F.conv2d(y, w, None, stride=(w.shape[2], 1))
    return y[:, :, 0].transpose(1, 2)


def conv_module(x, sd, pre, bn, drop):
    D = x.shape[-1]
    h = O._ln(x, sd, pre + ".0").transpose(1, 2)
    h = site(h, "cm_h3")
    h = F.conv1d(h, sd[pre + ".2.conv.weight"], sd[pre + ".2.conv.bias"])
    h = site(h, "cm_g")
    h = h[:, :D] * torch.sigmoid(h[:, D:])
    h = site(h, "cm_ga")
    w = sd[pre + ".4.conv.weight"]
    h = F.conv1d(h, w, None, padding=(w.shape[-1] - 1) // 2, groups=D)
    h = site(h, "cm_cv")
    h = O._bn(h, sd, pre + ".5", bn)
    h = h * torch.sigmoid(h)
    h = site(h, "cm_z")
    h = F.conv1d(h, sd[pre + ".7.conv.weight"], sd[pre + ".7.conv.bias"])
    return drop(h).transpose(1, 2)


O.cnn_stem, O.conv_module = cnn_stem, conv_module
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 64
torch.set_num_threads(8)
x = O.preprocess(O.synthetic_waveforms(nb, (nt + 1) * 256, 2, seed=16))
c = "spat_encoder.embed.layers.2.sequential.2.module.sequential"
f = "spat_encoder.embed.layers.2.sequential.0.module.sequential"
keys = [c + ".0.weight", c + ".2.conv.weight", c + ".4.conv.weight", c + ".5.weight", c + ".7.conv.weight", f + ".1.linear.weight", f + ".4.linear.weight",
        "spat_encoder.patch_embed.6.weight", "decoder.proj.0.weight"]


def run(sites):
    SITES.clear(); SITES.update(sites)
    sd = O.synthetic_state_dict(7)
    for k in keys:
        sd[k].requires_grad_(True)
    random.seed(400000003)
    pidx, cidx = O.draw_masks(nb, nt, nt // 2, 2)
    loss, _, _ = O.pretrain_forward(x, sd, pidx, cidx, training=True)
    loss.backward()
    return {k: sd[k].grad.clone() for k in keys}


ref = run({})
T, Fa = True, False
cases = {"stem fwd only": {"stem_y": (T, Fa), "stem_z": (T, Fa)}, "cm fwd all": {k: (T, Fa) for k in ("cm_h3", "cm_g", "cm_ga", "cm_cv", "cm_z")},
         "cm bwd all": {k: (Fa, T) for k in ("cm_h3", "cm_g", "cm_ga", "cm_cv", "cm_z")},
         "cm fwd cv": {"cm_cv": (T, Fa)}, "cm fwd g": {"cm_g": (T, Fa)}, "cm fwd ga": {"cm_ga": (T, Fa)}, "cm fwd z": {"cm_z": (T, Fa)}, "cm fwd h3": {"cm_h3": (T, Fa)},
         "cm bwd z": {"cm_z": (Fa, T)}, "cm bwd cv": {"cm_cv": (Fa, T)}, "cm bwd ga": {"cm_ga": (Fa, T)}, "cm bwd g": {"cm_g": (Fa, T)}}
print(f"nb {nb} nt {nt}")
print(" " * 16 + "  ".join(k.split("sequential.")[-1][:12].rjust(12) if "layers" in k else k[-12:].rjust(12) for k in keys))
for name, sites in cases.items():
    got = run(sites)
    print(name.ljust(16) + "  ".join(f"{float((got[k] - ref[k]).norm() / ref[k].norm()):12.1e}" for k in keys), flush=True)
