#!/usr/bin/env python
"""BASELINE.json configs[4] on one GPU: 16.4 s clips (262,400 samples, nt = 1024), 32 clips per GPU, bf16 pre-training step.
    python scripts/long_clip_bench.py [--batch 32] [--steps 5]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sarssl_b200 import ops  # noqa: E402
from sarssl_b200.learner import STFTLearner  # noqa: E402
from sarssl_b200.model import SARSSL  # noqa: E402
from sarssl_b200.optim import FusedAdam  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--nt", type=int, default=1024)
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(1)
model = SARSSL(sig_shape=(256, args.nt, 2, 2), device=dev)
model.to(dev)
model.set_compute_dtype(torch.bfloat16)
model.set_dropout(0.1)
model.rng_state = ops.mt_seed(400000001)
model.train()
L = STFTLearner(model, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=1, fs=16000, task=None, ch_mode="M")
L.device = dev
sig = 0.1 * torch.randn(args.batch, (args.nt + 1) * 256, 2, device=dev)
opt = FusedAdam(model, lr=1e-3)


def step():
    x, = L.data_preprocess(sig)
    loss, _, _ = model(x)
    loss.backward()
    opt.step(1e-3, grad_scale=1.0, zero_grad=True)
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
print("nt=%d batch=%d: %.2f ms/step, %.1f clips/s (%.1f s of audio per clip), loss %.4f" % (args.nt, args.batch, ms, args.batch / ms * 1e3, (args.nt + 1) * 256 / 16000, float(loss)))
