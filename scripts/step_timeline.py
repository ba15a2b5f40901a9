#!/usr/bin/env python
"""Kernel timeline of the pre-training step through CUPTI (torch.profiler): per-kernel totals with real (not serialised)
durations, GPU busy time against the step's wall time, and the idle gaps that show where the host falls behind.

    python scripts/step_timeline.py [--batch 128] [--steps 2] [--out profiles/xxx.txt]

A number from this script is a diagnosis, never a bench value (CUPTI adds a few microseconds per launch on the host).
"""
import argparse
import collections
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from sarssl_b200 import ops  # noqa: E402
from sarssl_b200.learner import STFTLearner  # noqa: E402
from sarssl_b200.model import SARSSL  # noqa: E402
from sarssl_b200.optim import FusedAdam  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--nt", type=int, default=256, help="frames per clip (1024 = the long-clip configuration)")
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--out", default=None)
    ap.add_argument("--graph", action="store_true", help="replay the step as a CUDA graph when the engine supports it")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.manual_seed(1)
    model = SARSSL(sig_shape=(256, args.nt, 2, 2), device=dev)
    model.to(dev)
    model.set_compute_dtype(torch.bfloat16 if args.dtype == "bf16" else torch.float32)
    model.set_dropout(0.1)
    model.rng_state = ops.mt_seed(400000001)
    model.train()
    learner = STFTLearner(model, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=1, fs=16000, task=None, ch_mode="M")
    learner.device = dev
    sig = 0.1 * torch.randn(args.batch, (args.nt + 1) * 256, 2, device=dev)
    opt = FusedAdam(model, lr=1e-3)

    def step():
        x, = learner.data_preprocess(sig)
        loss, diff, _ = model(x)
        loss.backward()
        opt.step(1e-3, grad_scale=1.0, zero_grad=True)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    clean_ms = e0.elapsed_time(e1) / args.steps
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in ev), key=lambda t: t[0])
    agg = collections.OrderedDict()
    busy = 0.0
    gaps = collections.Counter()
    last_end, last_name = None, None
    for s, t, name in ks:
        name = re.sub(r"\(.*", "", name)
        name = re.sub(r"^void ", "", name).replace("sarssl::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t - s
        if last_end is not None and s > last_end:
            gaps[last_name] += s - last_end
        if last_end is None or t > last_end:
            busy += t - max(s, last_end if last_end is not None else s)
            last_end, last_name = t, name
    wall = ks[-1][1] - ks[0][0]
    n = args.steps
    out = ["batch %d, %d steps: step without profiler %.2f ms; under CUPTI wall %.2f ms/step, GPU busy %.2f ms/step (%.1f %% idle)"
           % (args.batch, n, clean_ms, wall / 1e3 / n, busy / 1e3 / n, 100 * (1 - busy / wall)),
           "%-84s %8s %11s %7s" % ("kernel", "launches", "us/step", "share")]
    tot = sum(a[1] for a in agg.values())
    for name, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%-84s %8d %11.1f %6.1f%%" % (name[:84], c // n if c % n == 0 else c, us / n, 100 * us / tot))
    out.append("%-84s %8d %11.1f" % ("TOTAL (sum of kernel durations)", sum(a[0] for a in agg.values()) // n, tot / n))
    out.append("")
    out.append("idle time following each kernel (us/step, top 12):")
    for name, g in gaps.most_common(12):
        out.append("  %-82s %11.1f" % (name[:82], g / n))
    txt = "\n".join(out)
    print(txt)
    if args.out:
        with open(args.out, "w") as f:
            f.write(txt + "\n")


if __name__ == "__main__":
    main()
