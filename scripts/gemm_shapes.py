#!/usr/bin/env python
"""Per-shape timing of every GEMM the pre-training step launches (diagnosis only: each call is bracketed by CUDA events and a sync).

    python scripts/gemm_shapes.py [--batch 128] [--out profiles/xxx.txt]
"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sarssl_b200 import kernels, ops  # noqa: E402
from sarssl_b200.learner import STFTLearner  # noqa: E402
from sarssl_b200.model import SARSSL  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.manual_seed(1)
    model = SARSSL(sig_shape=(256, 256, 2, 2), device=dev)
    model.to(dev)
    model.set_compute_dtype(torch.bfloat16)
    model.set_dropout(0.1)
    model.rng_state = ops.mt_seed(400000001)
    model.train()
    learner = STFTLearner(model, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=1, fs=16000, task=None, ch_mode="M")
    learner.device = dev
    sig = 0.1 * torch.randn(args.batch, 65792, 2, device=dev)

    def step():
        x, = learner.data_preprocess(sig)
        loss, diff, _ = model(x)
        loss.backward()

    for _ in range(2):
        step()
    rec = collections.OrderedDict()
    orig = kernels.KernelSet.gemm

    def timed(self, A, B, Cmat, M, N, K, sA, sB, ldc, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(self, A, B, Cmat, M, N, K, sA, sB, ldc, **kw)
        e1.record()
        torch.cuda.synchronize()
        nb = kw.get("batch", (1, 1))
        key = (M, N, K, nb[0] * nb[1], "A_mn" if sA[0] == 1 and M > 1 else "A_k", "B_n" if sB[0] == 1 and N > 1 else "B_k",
               str(Cmat.dtype).replace("torch.", ""), "acc" if kw.get("accumulate") else "", "epi" if (kw.get("bias") is not None or kw.get("resid") is not None or kw.get("act", 0)) else "")
        r = rec.setdefault(key, [0, 0.0])
        r[0] += 1
        r[1] += e0.elapsed_time(e1)

    kernels.KernelSet.gemm = timed
    step()
    kernels.KernelSet.gemm = orig
    out = ["%-7s %-6s %-7s %-6s %-5s %-4s %-9s %-4s %-4s %6s %10s %10s" % ("M", "N", "K", "batch", "A", "B", "C", "acc", "epi", "calls", "ms total", "TFLOP/s")]
    tot = 0.0
    for key, (n, ms) in sorted(rec.items(), key=lambda kv: -kv[1][1]):
        M, N, K, nb = key[:4]
        tf = 2.0 * M * N * K * nb * n / (ms * 1e-3) / 1e12
        tot += ms
        out.append("%-7d %-6d %-7d %-6d %-5s %-4s %-9s %-4s %-4s %6d %10.3f %10.1f" % (*key, n, ms, tf))
    out.append("total %.3f ms" % tot)
    txt = "\n".join(out)
    print(txt)
    if args.out:
        open(args.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
