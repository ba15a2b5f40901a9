#!/usr/bin/env python
"""Time the memory-bound stem kernels alone (CUDA events, inputs far larger than L2) and print their effective HBM bandwidth.

    python scripts/stem_bench.py [--clips 128] [--only name]

Algorithmic bytes: every tensor the kernel must read or write once (bf16 activations, [P][64] "wide", [P][4] "narrow").
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sarssl_b200.kernels import KernelSet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=128)
    ap.add_argument("--only", default=None)
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    k = KernelSet(dev, torch.bfloat16)
    B, T, F = args.clips, 256, 256
    P = B * T * F
    g = torch.Generator(device=dev).manual_seed(1)
    wide = [torch.randn(P, 64, device=dev, generator=g).to(torch.bfloat16) for _ in range(3)]
    narrow = torch.randn(P, 4, device=dev, generator=g).to(torch.bfloat16)
    narrow2 = torch.empty_like(narrow)
    gamma, beta = torch.ones(64, device=dev), torch.zeros(64, device=dev)
    rm, rv, nbt = torch.zeros(64, device=dev), torch.ones(64, device=dev), torch.zeros(1, dtype=torch.long, device=dev)
    stats = k.bn_stats(wide[0], P, 64, gamma, beta, rm, rv, nbt, True)
    dg, db = torch.zeros(64, device=dev), torch.zeros(64, device=dev)
    w64x4 = torch.randn(64, 4, device=dev, generator=g) * 0.3
    w4x64 = torch.randn(4, 64, device=dev, generator=g) * 0.1
    dw = torch.zeros(64, 4, device=dev)
    W, N = 2.0 * P * 64, 2.0 * P * 4
    cases = [
        ("bn_stats (reduce<0>)", lambda: k.bn_stats(wide[0], P, 64, gamma, beta, rm, rv, nbt, True), W),
        ("bn_act_fwd", lambda: k.bn_act_fwd(wide[0], stats, 1, wide[1], P, 64), 2 * W),
        ("bn_act_bwd (reduce<1> + apply, in place)", lambda: k.bn_act_bwd(wide[2], wide[0], stats, 1, wide[2], dg, db, P, 64), 5 * W),
        ("stem_expand (narrow -> wide)", lambda: k.stem_expand(narrow, 0, None, None, w64x4, wide[1], P, F, T), W + N),
        ("stem_reduce (bn+relu(wide) -> narrow)", lambda: k.stem_reduce(wide[0], stats, w4x64, narrow2, P), W + N),
        ("stem_pw_wgrad (bn+relu(wide), narrow)", lambda: k.stem_pw_wgrad(wide[0], stats, narrow, 0, None, None, dw, False, P, F, T), W + N),
        ("stem_pw_wgrad (wide, narrow)", lambda: k.stem_pw_wgrad(wide[0], None, narrow, 0, None, None, dw, False, P, F, T), W + N),
    ]
    z = torch.relu(wide[1].float()).to(torch.bfloat16)
    patches = torch.randn(P, 4, device=dev, generator=g)
    dgam, dbet, dw2 = torch.zeros(64, device=dev), torch.zeros(64, device=dev), torch.zeros(4, 64, device=dev)
    cases += [
        ("stem_head_bwd (dz, z, fp32 patches)", lambda: k.stem_head_bwd(wide[2], z, stats, patches, 3, None, None, w64x4, dgam, dbet, dw, P, F, T), 2 * W + 2 * N),
        ("stem_tail_bwd (y, dq -> dy)", lambda: k.stem_tail_bwd(wide[0], stats, narrow, w4x64, dgam, dbet, dw2, wide[1], P), 3 * W + 2 * N),
    ]
    for name, fn, nbytes in cases:
        if args.only and args.only not in name:
            continue
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        print("%-44s %8.3f ms  %7.0f GB/s  (%.2f GB)" % (name, ms, nbytes / ms / 1e6, nbytes / 1e9))


if __name__ == "__main__":
    main()
