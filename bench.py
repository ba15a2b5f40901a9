#!/usr/bin/env python
"""Benchmark of the SAR-SSL pre-training hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py --gpus 1 --steps 20 --warmup 5                      # our arm (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1      # reference arm: CPU restatement on host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...    # one rank per GPU, weak scaling over clips

One JSON line on stdout (rank 0).  metric = clips/s (BASELINE.json); a "step" is one pass of the hot path over one
batch of synthetic clips.  `value` is measured with inputs resident in HBM; `e2e` goes through the reference-facing
Python API from pinned host memory (H2D of the waveforms and D2H of the loss inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NSAMPLE = 65792            # 4.112 s @ 16 kHz -> 256 frames (opt.py:19, run_pretrain.py:67-72)
STFT_BYTES_PER_CLIP = NSAMPLE * 2 * 4 + 2 * 256 * 256 * 2 * 4           # 1,574,912 B
LOSS_BYTES_PER_CLIP = 128 * 256 * 2 * 4 + 128 * 256 * 4 * 4 + 256 * 1024 * 4   # 1,835,008 B fp32 pred
FLOP_PER_CLIP = 86.53e9    # fwd+bwd, 3 x forward convention


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d.get("bf16_tflops", 1590.0)),
                    "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", 1400.0)), "source": "measured"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons}


# --------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU restatement of the reference (oracle/) on the host cores
# --------------------------------------------------------------------------------------------------------------

def cpu_frontend_loss(nb, steps, warmup, threads):
    import torch
    from oracle import sarssl_oracle as O
    torch.set_num_threads(threads)
    sig = O.synthetic_waveforms(nb, NSAMPLE, 2, seed=1234)
    pred = torch.randn(nb, 256, 1024, generator=torch.Generator().manual_seed(1)).requires_grad_(True)
    import random
    random.seed(400000001)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        x = O.preprocess(sig)
        vec = x.permute(0, 3, 2, 4, 1)
        pidx, cidx = O.draw_masks(nb, 256, 128, 2)
        loss, diff = O.masked_loss(pred.view(nb, 256, 256, 2, 2), vec, pidx, cidx)
        loss.backward()
        pred.grad = None
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return nb * len(times) / sum(times)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    nb = 64
    v = cpu_frontend_loss(nb, args.steps, args.warmup, cores)
    sample = f"{nb} clips/step x {args.steps} steps of the frontend+loss workload (oracle port of the reference, torch CPU fp32)"
    line = {"impl": "reference", "metric": "pretrain_clips_per_s", "value": v, "unit": "clips/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * nb / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, per_gpu_batch=nb),
            "cpu_baseline": {"value": v, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, per_gpu_batch):
    return {"workload": "stft_frontend+masked_recon_loss fwd+bwd, 2-mic, 65792 samples (4.112 s @16 kHz), "
                        f"batch {per_gpu_batch}/GPU (BASELINE.json configs[1])",
            "per_gpu_batch": per_gpu_batch, "nsample": NSAMPLE, "nmic": 2, "nt": 256, "nf": 256,
            "l2": "inputs (539 MB/step) larger than the 126 MB L2; no explicit flush", "parallelism": f"dp{args.gpus}"}


# --------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist
    from sarssl_b200 import ops
    from sarssl_b200.learner import STFTLearner
    from sarssl_b200.model import MaskedReconLoss

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nb = args.batch
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    sig = 0.1 * torch.randn(nb, NSAMPLE, 2, device=dev, generator=g)
    pred = torch.randn(nb, 256, 1024, device=dev, generator=g)
    state = ops.mt_seed(400000001 + rank)
    host_sig = torch.empty(nb, NSAMPLE, 2, dtype=torch.float32).pin_memory()
    host_sig.copy_(sig)

    def draw():
        pidx, cidx, flag = ops.draw_masks(state, nb, 256, 128, 2)
        return torch.from_numpy(flag).pin_memory().to(dev, non_blocking=True), \
            torch.from_numpy(cidx.astype("int32")).pin_memory().to(dev, non_blocking=True)

    patches = torch.empty(nb, 256, 256, 2, 2, device=dev)
    dpred = torch.empty_like(pred)
    out2 = torch.empty(2, device=dev)

    def step_resident():
        flag, cidx = draw()
        ops.stft_frontend(sig, out=patches)
        ops.masked_loss(pred, patches, flag, cidx, 128, out2=out2, dpred=dpred)

    learner = STFTLearner(None, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=1, fs=16000, task=None, ch_mode="M")
    learner.device = dev
    loss_mod = MaskedReconLoss(nmasked_patch=128, npatch=256, device=dev, rng_state=state)

    def step_e2e():
        x, = learner.data_preprocess(host_sig)                  # H2D of the waveforms happens inside (learner.py:533)
        loss, diff, _ = loss_mod(pred, x)
        return float(loss)                                       # D2H of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(step_resident, args.steps)
    sampler.stop_flag = True
    ops.stft_frontend_check(dev)
    clips = nb * world * args.steps
    value = clips / (ms * 1e-3)

    # dominant kernel alone (the fused STFT front-end), CUDA events on the launching stream
    flag, cidx = draw()
    for _ in range(3):
        ops.stft_frontend(sig, out=patches)
    k_ms = timed(lambda: ops.stft_frontend(sig, out=patches), args.steps) / args.steps
    l_ms = timed(lambda: ops.masked_loss(pred, patches, flag, cidx, 128, out2=out2, dpred=dpred), args.steps) / args.steps
    pk = peaks()
    achieved = nb * STFT_BYTES_PER_CLIP / (k_ms * 1e-3) / 1e9
    loss_gbs = nb * LOSS_BYTES_PER_CLIP / (l_ms * 1e-3) / 1e9

    for _ in range(2):
        step_e2e()
    e2e_steps = max(2, min(args.steps, 10))
    e2e_ms = timed(step_e2e, e2e_steps)
    e2e_value = nb * world * e2e_steps / (e2e_ms * 1e-3)

    if rank == 0:
        cores = os.cpu_count() or 1
        cpu_nb = 64
        cpu_v = cpu_frontend_loss(cpu_nb, 3, 1, cores) if world == 1 else None
        line = {"metric": "pretrain_clips_per_s", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, nb),
                "clocks": sampler.summary(),
                "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": int(host_sig.numel() * 4 + nb * 256 + nb * 4),
                        "d2h_bytes_per_step": 4},
                "gpu_launches": 2 * args.steps,
                "roofline": {"bound": "hbm", "kernel": "stft_frontend_fused_kernel", "achieved": achieved, "peak": pk["hbm_gbs"],
                             "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": None, "peak_source": pk["source"],
                             "kernel_ms": k_ms, "algorithmic_bytes_per_launch": nb * STFT_BYTES_PER_CLIP,
                             "secondary": {"kernel": "masked_loss_kernel", "achieved": loss_gbs, "frac": loss_gbs / pk["hbm_gbs"],
                                           "kernel_ms": l_ms, "algorithmic_bytes_per_launch": nb * LOSS_BYTES_PER_CLIP}}}
        if cpu_v is not None:
            line["cpu_baseline"] = {"value": cpu_v, "unit": "clips/s", "cores": cores, "kind": "port",
                                    "sample": f"{cpu_nb} clips/step x 3 steps of the same workload, oracle port (torch CPU fp32)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="clips per GPU per step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
