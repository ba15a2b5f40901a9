#!/usr/bin/env python
"""Benchmark of the SAR-SSL pre-training hot path on B200 (contract: task statement / DESIGN.md section 6).

    python bench.py --gpus 1 --steps 5 --warmup 3                       # our arm (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1      # reference arm: CPU restatement on the host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...    # one rank per GPU, weak scaling over clips

One JSON line on stdout (rank 0).  metric = pre-training clips/s (BASELINE.json); a "step" is one full training step
(STFT front-end -> masks -> MC-Conformer forward -> masked reconstruction loss -> backward -> Adam) over one micro-batch of
synthetic 2-microphone clips per GPU.  `value` has the waveforms resident in HBM; `e2e` goes through the reference-facing
`STFTLearner.pretrain_epoch` from pinned host memory (H2D of the waveforms + D2H of the loss inside the timed region).
The front-end + loss sub-path (BASELINE.json configs[1]) is reported in the same line under "frontend"; the other configurations of
BASELINE.json (fine-tuning step configs[3], long clips configs[4]) ride along as short sub-runs under "other_configs" (or run alone:
--workload finetune | longclip); "torch_eager_b200" is the real bar on the same GPU: the reference algorithm in plain PyTorch eager
(cuDNN / cuBLAS / cuFFT), fp32+TF32 as shipped (run_pretrain.py:33-34) and under autocast(bf16).
--accum-steps A runs BASELINE.json configs[2] literally (global batch = micro-batch x A x GPUs, one all-reduce + Adam per A micro-batches).
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NSAMPLE = 65792            # 4.112 s @ 16 kHz -> 256 frames (opt.py:19, run_pretrain.py:67-72)
NT, NF = 256, 256
STFT_BYTES_PER_CLIP = NSAMPLE * 2 * 4 + 2 * 256 * 256 * 2 * 4                  # 1,574,912 B
LOSS_BYTES_PER_CLIP = 128 * 256 * 2 * 4 + 128 * 256 * 4 * 4 + 256 * 1024 * 4   # 1,835,008 B fp32 pred
FLOP_PER_CLIP = 86.53e9                                                          # fwd+bwd, 3 x forward convention
CONV3_FLOP_PER_CLIP = 2.0 * 256 * 256 * 64 * 576                                # one 3x3 64->64 conv over a 256x256 map


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


def workload_config(world, per_gpu_batch, dtype, accum_steps=1):
    return {"workload": "MC-Conformer pre-training step (STFT front-end + masks + fwd + masked recon loss + bwd + Adam), 2-mic, 65792 samples "
                        f"(4.112 s @16 kHz), micro-batch {per_gpu_batch}/GPU, {dtype} (BASELINE.json configs[2])",
            "per_gpu_batch": per_gpu_batch, "accum_steps": accum_steps, "global_batch": per_gpu_batch * world * accum_steps, "nsample": NSAMPLE, "nmic": 2,
            "nt": NT, "nf": NF, "dropout": 0.1, "optimizer": "Adam",
            "l2": "activations per step (tens of GB) far exceed the 126 MB L2; no explicit flush", "parallelism": f"dp{world}"}


def measured_traffic(kernel, clips):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json,
    written by scripts/ncu_summary.py from the .ncu-rep of the same kernel build), scaled to `clips` per launch; None when no capture is recorded."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kernel]
        return rec["dram_bytes_per_clip"] * clips, rec["source"]
    except Exception:
        return None, "no ncu capture recorded for this kernel build (profiles/traffic.json)"


# --------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU restatement of the reference (oracle/) on the host cores
# --------------------------------------------------------------------------------------------------------------

def cpu_pretrain_clips_per_s(nb, steps, warmup, threads):
    """Full training step of the oracle (= the reference's algorithm: STFT, masks, fwd, loss, autograd bwd, Adam) on CPU, dropout on."""
    import torch
    from oracle import sarssl_oracle as O
    torch.set_num_threads(threads)
    sig = O.synthetic_waveforms(nb, NSAMPLE, 2, seed=1234)
    sd = O.synthetic_state_dict(7)
    names = [k for k, v in sd.items() if v.is_floating_point() and "running" not in k and not k.endswith(".pe")]
    for k in names:
        sd[k].requires_grad_(True)
    m = [torch.zeros_like(sd[k]) for k in names]
    v = [torch.zeros_like(sd[k]) for k in names]
    random.seed(400000001)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        x = O.preprocess(sig)
        pidx, cidx = O.draw_masks(nb, NT, NT // 2, 2)
        loss, diff, _ = O.pretrain_forward(x, sd, pidx, cidx, training=True, dropout_p=0.1)
        loss.backward()
        with torch.no_grad():
            O.adam_step([sd[k] for k in names], [sd[k].grad for k in names], m, v, i + 1, 1e-3)
        for k in names:
            sd[k].grad = None
        float(loss)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return nb * len(times) / sum(times)


def reference_pretrain_clips_per_s(nb, steps, warmup, threads):
    """The REAL reference (`STFTLearner.pretrain_epoch`, code/learner.py:76-131, per-item Python loops and all) on the host cores, when its
    source tree is importable (build container: /root/reference/code; it cannot travel to the GPU box).  None when absent."""
    try:
        from oracle import ref_shim
        if ref_shim.reference_root() is None:
            return None
        import torch
        rm, rl, _, ru = ref_shim.load_reference()
    except Exception:
        return None
    from oracle import sarssl_oracle as O
    torch.set_num_threads(threads)
    ru.set_seed(1)
    net = rm.SARSSL(sig_shape=(NF, NT, 2, 2), pretrain=True, device="cpu")
    L = rl.STFTLearner(net, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=1, fs=16000, task=None, ch_mode="M")
    L.cpu()
    sig = O.synthetic_waveforms(nb, NSAMPLE, 2, seed=1234)
    L.pretrain_epoch([[sig]] * warmup, lr=1e-3, epoch=1)
    t0 = time.perf_counter()
    L.pretrain_epoch([[sig]] * steps, lr=1e-3, epoch=1)
    return nb * steps / (time.perf_counter() - t0)


def cpu_arm(nb, steps, warmup, cores):
    """(clips/s, kind, sample): the real reference when importable (kind "reference"), else the oracle port (kind "port")."""
    v = reference_pretrain_clips_per_s(nb, steps, warmup, cores)
    if v is not None:
        return v, "reference", (f"{nb} clips/step x {steps} steps (+{warmup} warm-up) of the unmodified reference's STFTLearner.pretrain_epoch "
                                "(torch CPU fp32, dropout on, its own per-item loops)")
    v = cpu_pretrain_clips_per_s(nb, steps, warmup, cores)
    return v, "port", (f"{nb} clips/step x {steps} steps (+{warmup} warm-up) of the same training step, oracle port of the reference (torch CPU fp32, "
                       "dropout on); the reference source tree is not present on this box. Build-container calibration (8 cores): see DESIGN.md section 6")


def run_reference_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    nb = 8                                                   # BASELINE.json configs[0]: the reference's own CPU-runnable case
    steps, warmup = max(1, min(args.steps, 30)), max(1, min(args.warmup, 3))     # 8 clips/step at ~7 clips/s: 30 steps stay under a minute
    v, kind, sample = cpu_arm(nb, steps, warmup, cores)
    line = {"impl": "reference", "metric": "pretrain_clips_per_s", "value": v, "unit": "clips/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": 1e3 * nb / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            # the same workload name as our arm (the driver matches the two lines); each timed step is a bounded sample of it
            "config": workload_config(args.gpus, args.batch, args.dtype, args.accum_steps),
            "reference_step": f"{nb} clips per step, fp32, all {cores} host cores, CPU ({kind}); this arm uses no GPU",
            "cpu_baseline": {"value": v, "unit": "clips/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------
# the real bar on the same GPU: the reference algorithm in plain PyTorch eager (cuDNN / cuBLAS / cuFFT)
# --------------------------------------------------------------------------------------------------------------

def torch_eager_b200(dev, nb, steps=2, warmup=1):
    """SURVEY.md 2.2 / BASELINE.md section 4 step 6: the oracle's restatement of the reference (plain torch ops + autograd + torch.optim.Adam)
    on one B200, once in fp32 with TF32 enabled (what run_pretrain.py:33-34 ships) and once under torch.autocast(bfloat16) (the reference's
    --use-amp route in bf16).  An informational baseline: nothing of sarssl_b200 runs here."""
    import torch
    from oracle import sarssl_oracle as O
    out = {"what": "reference algorithm (oracle restatement) in PyTorch eager on this GPU: STFT + masks + fwd + loss + autograd bwd + torch.optim.Adam, dropout 0.1"}
    flags = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    try:
        for mode in ("fp32_tf32", "autocast_bf16"):
            b = nb
            while b >= 8:
                try:
                    sd = {k: v.to(dev) for k, v in O.synthetic_state_dict(7).items()}
                    names = [k for k, v in sd.items() if v.is_floating_point() and "running" not in k and not k.endswith(".pe")]
                    for k in names:
                        sd[k].requires_grad_(True)
                    opt = torch.optim.Adam([sd[k] for k in names], lr=1e-3)
                    sig = 0.1 * torch.randn(b, NSAMPLE, 2, device=dev)
                    random.seed(400000001)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    for i in range(warmup + steps):
                        if i == warmup:
                            torch.cuda.synchronize(dev)
                            e0.record()
                        x = O.preprocess(sig)
                        pidx, cidx = O.draw_masks(b, NT, NT // 2, 2)
                        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "autocast_bf16")):
                            loss, diff, _ = O.pretrain_forward(x, sd, pidx.to(dev), cidx.to(dev), training=True, dropout_p=0.1)
                        loss.backward()
                        opt.step()
                        opt.zero_grad(set_to_none=True)
                    e1.record()
                    torch.cuda.synchronize(dev)
                    out[mode] = {"clips_per_s": b * steps / (e0.elapsed_time(e1) * 1e-3), "batch": b, "steps": steps, "loss": float(loss)}
                    break
                except torch.OutOfMemoryError:
                    b //= 2
                finally:
                    sd = opt = sig = x = loss = diff = None
                    torch.cuda.empty_cache()
            if mode not in out:
                out[mode] = {"clips_per_s": None, "note": "out of memory at every batch tried"}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = flags
    return out


# --------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist
    from sarssl_b200 import ops
    from sarssl_b200.kernels import KernelSet
    from sarssl_b200.learner import STFTLearner
    from sarssl_b200.model import SARSSL, MaskedReconLoss
    from sarssl_b200.optim import FusedAdam

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    nb = args.batch
    warmup = max(args.warmup, 3)

    timed = make_timed(world, dev)       # barrier + synchronize on both sides, CUDA events on the launching stream, max over ranks

    # ---------------- model + learner (identical weights on every rank)
    torch.manual_seed(1)
    model = SARSSL(sig_shape=(NF, NT, 2, 2), device=dev)
    model.to(dev)
    model.set_compute_dtype(dtype)
    model.set_dropout(0.1)
    model.rng_state = ops.mt_seed(400000001 + rank)
    model.train()
    learner = STFTLearner(model, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=1, fs=16000, task=None, ch_mode="M")
    learner.device = dev
    if world > 1:
        learner.mul_gpu()
    sync = getattr(learner, "grad_sync", None)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    sig = 0.1 * torch.randn(nb, NSAMPLE, 2, device=dev, generator=g)
    host_sig = torch.empty(nb, NSAMPLE, 2, dtype=torch.float32).pin_memory()
    host_sig.copy_(sig)
    opt = FusedAdam(model, lr=1e-3)
    losses = []

    accum = max(args.accum_steps, 1)

    def step_eager():
        """One optimizer step = `accum` micro-batches (forward + backward each), then one all-reduce + Adam."""
        for a in range(accum):
            if sync is not None:
                sync.defer = a + 1 < accum
            x, = learner.data_preprocess(sig)
            loss, diff, _ = model(x)
            loss.backward()
        scale = (sync.all_reduce() if sync is not None else 1.0) / accum
        opt.step(1e-3, grad_scale=scale, zero_grad=True)
        losses.append(loss.detach())

    graph_state = {"step": None}

    def step_resident():
        """The same step as the library runs it inside pretrain_epoch: replayed as one CUDA graph after the first eager step (--no-graph: eager)."""
        g = graph_state["step"]
        if g is None:
            return step_eager()
        try:
            loss, _, _ = g.run(sig, 1e-3)
        except RuntimeError as e:                 # a step that cannot be recorded on this box: say so and measure the eager step
            if g.graph is not None:
                raise
            sys.stderr.write(f"bench: CUDA-graph capture failed ({e}); running the step eagerly\n")
            graph_state["step"] = None
            return step_eager()
        losses.append(loss)

    step_eager()
    if accum == 1 and not args.no_graph:
        learner._eager_pretrain_steps = 1
        graph_state["step"] = learner.graphed_pretrain_step(sig, opt)

    for _ in range(warmup):
        step_resident()
    eng = model.engine
    l0 = eng.k.launches + opt.k.launches
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(step_resident, args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=10)              # an nvidia-smi query still in flight would stall the end-to-end region that follows
    gstep = graph_state["step"]
    launches = (eng.k.launches + opt.k.launches - l0 + args.steps) if gstep is None else args.steps * (gstep.launches_per_replay + 1)       # + the front-end kernel
    ops.stft_frontend_check(dev)
    value = nb * world * accum * args.steps / (ms * 1e-3)
    loss_first, loss_last = float(losses[0]), float(losses[-1])

    # ---------------- end to end through the reference-facing API, host buffers
    e2e_steps = max(args.steps, 30) * accum       # one epoch call; long enough that its fixed costs (fresh Adam, first un-overlapped copy, final read-back) stay small
    learner.pretrain_epoch([[host_sig]] * e2e_steps, lr=1e-3, epoch=1, accum_steps=accum)          # warm-up epoch of the same length (allocator / pinned pools at their steady size)
    e2e_ms = timed(lambda: learner.pretrain_epoch([[host_sig]] * e2e_steps, lr=1e-3, epoch=1, accum_steps=accum), 1)
    e2e_value = nb * world * e2e_steps / (e2e_ms * 1e-3)

    # ---------------- the input pipeline in front of the path, from REAL WAV files: loader alone, then through pretrain_epoch
    pipeline = None
    if world == 1 and not args.no_input_pipeline:
        pipeline = bench_input_pipeline(learner, nb, timed)

    # ---------------- dominant kernel of the step alone: 3x3 conv 64->64 (69 % of the FLOPs), CUDA events on the launching stream
    k = KernelSet(dev, dtype)
    cb = min(nb, 64)
    xin = torch.randn(cb * NT * NF, 64, device=dev, generator=g).to(dtype)
    wpk = (torch.randn(64, 9, 64, device=dev, generator=g) / 24.0).to(dtype)
    yout = torch.empty_like(xin)
    conv = (lambda: k.conv3x3_tc(xin, wpk, yout, cb, NT, NF)) if dtype == torch.bfloat16 else (lambda: k.conv3x3(xin, None, wpk, yout, cb, NT, NF))
    for _ in range(3):
        conv()
    conv_ms = timed(conv, 5) / 5
    pk = peaks()
    conv_traffic, conv_traffic_src = measured_traffic("conv3x3_tc_kernel", cb)
    conv_tflops = cb * CONV3_FLOP_PER_CLIP / (conv_ms * 1e-3) / 1e12
    del xin, yout

    # ---------------- front-end + loss sub-path (BASELINE.json configs[1], batch 1024)
    fb = args.frontend_batch
    fsig = 0.1 * torch.randn(fb, NSAMPLE, 2, device=dev, generator=g)
    fpred = torch.randn(fb, NT, 4 * NF, device=dev, generator=g)
    patches = torch.empty(fb, NT, NF, 2, 2, device=dev)
    dpred, out2 = torch.empty_like(fpred), torch.empty(2, device=dev)
    st = ops.mt_seed(7)
    pidx, cidx, flag = ops.draw_masks(st, fb, NT, NT // 2, 2)
    flag_d, cidx_d = torch.from_numpy(flag).to(dev), torch.from_numpy(cidx.astype("int32")).to(dev)
    fe = lambda: ops.stft_frontend(fsig, out=patches)
    ls = lambda: ops.masked_loss(fpred, patches, flag_d, cidx_d, NT // 2, out2=out2, dpred=dpred)
    for _ in range(3):
        fe(); ls()
    fe_ms = timed(fe, 10) / 10
    ls_ms = timed(ls, 10) / 10
    both_ms = timed(lambda: (fe(), ls()), 10) / 10
    fe_gbs = fb * STFT_BYTES_PER_CLIP / (fe_ms * 1e-3) / 1e9
    ls_gbs = fb * LOSS_BYTES_PER_CLIP / (ls_ms * 1e-3) / 1e9

    # ---------------- the other configurations of BASELINE.json, short sub-runs (every rank takes part: weak scaling like the headline)
    extras = {}
    if not args.no_other_configs:
        del fsig, fpred, patches, dpred
        torch.cuda.empty_cache()
        extras["longclip"] = bench_pretrain_shape(dev, world, timed, nt=LC_NT, nb=args.longclip_batch, steps=3, dtype=dtype, sync_learner=None,
                                                  label="configs[4]: 16.4 s clips (262,400 samples, nt = 1024), bf16 pre-training step", use_graph=not args.no_graph)
        extras["finetune"] = bench_finetune_core(dev, world, rank, timed, nb=64, steps=10, dtype=dtype, use_graph=not args.no_graph)

    if rank == 0:
        step_tflops = value * FLOP_PER_CLIP / 1e12 / world
        line = {"metric": "pretrain_clips_per_s", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
                "data": "synthetic", "config": workload_config(world, nb, args.dtype, accum), "clocks": sampler.summary(),
                "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": int(host_sig.numel() * 4 + nb * NT + nb * 4) * accum,
                        "d2h_bytes_per_step": 8},
                "gpu_launches": int(launches), "tensor_core_gemm_launches": int(eng.k.tc_launches),
                "cuda_graph": gstep is not None,
                "loss_first_step": loss_first, "loss_last_step": loss_last,
                "step_tflops_per_gpu": step_tflops, "step_frac_of_bf16_sustained": step_tflops / pk["bf16_tflops_sustained"],
                "roofline": {"bound": "tensor", "kernel": "conv3x3_tc_kernel (3x3 conv 64->64 implicit GEMM, forward)" if dtype == torch.bfloat16 else "conv3x3_kernel (CUDA cores, fp32)", "achieved": conv_tflops,
                             "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": conv_tflops / pk["bf16_tflops"],
                             "traffic": conv_traffic, "traffic_source": conv_traffic_src,
                             "algorithmic_bytes_per_launch": cb * 2 * NT * NF * 64 * 2,        # read + write one bf16 64-channel map
                             "peak_source": pk["source"], "kernel_ms": conv_ms, "algorithmic_flops_per_launch": cb * CONV3_FLOP_PER_CLIP,
                             "clips_per_launch": cb},
                "frontend": {"workload": f"stft_frontend + masked_recon_loss fwd+bwd, batch {fb} (BASELINE.json configs[1])",
                             "clips_per_s": fb / (both_ms * 1e-3),
                             "stft": {"bound": "hbm", "kernel": "stft_frontend_warp2_kernel", "achieved": fe_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                      "frac": fe_gbs / pk["hbm_gbs"], "kernel_ms": fe_ms, "algorithmic_bytes_per_launch": fb * STFT_BYTES_PER_CLIP,
                                      "traffic": measured_traffic("stft_frontend_warp2_kernel", fb)[0]},
                             "loss": {"bound": "hbm", "kernel": "masked_loss_kernel", "achieved": ls_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                      "frac": ls_gbs / pk["hbm_gbs"], "kernel_ms": ls_ms, "algorithmic_bytes_per_launch": fb * LOSS_BYTES_PER_CLIP}}}
        line["other_configs"] = extras
        if pipeline is not None:
            line["input_pipeline"] = pipeline
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            cpu_v, kind, sample = cpu_arm(8, 2, 1, cores)
            line["cpu_baseline"] = {"value": cpu_v, "unit": "clips/s", "cores": cores, "kind": kind, "sample": sample}
        if world == 1 and not args.no_torch_eager:
            del model, learner, opt, sig
            torch.cuda.empty_cache()
            line["torch_eager_b200"] = torch_eager_b200(dev, nb)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------------------
# input pipeline from real WAV files (FixMicSigDataset + DataLoader of run_pretrain.py:175-199)
# --------------------------------------------------------------------------------------------------------------

def write_synthetic_wavs(root, nfiles, nsample, nch=2, fs=16000, seed=0):
    """16-bit PCM RIFF/WAVE files of seeded noise (the format the reference's simulated data set is saved in)."""
    import struct
    import numpy as np
    rng = np.random.default_rng(seed)
    os.makedirs(root, exist_ok=True)
    base = np.round(rng.standard_normal((nsample + nfiles, nch)) * 3000).astype("<i2")
    hdr = struct.pack("<4sI4s4sIHHIIHH4sI", b"RIFF", 36 + nsample * nch * 2, b"WAVE", b"fmt ", 16, 1, nch, fs, fs * nch * 2, nch * 2, 16, b"data", nsample * nch * 2)
    for i in range(nfiles):
        with open(os.path.join(root, f"clip{i:05d}.wav"), "wb") as f:
            f.write(hdr)
            f.write(base[i:i + nsample].tobytes())          # every file a different window of the same noise: cheap to generate, all distinct


def bench_input_pipeline(learner, nb, timed, nfiles=1024, workers=None):
    import shutil
    import tempfile
    from sarssl_b200.data import FixMicSigDataset, WaveformBatchLoader
    shm = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    root = tempfile.mkdtemp(prefix="sarssl_wavs_", dir=shm)
    try:
        write_synthetic_wavs(root, nfiles, NSAMPLE)
        workers = workers or min(os.cpu_count() or 8, 32)
        ds = FixMicSigDataset(data_dir=root, fs=16000, load_anno=False, dataset_sz=None)
        loader = WaveformBatchLoader(ds, batch_size=nb, shuffle=True, num_workers=workers, pin_memory=True, drop_last=True, prefetch=3)
        for _ in loader:                                    # warm-up epoch: page cache, pinned-memory pool
            pass
        epochs = 3
        t0 = time.perf_counter()
        n = 0
        for e in range(epochs):
            loader.set_epoch(e + 1)
            for (b,) in loader:
                n += b.shape[0]
        dt = time.perf_counter() - t0
        loader.set_epoch(9)
        learner.pretrain_epoch(loader, lr=1e-3, epoch=1)   # warm-up epoch through the training call
        steps = len(loader)
        ms = timed(lambda: learner.pretrain_epoch(loader, lr=1e-3, epoch=1), 1)
        return {"what": f"{nfiles} real 16-bit PCM WAV files (2 ch x {NSAMPLE} samples) on {'tmpfs' if shm else 'local disk'} -> FixMicSigDataset -> WaveformBatchLoader "
                        f"(native batch decoder, {workers} threads, pinned batches of {nb}) [-> H2D on a side stream -> STFTLearner.pretrain_epoch]",
                "loader_clips_per_s": n / dt, "loader_f32_gb_per_s": n * NSAMPLE * 2 * 4 / dt / 1e9, "loader_pcm16_gb_per_s": n * NSAMPLE * 2 * 2 / dt / 1e9,
                "decoder_threads": workers, "host_cores": os.cpu_count(),
                "train_from_wav_clips_per_s": nb * steps / (ms * 1e-3), "train_from_wav_steps": steps}
    finally:
        shutil.rmtree(root, ignore_errors=True)


# --------------------------------------------------------------------------------------------------------------
# the other configurations of BASELINE.json: long clips (configs[4]) and the downstream fine-tuning step (configs[3])
# --------------------------------------------------------------------------------------------------------------
LC_NT = 1024                                # 16.4 s clips: 262,400 samples -> 1024 frames
LC_FLOP_PER_CLIP = 361.5e9                  # fwd+bwd at nt = 1024
FT_NT = 64                                  # TDOA fine-tuning uses 1.04 s clips -> 64 frames (run_downstream.py:71-84)
FT_NSAMPLE = (FT_NT + 1) * 256


def make_timed(world, dev):
    import torch
    import torch.distributed as dist

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms
    return timed


def bench_pretrain_shape(dev, world, timed, nt, nb, steps, dtype, sync_learner, label, use_graph=True):
    """A pre-training step at another clip length (same path as the headline: front-end, masks, fwd, loss, bwd, all-reduce, Adam)."""
    import torch
    from sarssl_b200 import ops
    from sarssl_b200.learner import STFTLearner
    from sarssl_b200.model import SARSSL
    from sarssl_b200.optim import FusedAdam
    rank = int(os.environ.get("RANK", "0"))
    torch.manual_seed(1)
    model = SARSSL(sig_shape=(NF, nt, 2, 2), device=dev)
    model.to(dev)
    model.set_compute_dtype(dtype)
    model.set_dropout(0.1)
    model.rng_state = ops.mt_seed(400000101 + rank)
    model.train()
    L = STFTLearner(model, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=1, fs=16000, task=None, ch_mode="M")
    L.device = dev
    if world > 1:
        L.mul_gpu()
    sync = getattr(L, "grad_sync", None)
    sig = 0.1 * torch.randn(nb, (nt + 1) * 256, 2, device=dev)
    host_sig = sig.cpu().pin_memory()
    opt = FusedAdam(model, lr=1e-3)

    def step_eager():
        x, = L.data_preprocess(sig)
        loss, _, _ = model(x)
        loss.backward()
        opt.step(1e-3, grad_scale=sync.all_reduce() if sync is not None else 1.0, zero_grad=True)

    for _ in range(3):
        step_eager()
    L._eager_pretrain_steps = 1
    gstep = L.graphed_pretrain_step(sig, opt) if use_graph else None

    def step():
        if gstep is None:
            return step_eager()
        gstep.run(sig, 1e-3)

    step()
    l0 = model.engine.k.launches + opt.k.launches
    ms = timed(step, steps)
    launches = (model.engine.k.launches + opt.k.launches - l0 + steps) if gstep is None else steps * (gstep.launches_per_replay + 1)
    e2e_steps = max(steps, 4)
    L.pretrain_epoch([[host_sig]] * 2, lr=1e-3, epoch=1)
    e2e_ms = timed(lambda: L.pretrain_epoch([[host_sig]] * e2e_steps, lr=1e-3, epoch=1), 1)
    v = nb * world * steps / (ms * 1e-3)
    pk = peaks()
    flop = LC_FLOP_PER_CLIP * (nt / LC_NT) if nt != NT else FLOP_PER_CLIP
    return {"workload": f"{label}, {nb} clips/GPU", "metric": "pretrain_clips_per_s", "value": v, "unit": "clips/s", "n_gpus": world, "steps": steps,
            "ms_per_step": ms / steps, "per_gpu_batch": nb, "nt": nt, "dtype": "bf16" if dtype == torch.bfloat16 else "f32",
            "e2e": {"value": nb * world * e2e_steps / (e2e_ms * 1e-3), "unit": "clips/s", "h2d_bytes_per_step": int(host_sig.numel() * 4 + nb * nt + nb * 4),
                    "d2h_bytes_per_step": 8},
            "gpu_launches": int(launches), "cuda_graph": gstep is not None,
            "roofline": {"bound": "tensor", "achieved": v / world * flop / 1e12, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": v / world * flop / 1e12 / pk["bf16_tflops_sustained"], "what": "whole step: algorithmic fwd+bwd FLOPs per clip x clips/s per GPU "
                         "against the sustained bf16 peak", "flop_per_clip": flop, "peak_source": pk["source"]}}


def cpu_finetune_clips_per_s(nb, steps, warmup, threads):
    import torch
    from oracle import sarssl_oracle as O
    torch.set_num_threads(threads)
    sig = O.synthetic_waveforms(nb, FT_NSAMPLE, 2, seed=1234)
    tar = torch.linspace(-1.0, 1.0, nb)[:, None]
    sd = O.synthetic_state_dict(7, pretrain=False, dembed_ds=768)
    names = [k for k, v in sd.items() if v.is_floating_point() and "running" not in k and not k.endswith(".pe")]
    for k in names:
        sd[k].requires_grad_(True)
    m = [torch.zeros_like(sd[k]) for k in names]
    v = [torch.zeros_like(sd[k]) for k in names]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        pred, _ = O.downstream_forward(O.preprocess(sig), sd, "spec_spat", training=True, dropout_p=0.1)
        loss = torch.nn.functional.mse_loss(pred, tar)
        loss.backward()
        with torch.no_grad():
            O.adam_step([sd[k] for k in names], [sd[k].grad if sd[k].grad is not None else torch.zeros_like(sd[k]) for k in names], m, v, i + 1, 1e-5)
        for k in names:
            sd[k].grad = None
        float(loss)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return nb * len(times) / sum(times)


def bench_finetune_core(dev, world, rank, timed, nb, steps, dtype, use_graph=True):
    """configs[3]: downstream fine-tuning step (TDOA head, MSE, Adam) on 1.04 s clips."""
    import torch
    from sarssl_b200.learner import STFTLearner
    from sarssl_b200.model import SARSSL
    from sarssl_b200.optim import FusedAdam
    model = SARSSL(sig_shape=(NF, FT_NT, 2, 2), pretrain=False, device=dev)
    model.to(dev)
    model.set_compute_dtype(dtype)
    model.train()
    L = STFTLearner(model, 512, 0.5, 512, 1, 16000, task="TDOA")
    L.device = dev
    if world > 1:
        L.mul_gpu()
    sync = getattr(L, "grad_sync", None)
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    sig = 0.1 * torch.randn(nb, FT_NSAMPLE, 2, device=dev, generator=g)
    labels = torch.linspace(-5e-5, 5e-5, nb, device=dev)
    host_sig, host_lab = sig.cpu().pin_memory(), labels.cpu().pin_memory()
    opt = FusedAdam(model, lr=1e-5)

    def step_eager():
        x, tar = L.data_preprocess(sig, {"TDOA": labels})
        pred, _ = model(x)
        L.loss(pred_batch=pred, gt_batch=tar).backward()
        opt.step(1e-5, grad_scale=sync.all_reduce() if sync is not None else 1.0)

    for _ in range(3):
        step_eager()
    L._eager_finetune_steps = 1
    gstep = L.graphed_finetune_step(sig, opt) if use_graph else None          # the step as train_epoch runs it: one CUDA graph after the first eager step
    tar_dev = L.get_tar_batch(labels)

    def step():
        if gstep is None:
            return step_eager()
        gstep.run(sig, tar_dev, 1e-5)

    step()
    l0 = model.engine.k.launches + opt.k.launches
    ms = timed(step, steps)
    launches = (model.engine.k.launches + opt.k.launches - l0 + steps) if gstep is None else steps * (gstep.launches_per_replay + 1)      # + the front-end kernel
    # end to end through the reference-facing call with host buffers; an epoch re-creates Adam (like the reference), so it is warmed up once
    # and timed over enough steps that this per-epoch cost does not dominate a 7 ms step
    e2e_steps = max(steps, 20)
    L.train_epoch([(host_sig, {"TDOA": host_lab})] * 2, lr=1e-5)
    e2e_ms = timed(lambda: L.train_epoch([(host_sig, {"TDOA": host_lab})] * e2e_steps, lr=1e-5), 1)
    return {"workload": f"configs[3]: downstream fine-tune step (TDOA head, MSE, Adam), 2-mic, {FT_NSAMPLE} samples (1.04 s), batch {nb}/GPU",
            "metric": "finetune_clips_per_s", "value": nb * world * steps / (ms * 1e-3), "unit": "clips/s", "n_gpus": world, "steps": steps,
            "ms_per_step": ms / steps, "per_gpu_batch": nb, "nt": FT_NT, "dtype": "bf16" if dtype == torch.bfloat16 else "f32",
            "e2e": {"value": nb * world * e2e_steps / (e2e_ms * 1e-3), "unit": "clips/s", "h2d_bytes_per_step": int(host_sig.numel() * 4 + nb * 4),
                    "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "cuda_graph": gstep is not None}


def run_sub_workload(args):
    """--workload finetune | longclip as a stand-alone JSON line."""
    import torch
    import torch.distributed as dist
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            cores = os.cpu_count() or 1
            v = cpu_finetune_clips_per_s(8, max(1, min(args.steps, 4)), 1, cores) if args.workload == "finetune" else None
            print(json.dumps({"impl": "reference", "metric": args.workload + "_clips_per_s", "value": v, "unit": "clips/s", "n_gpus": args.gpus, "higher_is_better": True,
                              "cpu_baseline": {"value": v, "unit": "clips/s", "cores": cores, "kind": "port", "sample": "8 clips/step, oracle port, torch CPU fp32"},
                              "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    timed = make_timed(world, dev)
    if args.workload == "finetune":
        line = bench_finetune_core(dev, world, rank, timed, nb=args.batch if args.batch != 256 else 64, steps=args.steps, dtype=dtype, use_graph=not args.no_graph)   # configs[3]: batch 512 over 8 GPUs
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            line["cpu_baseline"] = {"value": cpu_finetune_clips_per_s(8, 2, 1, cores), "unit": "clips/s", "cores": cores, "kind": "port",
                                    "sample": "8 clips/step x 2 steps, oracle port (torch CPU fp32, dropout on)"}
    else:
        line = bench_pretrain_shape(dev, world, timed, nt=LC_NT, nb=args.longclip_batch, steps=args.steps, dtype=dtype, sync_learner=None,
                                    label="configs[4]: 16.4 s clips (262,400 samples, nt = 1024), bf16 pre-training step", use_graph=not args.no_graph)
    line.update({"warmup": 3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic", "config": {"workload": line["workload"], "parallelism": f"dp{world}"}})
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    # stdout carries exactly one JSON line: libraries that chat on file descriptor 1 (NCCL prints its version banner there when a communicator
    # comes up) are pointed at stderr for the whole run, and the line is written to the real stdout at the end
    global print
    sys.stdout.flush()
    real_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    builtin_print = print

    def print(*a, **kw):                      # (every print in this file is a result line)
        kw.pop("flush", None)
        builtin_print(*a, file=real_out, flush=True, **kw)

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="clips per GPU per step (micro-batch)")
    ap.add_argument("--accum-steps", type=int, default=1, help="micro-batches per optimizer step (configs[2]: global batch 2048 = 256 x accum x GPUs)")
    ap.add_argument("--frontend-batch", type=int, default=1024)
    ap.add_argument("--longclip-batch", type=int, default=32, help="clips per GPU of the long-clip sub-run (configs[4]: 256 over 8 GPUs)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-eager", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--no-input-pipeline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the training step eagerly (one launch per kernel) instead of replaying it as a CUDA graph")
    ap.add_argument("--workload", default="pretrain", choices=["pretrain", "finetune", "longclip"],
                    help="pretrain = headline (configs[2]); finetune = configs[3]; longclip = configs[4]")
    args = ap.parse_args()
    if args.workload != "pretrain":
        run_sub_workload(args)
    elif args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
