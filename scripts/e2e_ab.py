"""Where the end-to-end epoch loses time against back-to-back graph replays: pretrain_epoch on (a) device-resident batches, (b) pinned host batches (H2D on the
side stream), against (c) the bare replay loop.  B = 256, nt = 256, bf16."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sarssl_b200 import ops
from sarssl_b200.learner import STFTLearner
from sarssl_b200.model import SARSSL
dev = torch.device("cuda", 0)
nb, n = 256, 30
model = SARSSL(sig_shape=(256, 256, 2, 2), device=dev)
model.set_compute_dtype(torch.bfloat16)
model.set_dropout(0.1)
model.rng_state = ops.mt_seed(400000001)
model.train()
L = STFTLearner(model, win_len=512, win_shift_ratio=0.5, nfft=512, fre_used_ratio=1, fs=16000, task=None, ch_mode="M")
L.device = dev
sig = 0.1 * torch.randn(nb, 65792, 2, device=dev)
host = torch.empty(nb, 65792, 2, dtype=torch.float32).pin_memory()
host.copy_(sig)
def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n
L.pretrain_epoch([[sig]] * 3, lr=1e-3, epoch=1)
g = L.graphed_pretrain_step(sig, L._epoch_optimizer)
def loop():
    for _ in range(n):
        g.run(sig, 1e-3)
L.pretrain_epoch([[host]] * 4, lr=1e-3, epoch=1)
for rep in range(2):
    for name, fn in (("bare graph replays", loop), ("pretrain_epoch, pinned host batches", lambda: L.pretrain_epoch([[host]] * n, lr=1e-3, epoch=1)),
                     ("pretrain_epoch, device batches", lambda: L.pretrain_epoch([[sig]] * n, lr=1e-3, epoch=1))):
        ev, wall = timed(fn)
        print(f"{name:36s}: {ev:.2f} ms/step (events), {wall:.2f} ms/step (wall)")
g = L.graphed_pretrain_step(sig, L._epoch_optimizer)
# host-side cost of one step's staging (everything before graph.replay)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(n):
    g._stage(1e-3)
print(f"host staging per step: {(time.perf_counter() - t0) * 1e3 / n:.2f} ms")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(n):
    t = host.to(dev, non_blocking=True)
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"H2D 134.8 MB: host-side call {(t1 - t0) * 1e3 / n:.2f} ms, completed {(time.perf_counter() - t0) * 1e3 / n:.2f} ms each")
# isolation: the same replay loop with (B) an independent H2D per step on a side stream, (C) double-buffered H2D feeding the step (no allocator involved)
side = torch.cuda.Stream()
bufs = [torch.empty_like(sig), torch.empty_like(sig)]
def loop_b():
    for _ in range(n):
        with torch.cuda.stream(side):
            bufs[0].copy_(host, non_blocking=True)
        g.run(sig, 1e-3)
def loop_d():
    for _ in range(n):
        g.run(sig, 1e-3)
        with torch.cuda.stream(side):
            bufs[0].copy_(host, non_blocking=True)
def loop_c():
    evs = [None, None]
    with torch.cuda.stream(side):
        bufs[0].copy_(host, non_blocking=True); evs[0] = torch.cuda.Event(); evs[0].record(side)
    done = [None, None]
    for i in range(n):
        j = (i + 1) & 1
        with torch.cuda.stream(side):
            if done[j] is not None:
                side.wait_event(done[j])           # the step that read this buffer has finished
            bufs[j].copy_(host, non_blocking=True); evs[j] = torch.cuda.Event(); evs[j].record(side)
        torch.cuda.current_stream().wait_event(evs[i & 1])
        g.run(bufs[i & 1], 1e-3)
        done[i & 1] = torch.cuda.Event(); done[i & 1].record()
for name, fn in (("replays only", loop), ("replays + independent H2D", loop_b), ("replays, H2D enqueued after the launch", loop_d), ("replays fed by double-buffered H2D", loop_c), ("replays only", loop)):
    fn(); ev, wall = timed(fn)
    print(f"{name:36s}: {ev:.2f} ms/step")
