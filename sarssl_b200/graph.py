"""The pre-training step as ONE CUDA graph (masks -> forward -> loss -> backward -> [gradient all-reduce] -> Adam).

The step is ~450 kernel launches; queued eagerly they leave 5 % of a 42 ms step (and more than half of the 7 ms fine-tuning step) to
launch latency and inter-kernel gaps.  Everything that changes from step to step lives in device memory, so the captured launch
arguments can stay frozen:
  * input patches: the STFT front-end (outside the graph) writes straight into the static patch buffer;
  * frame / channel masks: drawn on the host exactly as in eager mode (CPython MT19937 stream) and copied into static device buffers;
  * dropout: every kernel adds a device word - step_part(step number) - to its by-value per-call-site seed (sarssl_gemm_args.seed_dev,
    ...), so a replay draws the masks the eager step with the same number would draw;
  * Adam: the bias-corrected step size and sqrt(1 - beta2^t) come from two device floats (sarssl_adam_step's hyper_dev).
One small pinned host buffer per step carries {seed, Adam scalars, channel indices, frame flags} to the device in one copy.
Eager and replayed steps are interchangeable (tests/test_graph_gpu.py compares them bit for bit with dropout on)."""
import numpy as np
import torch

from . import ops
from .engine import step_part


class GraphedPretrainStep:
    def __init__(self, learner, optimizer, nb, nsample):
        self.learner, self.model, self.opt = learner, learner.model, optimizer
        m = self.model
        self.dev = m.store.flat.device
        nf, nt = m.sig_shape[:2]
        self.nb, self.nt, self.nf, self.nsample = nb, nt, nf, nsample
        self.patches = torch.empty(nb, nt, nf, 2, 2, dtype=torch.float32, device=self.dev)
        # one staging image: [seed u64 | adam 2 x f32 | ch nb x i32 | flag nb*nt x u8]
        self.o_ch, self.o_flag = 16, 16 + 4 * nb
        self.nbytes = (self.o_flag + nb * nt + 15) // 16 * 16
        self.dbuf = torch.zeros(self.nbytes, dtype=torch.uint8, device=self.dev)
        self.seed_dev = self.dbuf[0:8].view(torch.int64)
        self.hyper_dev = self.dbuf[8:16].view(torch.float32)
        self.ch_dev = self.dbuf[self.o_ch:self.o_flag].view(torch.int32)
        self.flag_dev = self.dbuf[self.o_flag:self.o_flag + nb * nt].view(nb, nt)
        self.graph, self.out, self.vis = None, None, None
        self.launches_per_replay = 0          # kernel launches recorded in the graph (what one replay executes)

    # ---- host side of one step: everything that varies goes into one pinned image
    def _stage(self, lr):
        m, eng = self.model, self.model._engine()
        h = torch.empty(self.nbytes, dtype=torch.uint8, pin_memory=True)
        hn = h.numpy()
        hn[0:8].view(np.int64)[0] = step_part(eng.step_seed + 1)
        self.opt.hyper_host(h[8:16].view(torch.float32), self.opt.t + 1, lr)
        pidx, cidx, flag = m.patch_mask.draw_host(self.nb, self.nt, 2, m.rng_state, dp=m.dp)
        hn[self.o_ch:self.o_flag].view(np.int32)[:] = cidx.reshape(-1)
        hn[self.o_flag:self.o_flag + self.nb * self.nt] = flag.reshape(-1)
        self.dbuf.copy_(h, non_blocking=True)
        return pidx, cidx

    def _body(self, lr):
        m = self.model
        x = self.patches.permute(0, 4, 2, 1, 3)                     # the reference's (nb, 2, nf, nt, 2) view of the static buffer
        loss, diff, vis = m(x, _static_masks=(self.flag_dev, self.ch_dev))
        loss.backward()
        sync = getattr(self.learner, "grad_sync", None)
        scale = sync.all_reduce() if sync is not None else 1.0
        self.opt.step(lr, grad_scale=scale, zero_grad=True, hyper_dev=self.hyper_dev)
        return loss.detach(), diff.detach(), vis

    def capture(self, lr):
        """Record the step.  Nothing executes; the host-side counters the body advances are put back."""
        eng = self.model._engine()
        t0, s0 = self.opt.t, eng.step_seed
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        eng.use_device_seed(self.seed_dev)          # only while recording: the pointer is baked into the captured launches, eager steps stay by-value
        l0 = eng.k.launches + self.opt.k.launches
        try:
            with torch.cuda.graph(g):
                self.out = self._body(lr)
            self.launches_per_replay = eng.k.launches + self.opt.k.launches - l0
        finally:
            eng.use_device_seed(None)
            self.opt.t, eng.step_seed = t0, s0
        self.graph = g

    def run(self, sig, lr):
        """One training step on waveforms `sig` (nb, nsample, 2) already on the device.  Returns (loss, diff, vis): device scalars + the views."""
        ops.stft_frontend(sig, out=self.patches)
        pidx, cidx = self._stage(lr)
        eng = self.model._engine()
        if self.graph is None:
            self.capture(lr)
        self.graph.replay()
        eng.step_seed += 1
        self.opt.t += 1
        loss, diff, vis = self.out
        vis.mask_patch_idx, vis.mask_ch_idx = torch.from_numpy(pidx), torch.from_numpy(cidx)
        return loss.clone(), diff.clone(), vis

    def release(self):
        self.graph, self.out = None, None


class GraphedFinetuneStep:
    """The downstream fine-tuning step (learner.py:170-222: forward, MSE loss, backward, [all-reduce], Adam) as one CUDA graph.  475 launches
    for 7 ms of device work make the eager step launch-bound; per-step inputs: patches (front-end output), targets, dropout seed, Adam scalars."""

    def __init__(self, learner, optimizer, nb, tar_shape):
        self.learner, self.model, self.opt = learner, learner.model, optimizer
        m = self.model
        self.dev = m.store.flat.device
        nf, nt = m.sig_shape[:2]
        self.patches = torch.empty(nb, nt, nf, 2, 2, dtype=torch.float32, device=self.dev)
        self.tar = torch.empty(tar_shape, dtype=torch.float32, device=self.dev)
        self.dbuf = torch.zeros(16, dtype=torch.uint8, device=self.dev)
        self.seed_dev = self.dbuf[0:8].view(torch.int64)
        self.hyper_dev = self.dbuf[8:16].view(torch.float32)
        self.graph, self.out = None, None
        self.launches_per_replay = 0

    def _body(self, lr):
        L = self.learner
        pred, emb = self.model(self.patches.permute(0, 4, 2, 1, 3))
        loss = L.loss(pred_batch=pred, gt_batch=self.tar)
        loss.backward()
        sync = getattr(L, "grad_sync", None)
        scale = sync.all_reduce() if sync is not None else 1.0
        self.opt.step(lr, grad_scale=scale, zero_grad=True, hyper_dev=self.hyper_dev)
        return loss.detach(), pred.detach(), emb

    def capture(self, lr):
        eng = self.model._engine()
        t0, s0 = self.opt.t, eng.step_seed
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        eng.use_device_seed(self.seed_dev)
        l0 = eng.k.launches + eng.k32.launches + self.opt.k.launches
        try:
            with torch.cuda.graph(g):
                self.out = self._body(lr)
            self.launches_per_replay = eng.k.launches + eng.k32.launches + self.opt.k.launches - l0
        finally:
            eng.use_device_seed(None)
            self.opt.t, eng.step_seed = t0, s0
        self.graph = g

    def run(self, sig, tar, lr):
        """sig (nb, nsample, 2) on the device, tar = get_tar_batch(labels) on the device.  Returns (loss, pred, embed)."""
        ops.stft_frontend(sig, out=self.patches)
        self.tar.copy_(tar, non_blocking=True)
        eng = self.model._engine()
        h = torch.empty(16, dtype=torch.uint8, pin_memory=True)
        h.numpy()[0:8].view(np.int64)[0] = step_part(eng.step_seed + 1)
        self.opt.hyper_host(h[8:16].view(torch.float32), self.opt.t + 1, lr)
        self.dbuf.copy_(h, non_blocking=True)
        if self.graph is None:
            self.capture(lr)
        self.graph.replay()
        eng.step_seed += 1
        self.opt.t += 1
        loss, pred, emb = self.out
        return loss.clone(), pred, emb

    def release(self):
        self.graph, self.out = None, None
