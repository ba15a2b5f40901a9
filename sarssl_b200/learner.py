"""Mirror of the reference's training-step surface (code/learner.py): Learner / STFTLearner.

Same constructor and method names; the step itself is re-designed: the waveform batch goes through one
fused front-end kernel, the model is a single fused forward/backward schedule, Adam is one multi-tensor kernel, and
the loss values are read back once per step."""
from abc import ABC

import numpy as np
import torch

from . import ops
from ._lib import SarsslError
from .checkpoint import CheckpointMixin


class Learner(CheckpointMixin, ABC):
    """learner.py:13-50."""

    def __init__(self, model):
        self.model = model
        self.max_score = -np.inf
        self.early_stop_counter = 0
        self.use_amp = False
        self.start_epoch = 1
        self.device = "cuda"

    def cuda(self):
        if self.model is not None:
            self.model.cuda()
        self.device = "cuda"

    def cpu(self):
        raise SarsslError("sarssl_b200 has no CPU path (Learner.cpu of the reference, learner.py:40-44, is not supported)")

    def amp(self):
        """learner.py:46-50.  Mixed precision here means bf16 storage for activations with fp32 master weights and fp32
        accumulation; no loss scaling is needed, so `scaler` is a disabled GradScaler kept for attribute compatibility."""
        self.use_amp = True
        self.scaler = torch.amp.GradScaler("cuda", enabled=False)
        if self.model is not None and hasattr(self.model, "set_compute_dtype"):
            self.model.set_compute_dtype(torch.bfloat16)


    def mul_gpu(self):
        """learner.py:25-31 wraps the model in nn.DataParallel (one process, per-step weight broadcast).  Here: one process per
        GPU (torchrun); gradients are all-reduced in contiguous buckets of the flat gradient arena over NCCL, overlapped with
        the rest of backward (see sarssl_b200/parallel.py).  Without an initialised process group this is a no-op."""
        from .parallel import GradientSync
        self.grad_sync = GradientSync.create(self.model)

    def device_batches(self, dataset):
        """Input pipeline step in front of the path (the reference leaves it to DataLoader + a blocking `.to(device)` inside
        data_preprocess, learner.py:530): host tensors of batch i+1 are copied to the device on a side stream while batch i is
        being computed (the copy is enqueued right after the consumer has launched batch i), so the H2D transfer (134 MB for 256 clips)
        never sits on the compute stream.  Items keep their structure
        (list / tuple / dict of tensors); pinned host tensors copy asynchronously, pageable ones still overlap with the device work
        already queued.  Device tensors pass through untouched."""
        dev = torch.device(self.device)
        if dev.type != "cuda":
            yield from dataset
            return
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        if getattr(self, "_copy_stream", None) is None or self._copy_stream.device != dev:
            self._copy_stream = torch.cuda.Stream(device=dev)
        copy_stream = self._copy_stream

        def move(obj, moved):
            if torch.is_tensor(obj):
                if obj.device.type == "cuda":
                    return obj
                t = obj.to(dev, non_blocking=True)
                moved.append(t)
                return t
            if isinstance(obj, dict):
                return {k: move(v, moved) for k, v in obj.items()}
            if isinstance(obj, (list, tuple)):
                return type(obj)(move(v, moved) for v in obj)
            return obj

        def stage(item):
            moved = []
            with torch.cuda.stream(copy_stream):
                out = move(item, moved)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
            return out, moved, ev

        # Order matters on the device: a CUDA-graph launch uploads its launch data through the same copy engine, so a 134 MB host-to-device copy enqueued
        # just BEFORE the step's launch delays that step by the copy's whole duration (measured: +2.4 ms per step, scripts/e2e_ab.py).  The next batch
        # is therefore staged AFTER the consumer has launched its work on the current one (when it asks for the next item): the copy then runs while
        # the device computes.
        it = iter(dataset)
        try:
            cur = stage(next(it))
        except StopIteration:
            return
        while cur is not None:
            out, moved, ev = cur
            compute = torch.cuda.current_stream(dev)
            compute.wait_event(ev)
            for t in moved:
                t.record_stream(compute)         # allocated on the copy stream, consumed on the compute stream
            yield out
            del out, moved
            try:
                cur = stage(next(it))            # the consumer's work on the previous batch is already queued: this copy overlaps it
            except StopIteration:
                cur = None

    def pretrain_epoch(self, dataset, lr=0.0001, epoch=None, return_diff=True, accum_steps=1, use_graph=None):
        """learner.py:76-131: one epoch of pre-training.  A fresh Adam (moments reset) per epoch like the reference; loss / diff
        of every step are kept on the device and read back once at the end of the epoch (the reference syncs 3x per step).

        accum_steps > 1: gradient accumulation - `accum_steps` consecutive micro-batches form one optimizer step (their gradients
        add up in the flat gradient arena, 1/accum_steps is folded into the Adam kernel, and under data parallelism the all-reduce
        runs once per optimizer step instead of once per micro-batch).  This is how 1 / 2 / 4 GPUs reach BASELINE.json's global batch
        of 2048 with 256-clip micro-batches; BatchNorm statistics stay per micro-batch, as under the reference's
        DataParallel replicas.  A trailing partial group is applied with its real count.

        use_graph (None = automatic, False = never): after the first (eager) step, batches of an unchanged shape are replayed as ONE CUDA graph
        (graph.py) - same arithmetic, same masks, same dropout stream as the eager step, without the ~450 per-step launches."""
        from .optim import FusedAdam
        self.model.train()
        optimizer = getattr(self, "_epoch_optimizer", None)
        if optimizer is None or optimizer.model is not self.model or optimizer.m.device != self.model.store.flat.device:
            optimizer = self._epoch_optimizer = FusedAdam(self.model, lr=lr)      # one set of moment buffers, reset every epoch: a captured graph keeps their addresses
        optimizer.reset()
        optimizer.zero_grad()
        sync = getattr(self, "grad_sync", None)
        accum_steps = max(int(accum_steps), 1)
        log, vis_batch, pending = [], None, 0

        def apply(count):
            scale = 1.0 / count
            if sync is not None:
                scale *= sync.all_reduce()
            optimizer.step(lr, grad_scale=scale, zero_grad=True)

        for batch_idx, (mic_sig_batch,) in enumerate(self.device_batches(dataset)):
            if accum_steps == 1 and use_graph is not False:
                step = self.graphed_pretrain_step(mic_sig_batch, optimizer)
                if step is not None:
                    try:
                        loss_batch, diff_batch, vis_batch = step.run(mic_sig_batch.to(self.device, non_blocking=True), lr)
                    except RuntimeError as e:            # a capture that cannot be recorded: remember it and stay eager (nothing has executed)
                        if step.graph is not None:
                            raise
                        self._graph_failed = True
                        import warnings
                        warnings.warn(f"sarssl_b200: CUDA-graph capture of the training step failed, running eagerly ({e})")
                    else:
                        log.append(torch.stack([loss_batch, diff_batch]))
                        continue
            in_batch, = self.data_preprocess(mic_sig_batch, None)
            if sync is not None:
                sync.defer = pending + 1 < accum_steps      # only the last micro-batch of a group announces its buckets
            loss_batch, diff_batch, vis_batch = self.model(in_batch)
            loss_batch.backward()
            pending += 1
            if pending == accum_steps:
                apply(pending)
                pending = 0
            self._eager_pretrain_steps = getattr(self, "_eager_pretrain_steps", 0) + 1
            log.append(torch.stack([loss_batch.detach(), diff_batch.detach()]))
        if pending:
            apply(pending)
        vals = torch.stack(log).mean(0).tolist() if log else [0.0, 0.0]      # the epoch's only device -> host read
        self.check_frontend()
        if return_diff:
            return vals[0], vals[1], vis_batch
        return vals[0]

    def graphed_pretrain_step(self, sig, optimizer):
        """The CUDA-graph runner (graph.py) for waveform batches shaped like `sig`, or None when this step must run eagerly: the very first
        step (lazy initialisation - cached tables, kernel attributes - must happen outside a capture), a model outside the pre-training
        configuration the graph covers, or a capture that failed once."""
        m = self.model
        dev = torch.device(self.device)
        if dev.type != "cuda" or not getattr(m, "pretrain", False) and not getattr(m, "pretrain_frozen_encoder", False):
            return None
        if not torch.is_tensor(sig) or sig.dim() != 3 or sig.shape[-1] != 2 or getattr(self, "fre_used_ratio", 1) != 1 or not m.training:
            return None
        if getattr(self, "_eager_pretrain_steps", 0) < 1 or getattr(self, "_graph_failed", False):
            return None
        eng = m._engine()
        key = (tuple(sig.shape), id(eng), id(optimizer), m.compute_dtype, m.dropout_p, getattr(self, "grad_sync", None) is not None,
               tuple(not p.requires_grad for p in m.store.params.values()))
        cache = self.__dict__.setdefault("_graph_steps", {})
        step = cache.get(key)
        if step is None:
            from .graph import GraphedPretrainStep
            for old in cache.values():           # one live graph: its private pool holds a whole step's activations
                old.release()
            cache.clear()
            step = cache[key] = GraphedPretrainStep(self, optimizer, sig.shape[0], sig.shape[1])
        return step

    def graphed_finetune_step(self, sig, optimizer):
        """The CUDA-graph runner of the fine-tuning step for batches shaped like `sig`, or None when the step must run eagerly."""
        m = self.model
        dev = torch.device(self.device)
        if dev.type != "cuda" or getattr(m, "pretrain", True) or getattr(m, "pretrain_frozen_encoder", False) or not m.training:
            return None
        if not torch.is_tensor(sig) or sig.dim() != 3 or sig.shape[-1] != 2 or getattr(self, "fre_used_ratio", 1) != 1:
            return None
        if getattr(self, "_eager_finetune_steps", 0) < 1 or getattr(self, "_graph_failed", False):
            return None
        eng = m._engine()
        key = ("ft", tuple(sig.shape), id(eng), id(optimizer), m.compute_dtype, m.dropout_p, getattr(self, "grad_sync", None) is not None,
               tuple(not p.requires_grad for p in m.store.params.values()))
        cache = self.__dict__.setdefault("_graph_steps", {})
        step = cache.get(key)
        if step is None:
            from .graph import GraphedFinetuneStep
            for old in cache.values():
                old.release()
            cache.clear()
            step = cache[key] = GraphedFinetuneStep(self, optimizer, sig.shape[0], (sig.shape[0], 1))
        return step

    def check_frontend(self):
        """Raise if the fused front-end kernel's per-clip rendezvous ever timed out during the epoch (it would have written
        patches scaled with an incomplete mean).  Reads one flag; called right after the epoch's loss read-back, when the
        device is idle anyway."""
        if torch.device(self.device).type == "cuda":
            dev = torch.device(self.device)
            if dev.index is None:
                dev = torch.device("cuda", torch.cuda.current_device())
            ops.stft_frontend_check(dev)

    def pretest_epoch(self, dataset, return_diff=True, return_eval=False):
        """learner.py:133-167: eval-mode pass (BatchNorm running statistics, no dropout, masks still random)."""
        self.model.eval()
        log, vis_batch = [], None
        with torch.no_grad():
            for data in self.device_batches(dataset):
                in_batch, = self.data_preprocess(data[0], None)
                loss_batch, diff_batch, vis_batch = self.model(in_batch)
                log.append(torch.stack([loss_batch, diff_batch]))
        vals = torch.stack(log).mean(0).tolist() if log else [0.0, 0.0]
        self.check_frontend()
        if return_diff:
            if return_eval:      # metrics of the last batch only, like the reference (learner.py:160-162)
                result = self.pretrain_evaluate(pred_batch=vis_batch["pred"], gt_batch=vis_batch["tar"], mask_batch=vis_batch["mask"])
                return vals[0], vals[1], vis_batch, result
            return vals[0], vals[1], vis_batch
        return vals[0]


class STFTLearner(Learner):
    """learner.py:488-572."""

    def __init__(self, model, win_len, win_shift_ratio, nfft, fre_used_ratio, fs, mel_scale=False, task=None, ch_mode="M"):
        super().__init__(model)
        if mel_scale:
            raise SarsslError("STFTLearner: mel_scale=True (torchaudio MelScale on the CPU, learner.py:506-513,548) is not built")
        if fre_used_ratio not in (1, 0.5) or ch_mode not in ("M", "MM"):
            raise SarsslError("STFTLearner: fre_used_ratio in {1, 0.5} and ch_mode in {'M', 'MM'} (learner.py:514-520, utils_module.py:124-148)")
        self.fre_used_ratio = fre_used_ratio
        self.win_len, self.win_shift_ratio, self.nfft, self.fs = win_len, win_shift_ratio, nfft, fs
        self.ch_mode, self.task = ch_mode, task
        from .modules import STFT, ISTFT
        self.stft = STFT(win_len=win_len, win_shift_ratio=win_shift_ratio, nfft=nfft)
        self.istft = ISTFT(win_len=win_len, win_shift_ratio=win_shift_ratio, nfft=nfft, inv=False)
        self.fre_range_used = range(1, int(nfft / 2 * fre_used_ratio) + 1, 1) if fre_used_ratio == 1 else range(0, int(nfft / 2 * fre_used_ratio), 1)

    def data_preprocess(self, mic_sig_batch=None, gt_batch=None, eps=1e-6):
        """learner.py:525-572.  mic_sig_batch (nb, nsample, nch) (host or device) -> [ (nb*(nch-1), 2, nf, nt, 2) f32 ].
        The result is a permuted view of patch-layout storage (see modules.as_patch_layout)."""
        data = []
        if mic_sig_batch is not None:
            sig = mic_sig_batch.to(self.device, non_blocking=True)
            if self.fre_used_ratio == 1 and (self.ch_mode == "M" or sig.shape[-1] == 2):          # the pre-training configuration: fused kernel
                patches = ops.stft_frontend(sig, eps=eps, win_len=self.win_len, hop=int(self.win_len * self.win_shift_ratio), nfft=self.nfft)
            else:                                                    # ch_mode 'MM' with > 2 microphones / lower half of the spectrum: generic route
                first, nbins = (1, self.nfft // 2) if self.fre_used_ratio == 1 else (0, self.nfft // 4)
                patches = ops.stft_frontend_ex(sig, eps=eps, win_len=self.win_len, hop=int(self.win_len * self.win_shift_ratio), nfft=self.nfft,
                                               all_pairs=self.ch_mode == "MM", first_bin=first, nbins=nbins)
            data += [patches.permute(0, 4, 2, 1, 3)]
        if gt_batch is not None:
            gt = gt_batch[self.task].to(self.device, non_blocking=True)
            data += [self.get_tar_batch(gt_batch=gt)]
        return data

    # ---- downstream fine-tuning surface
    def get_tar_batch(self, gt_batch):
        """learner.py:620-631: (nbatch,) labels -> (nbatch, 1) regression targets (TDOA in samples at 16 kHz)."""
        if self.task == "TDOA":
            return gt_batch[:, None] * 16000
        if self.task in ("DRR", "C50", "T60", "ABS"):
            return gt_batch[:, None]
        raise SarsslError(f"Task mode unrecognized: {self.task}")

    def loss(self, pred_batch, gt_batch):
        """learner.py:633-642: MSE between the (nbatch, 1) prediction and target."""
        return torch.nn.functional.mse_loss(pred_batch.contiguous(), gt_batch.contiguous().detach().float())

    def evaluate(self, pred_batch, gt_batch):
        """learner.py:644-653: mean absolute error."""
        return torch.mean(torch.abs(pred_batch.contiguous().detach() - gt_batch.contiguous().detach()))

    def train_epoch(self, dataset, lr=0.0001, epoch=None, return_metric=False, use_graph=None):
        """learner.py:170-222: one fine-tuning epoch; dataset yields (mic_sig_batch, {task: labels}).  After the first (eager) step, batches of an
        unchanged shape are replayed as one CUDA graph (graph.py; use_graph=False: never)."""
        from .optim import FusedAdam
        self.model.train()
        optimizer = getattr(self, "_epoch_optimizer", None)
        if optimizer is None or optimizer.model is not self.model or optimizer.m.device != self.model.store.flat.device:
            optimizer = self._epoch_optimizer = FusedAdam(self.model, lr=lr)      # one set of moment buffers, reset every epoch (a captured graph keeps their addresses)
        optimizer.reset()
        optimizer.zero_grad()
        sync = getattr(self, "grad_sync", None)
        losses, metrics = [], []
        for mic_sig_batch, gt_batch in self.device_batches(dataset):
            step = self.graphed_finetune_step(mic_sig_batch, optimizer) if use_graph is not False else None
            if step is not None:
                tar_batch = self.get_tar_batch(gt_batch=gt_batch[self.task].to(self.device))
                try:
                    loss_batch, pred_batch, embed_batch = step.run(mic_sig_batch.to(self.device, non_blocking=True), tar_batch, lr)
                except RuntimeError as e:
                    if step.graph is not None:
                        raise
                    self._graph_failed = True
                    import warnings
                    warnings.warn(f"sarssl_b200: CUDA-graph capture of the fine-tuning step failed, running eagerly ({e})")
                else:
                    losses.append(loss_batch)
                    if return_metric:
                        metrics.append(self.evaluate(pred_batch=pred_batch, gt_batch=tar_batch))
                    continue
            in_batch, tar_batch = self.data_preprocess(mic_sig_batch, gt_batch)
            pred_batch, embed_batch = self.model(in_batch)
            loss_batch = self.loss(pred_batch=pred_batch, gt_batch=tar_batch)
            loss_batch.backward()
            scale = sync.all_reduce() if sync is not None else 1.0
            optimizer.step(lr, grad_scale=scale, zero_grad=True)
            self._eager_finetune_steps = getattr(self, "_eager_finetune_steps", 0) + 1
            losses.append(loss_batch.detach())
            if return_metric:
                metrics.append(self.evaluate(pred_batch=pred_batch, gt_batch=tar_batch))
        loss = float(torch.stack(losses).mean()) if losses else 0.0
        self.check_frontend()
        if return_metric:
            return loss, (torch.stack(metrics).mean() if metrics else torch.zeros(()))
        return loss

    def test_epoch(self, dataset, return_metric=False, return_vis=False):
        """learner.py:224-269."""
        self.model.eval()
        losses, metrics, embed, gt = [], [], [], []
        with torch.no_grad():
            for mic_sig_batch, gt_batch in self.device_batches(dataset):
                in_batch, tar_batch = self.data_preprocess(mic_sig_batch, gt_batch)
                pred_batch, embed_batch = self.model(in_batch)
                losses.append(self.loss(pred_batch=pred_batch, gt_batch=tar_batch))
                if return_metric:
                    metrics.append(self.evaluate(pred_batch=pred_batch, gt_batch=tar_batch))
                if return_vis:
                    embed.append(embed_batch)
                    gt.append(tar_batch)
        out = [float(torch.stack(losses).mean()) if losses else 0.0]
        if return_metric:
            out.append(torch.stack(metrics).mean())
        if return_vis:
            out.append({"embed": torch.cat(embed, dim=0), "label": torch.cat(gt, dim=0)})
        return out[0] if len(out) == 1 else tuple(out)

    def pretrain_evaluate(self, pred_batch, gt_batch, mask_batch):
        """learner.py:574-618: reconstruct the waveforms of prediction and target with the iSTFT (zero DC
        bin re-inserted, rectangular synthesis), normalise each by its global maximum, and report the reconstruction errors.
            pred_batch / gt_batch (nb, nf, nt, 2, 2), mask_batch (nb, nf, nt, 2) with 0 = masked
        Returns {'sig_pred', 'sig_tar', 'mse', 'mse_mask', 'mse_mask_ch', 'pesq', 'pesq_mask_ch'}.  PESQ comes from torchmetrics in the
        reference; when that third-party package is not installed the two PESQ entries are NaN."""
        import ctypes as C
        from ._lib import check, lib, ptr, stream_ptr
        L = lib()

        def patch_layout(t):                      # (nb, nf, nt, 2, 2) -> contiguous (nb, nt, nf, 2, 2)
            v = t.permute(0, 2, 1, 3, 4)
            return (v if v.is_contiguous() else v.contiguous()).float()

        pred, gt = patch_layout(pred_batch), patch_layout(gt_batch)
        nb, nt, nf = pred.shape[:3]
        dev = pred.device
        # compact masks from the dense one: a frame is masked where some entry is 0; the masked microphone is the one holding zeros
        zero = mask_batch == 0                                         # (nb, nf, nt, nch)
        flag = zero.any(dim=3).any(dim=1).to(torch.uint8).contiguous()         # (nb, nt)
        ch = zero.any(dim=2).any(dim=1).float().argmax(dim=1).to(torch.int32).contiguous()
        ws = torch.empty(max(L.sarssl_masked_loss_workspace_bytes(nb, nt), 4096), dtype=torch.uint8, device=dev)
        sums = torch.empty(2, dtype=torch.float32, device=dev)
        check(L.sarssl_eval_mse_sums(ptr(pred), ptr(gt), ptr(flag), ptr(ch), ptr(sums), nb, nt, nf, ptr(ws), ws.numel(), stream_ptr(dev)), "eval_mse_sums")
        sigs = []
        for spec in (pred, gt):
            sig = torch.empty(nb, (nt + 1) * 256, 2, dtype=torch.float32, device=dev)
            check(L.sarssl_istft_patches(ptr(spec), ptr(sig), nb, nt, stream_ptr(dev)), "istft_patches")
            check(L.sarssl_normalize_by_max(ptr(sig), sig.numel(), ptr(ws), ws.numel(), stream_ptr(dev)), "normalize_by_max")
            sigs.append(sig)
        nmasked_elems = flag.sum().float() * nf * 2                     # sum(1 - mask_dense): masked frames x bins x re/im
        result = {"sig_pred": sigs[0], "sig_tar": sigs[1], "mse": sums[0] / pred.numel(), "mse_mask": sums[1] / nmasked_elems,
                  "mse_mask_ch": sums[1] / (nb * nt * nf * 2)}
        try:
            from torchmetrics.functional.audio.pesq import perceptual_evaluation_speech_quality as pesq_fn
        except Exception:
            pesq_fn = None
        pesq = torch.full((nb, 2), float("nan"))
        if pesq_fn is not None:
            for b in range(nb):
                for c in range(2):
                    pesq[b, c] = pesq_fn(sigs[0][b, :, c].cpu(), sigs[1][b, :, c].cpu(), 16000, "wb")
        result["pesq"] = pesq
        result["pesq_mask_ch"] = pesq[torch.arange(nb), ch.cpu().long()]
        return result
