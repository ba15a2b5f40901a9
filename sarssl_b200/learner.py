"""Mirror of the reference's training-step surface (code/learner.py): Learner / STFTLearner.

Same constructor and method names; the step itself is re-designed: the waveform batch goes through one
fused front-end kernel, the model is a single fused forward/backward schedule, Adam is one multi-tensor kernel, and
the loss values are read back once per step."""
from abc import ABC

import numpy as np
import torch

from . import ops
from ._lib import SarsslError


class Learner(ABC):
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


class STFTLearner(Learner):
    """learner.py:488-572."""

    def __init__(self, model, win_len, win_shift_ratio, nfft, fre_used_ratio, fs, mel_scale=False, task=None, ch_mode="M"):
        super().__init__(model)
        if mel_scale or fre_used_ratio != 1 or ch_mode != "M":
            raise SarsslError("STFTLearner: only mel_scale=False, fre_used_ratio=1, ch_mode='M' (the pre-training configuration, "
                              "run_pretrain.py:67-72,208) are on the hot path")
        self.win_len, self.win_shift_ratio, self.nfft, self.fs = win_len, win_shift_ratio, nfft, fs
        self.ch_mode, self.task = ch_mode, task
        from .modules import STFT, ISTFT
        self.stft = STFT(win_len=win_len, win_shift_ratio=win_shift_ratio, nfft=nfft)
        self.istft = ISTFT(win_len=win_len, win_shift_ratio=win_shift_ratio, nfft=nfft, inv=False)
        self.fre_range_used = range(1, int(nfft / 2 * fre_used_ratio) + 1, 1)

    def data_preprocess(self, mic_sig_batch=None, gt_batch=None, eps=1e-6):
        """learner.py:525-572.  mic_sig_batch (nb, nsample, nch) (host or device) -> [ (nb*(nch-1), 2, nf, nt, 2) f32 ].
        The result is a permuted view of patch-layout storage (see modules.as_patch_layout)."""
        data = []
        if mic_sig_batch is not None:
            sig = mic_sig_batch.to(self.device, non_blocking=True)
            patches = ops.stft_frontend(sig, eps=eps, win_len=self.win_len, hop=int(self.win_len * self.win_shift_ratio), nfft=self.nfft)
            data += [patches.permute(0, 4, 2, 1, 3)]
        if gt_batch is not None:
            raise SarsslError("data_preprocess(gt_batch=...) belongs to the downstream path (SURVEY.md 8(f) row 1), not built yet")
        return data
