"""Fused Adam over the model's flat parameter arena (torch.optim.Adam(lr, betas=(0.9, 0.999), weight_decay=0) as created
once per epoch at learner.py:83): one kernel updates all 17.5 M parameters, rescales the gradient (1/world_size after the
all-reduce, 1/accum_steps under gradient accumulation), and clears it for the next step.

Frozen parameters (`requires_grad == False`: `load_checkpoint_best(param_frozen=True)`, learner.py:441-446, the linear-evaluation
flow of run_downstream.py:256, or the substring freezing of run_pretrain.py:364-369) are skipped like torch.optim.Adam skips them:
their gradient ranges are cleared before the update, so their moments stay zero and the update is exactly zero (bit-identical
weights)."""
import torch

from .kernels import KernelSet


class FusedAdam:
    def __init__(self, model, lr=1e-3):
        self.model, self.lr = model, lr
        st = model.store
        self.m = torch.zeros_like(st.flat)
        self.v = torch.zeros_like(st.flat)
        self.t = 0
        self.k = KernelSet(st.flat.device, torch.float32)
        self._frozen_sig, self._frozen_ranges = None, []

    def frozen_ranges(self):
        """Merged [(offset, numel)] arena ranges of the parameters with requires_grad == False (cached on the flag pattern)."""
        st = self.model.store
        sig = tuple(not p.requires_grad for p in st.params.values())
        if sig != self._frozen_sig:
            spans = sorted((st.offsets[k][0], (st.offsets[k][1] + 3) // 4 * 4) for k, p in st.params.items() if not p.requires_grad)
            merged = []
            for o, n in spans:
                if merged and merged[-1][0] + merged[-1][1] == o:
                    merged[-1][1] += n
                else:
                    merged.append([o, n])
            self._frozen_sig, self._frozen_ranges = sig, [(o, n) for o, n in merged]
        return self._frozen_ranges

    def step(self, lr=None, grad_scale=1.0, zero_grad=True, hyper_dev=None):
        """hyper_dev: two device floats {lr / (1 - beta1^t), sqrt(1 - beta2^t)} (hyper_host fills their host image) used INSTEAD of the values
        derived from `lr` and the step count - the form a step replayed as a CUDA graph needs (its launch arguments are frozen)."""
        st = self.model.store
        self.t += 1
        for o, n in self.frozen_ranges():
            self.k.fill(st.grad[o:o + n], 0.0)
        self.k.adam(st.flat, st.grad, self.m, self.v, None, st.total, self.t, float(self.lr if lr is None else lr), grad_scale, zero_grad,
                    hyper_dev=hyper_dev)

    def hyper_host(self, out2, step, lr):
        """out2: float32 CPU tensor of 2 elements <- the bias-corrected step size and sqrt(1 - beta2^step) of `step`, rounded like the kernel's."""
        from ._lib import check
        import ctypes as C
        check(self.k.L.sarssl_adam_hyper_host(C.c_void_p(out2.data_ptr()), int(step), float(lr), 0.9, 0.999), "sarssl_adam_hyper_host")

    def reset(self):
        """A fresh optimizer (the reference creates a new Adam every epoch, learner.py:83) in the same buffers (a captured graph keeps their addresses)."""
        self.m.zero_()
        self.v.zero_()
        self.t = 0

    def zero_grad(self):
        self.k.fill(self.model.store.grad, 0.0)
