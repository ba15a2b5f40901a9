"""Fused Adam over the model's flat parameter arena (torch.optim.Adam(lr, betas=(0.9, 0.999), weight_decay=0) as created
once per epoch at learner.py:83): one kernel updates all 17.5 M parameters, rescales the gradient (1/world_size after the
all-reduce), and clears it for the next step."""
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

    def step(self, lr=None, grad_scale=1.0, zero_grad=True):
        st = self.model.store
        self.t += 1
        self.k.adam(st.flat, st.grad, self.m, self.v, None, st.total, self.t, float(self.lr if lr is None else lr), grad_scale, zero_grad)

    def zero_grad(self):
        self.k.fill(self.model.store.grad, 0.0)
