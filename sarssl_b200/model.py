"""Mirror of the reference's model surface (code/model.py) - pre-training path."""
import torch
import torch.nn as nn

from . import ops
from .modules import PatchMask, as_patch_layout


class _MaskedLossFn(torch.autograd.Function):
    """loss/diff of model.py:721-747 with the gradient produced by the same kernel launch."""

    @staticmethod
    def forward(ctx, pred, patches, frame_flag, ch_idx, nmasked):
        want = pred.requires_grad
        out2, dpred = ops.masked_loss(pred.detach(), patches, frame_flag, ch_idx, nmasked, want_grad=want)
        ctx.dpred, ctx.flag = dpred, frame_flag
        ctx.mark_non_differentiable(out2)
        loss = out2[0].clone()
        ctx.save_for_backward()
        return loss, out2[1]

    @staticmethod
    def backward(ctx, g_loss, g_diff):
        from ._lib import check, lib, ptr, stream_ptr, dtype_code
        d = ctx.dpred
        nb, nt, w = d.shape
        g = g_loss.reshape(1).float().contiguous()
        check(lib().sarssl_scale_masked_rows(ptr(d), dtype_code(d), ptr(ctx.flag), ptr(g), nb, nt, w // 4, stream_ptr(d.device)),
              "sarssl_scale_masked_rows")
        return d, None, None, None, None


class MaskedReconLoss(nn.Module):
    """The masking + loss tail of SARSSL.forward (model.py:528,585-592) as a stand-alone module: draws the frame / channel
    masks (PatchMask semantics) and evaluates the masked cross-channel reconstruction loss of a prediction.
    Used directly by the front-end + loss benchmark (BASELINE.json configs[1])."""

    def __init__(self, nmasked_patch, npatch, device, rng_state=None):
        super().__init__()
        self.patch_mask = PatchMask("T", nmasked_patch, [1, npatch], device)
        self.rng_state = rng_state

    def forward(self, pred, x):
        patches = as_patch_layout(x)
        nb, nt, nf = patches.shape[:3]
        pidx, cidx, flag = self.patch_mask.draw(nb, nt, 2, self.rng_state)
        loss, diff = _MaskedLossFn.apply(pred.reshape(nb, nt, nf * 4), patches, flag, cidx, self.patch_mask.nmasked_patch)
        return loss, diff, {"mask_patch_idx": pidx, "mask_ch_idx": cidx}
