"""Mirror of the reference's basic modules (common/utils_module.py) on top of the sm_100a kernels.

Same class names, constructor arguments, call signatures and result shapes as the reference; the data movement is
redesigned: STFT results are stored frame-major and handed back as permuted *views* with the reference's index order,
PatchSplit / PatchRecover are pure index maps (zero-copy views), PatchMask keeps masks compact and only materialises
the three dense tensors because its reference signature returns them."""
import torch
import torch.nn as nn

from . import ops
from ._lib import SarsslError, check, lib, ptr, stream_ptr


class STFT(nn.Module):
    """common/utils_module.py:28-72.  signal (nb, nsample, nch) -> stft (nb, nf, nt, nch) complex64."""

    def __init__(self, win_len, win_shift_ratio, nfft, win="hann", inv=False):
        super().__init__()
        self.win_len, self.win_shift_ratio, self.nfft, self.win, self.inv = win_len, win_shift_ratio, nfft, win, inv
        if win != "hann" or inv:
            raise SarsslError("sarssl_b200.STFT implements the hot-path configuration only: win='hann', inv=False")

    def forward(self, signal):
        hop = int(self.win_len * self.win_shift_ratio)
        spec = ops.stft_spectrum(signal, self.win_len, hop, self.nfft)        # (nb, nt, nf, nch) storage
        return spec.permute(0, 2, 1, 3)                                       # reference index order, zero-copy


class ISTFT(nn.Module):
    """common/utils_module.py:75-113 (inv=False).  stft (nb, nf, nt, nch) complex64 -> signal (nb, (nt+1)*hop, nch)."""

    def __init__(self, win_len, win_shift_ratio, nfft, inv=False):
        super().__init__()
        self.win_len, self.win_shift_ratio, self.nfft, self.inv = win_len, win_shift_ratio, nfft, inv
        if inv:
            raise SarsslError("sarssl_b200.ISTFT implements inv=False only (the branch the learner uses, learner.py:500-505)")

    def forward(self, stft):
        return ops.istft(stft, self.win_len, int(self.win_len * self.win_shift_ratio), self.nfft)


class PatchSplit(nn.Module):
    """common/utils_module.py:175-207 for frame patches (patch_shape = (nf, 1)): a pure index map
    vec[b, t, f, r, m] = data[b, f, t, r, m], returned as a view."""

    def __init__(self, patch_shape, f_first=False):
        super().__init__()
        self.patch_shape, self.f_first = patch_shape, f_first

    def forward(self, data):
        if self.f_first or self.patch_shape[1] != 1 or data.shape[1] != self.patch_shape[0]:
            raise SarsslError("PatchSplit: only frame patches (patch_shape == (nf, 1), patch_mode 'T') are on the hot path")
        if data.dim() == 4:
            return data.permute(0, 2, 1, 3)
        return data.permute(0, 2, 1, 3, 4)


class PatchRecover(nn.Module):
    """common/utils_module.py:210-244 for frame patches: inverse index map of PatchSplit (view)."""

    def __init__(self, output_shape, patch_shape, f_first=False):
        super().__init__()
        self.output_shape, self.patch_shape, self.f_first = output_shape, patch_shape, f_first

    def forward(self, data):
        if self.f_first or self.patch_shape[1] != 1:
            raise SarsslError("PatchRecover: only frame patches (patch_shape == (nf, 1)) are on the hot path")
        if data.dim() == 4:
            return data.permute(0, 2, 1, 3)
        return data.permute(0, 2, 1, 3, 4)


class PatchMask(nn.Module):
    """common/utils_module.py:247-272 with patch_mode 'T'.  Consumes the process-global python `random` stream exactly
    like the reference (random.sample then random.randint per item) through the C++ MT19937 restatement."""

    def __init__(self, patch_mode, nmasked_patch, npatch_shape, device):
        super().__init__()
        if patch_mode != "T":
            raise SarsslError("PatchMask: only patch_mode 'T' (frame masking, the shipped default) is implemented")
        self.patch_mode, self.nmasked_patch, self.npatch_shape, self.device = patch_mode, nmasked_patch, npatch_shape, device

    def draw_host(self, nbatch, npatch, nmic, rng_state=None, dp=None):
        """Compact masks as numpy arrays: (patch_idx int64 (nb, nmasked), ch_idx int64 (nb, 1), frame_flag uint8 (nb, npatch)).
        dp = (rank, world): every rank draws the masks of the GLOBAL batch (nbatch * world items, same seed everywhere) and keeps
        rows [rank*nbatch, (rank+1)*nbatch) - bit-exact with the single-process reference on the concatenated batch."""
        rank, world = dp if dp is not None else (0, 1)
        total = nbatch * world
        if rng_state is None:
            pidx, cidx, flag = ops.draw_masks_python_stream(total, npatch, self.nmasked_patch, nmic)
        else:
            pidx, cidx, flag = ops.draw_masks(rng_state, total, npatch, self.nmasked_patch, nmic)
        if world > 1:
            sl = slice(rank * nbatch, (rank + 1) * nbatch)
            pidx, cidx, flag = pidx[sl].copy(), cidx[sl].copy(), flag[sl].copy()
        return pidx, cidx, flag

    def draw(self, nbatch, npatch, nmic, rng_state=None, dp=None):
        """draw_host on the device: (patch_idx int64 (nb, nmasked), ch_idx int32 (nb,), frame_flag uint8 (nb, npatch))."""
        pidx, cidx, flag = self.draw_host(nbatch, npatch, nmic, rng_state, dp)
        dev = self.device
        to = lambda a: torch.from_numpy(a).pin_memory().to(dev, non_blocking=True) if str(dev) != "cpu" else torch.from_numpy(a)
        return to(pidx), to(cidx.astype("int32")), to(flag)

    def forward(self, data_shape):
        nbatch, npatch, dpatch, _, nmic = data_shape
        pidx, cidx, flag = self.draw(nbatch, npatch, nmic)
        shape = (nbatch, npatch, dpatch, nmic)
        mask = torch.empty(shape, device=self.device)
        mask_patch = torch.empty(shape, device=self.device)
        mask_ch = torch.empty(shape, device=self.device)
        check(lib().sarssl_expand_masks(ptr(flag), ptr(cidx), ptr(mask), ptr(mask_patch), ptr(mask_ch), nbatch, npatch, dpatch, nmic,
                                        stream_ptr(mask.device)), "sarssl_expand_masks")
        return mask, mask_patch, mask_ch, pidx, cidx.long()[:, None]


def as_patch_layout(x):
    """x indexed like the reference's model input (nb, 2, nf, nt, 2).  Returns the (nb, nt, nf, 2, 2) patch-layout tensor:
    zero-copy when x is the view our front-end returns, one transposition kernel otherwise."""
    if x.dim() != 5 or x.shape[1] != 2 or x.shape[4] != 2:
        raise SarsslError(f"expected (nb, 2, nf, nt, 2), got {tuple(x.shape)}")
    v = x.permute(0, 3, 2, 4, 1)
    if v.is_contiguous():
        return v
    x = x.contiguous()
    nb, _, nf, nt, _ = x.shape
    out = torch.empty((nb, nt, nf, 2, 2), dtype=torch.float32, device=x.device)
    check(lib().sarssl_to_patch_layout(ptr(x), ptr(out), nb, nf, nt, stream_ptr(x.device)), "sarssl_to_patch_layout")
    return out
