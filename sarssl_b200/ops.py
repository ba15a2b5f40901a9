"""Thin Python wrappers over the C ABI: allocate outputs with torch (device-memory plumbing only), pass raw pointers.

Every function here launches hand-written sm_100a kernels from libsarssl_b200.so; none has a torch/CPU fallback."""
import ctypes as C
import random as _pyrandom

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr, require_cuda

WIN, HOP, NFFT, NBINS = 512, 256, 512, 257


def num_frames(nsample, win_len=WIN, hop=HOP):
    return int(lib().sarssl_stft_num_frames(int(nsample), int(win_len), int(hop)))


def _f32c(t, name):
    require_cuda(t, name)
    if t.dtype != torch.float32:
        raise _lib.SarsslError(f"{name} must be float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def stft_spectrum(signal, win_len=WIN, hop=HOP, nfft=NFFT):
    """(nb, nsample, nch) f32 -> frame-major complex64 (nb, nt, 257, nch).  common/utils_module.py:49-72"""
    signal = _f32c(signal, "signal")
    nb, nsample, nch = signal.shape
    nt = num_frames(nsample, win_len, hop)
    spec = torch.empty((nb, nt, nfft // 2 + 1, nch), dtype=torch.complex64, device=signal.device)
    check(lib().sarssl_stft_spectrum(ptr(signal), ptr(spec), nb, nsample, nch, win_len, hop, nfft, stream_ptr(signal.device)),
          "sarssl_stft_spectrum")
    return spec


_ws_cache = {}


def _workspace(device, nbytes, tag):
    key = (str(device), tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=device)      # zeroed: the front-end's sticky error flag lives here
        _ws_cache[key] = buf
    return buf


def stft_frontend(signal, eps=1e-6, win_len=WIN, hop=HOP, nfft=NFFT, force_generic=False, out=None):
    """data_preprocess (learner.py:525-572) in patch layout: (nb, nsample, nch) f32 -> (nb*(nch-1), nt, 256, 2, 2) f32."""
    signal = _f32c(signal, "signal")
    nb, nsample, nch = signal.shape
    nt = num_frames(nsample, win_len, hop)
    if out is None:
        out = torch.empty((nb * (nch - 1), nt, nfft // 2, 2, 2), dtype=torch.float32, device=signal.device)
    L = lib()
    variant = int(force_generic)       # 0: fastest applicable kernel (independent-warp kernel), 1: generic three-kernel path, 4: 64-lane-group kernel, 5: independent-warp kernel
    generic = variant == 1
    for _ in range(2):
        nbytes = L.sarssl_stft_workspace_bytes(nb, nsample, nch, int(generic))
        ws = _workspace(signal.device, nbytes, "stft")
        rc = L.sarssl_stft_frontend(ptr(signal), ptr(out), nb, nsample, nch, win_len, hop, nfft, float(eps), 1 if generic else variant, ptr(ws),
                                    ws.numel(), stream_ptr(signal.device))
        if rc == -2 and not generic:       # SARSSL_ERR_WORKSPACE: the generic path was selected on the host, retry with its size
            generic = True
            continue
        check(rc, "sarssl_stft_frontend")
        break
    return out


def stft_frontend_ex(signal, eps=1e-6, win_len=WIN, hop=HOP, nfft=NFFT, all_pairs=False, first_bin=1, nbins=256):
    """data_preprocess variants off the hot path (ch_mode 'MM', fre_used_ratio 0.5): (nb, nsample, nch) -> (items, nt, nbins, 2, 2) f32."""
    signal = _f32c(signal, "signal")
    nb, nsample, nch = signal.shape
    nt = num_frames(nsample, win_len, hop)
    items = nb * (nch * (nch - 1) // 2 if all_pairs else nch - 1)
    out = torch.empty((items, nt, nbins, 2, 2), dtype=torch.float32, device=signal.device)
    L = lib()
    ws = _workspace(signal.device, L.sarssl_stft_workspace_bytes(nb, nsample, nch, 1), "stft")
    check(L.sarssl_stft_frontend_ex(ptr(signal), ptr(out), nb, nsample, nch, win_len, hop, nfft, float(eps), int(all_pairs), int(first_bin), int(nbins),
                                    ptr(ws), ws.numel(), stream_ptr(signal.device)), "sarssl_stft_frontend_ex")
    return out


def stft_frontend_check(device):
    """Synchronises and raises if the fused kernel's clip rendezvous timed out in ANY launch since the workspace was created (the
    flag is sticky).  Learner checks it once per epoch, at the loss read-back."""
    ws = _ws_cache.get((str(device), "stft"))
    if ws is None:
        return
    flag = C.c_int(0)
    check(lib().sarssl_stft_frontend_error_flag(ptr(ws), C.byref(flag), stream_ptr(device)), "error_flag")
    if flag.value:
        raise _lib.SarsslError("stft_frontend: clip rendezvous timed out")


def istft(spec_bfkc_view, win_len=WIN, hop=HOP, nfft=NFFT):
    """spec: complex64 tensor indexed (nb, nf, nt, nch) with arbitrary strides -> (nb, (nt+1)*hop, nch) f32.
    common/utils_module.py:91-113"""
    require_cuda(spec_bfkc_view, "stft")
    if spec_bfkc_view.dtype != torch.complex64:
        raise _lib.SarsslError("istft expects complex64")
    nb, nf, nt, nch = spec_bfkc_view.shape
    if nf != nfft // 2 + 1:
        raise _lib.SarsslError(f"istft expects {nfft // 2 + 1} bins, got {nf}")
    sb, sk, st, sc = spec_bfkc_view.stride()
    sig = torch.empty((nb, (nt + 1) * hop, nch), dtype=torch.float32, device=spec_bfkc_view.device)
    check(lib().sarssl_istft(ptr(spec_bfkc_view), ptr(sig), nb, nt, nch, sb, st, sk, sc, win_len, hop, nfft,
                             stream_ptr(spec_bfkc_view.device)), "sarssl_istft")
    return sig


# ----------------------------------------------------------------------------------------------------------------
# mask indices (host, bit-exact with CPython random)
# ----------------------------------------------------------------------------------------------------------------

def mt_state_from_python(rng=None):
    st = (rng or _pyrandom).getstate()
    return np.array(st[1], dtype=np.uint32), st


def mt_state_to_python(state, template, rng=None):
    (rng or _pyrandom).setstate((template[0], tuple(int(v) for v in state), template[2]))


def mt_seed(seed):
    """State equal to random.seed(int(seed)) (key = 32-bit little-endian words of |seed|)."""
    seed = abs(int(seed))
    words = []
    while True:
        words.append(seed & 0xFFFFFFFF)
        seed >>= 32
        if seed == 0:
            break
    key = np.array(words, dtype=np.uint32)
    state = np.zeros(625, dtype=np.uint32)
    check(lib().sarssl_mt19937_seed_host(state.ctypes.data_as(C.c_void_p), key.ctypes.data_as(C.c_void_p), len(words)), "mt_seed")
    return state


def draw_masks(state, nb, npatch, nmasked, nmic=2):
    """Advance `state` (np.uint32[625]) exactly like the reference's per-item random.sample + random.randint.
    Returns numpy (patch_idx int64 (nb, nmasked), ch_idx int64 (nb,), frame_flag uint8 (nb, npatch))."""
    pidx = np.empty((nb, nmasked), dtype=np.int64)
    cidx = np.empty((nb,), dtype=np.int64)
    flag = np.empty((nb, npatch), dtype=np.uint8)
    check(lib().sarssl_mt19937_draw_masks_host(state.ctypes.data_as(C.c_void_p), nb, npatch, nmasked, nmic,
                                               pidx.ctypes.data_as(C.c_void_p), cidx.ctypes.data_as(C.c_void_p),
                                               flag.ctypes.data_as(C.c_void_p)), "draw_masks")
    return pidx, cidx, flag


def draw_masks_python_stream(nb, npatch, nmasked, nmic=2, rng=None):
    """Drop-in behaviour: consume the process-global python `random` stream (what PatchMask.forward does in the
    reference, common/utils_module.py:263-267) through the C++ generator and write the advanced state back."""
    state, template = mt_state_from_python(rng)
    out = draw_masks(state, nb, npatch, nmasked, nmic)
    mt_state_to_python(state, template, rng)
    return out


# ----------------------------------------------------------------------------------------------------------------
# masked reconstruction loss
# ----------------------------------------------------------------------------------------------------------------

def masked_loss(pred, patches, frame_flag, ch_idx, nmasked, want_grad=True, out2=None, dpred=None):
    """pred (nb, nt, nf*4) f32|bf16, patches (nb, nt, nf, 2, 2) f32, frame_flag (nb, nt) uint8, ch_idx (nb,) int32.
    Returns (out2 [loss, diff] f32 device tensor, dpred or None).  model.py:585-592,721-747"""
    require_cuda(pred, "pred")
    nb, nt = patches.shape[0], patches.shape[1]
    nf = patches.shape[2]
    assert pred.is_contiguous() and patches.is_contiguous() and patches.dtype == torch.float32
    assert frame_flag.dtype == torch.uint8 and ch_idx.dtype == torch.int32
    if out2 is None:
        out2 = torch.empty(2, dtype=torch.float32, device=pred.device)
    if want_grad and dpred is None:
        dpred = torch.empty_like(pred)
    L = lib()
    ws = _workspace(pred.device, L.sarssl_masked_loss_workspace_bytes(nb, nt), "loss")
    check(L.sarssl_masked_loss(ptr(pred), _lib.dtype_code(pred), ptr(patches), ptr(frame_flag), ptr(ch_idx), ptr(out2),
                               ptr(dpred) if want_grad else C.c_void_p(0), nb, nt, nf, int(nmasked), ptr(ws), ws.numel(),
                               stream_ptr(pred.device)), "sarssl_masked_loss")
    return out2, (dpred if want_grad else None)
