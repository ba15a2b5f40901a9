"""ctypes loader for libsarssl_b200.so (the C ABI declared in include/sarssl_b200.h).

There is no fallback of any kind: if the shared library is missing the import of any compute op raises, and every
entry point that launches a kernel raises SarsslError on a non-zero return code."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsarssl_b200.so")

F32, BF16 = 0, 1


class SarsslError(RuntimeError):
    pass


_lib = None

_vp, _i, _ll, _f, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/sarssl_b200.h one to one (tests/test_abi.py checks the header)
PROTOTYPES = {
    "sarssl_version": (_i, []),
    "sarssl_last_error": (C.c_char_p, []),
    "sarssl_stft_num_frames": (_i, [_ll, _i, _i]),
    "sarssl_stft_spectrum": (_i, [_vp, _vp, _i, _ll, _i, _i, _i, _i, _vp]),
    "sarssl_stft_workspace_bytes": (_sz, [_i, _ll, _i, _i]),
    "sarssl_stft_frontend": (_i, [_vp, _vp, _i, _ll, _i, _i, _i, _i, _f, _i, _vp, _sz, _vp]),
    "sarssl_stft_frontend_error_flag": (_i, [_vp, C.POINTER(_i), _vp]),
    "sarssl_istft": (_i, [_vp, _vp, _i, _i, _i, _ll, _ll, _ll, _ll, _i, _i, _i, _vp]),
    "sarssl_mt19937_seed_host": (_i, [_vp, _vp, _i]),
    "sarssl_mt19937_draw_masks_host": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "sarssl_expand_masks": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "sarssl_to_patch_layout": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "sarssl_masked_loss_workspace_bytes": (_sz, [_i, _i]),
    "sarssl_masked_loss": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "sarssl_scale_masked_rows": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _vp]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SarsslError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(sarssl_b200 has no CPU or PyTorch fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().sarssl_last_error().decode(errors="replace")
        raise SarsslError(f"{what} failed with code {rc}: {msg}")


def stream_ptr(device=None):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def dtype_code(t):
    import torch
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise SarsslError(f"unsupported dtype {t.dtype}")


def require_cuda(t, name):
    if not t.is_cuda:
        raise SarsslError(f"{name} must be a CUDA tensor: sarssl_b200 has no CPU path (got device {t.device})")
