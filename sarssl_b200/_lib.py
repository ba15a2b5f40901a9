"""ctypes loader for libsarssl_b200.so.  Prototypes are parsed from include/sarssl_b200.h, the single source of truth
for the C ABI.  There is no fallback of any kind: if the shared library is missing, importing any compute op raises, and
every entry point that launches a kernel raises SarsslError on a non-zero return code."""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsarssl_b200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "sarssl_b200.h")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_SWISH = 0, 1, 2


class SarsslError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    """struct sarssl_gemm_args (include/sarssl_b200.h)."""
    _fields_ = [("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p), ("pre_out", C.c_void_p), ("resid", C.c_void_p), ("bias", C.c_void_p),
                ("sAm", C.c_longlong), ("sAk", C.c_longlong), ("sAb1", C.c_longlong), ("sAb2", C.c_longlong),
                ("sBn", C.c_longlong), ("sBk", C.c_longlong), ("sBb1", C.c_longlong), ("sBb2", C.c_longlong),
                ("ldc", C.c_longlong), ("sCb1", C.c_longlong), ("sCb2", C.c_longlong), ("ldr", C.c_longlong),
                ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("nb1", C.c_int), ("nb2", C.c_int),
                ("ab_dtype", C.c_int), ("c_dtype", C.c_int), ("act", C.c_int), ("accumulate", C.c_int),
                ("alpha", C.c_float), ("beta", C.c_float),
                ("drop_p", C.c_float), ("drop_seed", C.c_ulonglong),
                ("a_drop_p", C.c_float), ("a_drop_seed", C.c_ulonglong), ("seed_dev", C.c_void_p)]


_SCALARS = {"int": C.c_int, "long long": C.c_longlong, "float": C.c_float, "double": C.c_double, "size_t": C.c_size_t,
            "unsigned long long": C.c_ulonglong, "cudaStream_t": C.c_void_p}


def _ctype(decl):
    decl = decl.strip()
    if decl in ("void", ""):
        return None
    if "*" in decl:
        return C.c_char_p if decl.replace(" ", "") == "constchar*" else C.c_void_p
    words = decl.replace("const ", "").split()
    if len(words) > 1 and " ".join(words) not in _SCALARS:
        words = words[:-1]                       # drop the parameter name
    key = " ".join(words)
    if key not in _SCALARS:
        raise SarsslError(f"cannot map C type '{decl}'")
    return _SCALARS[key]


def parse_header(path=HEADER):
    """{name: (restype, [argtypes])} for every `sarssl_*` function declared in the public header."""
    src = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
    src = re.sub(r"typedef struct sarssl_gemm_args \{.*?\} sarssl_gemm_args;", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z_ \*]*?)\b(sarssl_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3)
        restype = _ctype(ret if "*" in ret else ret + " x") if ret != "const char*" else C.c_char_p
        if ret.replace(" ", "") == "constchar*":
            restype = C.c_char_p
        argtypes = [t for t in (_ctype(a) for a in args.split(",")) if t is not None]
        protos[name] = (restype, argtypes)
    return protos


_lib = None
PROTOTYPES = None


def lib():
    global _lib, PROTOTYPES
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SarsslError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(sarssl_b200 has no CPU or PyTorch fallback)")
        l = C.CDLL(LIB_PATH)
        PROTOTYPES = parse_header()
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)          # AttributeError here = header declares something the library lacks
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
    dt = t if isinstance(t, torch.dtype) else t.dtype
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise SarsslError(f"unsupported dtype {dt}")


def require_cuda(t, name):
    if not t.is_cuda:
        raise SarsslError(f"{name} must be a CUDA tensor: sarssl_b200 has no CPU path (got device {t.device})")
