"""In-tree build of libsarssl_b200.so (nvcc, sm_100a only).  No torch headers, no JIT cache: the .so sits next to this
file so it travels to the GPU box with the repo snapshot."""
import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libsarssl_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--use_fast_math=false",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _nccl_dirs():
    """NCCL headers + library shipped inside the torch wheel (nvidia/nccl)."""
    try:
        import nvidia.nccl as n
        base = list(n.__path__)[0]
        inc, lib = os.path.join(base, "include"), os.path.join(base, "lib")
        if os.path.exists(os.path.join(inc, "nccl.h")):
            return inc, lib
    except Exception:
        pass
    return None, None


def _stale(src, obj):
    if not os.path.exists(obj):
        return True
    newest_dep = max(os.path.getmtime(p) for p in [src] + glob.glob(os.path.join(CSRC, "*.cuh")) +
                     glob.glob(os.path.join(HERE, "..", "include", "*.h")) + [os.path.abspath(__file__)])
    return os.path.getmtime(obj) < newest_dep


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))
    inc, nccl_lib = _nccl_dirs()
    extra = ["-I", inc] if inc else []
    jobs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s) + ".o")
        if force or _stale(s, o):
            flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
            jobs.append((s, [nvcc, "-c", s, "-o", o] + flags + extra))
    logs = {}

    def run(job):
        s, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return s, r.returncode, r.stdout + r.stderr

    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, rc, log in ex.map(run, jobs):
            logs[s] = log
            if rc != 0:
                raise RuntimeError(f"nvcc failed for {s}:\n{log}")
            if verbose:
                print(log)
    objs = [os.path.join(OBJ, os.path.basename(s) + ".o") for s in srcs]
    if jobs or force or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"]
        if nccl_lib:
            so = sorted(glob.glob(os.path.join(nccl_lib, "libnccl.so*")))
            if so:
                cmd += ["-L", nccl_lib, "-l:" + os.path.basename(so[0]), "-Xlinker", "-rpath=" + nccl_lib]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(os.path.join(OBJ, "ptxas.log"), "a") as f:
        for s, log in logs.items():
            f.write(f"==== {s}\n{log}\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
