"""CPU: the C-ABI library loads, exports every symbol include/sarssl_b200.h declares, and the host-only entry points
(mask RNG, size queries, argument validation) behave.  No kernel is launched here."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest

from sarssl_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sarssl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sarssl_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    l = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(l, s), f"{s} declared in include/sarssl_b200.h but not exported"
    assert set(_lib.PROTOTYPES) == set(syms), set(_lib.PROTOTYPES) ^ set(syms)
    import subprocess
    exported = {ln.split()[-1] for ln in subprocess.run(["nm", "-D", _lib.LIB_PATH], capture_output=True, text=True).stdout.splitlines() if " T sarssl_" in ln}
    assert exported == set(syms), exported ^ set(syms)
    assert l.sarssl_version() >= 100


def test_num_frames_and_workspace_queries():
    assert ops.num_frames(65792) == 256 and ops.num_frames(262400) == 1024 and ops.num_frames(64000) == 249
    assert ops.num_frames(511) == 0 and ops.num_frames(512) == 1
    l = _lib.lib()
    small = l.sarssl_stft_workspace_bytes(4, 65792, 2, 0)
    big = l.sarssl_stft_workspace_bytes(4, 65792, 2, 1)
    assert 0 < small < big and big - small == 4 * 256 * 257 * 2 * 8


def test_bad_arguments_are_reported_not_crashed():
    l = _lib.lib()
    rc = l.sarssl_stft_spectrum(None, None, 1, 1024, 2, 512, 256, 512, None)
    assert rc == -1 and b"null" in l.sarssl_last_error()
    rc = l.sarssl_stft_spectrum(C.c_void_p(16), C.c_void_p(16), 1, 1024, 2, 400, 160, 512, None)
    assert rc == -3 and b"512" in l.sarssl_last_error()
    with pytest.raises(_lib.SarsslError):
        import torch
        ops.stft_spectrum(torch.zeros(1, 1024, 2))          # CPU tensor: there is no CPU path


@pytest.mark.parametrize("seed,nb,npatch,nmasked,nmic", [
    (400000001, 8, 256, 128, 2), (1, 3, 16, 8, 2), (2 ** 40 + 17, 2, 1024, 512, 2), (0, 4, 249, 124, 2),
    (5, 6, 100, 3, 2),           # n > setsize: random.sample's rejection-set branch
    (6, 3, 10000, 5, 4),         # set branch, wider _randbelow, 4 microphones
    (7, 2, 64, 64, 1),           # everything masked, single microphone
])
def test_mask_stream_bit_exact_with_cpython(seed, nb, npatch, nmasked, nmic):
    rng = random.Random(seed)
    want_p = np.array([rng.sample(range(npatch), nmasked) + [rng.randint(0, nmic - 1)] for _ in range(nb)], dtype=np.int64)
    state = ops.mt_seed(seed)
    p, c, flag = ops.draw_masks(state, nb, npatch, nmasked, nmic)
    assert np.array_equal(p, want_p[:, :-1]) and np.array_equal(c, want_p[:, -1])
    assert flag.sum() == nb * nmasked and all(flag[b, p[b]].all() for b in range(nb))
    # the advanced state equals python's: the next draws agree too
    st = rng.getstate()
    assert np.array_equal(np.array(st[1], dtype=np.uint32), state)


def test_mask_stream_golden_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "mask_streams.npz"))
    for key in g.files:
        if not key.startswith("p/"):
            continue
        parts = key.split("/")
        seed, nb, npatch = int(parts[1]), int(parts[2]), int(parts[3])
        nmic = 5 if key.endswith("nmic5") else 2
        p, c, _ = ops.draw_masks(ops.mt_seed(seed), nb, npatch, npatch // 2, nmic)
        assert np.array_equal(p, g[key]) and np.array_equal(c, g["c/" + key[2:]][:, 0])


def test_global_python_stream_roundtrip():
    random.seed(400000003)
    want = [(random.sample(range(256), 128), random.randint(0, 1)) for _ in range(4)]
    tail = random.random()
    random.seed(400000003)
    p, c, _ = ops.draw_masks_python_stream(4, 256, 128, 2)
    assert [list(r) for r in p] == [w[0] for w in want] and list(c) == [w[1] for w in want]
    assert random.random() == tail


def test_rejects_impossible_sample():
    state = ops.mt_seed(1)
    with pytest.raises(_lib.SarsslError):
        ops.draw_masks(state, 1, 8, 9, 2)
