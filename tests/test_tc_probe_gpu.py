"""GPU hardware probe: UMMA descriptors that start a whole number of 128-byte rows inside a swizzle atom (what lets one TMA
box serve the three horizontal taps of the 3x3 convolution).  Records which descriptor encoding the hardware honours."""
import pytest
import torch

from sarssl_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def run(mode, row_off, use_bo):
    g = torch.Generator(device=DEV).manual_seed(5 + mode)
    if mode == 0:
        A = torch.randn(136, 64, device=DEV, generator=g).bfloat16()
        B = torch.randn(64, 64, device=DEV, generator=g).bfloat16()
        want = A[row_off:row_off + 128].float() @ B.float().T
    else:
        A = torch.randn(72, 128, device=DEV, generator=g).bfloat16()
        B = torch.randn(64, 64, device=DEV, generator=g).bfloat16()
        want = A[row_off:row_off + 64].float().T @ B.float()
    D = torch.empty(128, 64, device=DEV)
    _lib.check(_lib.lib().sarssl_probe_umma_row_offset(_lib.ptr(A), _lib.ptr(B), _lib.ptr(D), mode, row_off, int(use_bo), _lib.stream_ptr(DEV)), "probe")
    torch.cuda.synchronize()
    return float((D - want).norm() / want.norm())


def test_row_offset_descriptors():
    report = {}
    for mode in (0, 1):
        for use_bo in (0, 1):
            report[(mode, use_bo)] = [run(mode, r, use_bo) for r in range(0, 4)]
    print("\nUMMA row-offset probe (relative error per row_off 0..3):")
    for k, v in report.items():
        print("  mode %d base_offset_field %d :" % k, " ".join("%.2e" % e for e in v))
    for mode in (0, 1):
        assert report[(mode, 0)][0] < 1e-5                      # aligned start must always work
    # at least one encoding must serve shifted starts for each operand layout
    for mode in (0, 1):
        assert all(e < 1e-5 for e in report[(mode, 0)]) or all(e < 1e-5 for e in report[(mode, 1)]), report
