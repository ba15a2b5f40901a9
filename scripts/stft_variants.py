"""Time the STFT front-end variants (GPU box): 4 = fused kernel (64-lane FFT groups), 3 = the same with the load / rendezvous / store pipelined inside the CTA (default), 2 = experimental warp-worker kernel, 5 = independent warps with the rendezvous one frame behind, 1 = generic 3-kernel path."""
import sys, torch
sys.path.insert(0, '/root/repo')
from sarssl_b200 import ops
nb, ns = 1024, 65792
sig = 0.1 * torch.randn(nb, ns, 2, device='cuda')
out = torch.empty(nb, 256, 256, 2, 2, device='cuda')
ref = None
for variant in (4, 5, 0, 1):
    for _ in range(3): ops.stft_frontend(sig, out=out, force_generic=variant)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.stft_frontend(sig, out=out, force_generic=variant)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    if ref is None: ref = out.clone()
    err = float((out - ref).norm() / ref.norm())
    print('variant %d: %.3f ms  %.0f GB/s algorithmic (%.1f%% of 6650)  rel diff vs variant 4: %.2e' % (variant, ms, nb * 1574912 / ms / 1e6, nb * 1574912 / ms / 1e6 / 66.5, err))
