"""One launch set of a single STFT front-end variant at the configs[1] size (for ncu): python scripts/stft_one.py <variant> [nb]"""
import sys, torch
sys.path.insert(0, '/root/repo')
from sarssl_b200 import ops
variant = int(sys.argv[1]); nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
sig = 0.1 * torch.randn(nb, 65792, 2, device='cuda')
out = torch.empty(nb, 256, 256, 2, 2, device='cuda')
for _ in range(3): ops.stft_frontend(sig, out=out, force_generic=variant)
torch.cuda.synchronize()
