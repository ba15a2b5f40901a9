"""A few launches of the stem's tensor-core 3x3 conv kernels on 64 clips (for ncu): forward plain, forward with the fused input BatchNorm + ReLU, weight gradient."""
import sys, torch
sys.path.insert(0, '/root/repo')
from sarssl_b200.kernels import KernelSet
dev = torch.device('cuda', 0)
k = KernelSet(dev, torch.bfloat16)
B, H, W = 64, 256, 256
x = torch.randn(B, H, W, 64, device=dev).bfloat16()
dy = (torch.randn(B, H, W, 64, device=dev) * 1e-3).bfloat16()
wp = (torch.randn(64, 9, 64, device=dev) / 24).bfloat16()
out = torch.empty_like(x)
stats = torch.cat([torch.zeros(128, device=dev), torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.3])
dw = torch.empty(64, 9, 64, device=dev)
for _ in range(3):
    k.conv3x3_tc(x, wp, out, B, H, W)
    k.conv3x3_tc(x, wp, out, B, H, W, in_stats=stats)
    k.conv3x3_wgrad_tc(dy, x, dw, B, H, W)
    k.conv3x3_wgrad_tc(dy, x, dw, B, H, W, in_stats=stats)
torch.cuda.synchronize()
