"""CUDA-event timing of the stem's tensor-core 3x3 conv kernels alone (64 clips, 256 x 256 x 64 ch bf16 maps; inputs rotate over buffers larger than L2)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from sarssl_b200.kernels import KernelSet
dev = torch.device('cuda', 0)
k = KernelSet(dev, torch.bfloat16)
B, H, W = 64, 256, 256
xs = [torch.randn(B, H, W, 64, device=dev).bfloat16() for _ in range(3)]
dy = (torch.randn(B, H, W, 64, device=dev) * 1e-3).bfloat16()
wp = (torch.randn(64, 9, 64, device=dev) / 24).bfloat16()
out = torch.empty_like(xs[0])
stats = torch.cat([torch.zeros(128, device=dev), torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.3])
dw = torch.empty(64, 9, 64, device=dev)
flops = 2.0 * B * H * W * 64 * 576
def run(name, fn, n=20):
    for i in range(3): fn(xs[i % 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(xs[i % 3])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name:34s} {ms*1e3:8.1f} us  {flops/ms/1e9:7.0f} TFLOP/s")
run("conv3x3 forward", lambda x: k.conv3x3_tc(x, wp, out, B, H, W))
run("conv3x3 forward + fused in BN/ReLU", lambda x: k.conv3x3_tc(x, wp, out, B, H, W, in_stats=stats))
run("conv3x3 weight gradient", lambda x: k.conv3x3_wgrad_tc(dy, x, dw, B, H, W))
run("conv3x3 weight gradient + fused in", lambda x: k.conv3x3_wgrad_tc(dy, x, dw, B, H, W, in_stats=stats))
