"""Which part of the fused GEMM epilogue costs what (M=32768, N=1024, K=256, bf16)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from sarssl_b200.kernels import KernelSet, ACT_SWISH, ACT_RELU
k = KernelSet('cuda', torch.bfloat16)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
M, N, K = 32768, 1024, 256
A = torch.randn(M, K, device='cuda').bfloat16(); B = torch.randn(N, K, device='cuda').bfloat16() / K ** 0.5
C = torch.empty(M, N, device='cuda', dtype=torch.bfloat16); bias = torch.randn(N, device='cuda'); pre = torch.empty_like(C)
R = torch.randn(M, N, device='cuda').bfloat16()
cases = [("plain", {}), ("bias", dict(bias=bias)), ("bias+relu", dict(bias=bias, act=ACT_RELU)), ("bias+swish", dict(bias=bias, act=ACT_SWISH)),
         ("bias+pre", dict(bias=bias, pre=pre)), ("bias+drop", dict(bias=bias, drop=(0.1, 5))), ("bias+resid", dict(bias=bias, resid=R, ldr=N)),
         ("bias+swish+pre", dict(bias=bias, act=ACT_SWISH, pre=pre)), ("bias+swish+drop", dict(bias=bias, act=ACT_SWISH, drop=(0.1, 5))),
         ("bias+swish+pre+drop", dict(bias=bias, act=ACT_SWISH, pre=pre, drop=(0.1, 5)))]
for name, kw in cases:
    ms = t(lambda: k.linear(A, B, C, M, N, K, **kw))
    print("%-24s %7.3f ms  %7.1f TFLOP/s" % (name, ms, 2.0 * M * N * K / ms / 1e9))
