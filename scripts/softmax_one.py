"""A few launches of the score softmax / shift kernels (for ncu): python scripts/softmax_one.py [T] [B]"""
import sys, math, torch
sys.path.insert(0, '/root/repo')
from sarssl_b200.kernels import KernelSet
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
H = 4
k = KernelSet(torch.device('cuda', 0), torch.bfloat16)
content = torch.randn(B, H, T, T, device='cuda').bfloat16(); pos = torch.randn(H, B, T, T, device='cuda').bfloat16()
prob = torch.empty_like(content); attn = torch.empty_like(content); dpos = torch.empty_like(pos)
dattn = torch.randn(B, H, T, T, device='cuda').bfloat16()
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
nbytes = content.numel() * 2
print("T=%d B=%d: softmax fwd %.3f ms (%.0f GB/s of 4 tensors)  bwd+unshift %.3f ms (%.0f GB/s of 4 tensors)" % (
    T, B, (a := t(lambda: k.attn_softmax_fwd(content, pos, prob, attn, B, H, T, 1 / math.sqrt(256), (0.1, 7)))), 4 * nbytes / a / 1e6,
    (b := t(lambda: k.attn_softmax_bwd(dattn, prob, dpos, B, H, T, 1 / math.sqrt(256), (0.1, 7)))), 4 * nbytes / b / 1e6))
