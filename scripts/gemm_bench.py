"""Micro-benchmark of sarssl_gemm_tc at the model's shapes (GPU box)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from sarssl_b200.kernels import KernelSet, ACT_SWISH
k = KernelSet('cuda', torch.bfloat16)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
M = 32768
rows = []
ONLY = sys.argv[1] if len(sys.argv) > 1 else None          # substring filter on the case name; "fwd" as 2nd arg skips dgrad / wgrad
FWD_ONLY = len(sys.argv) > 2 and sys.argv[2] == "fwd"
for name, N, K, kw in [("ffn1 spec (bias+swish+drop, pre)", 2048, 512, "ffn1"), ("ffn2 spec (bias+drop+resid)", 512, 2048, "ffn2"), ("qkv spec", 1536, 512, None),
                       ("ffn1 spat", 1024, 256, "ffn1"), ("ffn2 spat", 256, 1024, "ffn2"), ("decoder0 (relu)", 3072, 768, None), ("decoder2", 1024, 3072, None),
                       ("patch embed spec", 512, 1024, None)]:
    if ONLY and ONLY not in name:
        continue
    A = torch.randn(M, K, device='cuda').bfloat16(); B = torch.randn(N, K, device='cuda').bfloat16() / K ** 0.5
    C = torch.empty(M, N, device='cuda', dtype=torch.bfloat16); bias = torch.randn(N, device='cuda')
    if kw == "ffn1":
        pre = torch.empty_like(C); fn = lambda: k.linear(A, B, C, M, N, K, bias=bias, act=ACT_SWISH, pre=pre, drop=(0.1, 5))
    elif kw == "ffn2":
        R = torch.randn(M, N, device='cuda').bfloat16(); fn = lambda: k.linear(A, B, C, M, N, K, bias=bias, resid=R, ldr=N, beta=0.5, drop=(0.1, 5))
    else:
        fn = lambda: k.linear(A, B, C, M, N, K, bias=bias)
    ms = t(fn); rows.append((name, M, N, K, ms, 2.0 * M * N * K / ms / 1e9))
    if FWD_ONLY:
        continue
    # dgrad (A K-major, W MN-major) and wgrad (both MN-major, fp32 accumulate, split-K)
    dX = torch.empty(M, K, device='cuda', dtype=torch.bfloat16)
    ms = t(lambda: k.linear_dgrad(C, B, dX, M, N, K)); rows.append(("  dgrad", M, K, N, ms, 2.0 * M * N * K / ms / 1e9))
    dW = torch.zeros(N, K, device='cuda')
    ms = t(lambda: k.linear_wgrad(C, A, dW, M, N, K)); rows.append(("  wgrad", N, K, M, ms, 2.0 * M * N * K / ms / 1e9))
for r in rows: print("%-36s M=%6d N=%5d K=%6d  %7.3f ms  %7.1f TFLOP/s" % r)
if ONLY:
    sys.exit(0)
# attention batched (B=128 clips, H=4, T=256, dh=128)
Bc, H, T, D = 128, 4, 256, 512; dh = D // H
qu = torch.randn(Bc * T, D, device='cuda').bfloat16(); qkv = torch.randn(Bc * T, 3 * D, device='cuda').bfloat16(); content = torch.empty(Bc, H, T, T, device='cuda', dtype=torch.bfloat16)
ms = t(lambda: k.gemm(qu, qkv, content, T, T, dh, (D, 1), (3 * D, 1), T, b_off=D, batch=(Bc, H), sAb=(T * D, dh), sBb=(T * 3 * D, dh), sCb=(H * T * T, T * T)))
print("attention content scores (batched)     %7.3f ms  %7.1f TFLOP/s" % (ms, 2.0 * Bc * H * T * T * dh / ms / 1e9))
