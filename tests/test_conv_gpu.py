"""GPU: the CNN-stem kernels.  CUDA-core 3x3 conv / wgrad vs torch (CPU fp32 conv2d); tcgen05 versions vs a plain PyTorch fp32
conv2d / its autograd weight gradient (TF32 off) on the same bf16 operands."""
import pytest
import torch
import torch.nn.functional as F

from sarssl_b200.kernels import KernelSet

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


torch.backends.cuda.matmul.allow_tf32 = False          # fp32 torch references must not run on TF32 tensor cores
torch.backends.cudnn.allow_tf32 = False


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def unpack(wp):      # [o][tap = kw*3 + kh][ci] -> torch (o, ci, kh[bin], kw[frame])
    return wp.reshape(64, 3, 3, 64).permute(0, 3, 2, 1).contiguous()


def conv_ref_fp32(x, wp):
    """x [B][H = frame][W = bin][64] (any float dtype), packed weights -> fp32 conv2d in our layout."""
    return F.conv2d(x.float().permute(0, 3, 2, 1), unpack(wp.float()), padding=1).permute(0, 3, 2, 1).contiguous()


def pack(w):        # reference (o, ci, kh[bin], kw[frame]) -> [o][tap = kw*3 + kh][ci]   (image is [frame][bin])
    return w.permute(0, 3, 2, 1).reshape(64, 9, 64).contiguous()


@pytest.mark.parametrize("B,H,W", [(2, 16, 256), (1, 7, 40), (3, 5, 130)])
def test_simt_conv3x3_and_wgrad_fp32_vs_torch(B, H, W):
    k = KernelSet(DEV, torch.float32)
    g = torch.Generator().manual_seed(B + H + W)
    x = torch.randn(B, H, W, 64, generator=g)                    # our layout [b][frame][bin][c]
    w = torch.randn(64, 64, 3, 3, generator=g) / 24
    scale, shift = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1
    stats = torch.cat([torch.zeros(128), scale, shift])
    z = torch.relu(x * scale + shift)
    ref = F.conv2d(z.permute(0, 3, 2, 1), w, padding=1).permute(0, 3, 2, 1).contiguous()      # torch image is (c, bin, frame)
    out = k.empty(B, H, W, 64)
    k.conv3x3(x.to(DEV), stats.to(DEV), pack(w).to(DEV), out, B, H, W)
    assert rel(out.cpu(), ref) < 1e-5
    dy = torch.randn(B, H, W, 64, generator=g)
    zz = z.permute(0, 3, 2, 1).clone().requires_grad_(True)
    ww = w.clone().requires_grad_(True)
    F.conv2d(zz, ww, padding=1).backward(dy.permute(0, 3, 2, 1))
    dwp = torch.empty(64, 9, 64, device=DEV)
    k.conv3x3_wgrad(dy.to(DEV), x.to(DEV), stats.to(DEV), dwp, B, H, W)
    assert rel(dwp.cpu(), pack(ww.grad)) < 1e-4
    # data gradient = the same kernel with the mirrored pack [ci][mirrored tap][o]
    wb = w.flip(2, 3).permute(1, 3, 2, 0).reshape(64, 9, 64).contiguous()
    dz = k.empty(B, H, W, 64)
    k.conv3x3(dy.to(DEV), None, wb.to(DEV), dz, B, H, W)
    assert rel(dz.cpu(), zz.grad.permute(0, 3, 2, 1)) < 1e-5


@pytest.mark.parametrize("B,H,W", [(1, 1, 128), (2, 16, 256), (1, 40, 256), (3, 33, 200), (2, 5, 64), (8, 256, 256)])
def test_tc_conv3x3_vs_torch_fp32(B, H, W):
    k = KernelSet(DEV, torch.bfloat16)
    g = torch.Generator(device=DEV).manual_seed(B * 7 + H + W)
    x = torch.randn(B, H, W, 64, device=DEV, generator=g).bfloat16()
    wp = (torch.randn(64, 9, 64, device=DEV, generator=g) / 24).bfloat16()
    out = torch.full((B, H, W, 64), 7.0, device=DEV, dtype=torch.bfloat16)
    ref = conv_ref_fp32(x, wp)
    gamma, beta = torch.rand(64, device=DEV) + 0.5, torch.randn(64, device=DEV)
    rm, rv, nbt = torch.zeros(64, device=DEV), torch.ones(64, device=DEV), torch.zeros((), dtype=torch.int64, device=DEV)
    stats = k.conv3x3_tc(x, wp, out, B, H, W, bn=(gamma, beta, rm, rv, nbt))
    torch.cuda.synchronize()
    assert rel(out.float(), ref) < 3e-3, rel(out.float(), ref)     # one bf16 rounding of the fp32 sums
    # fused BatchNorm statistics of the (bf16-rounded) conv output
    o = out.float().reshape(-1, 64)
    mean, var = o.mean(0), o.var(0, unbiased=False)
    assert torch.allclose(stats[:64], mean, rtol=1e-3, atol=1e-4) and torch.allclose(stats[64:128], torch.rsqrt(var + 1e-5), rtol=2e-3)
    assert int(nbt) == 1 and torch.allclose(rm, 0.1 * mean, rtol=1e-3, atol=1e-4)
    # without statistics (the data-gradient use) the epilogue stores straight from registers instead of staging a TMA store: same bits
    out2 = torch.full((B, H, W, 64), -3.0, device=DEV, dtype=torch.bfloat16)
    k.conv3x3_tc(x, wp, out2, B, H, W)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)


@pytest.mark.parametrize("B,H,W", [(1, 1, 64), (2, 16, 256), (1, 40, 256), (3, 33, 200), (2, 70, 64), (8, 256, 256)])
def test_tc_conv3x3_wgrad_vs_torch_fp32(B, H, W):
    k = KernelSet(DEV, torch.bfloat16)
    g = torch.Generator(device=DEV).manual_seed(B * 5 + H + W)
    z = torch.randn(B, H, W, 64, device=DEV, generator=g).bfloat16()
    dy = (torch.randn(B, H, W, 64, device=DEV, generator=g) / (B * H * W) ** 0.5).bfloat16()
    out = torch.empty(64, 9, 64, device=DEV)
    k.conv3x3_wgrad_tc(dy, z, out, B, H, W)
    torch.cuda.synchronize()
    w = torch.zeros(64, 64, 3, 3, device=DEV, requires_grad=True)
    F.conv2d(z.float().permute(0, 3, 2, 1), w, padding=1).backward(dy.float().permute(0, 3, 2, 1))
    assert rel(out, pack(w.grad)) < 1e-4, rel(out, pack(w.grad))


@pytest.mark.parametrize("B,H,W", [(1, 1, 64), (2, 16, 256), (1, 40, 256), (3, 33, 200), (2, 70, 64), (1, 5, 130), (8, 256, 256)])
def test_tc_conv3x3_fused_input_batchnorm_relu_is_bit_identical_to_the_materialised_route(B, H, W):
    """in_stats: the operand relu(scale * y + shift) is produced from the landed TMA boxes in shared memory (transform warps).  Forward conv,
    its fused output statistics and the weight gradient must equal - bit for bit - what the same kernels give on a materialised z (the
    arithmetic and rounding of bn_act_fwd_kernel), including the zero padding at the image borders (out-of-image pixels stay zero, they do
    NOT become relu(shift)) and ragged widths."""
    k = KernelSet(DEV, torch.bfloat16)
    g = torch.Generator(device=DEV).manual_seed(B * 11 + H + W)
    y = torch.randn(B, H, W, 64, device=DEV, generator=g).bfloat16()
    scale, shift = torch.rand(64, device=DEV, generator=g) + 0.5, torch.randn(64, device=DEV, generator=g) * 0.5 + 0.3     # shift > 0 mostly: padding bugs show
    stats = torch.cat([torch.zeros(128, device=DEV), scale, shift])
    z = torch.empty_like(y)
    k.bn_act_fwd(y, stats, 1, z, B * H * W, 64)
    wp = (torch.randn(64, 9, 64, device=DEV, generator=g) / 24).bfloat16()
    bn = lambda: (torch.rand(64, device=DEV) + 0.5, torch.randn(64, device=DEV), torch.zeros(64, device=DEV), torch.ones(64, device=DEV),
                  torch.zeros((), dtype=torch.int64, device=DEV))
    out_a, out_b = torch.empty_like(y), torch.empty_like(y)
    st_a = k.conv3x3_tc(z, wp, out_a, B, H, W, bn=bn())
    st_b = k.conv3x3_tc(y, wp, out_b, B, H, W, bn=bn(), in_stats=stats)
    torch.cuda.synchronize()
    assert torch.equal(out_a, out_b) and torch.equal(st_a[:128], st_b[:128])
    assert rel(out_b.float(), conv_ref_fp32(z, wp)) < 3e-3
    dy = (torch.randn(B, H, W, 64, device=DEV, generator=g) / (B * H * W) ** 0.5).bfloat16()
    dw_a, dw_b = torch.empty(64, 9, 64, device=DEV), torch.empty(64, 9, 64, device=DEV)
    k.conv3x3_wgrad_tc(dy, z, dw_a, B, H, W)
    k.conv3x3_wgrad_tc(dy, y, dw_b, B, H, W, in_stats=stats)
    torch.cuda.synchronize()
    assert torch.equal(dw_a, dw_b)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,W,mode", [(2, 16, 256, 0), (3, 5, 130, 0), (1, 3, 37, 0), (2, 16, 256, 1), (2, 16, 256, 2), (2, 7, 64, 3), (16, 64, 256, 1)])
def test_stem_pointwise_wgrad_vs_torch(dtype, B, H, W, mode):
    """dW[o][c] = sum_p f(wide[p][o]) * narrow[p][c] (the weight gradient of both 1x1 convs): every narrow-operand mode, with and without
    the fused BatchNorm+ReLU on the wide operand, ragged pixel counts.  The bf16 version runs on mma.sync with a split (value + residual)
    narrow operand, so fp32 patches are not rounded."""
    k = KernelSet(DEV, dtype)
    g = torch.Generator().manual_seed(B * 1000 + H * 10 + mode)
    P = B * H * W
    wide = torch.randn(P, 64, generator=g).to(dtype)
    scale, shift = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.3
    stats = torch.cat([torch.zeros(128), scale, shift])
    flag = ch = None
    if mode == 0:
        narrow = torch.randn(P, 4, generator=g).to(dtype)
        eff = narrow.float()
    else:
        narrow = torch.randn(P, 4, generator=g)                    # fp32 patches (re0, re1, im0, im1)
        eff = narrow.clone().view(B, H, W, 4)
        if mode != 3:
            flag = (torch.rand(B * H, generator=g) < 0.5).to(torch.uint8)
            ch = torch.randint(0, 2, (B,), generator=g, dtype=torch.int32)
            pm = flag.view(B, H).bool()
            for b in range(B):
                mc = int(ch[b])
                for h in range(H):
                    if mode == 1:
                        keep = 1 - mc if pm[b, h] else mc
                        eff[b, h, :, [0, 2] if keep else [1, 3]] = 0.0
                    elif pm[b, h]:
                        eff[b, h] = 0.0
        eff = eff.reshape(P, 4)
    dev = lambda t: None if t is None else t.to(DEV)
    for st in (None, stats):
        z = wide.float() if st is None else torch.relu(wide.float() * scale + shift)
        if st is not None and dtype == torch.bfloat16:
            z = z.to(dtype).float()                                 # the kernel rounds the activation like the materialised tensor
        ref = z.double().t() @ eff.double()
        dw = torch.zeros(64, 4, device=DEV)
        k.stem_pw_wgrad(dev(wide), dev(st), dev(narrow), mode, dev(flag), dev(ch), dw, False, P, W, H)
        assert rel(dw.cpu(), ref) < (2e-5 if dtype == torch.float32 else 1e-4), (st is not None)
        k.stem_pw_wgrad(dev(wide), dev(st), dev(narrow), mode, dev(flag), dev(ch), dw, True, P, W, H)       # accumulate
        assert rel(dw.cpu(), 2 * ref) < (2e-5 if dtype == torch.float32 else 1e-4)


def _bn_stats64(y, gamma, beta, eps=1e-5):
    mean, var = y.double().mean(0), y.double().var(0, unbiased=False)
    rstd = 1.0 / torch.sqrt(var + eps)
    scale = gamma.double() * rstd
    return torch.cat([mean, rstd, scale, beta.double() - mean * scale]).float()


@pytest.mark.parametrize("B,H,W,mode", [(2, 16, 256, 1), (2, 16, 256, 2), (3, 5, 130, 3), (1, 3, 37, 3), (8, 64, 256, 1)])
def test_stem_head_backward_fused_vs_autograd(B, H, W, mode):
    """conv1x1(4->64) + BatchNorm(batch statistics) + ReLU backward from ONE pass over (dz, z): dW, dgamma, dbeta vs torch autograd."""
    k = KernelSet(DEV, torch.bfloat16)
    g = torch.Generator().manual_seed(B * 100 + H + mode)
    P = B * H * W
    x = torch.randn(P, 4, generator=g) + 0.3
    flag = ch = None
    eff = x.clone().view(B, H, W, 4)
    if mode != 3:
        flag = (torch.rand(B * H, generator=g) < 0.5).to(torch.uint8)
        ch = torch.randint(0, 2, (B,), generator=g, dtype=torch.int32)
        pm = flag.view(B, H).bool()
        for b in range(B):
            mc = int(ch[b])
            for h in range(H):
                if mode == 1:
                    keep = 1 - mc if pm[b, h] else mc
                    eff[b, h, :, [0, 2] if keep else [1, 3]] = 0.0
                elif pm[b, h]:
                    eff[b, h] = 0.0
    eff = eff.reshape(P, 4)
    w = (torch.randn(64, 4, generator=g) * 0.5).requires_grad_(True)
    gamma = (torch.rand(64, generator=g) + 0.5).requires_grad_(True)
    beta = (torch.randn(64, generator=g) * 0.2).requires_grad_(True)
    dz = torch.randn(P, 64, generator=g).bfloat16()
    y = eff @ w.t()
    z = torch.relu(F.batch_norm(y, None, None, gamma, beta, training=True, eps=1e-5))
    z.backward(dz.float())
    stats = _bn_stats64(y.detach(), gamma.detach(), beta.detach())
    dw, dg, db = torch.zeros(64, 4, device=DEV), torch.zeros(64, device=DEV), torch.zeros(64, device=DEV)
    dev = lambda t: None if t is None else t.to(DEV)
    for rep in (1, 2):                                                       # the second call accumulates
        k.stem_head_bwd(dev(dz), dev(z.detach().bfloat16()), dev(stats), dev(x), mode, dev(flag), dev(ch), dev(w.detach()), dg, db, dw, P, W, H)
        assert rel(dw.cpu(), rep * w.grad) < 2e-3, rel(dw.cpu(), rep * w.grad)
        assert rel(dg.cpu(), rep * gamma.grad) < 2e-3 and rel(db.cpu(), rep * beta.grad) < 2e-3


@pytest.mark.parametrize("B,H,W", [(2, 16, 256), (3, 5, 130), (1, 3, 37), (8, 64, 256)])
def test_stem_tail_backward_fused_vs_autograd(B, H, W):
    """BatchNorm + ReLU + conv1x1(64->4) backward without materialising the P x 64 ReLU gradient: dy, dW, dgamma, dbeta vs torch autograd."""
    k = KernelSet(DEV, torch.bfloat16)
    g = torch.Generator().manual_seed(B * 100 + H)
    P = B * H * W
    y16 = (torch.randn(P, 64, generator=g) * 1.5 + 0.2).bfloat16()
    y = y16.float().requires_grad_(True)
    w = (torch.randn(4, 64, generator=g) * 0.2).requires_grad_(True)
    gamma = (torch.rand(64, generator=g) + 0.5).requires_grad_(True)
    beta = (torch.randn(64, generator=g) * 0.2).requires_grad_(True)
    dq = torch.randn(P, 4, generator=g).bfloat16()
    q = torch.relu(F.batch_norm(y, None, None, gamma, beta, training=True, eps=1e-5)) @ w.t()
    q.backward(dq.float())
    stats = _bn_stats64(y.detach(), gamma.detach(), beta.detach())
    dw, dg, db = torch.zeros(4, 64, device=DEV), torch.zeros(64, device=DEV), torch.zeros(64, device=DEV)
    dy = torch.empty(P, 64, device=DEV, dtype=torch.bfloat16)
    for rep in (1, 2):
        k.stem_tail_bwd(y16.to(DEV), stats.to(DEV), dq.to(DEV), w.detach().to(DEV), dg, db, dw, dy, P)
        assert rel(dy.float().cpu(), y.grad) < 6e-3, rel(dy.float().cpu(), y.grad)
        assert rel(dw.cpu(), rep * w.grad) < 1e-3, rel(dw.cpu(), rep * w.grad)
        assert rel(dg.cpu(), rep * gamma.grad) < 1e-3 and rel(db.cpu(), rep * beta.grad) < 1e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,W,mode", [(2, 16, 256, 1), (2, 16, 256, 2), (3, 5, 130, 3), (1, 3, 37, 3)])
def test_stem_first_layer_stats_from_input_moments(dtype, B, H, W, mode):
    """BatchNorm statistics of conv1x1(x) from the moments of x == statistics of the materialised conv output (fp32 tensor), including the
    running-stat update; conv+BN+ReLU in one pass == the three separate kernels."""
    k, k32 = KernelSet(DEV, dtype), KernelSet(DEV, torch.float32)
    g = torch.Generator().manual_seed(B * 100 + H + mode)
    P = B * H * W
    x = (torch.randn(P, 4, generator=g) + 0.4).to(DEV)
    flag = ch = None
    if mode != 3:
        flag = (torch.rand(B * H, generator=g) < 0.5).to(torch.uint8).to(DEV)
        ch = torch.randint(0, 2, (B,), generator=g, dtype=torch.int32).to(DEV)
    w = (torch.randn(64, 4, generator=g) * 0.5).to(DEV)
    gamma, beta = (torch.rand(64, generator=g) + 0.5).to(DEV), (torch.randn(64, generator=g) * 0.2).to(DEV)
    bn_a = (gamma, beta, torch.zeros(64, device=DEV), torch.ones(64, device=DEV), torch.zeros(1, dtype=torch.long, device=DEV))
    bn_b = (gamma, beta, torch.zeros(64, device=DEV), torch.ones(64, device=DEV), torch.zeros(1, dtype=torch.long, device=DEV))
    y = torch.empty(P, 64, device=DEV)
    k32.stem_expand(x, mode, flag, ch, w, y, P, W, H)
    ref = k32.bn_stats(y, P, 64, *bn_a, True)
    got = k.stem_input_bn_stats(x, mode, flag, ch, w, P, W, H, bn_b)
    assert rel(got[:64].cpu(), ref[:64].cpu()) < 1e-5 and rel(got[64:].cpu(), ref[64:].cpu()) < 1e-4
    assert rel(bn_b[2].cpu(), bn_a[2].cpu()) < 1e-5 and rel(bn_b[3].cpu(), bn_a[3].cpu()) < 1e-5 and int(bn_b[4]) == 1
    z_ref = torch.relu(y * ref[128:192] + ref[192:256])
    z = torch.empty(P, 64, device=DEV, dtype=dtype)
    k.stem_expand_bn_relu(x, mode, flag, ch, w, ref, z, P, W, H)
    assert rel(z.float().cpu(), z_ref.cpu()) < (1e-5 if dtype == torch.float32 else 4e-3)
