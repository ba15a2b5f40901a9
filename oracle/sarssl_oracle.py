"""TEST INFRASTRUCTURE ONLY - CPU restatement (plain PyTorch fp32) of the SAR-SSL pre-training hot path.

This file is the parity *checker* for the CUDA path.  Only `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product package
`sarssl_b200` never does (and fails loudly without its CUDA library).

It is a functional restatement (plain functions over a flat `state_dict`, autograd for the backward),
written from the behaviour of the reference, not a copy of its nn.Module classes.  Each function cites
the reference lines it follows (paths relative to /root/reference/code).

Pinning: the reference ships no tests / golden vectors, so the oracle is pinned by
running the reference itself in the build container: `oracle/make_golden.py` imports the real reference
(via `oracle/ref_shim.py`), feeds it seeded synthetic inputs + a seeded synthetic state_dict and stores
its outputs under tests/golden/; `tests/test_oracle.py` checks this file against those fixtures, and
`tests/test_oracle_vs_reference.py` re-checks live whenever /root/reference is present.
"""
import math
import random as _pyrandom

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# A1  STFT                                                   common/utils_module.py:49-72
# ----------------------------------------------------------------------------------------------

def hann_periodic(n, dtype=torch.float32):
    """torch.hann_window(n) default (periodic=True): 0.5 - 0.5 cos(2 pi k / n)."""
    k = torch.arange(n, dtype=torch.float64)
    return (0.5 - 0.5 * torch.cos(2.0 * math.pi * k / n)).to(dtype)


def num_frames(nsample, win_len, hop):
    # utils_module.py:59  floor((nsample - win_len) / hop + 1)
    return (nsample - win_len) // hop + 1


def stft(signal, win_len=512, win_shift_ratio=0.5, nfft=512):
    """signal (nb, nsample, nch) f32 -> (nb, nfft/2+1, nt, nch) complex64.

    Same numbers as torch.stft(center=False, onesided, normalized=False, periodic Hann): frames of
    win_len samples every hop samples, times the window, rFFT (utils_module.py:65-71)."""
    assert win_len == nfft, "the hot path uses win_len == nfft (run_pretrain.py:68-69)"
    hop = int(win_len * win_shift_ratio)
    nb, nsample, nch = signal.shape
    nt = num_frames(nsample, win_len, hop)
    w = hann_periodic(win_len, signal.dtype).to(signal.device)
    frames = signal.permute(0, 2, 1).unfold(-1, win_len, hop)[:, :, :nt]          # (nb, nch, nt, win)
    spec = torch.fft.rfft(frames * w, n=nfft, dim=-1)                             # (nb, nch, nt, nf)
    return spec.permute(0, 3, 2, 1).contiguous()                                  # (nb, nf, nt, nch)


def istft(spec, win_len=512, win_shift_ratio=0.5, nfft=512):
    """(nb, nf, nt, nch) c64 -> (nb, (nt+1)*hop, nch) f32.  utils_module.py:91-113 (inv=False branch):
    torch.istft with no window == rectangular synthesis window, overlap-add divided by the window
    envelope (sum of squared rectangular windows = number of overlapping frames)."""
    hop = int(win_len * win_shift_ratio)
    nb, nf, nt, nch = spec.shape
    frames = torch.fft.irfft(spec.permute(0, 3, 2, 1), n=nfft, dim=-1)[..., :win_len]   # (nb, nch, nt, win)
    nsample = (nt - 1) * hop + win_len
    out = torch.zeros(nb, nch, nsample, dtype=frames.dtype, device=spec.device)
    env = torch.zeros(nsample, dtype=frames.dtype, device=spec.device)
    for t in range(nt):
        out[:, :, t * hop:t * hop + win_len] += frames[:, :, t]
        env[t * hop:t * hop + win_len] += 1.0
    out = out / env
    return out.permute(0, 2, 1)[:, :(nt + 1) * hop].contiguous()


# ----------------------------------------------------------------------------------------------
# A2  data_preprocess                                        learner.py:525-572, utils_module.py:124-148
# ----------------------------------------------------------------------------------------------

def add_ch_to_batch(x, ch_mode="M"):
    """(nb, nch, ...) -> microphone pairs stacked on the batch axis (utils_module.py:124-148)."""
    nb, nch = x.shape[:2]
    if ch_mode == "M":          # (ref mic 0, mic j) for j = 1..nch-1
        pairs = [(0, j) for j in range(1, nch)]
    elif ch_mode == "MM":       # all pairs i<j, ordered by i then j
        pairs = [(i, j) for i in range(nch - 1) for j in range(i + 1, nch)]
    else:
        return x.clone()
    first = torch.tensor([p[0] for p in pairs])
    second = torch.tensor([p[1] for p in pairs])
    out = torch.stack([x[:, first], x[:, second]], dim=2)        # (nb, npair, 2, ...)
    return out.reshape((nb * len(pairs), 2) + tuple(x.shape[2:])).contiguous()


def preprocess(signal, win_len=512, win_shift_ratio=0.5, nfft=512, eps=1e-6, ch_mode="M", fre_used_ratio=1):
    """(nb, nsample, nch) f32 -> (nb*(nch-1), 2, nfft/2, nt, 2) f32.

    STFT, divide every channel by (mean over all 257 x nt bins of |X_ch0| + eps) (learner.py:539-542),
    pair microphones, split re/im, keep bins 1..nfft/2 (fre_used_ratio == 1, learner.py:515-516,551)."""
    X = stft(signal, win_len, win_shift_ratio, nfft).permute(0, 3, 1, 2)          # (nb, nch, nf, nt)
    mean_mag = X[:, 0].abs().reshape(X.shape[0], -1).mean(dim=1)
    X = X / (mean_mag + eps)[:, None, None, None]
    X = add_ch_to_batch(X, ch_mode)
    bins = slice(1, nfft // 2 + 1) if fre_used_ratio == 1 else slice(0, int(nfft / 2 * fre_used_ratio))     # learner.py:514-517
    return torch.view_as_real(X)[:, :, bins].contiguous()


# ----------------------------------------------------------------------------------------------
# A4  mask indices (host RNG)                                utils_module.py:255-272,305-308
# ----------------------------------------------------------------------------------------------

def draw_masks(nbatch, npatch, nmasked, nmic=2, rng=None):
    """Per item, in order: random.sample(range(npatch), nmasked) then random.randint(0, nmic-1).
    `rng` is a random.Random (or the global `random` module, which is what the reference uses).
    Returns (mask_patch_idx (nb, nmasked) int64, mask_ch_idx (nb, 1) int64)."""
    rng = rng or _pyrandom
    pidx = torch.empty(nbatch, nmasked, dtype=torch.int64)
    cidx = torch.empty(nbatch, 1, dtype=torch.int64)
    for b in range(nbatch):
        pidx[b] = torch.tensor(rng.sample(range(0, npatch), nmasked))
        cidx[b, 0] = rng.randint(0, nmic - 1)
    return pidx, cidx


def dense_masks(pidx, cidx, npatch, dpatch, nmic=2):
    """(mask_dense, mask_patch_dense, mask_ch_dense), each (nb, npatch, dpatch, nmic) f32; 0 = masked."""
    nb = pidx.shape[0]
    mp = torch.ones(nb, npatch, device=pidx.device)
    mp.scatter_(1, pidx, 0.0)
    mc = torch.ones(nb, nmic, device=pidx.device)
    mc.scatter_(1, cidx, 0.0)
    mask_patch = mp[:, :, None, None].expand(nb, npatch, dpatch, nmic)
    mask_ch = mc[:, None, None, :].expand(nb, npatch, dpatch, nmic)
    mask = 1.0 - (1.0 - mask_patch) * (1.0 - mask_ch)
    return mask.contiguous(), mask_patch.contiguous(), mask_ch.contiguous()


# ----------------------------------------------------------------------------------------------
# A6..A11  model pieces
# ----------------------------------------------------------------------------------------------

class BNState:
    """How BatchNorm layers behave: training=True -> batch statistics + running-stat update in `sd`."""

    def __init__(self, training=True, momentum=0.1, eps=1e-5):
        self.training, self.momentum, self.eps = training, momentum, eps


def _bn(x, sd, key, bn):
    out = F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"], sd[key + ".weight"],
                       sd[key + ".bias"], bn.training, bn.momentum, bn.eps)
    if bn.training:
        sd[key + ".num_batches_tracked"] += 1
    return out


def cnn_stem(img, sd, pre, bn, taps=None):
    """model.py:50-64.  img (nb, 4, nf, nt) -> (nb, nt, D).  All convs bias-free."""
    y = img
    for i, pad in ((0, 0), (3, 1), (6, 1), (9, 0)):
        y = F.conv2d(y, sd[f"{pre}.{i}.weight"], None, padding=pad)
        y = F.relu(_bn(y, sd, f"{pre}.{i + 1}", bn))
        if taps is not None:
            taps[f"{pre}.{i + 2}"] = y
    w = sd[f"{pre}.12.weight"]                                   # (D, 4, nf, 1): collapses the frequency axis
    y = F.conv2d(y, w, None, stride=(w.shape[2], 1))             # (nb, D, 1, nt)
    return y[:, :, 0].transpose(1, 2)


def _ln(x, sd, key):
    return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"], sd[key + ".bias"], 1e-5)


def _lin(x, sd, key, bias=True):
    return F.linear(x, sd[key + ".weight"], sd[key + ".bias"] if bias else None)


def feed_forward(x, sd, pre, drop):
    """conformer/feed_forward.py:39-54 (LN, Linear, Swish, Dropout, Linear, Dropout)."""
    h = _ln(x, sd, pre + ".0")
    h = _lin(h, sd, pre + ".1.linear")
    h = drop(h * torch.sigmoid(h))
    return drop(_lin(h, sd, pre + ".4.linear"))


def relative_shift(pos):
    """conformer/attention.py:105-113 as an explicit index map:
        out[i, j] = pos[i, T-1-i+j]   for j <= i
                  = 0                 for j == i+1
                  = pos[i+1, j-i-2]   for j >  i+1
    pos (..., T, T).  (The pad/view trick of the reference produces exactly this.)"""
    T = pos.shape[-1]
    i = torch.arange(T, device=pos.device).view(T, 1)
    j = torch.arange(T, device=pos.device).view(1, T)
    low = j <= i
    up = j > i + 1
    row = torch.where(up, i + 1, i).clamp(max=T - 1).expand(T, T)
    col = torch.where(low, T - 1 - i + j, j - i - 2).clamp(min=0, max=T - 1)
    g = pos[..., row, col]
    return g * (low | up).to(pos.dtype)


def rel_attention(x, sd, pre, nhead, drop):
    """conformer/attention.py:72-103,143-151.  x (nb, T, D)."""
    nb, T, D = x.shape
    dh = D // nhead
    h = _ln(x, sd, pre + ".layer_norm")
    a = pre + ".attention"
    q = _lin(h, sd, a + ".query_proj.linear").view(nb, T, nhead, dh)
    k = _lin(h, sd, a + ".key_proj.linear").view(nb, T, nhead, dh).permute(0, 2, 1, 3)
    v = _lin(h, sd, a + ".value_proj.linear").view(nb, T, nhead, dh).permute(0, 2, 1, 3)
    pe = sd[pre + ".positional_encoding.pe"][0, :T]                                    # (T, D)
    p = F.linear(pe, sd[a + ".pos_proj.linear.weight"]).view(T, nhead, dh).permute(1, 0, 2)  # (H, T, dh)
    content = torch.matmul((q + sd[a + ".u_bias"]).transpose(1, 2), k.transpose(2, 3))
    pos = torch.matmul((q + sd[a + ".v_bias"]).transpose(1, 2), p.transpose(1, 2))      # (nb, H, T, T)
    score = (content + relative_shift(pos)) / math.sqrt(D)        # sqrt(d_model), not d_head (attention.py:57,91)
    attn = drop(F.softmax(score, -1))
    ctx = torch.matmul(attn, v).transpose(1, 2).reshape(nb, T, D)
    return drop(_lin(ctx, sd, a + ".out_proj.linear"))


def conv_module(x, sd, pre, bn, drop):
    """conformer/convolution.py:136-149: LN, pointwise D->2D, GLU, depthwise k=31, BN1d, Swish, pointwise, Dropout."""
    D = x.shape[-1]
    h = _ln(x, sd, pre + ".0").transpose(1, 2)                                       # (nb, D, T)
    h = F.conv1d(h, sd[pre + ".2.conv.weight"], sd[pre + ".2.conv.bias"])
    h = h[:, :D] * torch.sigmoid(h[:, D:])
    w = sd[pre + ".4.conv.weight"]
    h = F.conv1d(h, w, None, padding=(w.shape[-1] - 1) // 2, groups=D)
    h = _bn(h, sd, pre + ".5", bn)
    h = h * torch.sigmoid(h)
    h = F.conv1d(h, sd[pre + ".7.conv.weight"], sd[pre + ".7.conv.bias"])
    return drop(h).transpose(1, 2)


def conformer_block(x, sd, pre, nhead, bn, drop, taps=None):
    """common/Conformer.py:59-88 with ResidualConnectionModule (modules.py:26-33)."""
    s = pre + ".sequential"
    x = x + 0.5 * feed_forward(x, sd, s + ".0.module.sequential", drop)
    if taps is not None: taps[s + ".0"] = x
    x = x + rel_attention(x, sd, s + ".1.module", nhead, drop)
    if taps is not None: taps[s + ".1"] = x
    x = x + conv_module(x, sd, s + ".2.module.sequential", bn, drop)
    if taps is not None: taps[s + ".2"] = x
    x = x + 0.5 * feed_forward(x, sd, s + ".3.module.sequential", drop)
    if taps is not None: taps[s + ".3"] = x
    x = _ln(x, sd, s + ".4")
    if taps is not None: taps[s + ".4"] = x
    return x


def embed_encoder(tokens, sd, pre, nlayer, nhead, bn, drop, taps=None):
    """model.py:194-222.  tokens (nb, nt, nf*4) in patch layout [t, (f, r, m)] -> (nb, nt, D)."""
    nb, nt, dim = tokens.shape
    nf = dim // 4
    img = tokens.view(nb, nt, nf, 4).permute(0, 3, 2, 1)          # (nb, c = r*2+m, nf, nt)
    e = cnn_stem(img, sd, pre + ".patch_embed", bn, taps)
    if taps is not None: taps[pre + ".patch_embed"] = e
    for l in range(nlayer):
        e = conformer_block(e, sd, f"{pre}.embed.layers.{l}", nhead, bn, drop, taps)
    return e


SPEC_LAYERS, SPAT_LAYERS, NHEAD = 1, 3, 4        # model.py:37-43,94


def pretrain_forward(x, sd, pidx, cidx, training=True, dropout_p=0.0, taps=None):
    """model.py:519-601 (pretrain branch, in_ver == 'separate').

    x (nb, 2, nf, nt, 2) f32; pidx (nb, nmasked) int64; cidx (nb, 1) int64.
    Returns (loss, diff, vis) with vis = {'mask', 'pred', 'tar'} exactly as the reference lays them out.
    BN running statistics inside `sd` are updated in place when training."""
    nb, nmic, nf, nt, _ = x.shape
    bn = BNState(training)
    drop = (lambda t: F.dropout(t, dropout_p, True)) if (training and dropout_p > 0) else (lambda t: t)
    vec = x.permute(0, 3, 2, 4, 1)                                # (nb, nt, nf, reim, mic)   model.py:524-525
    mask, mask_p, mask_c = dense_masks(pidx, cidx, nt, nf, nmic)
    mp, mc = mask_p[:, :, :, None, :], mask_c[:, :, :, None, :]
    spec_in = vec * (1 - mp) * mc + vec * mp * (1 - mc)           # model.py:541
    spat_in = vec * mp                                            # model.py:563
    e_spec = embed_encoder(spec_in.reshape(nb, nt, -1), sd, "spec_encoder", SPEC_LAYERS, NHEAD, bn, drop, taps)
    e_spat = embed_encoder(spat_in.reshape(nb, nt, -1), sd, "spat_encoder", SPAT_LAYERS, NHEAD, bn, drop, taps)
    e = torch.cat([e_spec, e_spat], dim=2)
    h = F.relu(_lin(e, sd, "decoder.proj.0"))
    pred = _lin(h, sd, "decoder.proj.2").view(nb, nt, nf, 2, nmic)                    # model.py:582,589
    if taps is not None: taps["decoder"] = pred
    loss, diff = masked_loss(pred, vec, pidx, cidx)
    vis = {"mask": mask.permute(0, 2, 1, 3).contiguous(),                             # (nb, nf, nt, nmic)
           "pred": pred.detach().permute(0, 2, 1, 3, 4).contiguous(),                 # (nb, nf, nt, 2, nmic)
           "tar": vec.detach().permute(0, 2, 1, 3, 4).contiguous()}
    return loss, diff, vis


def frozen_encoder_forward(x, sd, pidx, cidx, training=True, dropout_p=0.0):
    """model.py:603-666 (pretrain=False, pretrain_frozen_encoder=True): the spectral encoder only sees the un-masked channel of the
    masked frames (model.py:622), the spatial encoder the un-masked frames; `spec_spat_decoder` predicts the patches and
    gen_loss_spec(tar_maskch=True) (model.py:749-774) is the masked-channel MSE of gen_loss.  Returns (loss, loss * 0, vis)."""
    nb, nmic, nf, nt, _ = x.shape
    bn = BNState(training)
    drop = (lambda t: F.dropout(t, dropout_p, True)) if (training and dropout_p > 0) else (lambda t: t)
    vec = x.permute(0, 3, 2, 4, 1)
    mask, mask_p, mask_c = dense_masks(pidx, cidx, nt, nf, nmic)
    mp, mc = mask_p[:, :, :, None, :], mask_c[:, :, :, None, :]
    spec_in = vec * (1 - mp) * mc                                 # model.py:622
    spat_in = vec * mp                                            # model.py:628
    e_spec = embed_encoder(spec_in.reshape(nb, nt, -1), sd, "spec_encoder", SPEC_LAYERS, NHEAD, bn, drop)
    e_spat = embed_encoder(spat_in.reshape(nb, nt, -1), sd, "spat_encoder", SPAT_LAYERS, NHEAD, bn, drop)
    e = torch.cat([e_spec, e_spat], dim=2)
    pred = _lin(F.relu(_lin(e, sd, "spec_spat_decoder.proj.0")), sd, "spec_spat_decoder.proj.2").view(nb, nt, nf, 2, nmic)   # model.py:653-654
    loss, _ = masked_loss(pred, vec, pidx, cidx)                  # gen_loss_spec, tar_maskch=True
    vis = {"mask": mask.permute(0, 2, 1, 3).contiguous(), "pred": pred.detach().permute(0, 2, 1, 3, 4).contiguous(),
           "tar": vec.detach().permute(0, 2, 1, 3, 4).contiguous()}
    return loss, loss * 0.0, vis


def downstream_forward(x, sd, embed_use="spec_spat", training=True, dropout_p=0.0):
    """model.py:667-719 (pretrain=False, head 'mlp', dlabel 1, token 'all'): both encoders on the un-masked input, concatenation (or
    one of them), mean over time, LayerNorm + Linear.  x (nb, 2, nf, nt, 2) -> (pred (nb, 1), pooled embedding (nb, dembed))."""
    nb, nmic, nf, nt, _ = x.shape
    bn = BNState(training)
    drop = (lambda t: F.dropout(t, dropout_p, True)) if (training and dropout_p > 0) else (lambda t: t)
    tokens = x.permute(0, 3, 2, 4, 1).reshape(nb, nt, -1)
    e_spec = embed_encoder(tokens, sd, "spec_encoder", SPEC_LAYERS, NHEAD, bn, drop)
    e_spat = embed_encoder(tokens, sd, "spat_encoder", SPAT_LAYERS, NHEAD, bn, drop)
    e = {"spec_spat": torch.cat([e_spec, e_spat], dim=2), "spec": e_spec + 0.0, "spat": e_spat + 0.0}[embed_use]
    pooled = e.mean(dim=1)
    if "joint_head.0.weight" in sd:                # dlabel > 1 (model.py:501-507,709-710)
        pred = _lin(F.relu(_lin(_ln(pooled, sd, "joint_head.0"), sd, "joint_head.1")), sd, "joint_head.3")
    else:
        pred = _lin(_ln(pooled, sd, "mlp_head.0"), sd, "mlp_head.1")
    return pred, pooled


def multich_forward(x, sd, nmic_pair, training=True, dropout_p=0.0):
    """SARSSL_MultiCH.forward (model.py:793-821): the spatial encoder's time-mean embedding of every microphone pair (the inner
    SARSSL(pretrain=False, downstream_head='', downstream_embed='spat') runs both encoders, model.py:676-678), the pairs of one item
    concatenated, LayerNorm + Linear + ReLU + Linear.  x (nb*nmic_pair, 2, nf, nt, 2) -> (pred (nb, factor), embed (nb, nmic_pair*256))."""
    inner = {k[len("model_sch."):]: v for k, v in sd.items() if k.startswith("model_sch.")}
    n, nmic, nf, nt, _ = x.shape
    bn = BNState(training)
    drop = (lambda t: F.dropout(t, dropout_p, True)) if (training and dropout_p > 0) else (lambda t: t)
    tokens = x.permute(0, 3, 2, 4, 1).reshape(n, nt, -1)
    embed_encoder(tokens, inner, "spec_encoder", SPEC_LAYERS, NHEAD, bn, drop)            # run for its BatchNorm side effects only, like the reference
    e = embed_encoder(tokens, inner, "spat_encoder", SPAT_LAYERS, NHEAD, bn, drop) + 0.0
    emb = e.mean(dim=1).reshape(-1, nmic_pair * e.shape[-1])
    h = F.relu(_lin(_ln(emb, sd, "head_mch.0"), sd, "head_mch.1"))
    return _lin(h, sd, "head_mch.3"), emb


def mcconformer_forward(x, sd, training=False):
    """MCConformer.forward (model.py:824-912, both encoders): un-masked input -> encoders -> decoder -> PatchRecover.
    x (nb, 2, nf, nt, 2) -> data_pred (nb, nf, nt, 2, 2)."""
    nb, nmic, nf, nt, _ = x.shape
    bn = BNState(training)
    tokens = x.permute(0, 3, 2, 4, 1).reshape(nb, nt, -1)
    e = torch.cat([embed_encoder(tokens, sd, "spec_encoder", SPEC_LAYERS, NHEAD, bn, lambda t: t),
                   embed_encoder(tokens, sd, "spat_encoder", SPAT_LAYERS, NHEAD, bn, lambda t: t)], dim=2)
    pred = _lin(F.relu(_lin(e, sd, "decoder.proj.0")), sd, "decoder.proj.2").view(nb, nt, nf, 2, nmic)
    return pred.permute(0, 2, 1, 3, 4)


def masked_loss(pred, vec, pidx, cidx):
    """model.py:585-592,721-747.  pred/vec (nb, nt, nf, 2, nmic).
    loss = mean over (item, masked frame, bin, re/im) of (pred - target)^2 on the masked channel,
    diff = same mean of (masked-channel target - other-channel target)^2."""
    nb, nt, nf, _, nmic = vec.shape
    sel = F.one_hot(cidx[:, 0], nmic).to(vec.dtype)[:, None, None, None, :]          # 1 on the masked channel
    tar = (vec * sel).sum(-1).detach()
    other = (vec * (1 - sel)).sum(-1).detach()
    prd = (pred * sel).sum(-1)
    gi = pidx[:, :, None, None].expand(-1, -1, nf, 2)
    prd, tar, other = prd.gather(1, gi), tar.gather(1, gi), other.gather(1, gi)
    return ((prd - tar) ** 2).mean(), ((tar - other) ** 2).mean()


# ----------------------------------------------------------------------------------------------
# synthetic weights shared by reference / oracle / CUDA path in the parity tests
# ----------------------------------------------------------------------------------------------

def positional_table(d_model, max_len=10000):
    """conformer/embedding.py:31-38."""
    pe = torch.zeros(max_len, d_model)
    pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(0)


def state_dict_spec(nf=256, pretrain=True, dembed_ds=768, frozen=False, head="mlp", prefix="", nmic_pair=0, factor=1, dlabel=1):
    """[(key, shape, kind)] for the 214 state_dict entries of SARSSL(pretrain=True).
    frozen=True: SARSSL(pretrain=False, pretrain_frozen_encoder=True) - three decoders instead of one (model.py:470-481; the spatial
    decoder is built with the SPECTRAL width, as the reference does).
    kind: 'w' weight matrix / filter, 'b' bias, 'g' norm gain, 'rm' running mean, 'rv' running var,
    'n' num_batches_tracked, 'pe' positional table, 'uv' u/v bias."""
    out = []

    def stem(pre, D):
        for i, shp in ((0, (64, 4, 1, 1)), (3, (64, 64, 3, 3)), (6, (64, 64, 3, 3)), (9, (4, 64, 1, 1))):
            out.append((f"{pre}.{i}.weight", shp, "w"))
            c = shp[0]
            out.extend([(f"{pre}.{i + 1}.weight", (c,), "g"), (f"{pre}.{i + 1}.bias", (c,), "b"),
                        (f"{pre}.{i + 1}.running_mean", (c,), "rm"), (f"{pre}.{i + 1}.running_var", (c,), "rv"),
                        (f"{pre}.{i + 1}.num_batches_tracked", (), "n")])
        out.append((f"{pre}.12.weight", (D, 4, nf, 1), "w"))

    def ffn(pre, D):
        out.extend([(pre + ".0.weight", (D,), "g"), (pre + ".0.bias", (D,), "b"),
                    (pre + ".1.linear.weight", (4 * D, D), "w"), (pre + ".1.linear.bias", (4 * D,), "b"),
                    (pre + ".4.linear.weight", (D, 4 * D), "w"), (pre + ".4.linear.bias", (D,), "b")])

    def block(pre, D, H):
        s = pre + ".sequential"
        ffn(s + ".0.module.sequential", D)
        m = s + ".1.module"
        out.append((m + ".positional_encoding.pe", (1, 10000, D), "pe"))
        out.extend([(m + ".layer_norm.weight", (D,), "g"), (m + ".layer_norm.bias", (D,), "b")])
        a = m + ".attention"
        out.extend([(a + ".u_bias", (H, D // H), "uv"), (a + ".v_bias", (H, D // H), "uv")])
        for nm in ("query_proj", "key_proj", "value_proj"):
            out.extend([(f"{a}.{nm}.linear.weight", (D, D), "w"), (f"{a}.{nm}.linear.bias", (D,), "b")])
        out.append((a + ".pos_proj.linear.weight", (D, D), "w"))
        out.extend([(a + ".out_proj.linear.weight", (D, D), "w"), (a + ".out_proj.linear.bias", (D,), "b")])
        c = s + ".2.module.sequential"
        out.extend([(c + ".0.weight", (D,), "g"), (c + ".0.bias", (D,), "b"),
                    (c + ".2.conv.weight", (2 * D, D, 1), "w"), (c + ".2.conv.bias", (2 * D,), "b"),
                    (c + ".4.conv.weight", (D, 1, 31), "w"),
                    (c + ".5.weight", (D,), "g"), (c + ".5.bias", (D,), "b"),
                    (c + ".5.running_mean", (D,), "rm"), (c + ".5.running_var", (D,), "rv"),
                    (c + ".5.num_batches_tracked", (), "n"),
                    (c + ".7.conv.weight", (D, D, 1), "w"), (c + ".7.conv.bias", (D,), "b")])
        ffn(s + ".3.module.sequential", D)
        out.extend([(s + ".4.weight", (D,), "g"), (s + ".4.bias", (D,), "b")])

    for enc, D, nl in (("spec_encoder", 512, SPEC_LAYERS), ("spat_encoder", 256, SPAT_LAYERS)):
        stem(enc + ".patch_embed", D)
        for l in range(nl):
            block(f"{enc}.embed.layers.{l}", D, NHEAD)
    def decoder(name, din):
        out.extend([(name + ".proj.0.weight", (3 * 4 * nf, din), "w"), (name + ".proj.0.bias", (3 * 4 * nf,), "b"),
                    (name + ".proj.2.weight", (4 * nf, 3 * 4 * nf), "w"), (name + ".proj.2.bias", (4 * nf,), "b")])

    if pretrain:
        decoder("decoder", 768)
    elif frozen:
        decoder("spec_spat_decoder", 768)
        decoder("spec_decoder", 512)
        decoder("spat_decoder", 512)
    elif head == "mlp" and dlabel == 1:       # downstream head (model.py:495-500)
        out.extend([("mlp_head.0.weight", (dembed_ds,), "g"), ("mlp_head.0.bias", (dembed_ds,), "b"),
                    ("mlp_head.1.weight", (1, dembed_ds), "w"), ("mlp_head.1.bias", (1,), "b")])
    elif head == "mlp":                       # joint_head (model.py:501-507)
        d = dembed_ds
        out.extend([("joint_head.0.weight", (d,), "g"), ("joint_head.0.bias", (d,), "b"), ("joint_head.1.weight", (d, d), "w"), ("joint_head.1.bias", (d,), "b"),
                    ("joint_head.3.weight", (dlabel, d), "w"), ("joint_head.3.bias", (dlabel,), "b")])
    out = [(prefix + k, shp, kind) for k, shp, kind in out]
    if nmic_pair:             # SARSSL_MultiCH.head_mch (model.py:807-812): LayerNorm, Linear, ReLU, Linear over the concatenated pair embeddings
        d = 256 * nmic_pair
        out.extend([("head_mch.0.weight", (d,), "g"), ("head_mch.0.bias", (d,), "b"), ("head_mch.1.weight", (d, d), "w"), ("head_mch.1.bias", (d,), "b"),
                    ("head_mch.3.weight", (factor, d), "w"), ("head_mch.3.bias", (factor,), "b")])
    return out


def synthetic_state_dict(seed=7, nf=256, pretrain=True, dembed_ds=768, frozen=False, **spec_kw):
    """Seeded, non-degenerate weights (non-zero biases, non-unit gains) for parity tests.  Generated
    key by key from one torch.Generator so reference, oracle and CUDA path can all be loaded with it."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape, kind in state_dict_spec(nf, pretrain, dembed_ds, frozen, **spec_kw):
        if kind == "w":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            sd[key] = torch.randn(shape, generator=g) * (1.0 / math.sqrt(fan_in))
        elif kind == "b":
            sd[key] = 0.1 * torch.randn(shape, generator=g)
        elif kind == "g":
            sd[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "uv":
            sd[key] = 0.2 * torch.randn(shape, generator=g)
        elif kind == "rm":
            sd[key] = torch.zeros(shape)
        elif kind == "rv":
            sd[key] = torch.ones(shape)
        elif kind == "n":
            sd[key] = torch.zeros(shape, dtype=torch.int64)
        elif kind == "pe":
            sd[key] = positional_table(shape[2], shape[1])
    return sd


def synthetic_waveforms(nb, nsample, nch=2, seed=1234):
    """0.1*randn, channel j>0 = delayed, attenuated copy of channel 0 + noise so that
    `diff` is meaningful."""
    g = torch.Generator().manual_seed(seed)
    pad = max(8, 3 * nch)
    base = 0.1 * torch.randn(nb, nsample + pad, generator=g)
    chans = [base[:, pad:]]
    for j in range(1, nch):
        d = 3 * j
        chans.append(0.8 * base[:, pad - d:pad - d + nsample] + 0.02 * torch.randn(nb, nsample, generator=g))
    return torch.stack(chans, dim=-1).contiguous()


def fixture_sample_idx(key, n, k):
    """Deterministic pseudo-random sample of min(k, n) element indices of a tensor with n elements (sorted, unique), keyed by the
    tensor's name: the gradient fixtures store the reference's values at these positions instead of whole 17.5 M-element gradients
    (an evenly spaced sample would alias with the tap / channel structure of the convolution weights)."""
    import zlib
    import numpy as np
    if n <= k:
        return np.arange(n, dtype=np.int64)
    rng = np.random.default_rng(zlib.crc32(key.encode()))
    return np.sort(rng.choice(n, size=k, replace=False)).astype(np.int64)


def adam_step(params, grads, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam (weight_decay 0, amsgrad off) as used at learner.py:83: in-place on the lists."""
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    for p, g, mi, vi in zip(params, grads, m, v):
        mi.mul_(b1).add_(g, alpha=1 - b1)
        vi.mul_(b2).addcmul_(g, g, value=1 - b2)
        p.addcdiv_(mi, (vi.sqrt() / math.sqrt(bc2)).add_(eps), value=-lr / bc1)


def cosine_lr(epoch, nepoch=30, base=1e-3, warmup=1):
    """common/utils.py:108-139 as called at run_pretrain.py:226 (cosine, 1 warm-up epoch)."""
    prog = min(max((epoch - warmup) / float(nepoch - warmup), 0.0), 1.0)
    lr = base * 0.5 * (1.0 + math.cos(math.pi * prog))
    if warmup:
        lr *= min(1.0, epoch / warmup)
    return lr
