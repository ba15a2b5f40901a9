"""Fused forward / backward schedule of the SAR-SSL pre-training model (the B200-native replacement for autograd over
the reference's nn.Module graph, code/model.py:519-601).

The whole model is ONE explicit schedule of kernel launches over pre-packed weights: no per-op autograd nodes, no dense
masks, no permute/unfold copies, every layout chosen for the kernels:
  tokens        [B*T][D]            (frame-major, the Conformer's (batch, time, dim))
  stem images   [B][T][F][C]        (channel-last; the 4 input channels (re0, re1, im0, im1) ARE the patch layout)
  scores        [B][H][T][T] content / probabilities, [H][B][T][T] positional (so dP is one GEMM over all clips)
Backward is written by hand (same kernels: the GEMM with swapped strides, mirrored-tap convolutions, reduction kernels)
and accumulates parameter gradients straight into the flat fp32 gradient arena.
Reference lines are cited next to each step."""
import math

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_RELU, ACT_SWISH
from .kernels import KernelSet
from .params import CNN_CH, DW_K, NHEAD, SPAT_D, SPAT_LAYERS, SPEC_D, SPEC_LAYERS

ENCODERS = (("spec_encoder", SPEC_D, SPEC_LAYERS, 1, 0), ("spat_encoder", SPAT_D, SPAT_LAYERS, 2, SPEC_D))   # name, D, layers, mask mode, column in cat


def _site_seed(step_seed, site):
    """Dropout seed of call site `site` in step `step_seed`.  step_seed None (CUDA-graph replay): only the site part - the kernels add the step
    part, step_part(step_seed), from a device word, so a replayed graph draws fresh masks every step and eager / replayed steps agree bit for bit."""
    if step_seed is None:
        return site * 7919 + 12345
    return (step_seed * 1000003 + site * 7919 + 12345) & 0x7FFFFFFFFFFFFFFF


def step_part(step_seed):
    return step_seed * 1000003


class Engine:
    def __init__(self, store, device, dtype=torch.float32, dropout_p=0.1):
        self.store = store
        self.dev = torch.device(device)
        self.dropout_p = dropout_p
        self.set_dtype(dtype)
        self.step_seed = 0
        self._pe_cache = {}

    def _seed(self):
        """The step's dropout seed as the kernels' by-value argument: the running step number, or None while the step part lives on the device."""
        return None if self.k.seed_dev is not None else self.step_seed

    def use_device_seed(self, seed_dev):
        """seed_dev: a uint64 device tensor holding step_part(step number) (written before every replay), or None for by-value seeds."""
        self.k.seed_dev = seed_dev
        if self.k32 is not self.k:
            self.k32.seed_dev = seed_dev

    def set_dtype(self, dtype):
        self.dtype = dtype
        self.k = KernelSet(self.dev, dtype)
        self.k32 = self.k if dtype == torch.float32 else KernelSet(self.dev, torch.float32)      # the tiny downstream head runs in fp32
        self._pe_cache = {}

    # ------------------------------------------------------------------------------------------------ weights
    def _prepare_weights(self, nf):
        """Compute-dtype views / packed copies of the weights for this step."""
        st, k = self.store, self.k
        cw = st.compute_copy(self.dtype, k)
        W = {}

        def view(key, shape=None):
            o, n = st.offsets[key]
            if shape is not None:                      # e.g. the [3D, D] QKV matrix starting at query_proj.weight
                n = 1
                for d in shape:
                    n *= d
            v = cw[o:o + n]
            return v.view(shape if shape is not None else st.shapes[key])

        self._wview = view
        for enc, D, nl, _, _ in ENCODERS:
            pe = enc + ".patch_embed"
            for i in (3, 6):
                src = view(f"{pe}.{i}.weight")                                   # [o][ci][kh (bin)][kw (frame)]
                fwd = k.empty(CNN_CH, 9, CNN_CH)                                  # [o][tap = kw*3+kh][ci]
                k.permute4(src, fwd, (CNN_CH, 3, 3, CNN_CH), (576, 1, 3, 9))
                bwd = k.empty(CNN_CH, 9, CNN_CH)                                  # [ci][mirrored tap][o]
                k.permute4(src, bwd, (CNN_CH, 3, 3, CNN_CH), (9, -1, -3, 576), src_off=8)
                W[f"{pe}.{i}.fwd"], W[f"{pe}.{i}.bwd"] = fwd, bwd
            src = view(f"{pe}.12.weight")                                        # (D, 4, nf, 1) -> [D][nf*4]
            pk = k.empty(D, nf * 4)
            k.permute4(src, pk, (D, nf, 4, 1), (4 * nf, 1, nf, 0))
            W[f"{pe}.12.packed"] = pk
            w9 = st.p(f"{pe}.9.weight")                                          # (4, 64, 1, 1) fp32 -> transposed [64][4] for the data gradient
            w9t = torch.empty(CNN_CH, 4, dtype=torch.float32, device=self.dev)
            k.permute4(w9, w9t, (CNN_CH, 4, 1, 1), (1, CNN_CH, 0, 0))
            W[f"{pe}.9.T"] = w9t
        self.W = W

    def w(self, key, shape=None):
        return self._wview(key, shape)

    def _pe(self, key, T, D):
        ck = (key, T, self.dtype)
        if ck not in self._pe_cache:
            src = self.store.b(key)[0, :T].contiguous()
            if self.dtype == torch.float32:
                self._pe_cache[ck] = src
            else:
                dst = self.k.empty(T, D)
                self.k.cast(src, dst, T * D)
                self._pe_cache[ck] = dst
        return self._pe_cache[ck]

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self, patches, flag, ch, nmasked, training, want_grad, frozen=False, plain=False):
        """patches (B, T, F, 2, 2) f32; flag (B, T) uint8; ch (B,) int32.  Returns (out2 = [loss, diff], pred (B*T, F*4), saved).
        frozen: the pretrain_frozen_encoder branch (model.py:603-666) - spectral input mode 4, `spec_spat_decoder` instead of `decoder`.
        plain: MCConformer.forward (model.py:890-912) - un-masked input into both encoders, decoder, no masks and no loss (out2 is None;
        backward() then takes the gradient of the prediction from the caller)."""
        k, st = self.k, self.store
        B, T, F = patches.shape[:3]
        M, P = B * T, B * T * F
        self.step_seed += 1
        p_drop = self.dropout_p if training else 0.0
        self._prepare_weights(F)
        dec = "spec_spat_decoder" if frozen else "decoder"
        sv = {"B": B, "T": T, "F": F, "p_drop": p_drop, "seed": self._seed(), "flag": flag, "ch": ch, "patches": patches, "dec": dec,
              "frozen": frozen} if want_grad else None
        cat = k.empty(M, SPEC_D + SPAT_D)
        site = [0]
        # encoders whose parameters are all frozen keep no activations (their backward is skipped, see backward())
        enc_sv = sv if (sv is not None and any(p.requires_grad for key, p in st.params.items() if "_encoder." in key)) else None
        for enc, D, nl, mode, col in ENCODERS:
            if frozen and mode == 1:
                mode = 4
            if plain:
                mode = 3
            e = self._stem_fwd(enc, D, mode, patches, flag, ch, B, T, F, training, enc_sv)
            for l in range(nl):
                last = l == nl - 1
                e = self._block_fwd(f"{enc}.embed.layers.{l}", D, e, B, T, training, p_drop, site, enc_sv, out=(cat, col, SPEC_D + SPAT_D) if last else None)
        # decoder MLP                                                                             model.py:297-301,582
        dff = st.shapes[dec + ".proj.0.weight"][0]
        hdec = k.empty(M, dff)
        k.linear(cat, self.w(dec + ".proj.0.weight"), hdec, M, dff, SPEC_D + SPAT_D, bias=st.p(dec + ".proj.0.bias"), act=ACT_RELU)
        pred = k.empty(M, 4 * F)
        k.linear(hdec, self.w(dec + ".proj.2.weight"), pred, M, 4 * F, dff, bias=st.p(dec + ".proj.2.bias"))
        if plain:
            if want_grad:
                sv.update(cat=cat, hdec=hdec, dpred=None, plain=True)
            return None, pred, sv
        # masked reconstruction loss (+ dpred in the same launch)                                  model.py:585-592,721-747
        from . import ops
        out2, dpred = ops.masked_loss(pred.view(B, T, 4 * F), patches, flag, ch, nmasked, want_grad=want_grad)
        k.launches += 1
        if want_grad:
            sv.update(cat=cat, hdec=hdec, dpred=dpred.view(M, 4 * F))
        return out2, pred, sv

    def _bn(self, key, y, rows, C, training):
        st = self.store
        return self.k.bn_stats(y, rows, C, st.p(key + ".weight"), st.p(key + ".bias"), st.b(key + ".running_mean"), st.b(key + ".running_var"),
                               st.b(key + ".num_batches_tracked"), training)

    def _bn_args(self, key):
        st = self.store
        return (st.p(key + ".weight"), st.p(key + ".bias"), st.b(key + ".running_mean"), st.b(key + ".running_var"), st.b(key + ".num_batches_tracked"))

    def _stem_fwd(self, enc, D, mode, patches, flag, ch, B, T, F, training, sv):
        """model.py:50-64,203-208: 1x1 (4->64) BN ReLU, 3x3 BN ReLU, 3x3 BN ReLU, 1x1 (64->4) BN ReLU, (F x 1) patch conv -> (B*T, D)."""
        k, st = self.k, self.store
        pe = enc + ".patch_embed"
        M, P = B * T, B * T * F
        y1, y2, y3, z1, z2 = None, k.empty(P, CNN_CH), k.empty(P, CNN_CH), None, None
        if k.conv_tc:
            # tensor-core path: TMA feeds the MMA directly, so BatchNorm+ReLU outputs are materialised (bf16).  The first layer is linear
            # in 4 channels: its BatchNorm statistics come from the input's moments and conv+BN+ReLU is one pass; y1 is never stored.
            z1 = k.empty(P, CNN_CH)
            if training:
                s1 = k.stem_input_bn_stats(patches, mode, flag, ch, st.p(f"{pe}.0.weight"), P, F, T, self._bn_args(f"{pe}.1"))
            else:
                s1 = self._bn(f"{pe}.1", z1, P, CNN_CH, False)          # running statistics; the tensor argument is not read
            k.stem_expand_bn_relu(patches, mode, flag, ch, st.p(f"{pe}.0.weight"), s1, z1, P, F, T)
            s2 = k.conv3x3_tc(z1, self.W[f"{pe}.3.fwd"], y2, B, T, F, bn=self._bn_args(f"{pe}.4") if training else None)    # BN statistics fused in the epilogue
        else:                    # CUDA-core path applies BatchNorm+ReLU while loading the operand tile
            y1 = k.empty(P, CNN_CH)
            k.stem_expand(patches, mode, flag, ch, st.p(f"{pe}.0.weight"), y1, P, F, T)
            s1 = self._bn(f"{pe}.1", y1, P, CNN_CH, training)
            k.conv3x3(y1, s1, self.W[f"{pe}.3.fwd"], y2, B, T, F)
            s2 = None
        if s2 is None:
            s2 = self._bn(f"{pe}.4", y2, P, CNN_CH, training)
        if k.conv_tc:            # BatchNorm + ReLU of y2 is applied to the conv's operand tiles in shared memory: z2 is never stored
            s3 = k.conv3x3_tc(y2, self.W[f"{pe}.6.fwd"], y3, B, T, F, bn=self._bn_args(f"{pe}.7") if training else None, in_stats=s2)
        else:
            k.conv3x3(y2, s2, self.W[f"{pe}.6.fwd"], y3, B, T, F)
            s3 = None
        if s3 is None:
            s3 = self._bn(f"{pe}.7", y3, P, CNN_CH, training)
        y4 = k.empty(P, 4)
        k.stem_reduce(y3, s3, st.p(f"{pe}.9.weight"), y4, P)
        s4 = self._bn(f"{pe}.10", y4, P, 4, training)
        z4 = k.empty(M, 4 * F)
        k.bn_act_fwd(y4, s4, ACT_RELU, z4, P, 4)
        e = k.empty(M, D)
        k.linear(z4, self.W[f"{pe}.12.packed"], e, M, D, 4 * F)
        if sv is not None:
            sv[pe] = dict(y1=y1, s1=s1, y2=y2, s2=s2, y3=y3, s3=s3, y4=y4, s4=s4, z4=z4, z1=z1, z2=z2)
        return e

    def _ffn_fwd(self, pre, D, x, M, p_drop, site, rec):
        """feed_forward.py:39-54 inside ResidualConnectionModule(module_factor=0.5) (Conformer.py:60-67, modules.py:33)."""
        k, st = self.k, self.store
        h, mean, rstd = k.empty(M, D), k.empty(M, dtype=torch.float32), k.empty(M, dtype=torch.float32)
        k.layernorm_fwd(x, D, st.p(pre + ".0.weight"), st.p(pre + ".0.bias"), h, D, mean, rstd, M, D)
        u, s = k.empty(M, 4 * D), k.empty(M, 4 * D)
        da = (p_drop, _site_seed(rec["seed"], site[0])); site[0] += 1
        db = (p_drop, _site_seed(rec["seed"], site[0])); site[0] += 1
        k.linear(h, self.w(pre + ".1.linear.weight"), s, M, 4 * D, D, bias=st.p(pre + ".1.linear.bias"), act=ACT_SWISH, pre=u, drop=da)
        xo = k.empty(M, D)
        k.linear(s, self.w(pre + ".4.linear.weight"), xo, M, D, 4 * D, bias=st.p(pre + ".4.linear.bias"), resid=x, ldr=D, beta=0.5, drop=db)
        rec.update(x=x, h=h, mean=mean, rstd=rstd, u=u, s=s, da=da, db=db)
        return xo

    def _block_fwd(self, pre, D, x, B, T, training, p_drop, site, sv, out=None):
        """common/Conformer.py:59-88."""
        k, st = self.k, self.store
        M, H, dh = B * T, NHEAD, D // NHEAD
        s = pre + ".sequential"
        rec = {"seed": self._seed()}
        f1 = {"seed": self._seed()}
        x1 = self._ffn_fwd(s + ".0.module.sequential", D, x, M, p_drop, site, f1)
        # ---- relative-position MHSA                                                              attention.py:72-103,143-151
        m = s + ".1.module"
        a = m + ".attention"
        h2, mean2, rstd2 = k.empty(M, D), k.empty(M, dtype=torch.float32), k.empty(M, dtype=torch.float32)
        k.layernorm_fwd(x1, D, st.p(m + ".layer_norm.weight"), st.p(m + ".layer_norm.bias"), h2, D, mean2, rstd2, M, D)
        qkv = k.empty(M, 3 * D)
        k.linear(h2, self.w(a + ".query_proj.linear.weight", (3 * D, D)), qkv, M, 3 * D, D, bias=st.arena_view(st.flat, a + ".query_proj.linear.bias", 3 * D))
        pe_t = self._pe(m + ".positional_encoding.pe", T, D)
        pp = k.empty(T, D)                                                        # pos_proj(PE[:T]) - batch invariant, computed once
        k.linear(pe_t, self.w(a + ".pos_proj.linear.weight"), pp, T, D, D)
        qu, qv = k.empty(M, D), k.empty(M, D)
        k.add_head_bias(qkv, 3 * D, st.p(a + ".u_bias"), st.p(a + ".v_bias"), qu, qv, M, D)
        content, pos, prob = k.empty(B, H, T, T), k.empty(H, B, T, T), k.empty(B, H, T, T)
        k.gemm(qu, qkv, content, T, T, dh, (D, 1), (3 * D, 1), T, b_off=D, batch=(B, H), sAb=(T * D, dh), sBb=(T * 3 * D, dh), sCb=(H * T * T, T * T))
        k.gemm(qv, pp, pos, T, T, dh, (D, 1), (D, 1), T, batch=(B, H), sAb=(T * D, dh), sBb=(0, dh), sCb=(T * T, B * T * T))
        dp = (p_drop, _site_seed(self._seed(), site[0])); site[0] += 1
        do = (p_drop, _site_seed(self._seed(), site[0])); site[0] += 1
        attn = k.empty(B, H, T, T) if p_drop > 0 else None            # Dropout(prob), materialised so the context GEMMs run on tensor cores
        k.attn_softmax_fwd(content, pos, prob, attn, B, H, T, 1.0 / math.sqrt(D), dp)
        del content, pos
        if attn is None:
            attn = prob
        ctx = k.empty(M, D)
        k.gemm(attn, qkv, ctx, T, dh, T, (T, 1), (1, 3 * D), D, b_off=2 * D, batch=(B, H), sAb=(H * T * T, T * T), sBb=(T * 3 * D, dh), sCb=(T * D, dh))
        x2 = k.empty(M, D)
        k.linear(ctx, self.w(a + ".out_proj.linear.weight"), x2, M, D, D, bias=st.p(a + ".out_proj.linear.bias"), resid=x1, ldr=D, beta=1.0, drop=do)
        # ---- convolution module                                                                  convolution.py:136-149
        c = s + ".2.module.sequential"
        h3, mean3, rstd3 = k.empty(M, D), k.empty(M, dtype=torch.float32), k.empty(M, dtype=torch.float32)
        k.layernorm_fwd(x2, D, st.p(c + ".0.weight"), st.p(c + ".0.bias"), h3, D, mean3, rstd3, M, D)
        g = k.empty(M, 2 * D)
        k.linear(h3, self.w(c + ".2.conv.weight", (2 * D, D)), g, M, 2 * D, D, bias=st.p(c + ".2.conv.bias"))
        ga = k.empty(M, D)
        k.glu_fwd(g, ga, M, D)
        cv = k.empty(M, D)
        k.dwconv(ga, st.p(c + ".4.conv.weight"), cv, B, T, D, DW_K, False)
        sbn = self._bn(c + ".5", cv, M, D, training)
        z = k.empty(M, D)
        k.bn_act_fwd(cv, sbn, ACT_SWISH, z, M, D)
        dc = (p_drop, _site_seed(self._seed(), site[0])); site[0] += 1
        x3 = k.empty(M, D)
        k.linear(z, self.w(c + ".7.conv.weight", (D, D)), x3, M, D, D, bias=st.p(c + ".7.conv.bias"), resid=x2, ldr=D, beta=1.0, drop=dc)
        # ---- second half-step FFN + final LayerNorm
        f2 = {"seed": self._seed()}
        x4 = self._ffn_fwd(s + ".3.module.sequential", D, x3, M, p_drop, site, f2)
        mean5, rstd5 = k.empty(M, dtype=torch.float32), k.empty(M, dtype=torch.float32)
        if out is None:
            y = k.empty(M, D)
            k.layernorm_fwd(x4, D, st.p(s + ".4.weight"), st.p(s + ".4.bias"), y, D, mean5, rstd5, M, D)
        else:
            buf, col, ld = out                                                    # write straight into the concatenated decoder input
            k.layernorm_fwd(x4, D, st.p(s + ".4.weight"), st.p(s + ".4.bias"), buf, ld, mean5, rstd5, M, D, out_off=col)
            y = None
        if sv is not None:
            rec.update(f1=f1, f2=f2, x1=x1, h2=h2, mean2=mean2, rstd2=rstd2, qkv=qkv, pp=pp, qu=qu, qv=qv, prob=prob, attn=attn, ctx=ctx, dp=dp, do=do,
                       x2=x2, h3=h3, mean3=mean3, rstd3=rstd3, g=g, ga=ga, cv=cv, sbn=sbn, z=z, dc=dc, x3=x3, x4=x4, mean5=mean5, rstd5=rstd5)
            sv[pre] = rec
        return y

    # ------------------------------------------------------------------------------------------------ downstream branch
    def forward_downstream(self, patches, embed_use, training, want_grad, head="mlp"):
        """model.py:667-719 (pretrain=False): both encoders on the UN-masked input, concatenate, mean over time, then the head:
        'mlp' = LayerNorm + Linear(dembed, 1) (dlabel 1), 'joint' = joint_head (dlabel > 1, model.py:501-507), '' = none, an integer nmic_pair =
        SARSSL_MultiCH.head_mch (model.py:793-821).
        Returns (pred fp32, pooled (B, dembed) fp32, saved)."""
        k, k32, st = self.k, self.k32, self.store
        B, T, F = patches.shape[:3]
        M = B * T
        self.step_seed += 1
        p_drop = self.dropout_p if training else 0.0
        self._prepare_weights(F)
        sv = {"B": B, "T": T, "F": F, "seed": self._seed(), "patches": patches, "flag": None, "ch": None, "embed_use": embed_use} if want_grad else None
        Dc = SPEC_D + SPAT_D
        cat = k.empty(M, Dc)
        site = [0]
        for enc, D, nl, _, col in ENCODERS:                # (the reference runs both encoders whatever `downstream_embed` says, model.py:677-678)
            e = self._stem_fwd(enc, D, 3, patches, None, None, B, T, F, training, sv)
            for l in range(nl):
                last = l == nl - 1
                e = self._block_fwd(f"{enc}.embed.layers.{l}", D, e, B, T, training, p_drop, site, sv, out=(cat, col, Dc) if last else None)
        col, Dd = {"spec_spat": (0, Dc), "spec": (0, SPEC_D), "spat": (SPEC_D, SPAT_D)}[embed_use]
        pooled = torch.empty(B, Dd, dtype=torch.float32, device=self.dev)
        k.mean_pool_fwd(cat, Dc, pooled, B, T, Dd, x_off=col)
        k32.use_tc = False
        f32 = dict(dtype=torch.float32, device=self.dev)
        if head == "":                               # SARSSL(downstream_head=''): the pooled embedding is the prediction   model.py:705-706
            pred, hs = pooled, {}
        elif head == "mlp":                          # LayerNorm + Linear(dembed, 1)                                       model.py:495-500
            hn, mean, rstd = torch.empty_like(pooled), torch.empty(B, **f32), torch.empty(B, **f32)
            k32.layernorm_fwd(pooled, Dd, st.p("mlp_head.0.weight"), st.p("mlp_head.0.bias"), hn, Dd, mean, rstd, B, Dd)
            pred = torch.empty(B, 1, **f32)
            k32.linear(hn, st.p("mlp_head.1.weight"), pred, B, 1, Dd, bias=st.p("mlp_head.1.bias"))
            hs = dict(hn=hn, hmean=mean, hrstd=rstd)
        else:   # LayerNorm, Linear, ReLU, Linear: SARSSL_MultiCH.head_mch over the P pairs of an item (model.py:807-820) or joint_head, P = 1 (model.py:501-507)
            hk, P = ("joint_head", 1) if head == "joint" else ("head_mch", int(head))
            if B % P:
                raise _lib.SarsslError(f"SARSSL_MultiCH: batch of {B} clips is not a multiple of nmic_pair = {P}")
            nbm, Dh = B // P, P * Dd
            emb = pooled.view(nbm, Dh)
            factor = st.shapes[hk + ".3.weight"][0]
            hn, mean, rstd = torch.empty_like(emb), torch.empty(nbm, **f32), torch.empty(nbm, **f32)
            k32.layernorm_fwd(emb, Dh, st.p(hk + ".0.weight"), st.p(hk + ".0.bias"), hn, Dh, mean, rstd, nbm, Dh)
            h1 = torch.empty(nbm, Dh, **f32)
            k32.linear(hn, st.p(hk + ".1.weight"), h1, nbm, Dh, Dh, bias=st.p(hk + ".1.bias"), act=ACT_RELU)
            pred = torch.empty(nbm, factor, **f32)
            k32.linear(h1, st.p(hk + ".3.weight"), pred, nbm, factor, Dh, bias=st.p(hk + ".3.bias"))
            hs = dict(hn=hn, hmean=mean, hrstd=rstd, h1=h1, nbm=nbm, Dh=Dh, factor=factor, hk=hk)
        if want_grad:
            sv.update(pooled=pooled, col=col, Dd=Dd, head=head, **hs)
        return pred, pooled, sv

    def backward_downstream(self, sv, dpred, on_ready=None):
        """dpred fp32 = d loss / d pred ((B, 1) for the 'mlp' head, (B, dembed) without a head, (B / nmic_pair, factor) for head_mch)."""
        k, k32, st = self.k, self.k32, self.store
        B, T, F, col, Dd = sv["B"], sv["T"], sv["F"], sv["col"], sv["Dd"]
        M, Dc = B * T, SPEC_D + SPAT_D
        k32.use_tc = False
        head = sv["head"]
        f32 = dict(dtype=torch.float32, device=self.dev)
        if head == "":
            dpool = dpred.reshape(B, Dd).contiguous()
        elif head == "mlp":
            k32.linear_wgrad(dpred, sv["hn"], st.g("mlp_head.1.weight"), B, 1, Dd)
            # bias gradient = sum_b dpred[b]: a 1 x 1 GEMM over K = B against a stride-0 "ones" operand
            k32.gemm(dpred, torch.ones(1, 1, device=self.dev), st.g("mlp_head.1.bias").view(1, 1), 1, 1, B, (1, 1), (1, 0), 1, accumulate=True)
            dhn = torch.empty(B, Dd, **f32)
            k32.linear_dgrad(dpred, st.p("mlp_head.1.weight"), dhn, B, 1, Dd)
            dpool = torch.empty_like(dhn)
            k32.layernorm_bwd(dhn, Dd, sv["pooled"], Dd, sv["hmean"], sv["hrstd"], st.p("mlp_head.0.weight"), None, dpool, st.g("mlp_head.0.weight"),
                              st.g("mlp_head.0.bias"), B, Dd)
        else:
            nbm, Dh, factor, hk = sv["nbm"], sv["Dh"], sv["factor"], sv["hk"]
            ones = torch.ones(1, 1, device=self.dev)
            dpred = dpred.reshape(nbm, factor).contiguous()
            k32.linear_wgrad(dpred, sv["h1"], st.g(hk + ".3.weight"), nbm, factor, Dh)
            k32.gemm(dpred, ones, st.g(hk + ".3.bias").view(factor, 1), factor, 1, nbm, (1, factor), (1, 0), 1, accumulate=True)     # column sums of dpred
            dh1 = torch.empty(nbm, Dh, **f32)
            k32.linear_dgrad(dpred, st.p(hk + ".3.weight"), dh1, nbm, factor, Dh)
            k32.relu_bwd(dh1, sv["h1"], dh1, nbm * Dh)
            k32.linear_wgrad(dh1, sv["hn"], st.g(hk + ".1.weight"), nbm, Dh, Dh)
            k32.gemm(dh1, ones, st.g(hk + ".1.bias").view(Dh, 1), Dh, 1, nbm, (1, Dh), (1, 0), 1, accumulate=True)
            dhn = torch.empty(nbm, Dh, **f32)
            k32.linear_dgrad(dh1, st.p(hk + ".1.weight"), dhn, nbm, Dh, Dh)
            dpool = torch.empty(nbm, Dh, **f32)
            k32.layernorm_bwd(dhn, Dh, sv["pooled"].view(nbm, Dh), Dh, sv["hmean"], sv["hrstd"], st.p(hk + ".0.weight"), None, dpool,
                              st.g(hk + ".0.weight"), st.g(hk + ".0.bias"), nbm, Dh)
            dpool = dpool.view(B, Dd)
        dcat = k.empty(M, Dc)
        k.mean_pool_bwd(dpool, dcat, Dc, B, T, Dd, dx_off=col)
        ready = on_ready if on_ready is not None else (lambda name: None)
        ready("decoder")
        for enc, D, nl, mode, ecol in ENCODERS:
            used = ecol >= col and ecol < col + Dd             # encoders outside the pooled slice get no gradient (model.py:693-700)
            if used:
                d = None
                for l in reversed(range(nl)):
                    last = l == nl - 1
                    d = self._block_bwd(f"{enc}.embed.layers.{l}", D, sv, B, T, dout=(dcat, ecol, Dc) if last else (d, 0, D))
                    if enc == "spat_encoder" and l == 1:
                        ready("spat_blocks_1_2")
                self._stem_bwd(enc, D, 3, sv, d, B, T, F)
            elif enc == "spat_encoder":
                ready("spat_blocks_1_2")
            ready("spec_encoder" if enc == "spec_encoder" else "spat_stem_block_0")

    # ------------------------------------------------------------------------------------------------ backward
    def backward(self, sv, gscale=None, on_ready=None):
        """Accumulates d loss / d parameter into the gradient arena.  gscale: device scalar multiplying the loss gradient.
        on_ready(name) is called when a gradient bucket (sarssl_b200/parallel.py) is complete, so its all-reduce can start."""
        k, st = self.k, self.store
        B, T, F = sv["B"], sv["T"], sv["F"]
        M = B * T
        dpred = sv["dpred"]
        if gscale is not None and not sv.get("plain"):
            _lib.check(_lib.lib().sarssl_scale_masked_rows(_lib.ptr(dpred), _lib.dtype_code(dpred), _lib.ptr(sv["flag"]), _lib.ptr(gscale), B, T, F,
                                                           k.stream), "scale_masked_rows")
        dec = sv["dec"]
        dff = st.shapes[dec + ".proj.0.weight"][0]
        Dc = SPEC_D + SPAT_D
        # decoder
        k.linear_wgrad(dpred, sv["hdec"], st.g(dec + ".proj.2.weight"), M, 4 * F, dff)
        k.colsum(dpred, 4 * F, st.g(dec + ".proj.2.bias"), M, 4 * F)
        dh = k.empty(M, dff)
        k.linear_dgrad(dpred, self.w(dec + ".proj.2.weight"), dh, M, 4 * F, dff)
        k.relu_bwd(dh, sv["hdec"], dh, M * dff)
        k.linear_wgrad(dh, sv["cat"], st.g(dec + ".proj.0.weight"), M, dff, Dc)
        k.colsum(dh, dff, st.g(dec + ".proj.0.bias"), M, dff)
        ready = on_ready if on_ready is not None else (lambda name: None)
        ready("decoder")
        if not any(p.requires_grad for key, p in st.params.items() if "_encoder." in key):      # every encoder parameter is frozen (run_pretrain.py:364-371): nothing below the decoder needs a gradient
            for name in ("spec_encoder", "spat_blocks_1_2", "spat_stem_block_0"):
                ready(name)
            return
        dcat = k.empty(M, Dc)
        k.linear_dgrad(dh, self.w(dec + ".proj.0.weight"), dcat, M, dff, Dc)
        del dh
        for enc, D, nl, mode, col in ENCODERS:
            if sv["frozen"] and mode == 1:
                mode = 4
            if sv.get("plain"):
                mode = 3
            d = None
            for l in reversed(range(nl)):
                last = l == nl - 1
                d = self._block_bwd(f"{enc}.embed.layers.{l}", D, sv, B, T, dout=(dcat, col, Dc) if last else (d, 0, D))
                if enc == "spat_encoder" and l == 1:
                    ready("spat_blocks_1_2")
            self._stem_bwd(enc, D, mode, sv, d, B, T, F)
            ready("spec_encoder" if enc == "spec_encoder" else "spat_stem_block_0")

    def _ffn_bwd(self, pre, D, rec, dxo, M):
        """dxo = gradient w.r.t. the module output x + 0.5*FFN(x); returns the gradient w.r.t. x."""
        k, st = self.k, self.store
        dv = k.empty(M, D)
        k.scale_dropout(dxo, dv, M * D, 0.5, rec["db"])
        k.linear_wgrad(dv, rec["s"], st.g(pre + ".4.linear.weight"), M, D, 4 * D)
        k.colsum(dv, D, st.g(pre + ".4.linear.bias"), M, D)
        ds = k.empty(M, 4 * D)
        k.linear_dgrad(dv, self.w(pre + ".4.linear.weight"), ds, M, D, 4 * D)
        k.swish_bwd(ds, rec["u"], ds, M * 4 * D, rec["da"])
        k.linear_wgrad(ds, rec["h"], st.g(pre + ".1.linear.weight"), M, 4 * D, D)
        k.colsum(ds, 4 * D, st.g(pre + ".1.linear.bias"), M, 4 * D)
        dh = k.empty(M, D)
        k.linear_dgrad(ds, self.w(pre + ".1.linear.weight"), dh, M, 4 * D, D)
        dx = k.empty(M, D)
        k.layernorm_bwd(dh, D, rec["x"], D, rec["mean"], rec["rstd"], st.p(pre + ".0.weight"), dxo, dx, st.g(pre + ".0.weight"), st.g(pre + ".0.bias"), M, D)
        return dx

    def _block_bwd(self, pre, D, sv, B, T, dout):
        k, st = self.k, self.store
        rec = sv[pre]
        M, H, dh_ = B * T, NHEAD, D // NHEAD
        s = pre + ".sequential"
        dbuf, dcol, dld = dout
        dx4 = k.empty(M, D)
        k.layernorm_bwd(dbuf, dld, rec["x4"], D, rec["mean5"], rec["rstd5"], st.p(s + ".4.weight"), None, dx4, st.g(s + ".4.weight"), st.g(s + ".4.bias"),
                        M, D, dy_off=dcol)
        dx3 = self._ffn_bwd(s + ".3.module.sequential", D, rec["f2"], dx4, M)
        del dx4
        # ---- convolution module
        c = s + ".2.module.sequential"
        do = k.empty(M, D)
        k.scale_dropout(dx3, do, M * D, 1.0, rec["dc"])
        k.linear_wgrad(do, rec["z"], st.g(c + ".7.conv.weight").view(D, D), M, D, D)
        k.colsum(do, D, st.g(c + ".7.conv.bias"), M, D)
        dz = k.empty(M, D)
        k.linear_dgrad(do, self.w(c + ".7.conv.weight", (D, D)), dz, M, D, D)
        k.bn_act_bwd(dz, rec["cv"], rec["sbn"], ACT_SWISH, dz, st.g(c + ".5.weight"), st.g(c + ".5.bias"), M, D)      # dz now holds d(conv out)
        k.dwconv_wgrad(rec["ga"], dz, st.g(c + ".4.conv.weight").view(D, DW_K), B, T, D, DW_K)
        dga = do
        k.dwconv(dz, st.p(c + ".4.conv.weight"), dga, B, T, D, DW_K, True)
        dg = k.empty(M, 2 * D)
        k.glu_bwd(dga, rec["g"], dg, M, D)
        k.linear_wgrad(dg, rec["h3"], st.g(c + ".2.conv.weight").view(2 * D, D), M, 2 * D, D)
        k.colsum(dg, 2 * D, st.g(c + ".2.conv.bias"), M, 2 * D)
        dh3 = dz
        k.linear_dgrad(dg, self.w(c + ".2.conv.weight", (2 * D, D)), dh3, M, 2 * D, D)
        dx2 = k.empty(M, D)
        k.layernorm_bwd(dh3, D, rec["x2"], D, rec["mean3"], rec["rstd3"], st.p(c + ".0.weight"), dx3, dx2, st.g(c + ".0.weight"), st.g(c + ".0.bias"), M, D)
        del dx3, dg, dh3, dga, do, dz
        # ---- MHSA
        m = s + ".1.module"
        a = m + ".attention"
        qkv, prob = rec["qkv"], rec["prob"]
        do = k.empty(M, D)
        k.scale_dropout(dx2, do, M * D, 1.0, rec["do"])
        k.linear_wgrad(do, rec["ctx"], st.g(a + ".out_proj.linear.weight"), M, D, D)
        k.colsum(do, D, st.g(a + ".out_proj.linear.bias"), M, D)
        dctx = k.empty(M, D)
        k.linear_dgrad(do, self.w(a + ".out_proj.linear.weight"), dctx, M, D, D)
        dattn = k.empty(B, H, T, T)
        k.gemm(dctx, qkv, dattn, T, T, dh_, (D, 1), (3 * D, 1), T, b_off=2 * D, batch=(B, H), sAb=(T * D, dh_), sBb=(T * 3 * D, dh_), sCb=(H * T * T, T * T))
        dqkv = k.empty(M, 3 * D)
        # dV[b, j, h, :] = sum_i drop(prob)[b, h, i, j] * dctx[b, i, h, :]
        k.gemm(rec["attn"], dctx, dqkv, T, dh_, T, (1, T), (1, D), 3 * D, c_off=2 * D, batch=(B, H), sAb=(H * T * T, T * T), sBb=(T * D, dh_),
               sCb=(T * 3 * D, dh_))
        dpos = k.empty(H, B, T, T)
        k.attn_softmax_bwd(dattn, prob, dpos, B, H, T, 1.0 / math.sqrt(D), rec["dp"])                              # dattn now holds dscore
        dqu, dqv = do, dctx
        k.gemm(dattn, qkv, dqu, T, dh_, T, (T, 1), (1, 3 * D), D, b_off=D, batch=(B, H), sAb=(H * T * T, T * T), sBb=(T * 3 * D, dh_), sCb=(T * D, dh_))
        # dK[b, j, h, :] = sum_i dscore[b, h, i, j] * (q + u)[b, i, h, :]
        k.gemm(dattn, rec["qu"], dqkv, T, dh_, T, (1, T), (1, D), 3 * D, c_off=D, batch=(B, H), sAb=(H * T * T, T * T), sBb=(T * D, dh_), sCb=(T * 3 * D, dh_))
        k.gemm(dpos, rec["pp"], dqv, T, dh_, T, (T, 1), (1, D), D, batch=(B, H), sAb=(T * T, B * T * T), sBb=(0, dh_), sCb=(T * D, dh_))
        # dPproj[k, h*dh + d] = sum_{b,i} dpos[h][b][i][k] * (q + v)[b, i, h, d]   (one GEMM per head over all B*T rows; only H x 2 output
        # tiles with K = B*T, so it accumulates in fp32 through split-K)
        dpp32 = torch.zeros(T, D, dtype=torch.float32, device=self.dev)
        k.gemm(dpos, rec["qv"], dpp32, T, dh_, B * T, (1, T), (1, D), D, batch=(1, H), sAb=(0, B * T * T), sBb=(0, dh_), sCb=(0, dh_), accumulate=True)
        if self.dtype == torch.float32:
            dpp = dpp32
        else:
            dpp = k.empty(T, D)
            k.cast(dpp32, dpp, T * D)
        k.colsum(dqu, D, st.g(a + ".u_bias").view(D), M, D)
        k.colsum(dqv, D, st.g(a + ".v_bias").view(D), M, D)
        k.add2(dqu, D, dqv, D, dqkv, 3 * D, M, D)
        pe_t = self._pe(m + ".positional_encoding.pe", T, D)
        k.gemm(dpp, pe_t, st.g(a + ".pos_proj.linear.weight"), D, D, T, (1, D), (1, D), D, accumulate=True)
        k.linear_wgrad(dqkv, rec["h2"], st.arena_view(st.grad, a + ".query_proj.linear.weight", 3 * D * D), M, 3 * D, D)
        k.colsum(dqkv, 3 * D, st.arena_view(st.grad, a + ".query_proj.linear.bias", 3 * D), M, 3 * D)
        dh2 = dqu
        k.linear_dgrad(dqkv, self.w(a + ".query_proj.linear.weight", (3 * D, D)), dh2, M, 3 * D, D)
        dx1 = k.empty(M, D)
        k.layernorm_bwd(dh2, D, rec["x1"], D, rec["mean2"], rec["rstd2"], st.p(m + ".layer_norm.weight"), dx2, dx1, st.g(m + ".layer_norm.weight"),
                        st.g(m + ".layer_norm.bias"), M, D)
        del dx2, dattn, dpos, dqkv, dqu, dqv, do, dctx, dh2
        return self._ffn_bwd(s + ".0.module.sequential", D, rec["f1"], dx1, M)

    def _stem_bwd(self, enc, D, mode, sv, de, B, T, F):
        k, st = self.k, self.store
        pe = enc + ".patch_embed"
        r = sv[pe]
        M, P = B * T, B * T * F
        # patch conv (F x 1, stride F x 1) = GEMM over K = F*4
        dwp = torch.zeros(D, 4 * F, dtype=torch.float32, device=self.dev)               # accumulate=True selects the split-K path: 16-32 output tiles, K = B*T
        k.gemm(de, r["z4"], dwp, D, 4 * F, M, (1, D), (1, 4 * F), 4 * F, accumulate=True)
        k.permute4(dwp, st.g(f"{pe}.12.weight"), (D, 4, F, 1), (4 * F, 1, 4, 0), accumulate=True)
        dz4 = k.empty(M, 4 * F)
        k.linear_dgrad(de, self.W[f"{pe}.12.packed"], dz4, M, D, 4 * F)
        k.bn_act_bwd(dz4, r["y4"], r["s4"], ACT_RELU, dz4, st.g(f"{pe}.10.weight"), st.g(f"{pe}.10.bias"), P, 4)        # dz4 -> dy4
        # 1x1 conv 64 -> 4 and the BatchNorm + ReLU in front of it
        dz = k.empty(P, CNN_CH)
        if k.conv_tc:            # one fused path: the P x 64 gradient of the ReLU output is never stored
            k.stem_tail_bwd(r["y3"], r["s3"], dz4, st.p(f"{pe}.9.weight"), st.g(f"{pe}.7.weight"), st.g(f"{pe}.7.bias"), st.g(f"{pe}.9.weight"), dz, P)
        else:
            tmp = torch.empty(CNN_CH, 4, dtype=torch.float32, device=self.dev)
            k.stem_pw_wgrad(r["y3"], r["s3"], dz4, 0, None, None, tmp, False, P, F, T)
            k.permute4(tmp, st.g(f"{pe}.9.weight"), (4, CNN_CH, 1, 1), (1, 4, 0, 0), accumulate=True)
            k.stem_expand(dz4, 0, None, None, self.W[f"{pe}.9.T"], dz, P, F, T)
            k.bn_act_bwd(dz, r["y3"], r["s3"], ACT_RELU, dz, st.g(f"{pe}.7.weight"), st.g(f"{pe}.7.bias"), P, CNN_CH)       # -> dy3
        del dz4
        dz_prev = k.empty(P, CNN_CH)
        for i, y_in, s_in, y_bn, bn_key, z_in in ((6, r["y2"], r["s2"], r["y2"], f"{pe}.4", r["z2"]), (3, r["y1"], r["s1"], r["y1"], f"{pe}.1", r["z1"])):
            dwpk = torch.empty(CNN_CH, 9, CNN_CH, dtype=torch.float32, device=self.dev)
            if k.conv_tc:        # layer 6 reads y2 and applies BatchNorm + ReLU on the fly (z_in is None there); layer 3 reads the stored z1
                k.conv3x3_wgrad_tc(dz, z_in if z_in is not None else y_in, dwpk, B, T, F, in_stats=s_in if z_in is None else None)
            else:
                k.conv3x3_wgrad(dz, y_in, s_in, dwpk, B, T, F)
            k.permute4(dwpk, st.g(f"{pe}.{i}.weight"), (CNN_CH, CNN_CH, 3, 3), (576, 1, 64, 192), accumulate=True)
            if k.conv_tc:
                k.conv3x3_tc(dz, self.W[f"{pe}.{i}.bwd"], dz_prev, B, T, F)
            else:
                k.conv3x3(dz, None, self.W[f"{pe}.{i}.bwd"], dz_prev, B, T, F)
            if not (k.conv_tc and i == 3):
                k.bn_act_bwd(dz_prev, y_bn, s_in, ACT_RELU, dz_prev, st.g(bn_key + ".weight"), st.g(bn_key + ".bias"), P, CNN_CH)
            dz, dz_prev = dz_prev, dz
        # 1x1 conv 4 -> 64 (no input gradient needed)
        if k.conv_tc:            # BatchNorm + ReLU + weight gradient from one pass over (dz1, z1)
            k.stem_head_bwd(dz, r["z1"], r["s1"], sv["patches"], mode, sv["flag"], sv["ch"], st.p(f"{pe}.0.weight"), st.g(f"{pe}.1.weight"),
                            st.g(f"{pe}.1.bias"), st.g(f"{pe}.0.weight"), P, F, T)
        else:
            k.stem_pw_wgrad(dz, None, sv["patches"], mode, sv["flag"], sv["ch"], st.g(f"{pe}.0.weight").view(CNN_CH, 4), True, P, F, T)
