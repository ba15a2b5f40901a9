"""Parameter storage for the MC-Conformer: a module tree that reproduces the reference's `state_dict` keys exactly
(214 entries, e.g. `spec_encoder.embed.layers.0.sequential.1.module.attention.query_proj.linear.weight`),
backed by flat fp32 arenas.

The tree is only a *namespace*: every nn.Parameter is a view into one flat fp32 buffer (and its .grad a view into a
second, identically laid out buffer), so the optimizer is one fused kernel over the arena, the bf16 compute copy is one
cast, and the data-parallel gradient exchange is a handful of contiguous NCCL buckets.  The arena order is chosen for
the kernels (q/k/v projection weights adjacent so the fused QKV GEMM reads one [3D, D] matrix), the state_dict order is
the reference's."""
import math

import torch
import torch.nn as nn

SPEC_D, SPAT_D, SPEC_LAYERS, SPAT_LAYERS, NHEAD, CNN_CH, DW_K = 512, 256, 1, 3, 4, 64, 31


def state_dict_layout(nf=256, pretrain=True, dembed_ds=SPEC_D + SPAT_D, frozen=False, head="mlp", nmic_pair=0, factor=1, dlabel=1):
    """[(key, shape, kind)] in the reference's state_dict order.  pretrain: one `decoder`; frozen (pretrain_frozen_encoder, model.py:470-481):
    `spec_spat_decoder`, `spec_decoder`, `spat_decoder` (the last one built with the spectral width, as the reference does); otherwise the
    downstream head (`mlp_head`, or nothing for head '').  kind: conv / lin_x (xavier Linear wrapper,
    conformer/modules.py:36-49) / lin_k (default nn.Linear) / bias0 / bias_k / gain / beta / rm / rv / nbt / pe / uv."""
    out = []

    def stem(pre, D):
        for i, shp in ((0, (CNN_CH, 4, 1, 1)), (3, (CNN_CH, CNN_CH, 3, 3)), (6, (CNN_CH, CNN_CH, 3, 3)), (9, (4, CNN_CH, 1, 1))):
            out.append((f"{pre}.{i}.weight", shp, "conv"))
            c = shp[0]
            out.extend([(f"{pre}.{i + 1}.weight", (c,), "gain"), (f"{pre}.{i + 1}.bias", (c,), "beta"),
                        (f"{pre}.{i + 1}.running_mean", (c,), "rm"), (f"{pre}.{i + 1}.running_var", (c,), "rv"),
                        (f"{pre}.{i + 1}.num_batches_tracked", (), "nbt")])
        out.append((f"{pre}.12.weight", (D, 4, nf, 1), "conv"))

    def ffn(pre, D):
        out.extend([(pre + ".0.weight", (D,), "gain"), (pre + ".0.bias", (D,), "beta"),
                    (pre + ".1.linear.weight", (4 * D, D), "lin_x"), (pre + ".1.linear.bias", (4 * D,), "bias0"),
                    (pre + ".4.linear.weight", (D, 4 * D), "lin_x"), (pre + ".4.linear.bias", (D,), "bias0")])

    def block(pre, D):
        s = pre + ".sequential"
        ffn(s + ".0.module.sequential", D)
        m = s + ".1.module"
        out.append((m + ".positional_encoding.pe", (1, 10000, D), "pe"))
        out.extend([(m + ".layer_norm.weight", (D,), "gain"), (m + ".layer_norm.bias", (D,), "beta")])
        a = m + ".attention"
        out.extend([(a + ".u_bias", (NHEAD, D // NHEAD), "uv"), (a + ".v_bias", (NHEAD, D // NHEAD), "uv")])
        for nm in ("query_proj", "key_proj", "value_proj"):
            out.extend([(f"{a}.{nm}.linear.weight", (D, D), "lin_x"), (f"{a}.{nm}.linear.bias", (D,), "bias0")])
        out.append((a + ".pos_proj.linear.weight", (D, D), "lin_x"))
        out.extend([(a + ".out_proj.linear.weight", (D, D), "lin_x"), (a + ".out_proj.linear.bias", (D,), "bias0")])
        c = s + ".2.module.sequential"
        out.extend([(c + ".0.weight", (D,), "gain"), (c + ".0.bias", (D,), "beta"),
                    (c + ".2.conv.weight", (2 * D, D, 1), "conv"), (c + ".2.conv.bias", (2 * D,), "bias_k"),
                    (c + ".4.conv.weight", (D, 1, DW_K), "conv"),
                    (c + ".5.weight", (D,), "gain"), (c + ".5.bias", (D,), "beta"),
                    (c + ".5.running_mean", (D,), "rm"), (c + ".5.running_var", (D,), "rv"), (c + ".5.num_batches_tracked", (), "nbt"),
                    (c + ".7.conv.weight", (D, D, 1), "conv"), (c + ".7.conv.bias", (D,), "bias_k")])
        ffn(s + ".3.module.sequential", D)
        out.extend([(s + ".4.weight", (D,), "gain"), (s + ".4.bias", (D,), "beta")])

    for enc, D, nl in (("spec_encoder", SPEC_D, SPEC_LAYERS), ("spat_encoder", SPAT_D, SPAT_LAYERS)):
        stem(enc + ".patch_embed", D)
        for l in range(nl):
            block(f"{enc}.embed.layers.{l}", D)
    def decoder(name, din):                      # EmbedDecoder(model=['', 'fc'])   model.py:295-301
        dff = 3 * 4 * nf
        out.extend([(name + ".proj.0.weight", (dff, din), "lin_k"), (name + ".proj.0.bias", (dff,), "bias_k"),
                    (name + ".proj.2.weight", (4 * nf, dff), "lin_k"), (name + ".proj.2.bias", (4 * nf,), "bias_k")])

    if pretrain:
        decoder("decoder", SPEC_D + SPAT_D)
    elif frozen:
        decoder("spec_spat_decoder", SPEC_D + SPAT_D)
        decoder("spec_decoder", SPEC_D)
        decoder("spat_decoder", SPEC_D)
    elif head == "mlp" and dlabel == 1:        # downstream head: nn.Sequential(LayerNorm(dembed_ds), Linear(dembed_ds, 1))   model.py:495-500
        out.extend([("mlp_head.0.weight", (dembed_ds,), "gain"), ("mlp_head.0.bias", (dembed_ds,), "beta"),
                    ("mlp_head.1.weight", (1, dembed_ds), "lin_k"), ("mlp_head.1.bias", (1,), "bias_k")])
    elif head == "mlp":                        # joint_head: LayerNorm, Linear(d, d), ReLU, Linear(d, dlabel)               model.py:501-507
        d = dembed_ds
        out.extend([("joint_head.0.weight", (d,), "gain"), ("joint_head.0.bias", (d,), "beta"), ("joint_head.1.weight", (d, d), "lin_k"),
                    ("joint_head.1.bias", (d,), "bias_k"), ("joint_head.3.weight", (dlabel, d), "lin_k"), ("joint_head.3.bias", (dlabel,), "bias_k")])
    if nmic_pair:              # SARSSL_MultiCH.head_mch: LayerNorm, Linear, ReLU, Linear over the concatenated pair embeddings   model.py:807-812
        d = SPAT_D * nmic_pair
        out.extend([("head_mch.0.weight", (d,), "gain"), ("head_mch.0.bias", (d,), "beta"), ("head_mch.1.weight", (d, d), "lin_k"),
                    ("head_mch.1.bias", (d,), "bias_k"), ("head_mch.3.weight", (factor, d), "lin_k"), ("head_mch.3.bias", (factor,), "bias_k")])
    return out


PARAM_KINDS = ("conv", "lin_x", "lin_k", "bias0", "bias_k", "gain", "beta", "uv")


def _fan_in(shape):
    f = 1
    for s in shape[1:]:
        f *= s
    return f


def positional_table(d_model, max_len):
    """Sinusoidal table of conformer/embedding.py:31-38."""
    pe = torch.zeros(max_len, d_model)
    pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(0)


class _Node(nn.Module):
    """Namespace node of the parameter tree (no forward)."""


def _arena_order(layout):
    """Arena order = state_dict order, except that within each attention module the q/k/v weights (then their biases) are
    adjacent so `query_proj.linear.weight` starts a contiguous [3D, D] QKV matrix."""
    keys = [k for k, _, kind in layout if kind in PARAM_KINDS]
    out, done = [], set()
    for k in keys:
        if k in done:
            continue
        if k.endswith("attention.query_proj.linear.weight"):
            a = k[:-len("query_proj.linear.weight")]
            grp = [a + f"{n}_proj.linear.weight" for n in ("query", "key", "value")] + [a + f"{n}_proj.linear.bias" for n in ("query", "key", "value")]
            out.extend(grp)
            done.update(grp)
        else:
            out.append(k)
            done.add(k)
    return out


class ParamStore:
    """Builds the tree under `root`, owns the arenas, and resolves keys to tensors for the engine."""

    def __init__(self, root, nf=256, device="cpu", seed_generator=None, pretrain=True, dembed_ds=SPEC_D + SPAT_D, frozen=False, head="mlp",
                 nmic_pair=0, factor=1, tree_prefix="", dlabel=1):
        """tree_prefix: where the encoder / decoder / mlp_head keys hang in the module tree (SARSSL_MultiCH keeps them under `model_sch.`, so its
        state_dict reads `model_sch.spec_encoder...` like the reference's); the engine-facing keys stay un-prefixed."""
        self.layout = state_dict_layout(nf, pretrain, dembed_ds, frozen, head, nmic_pair, factor, dlabel)
        self.tree_prefix = tree_prefix
        self.shapes = {k: tuple(s) for k, s, _ in self.layout}
        self.kinds = {k: kind for k, _, kind in self.layout}
        self.order = _arena_order(self.layout)
        self.offsets, off = {}, 0
        for k in self.order:
            n = 1
            for s in self.shapes[k]:
                n *= s
            self.offsets[k] = (off, n)
            off += (n + 3) // 4 * 4                   # keep every tensor 16-byte aligned
        self.total = off
        self.root = root
        self.device = torch.device(device)
        self.flat = torch.zeros(self.total, dtype=torch.float32, device=self.device)
        self.grad = torch.zeros(self.total, dtype=torch.float32, device=self.device)
        self.flat_bf16 = None
        self.params, self.buffers = {}, {}
        self._build_tree()
        self.reset_parameters(seed_generator)

    # ---- tree
    def _node(self, path):
        cur = self.root
        for p in path:
            if p not in cur._modules:
                cur.add_module(p, _Node())
            cur = cur._modules[p]
        return cur

    def _view(self, buf, k):
        o, n = self.offsets[k]
        return buf[o:o + n].view(self.shapes[k])

    def _build_tree(self):
        for k, shape, kind in self.layout:
            *path, leaf = ((self.tree_prefix if not k.startswith("head_mch.") else "") + k).split(".")
            node = self._node(path)
            if kind in PARAM_KINDS:
                p = nn.Parameter(self._view(self.flat, k))
                p.grad = self._view(self.grad, k)
                node.register_parameter(leaf, p)
                self.params[k] = p
            else:
                if kind == "pe":
                    b = positional_table(shape[2], shape[1]).to(self.device)
                elif kind == "rv":
                    b = torch.ones(shape, device=self.device)
                elif kind == "nbt":
                    b = torch.zeros(shape, dtype=torch.int64, device=self.device)
                else:
                    b = torch.zeros(shape, device=self.device)
                node.register_buffer(leaf, b)
                self.buffers[k] = (node, leaf)

    def reset_parameters(self, generator=None):
        """Same distributions as the reference's constructors (not the same RNG stream): Conv/Linear default
        kaiming_uniform(a=sqrt 5) = U(+-1/sqrt(fan_in)); the conformer Linear wrapper uses xavier_uniform + zero bias
        (conformer/modules.py:44-47); u/v bias xavier_uniform (attention.py:66-67); norms 1/0."""
        with torch.no_grad():
            for k in self.order:
                kind, shape = self.kinds[k], self.shapes[k]
                t = torch.empty(shape)
                if kind in ("conv", "lin_k"):
                    b = 1.0 / math.sqrt(_fan_in(shape))
                    t.uniform_(-b, b, generator=generator)
                elif kind == "bias_k":
                    wshape = self.shapes[k[:-len("bias")] + "weight"]
                    b = 1.0 / math.sqrt(_fan_in(wshape))
                    t.uniform_(-b, b, generator=generator)
                elif kind in ("lin_x", "uv"):
                    b = math.sqrt(6.0 / (shape[0] + shape[1]))
                    t.uniform_(-b, b, generator=generator)
                elif kind == "gain":
                    t.fill_(1.0)
                else:
                    t.zero_()
                self.params[k].data.copy_(t)

    # ---- device moves keep the arena aliasing intact
    def to(self, device):
        device = torch.device(device)
        if device == self.device:
            return
        self.device = device
        self.flat = self.flat.to(device)
        self.grad = self.grad.to(device)
        self.flat_bf16 = None
        for k, p in self.params.items():
            p.data = self._view(self.flat, k)
            p.grad = self._view(self.grad, k)
        for k, (node, leaf) in self.buffers.items():
            node._buffers[leaf] = node._buffers[leaf].to(device)

    def reattach_grads(self):
        """If an external optimizer set .grad to None (zero_grad(set_to_none=True)), point it back at a zeroed arena view."""
        for k, p in self.params.items():
            g = p.grad
            o, n = self.offsets[k]
            if g is None or g.data_ptr() != self.grad.data_ptr() + 4 * o:
                v = self._view(self.grad, k)
                v.zero_()
                p.grad = v

    # ---- access
    def p(self, k):
        return self.params[k]

    def g(self, k):
        return self._view(self.grad, k)

    def b(self, k):
        node, leaf = self.buffers[k]
        return node._buffers[leaf]

    def arena_view(self, buf, first_key, numel):
        o, _ = self.offsets[first_key]
        return buf[o:o + numel]

    def compute_copy(self, dtype, kernels):
        """Arena in the compute dtype (fp32: the arena itself; bf16: refreshed copy, one cast kernel)."""
        if dtype == torch.float32:
            return self.flat
        if self.flat_bf16 is None or self.flat_bf16.device != self.flat.device:
            self.flat_bf16 = torch.empty(self.total, dtype=torch.bfloat16, device=self.flat.device)
        kernels.cast(self.flat, self.flat_bf16, self.total)
        return self.flat_bf16
