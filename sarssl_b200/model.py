"""Mirror of the reference's model surface (code/model.py) for the pre-training path: `SARSSL` with the same constructor
signature, `forward(x) -> (loss, diff, data_vis)` contract and `state_dict` keys, running on the fused engine."""
import torch
import torch.nn as nn

from . import _lib, ops
from .engine import Engine
from .modules import PatchMask, PatchRecover, PatchSplit, as_patch_layout
from .params import ParamStore


class _MaskedLossFn(torch.autograd.Function):
    """loss/diff of model.py:721-747 with the gradient produced by the same kernel launch."""

    @staticmethod
    def forward(ctx, pred, patches, frame_flag, ch_idx, nmasked):
        want = pred.requires_grad
        out2, dpred = ops.masked_loss(pred.detach(), patches, frame_flag, ch_idx, nmasked, want_grad=want)
        ctx.dpred, ctx.flag = dpred, frame_flag
        return out2[0].clone(), out2[1].clone()

    @staticmethod
    def backward(ctx, g_loss, g_diff):
        d = ctx.dpred
        nb, nt, w = d.shape
        g = g_loss.reshape(1).float().contiguous()
        _lib.check(_lib.lib().sarssl_scale_masked_rows(_lib.ptr(d), _lib.dtype_code(d), _lib.ptr(ctx.flag), _lib.ptr(g), nb, nt, w // 4,
                                                       _lib.stream_ptr(d.device)), "sarssl_scale_masked_rows")
        return d, None, None, None, None


class MaskedReconLoss(nn.Module):
    """The masking + loss tail of SARSSL.forward (model.py:528,585-592) as a stand-alone module: draws the frame / channel
    masks (PatchMask semantics) and evaluates the masked cross-channel reconstruction loss of a prediction.
    Used directly by the front-end + loss benchmark (BASELINE.json configs[1])."""

    def __init__(self, nmasked_patch, npatch, device, rng_state=None):
        super().__init__()
        self.patch_mask = PatchMask("T", nmasked_patch, [1, npatch], device)
        self.rng_state = rng_state

    def forward(self, pred, x):
        patches = as_patch_layout(x)
        nb, nt, nf = patches.shape[:3]
        pidx, cidx, flag = self.patch_mask.draw(nb, nt, 2, self.rng_state)
        loss, diff = _MaskedLossFn.apply(pred.reshape(nb, nt, nf * 4), patches, flag, cidx, self.patch_mask.nmasked_patch)
        return loss, diff, {"mask_patch_idx": pidx, "mask_ch_idx": cidx}


class _PretrainFn(torch.autograd.Function):
    """One autograd node for the whole model.  `anchor` is a dummy leaf that makes autograd call backward(); parameter
    gradients are accumulated directly into the flat gradient arena (p.grad are views of it), not returned."""

    @staticmethod
    def forward(ctx, anchor, model, patches, flag, ch, want):
        out2, pred, saved = model.engine.forward(patches, flag, ch, model.nmasked_patch, training=model.training, want_grad=want,
                                                 frozen=model.pretrain_frozen_encoder)
        ctx.model, ctx.saved = model, saved
        model._last = {"pred": pred}          # only the prediction (data_vis); the saved activations live in ctx until backward frees them
        return out2[0].clone(), out2[1].clone()

    @staticmethod
    def backward(ctx, g_loss, g_diff):
        if ctx.saved is None:
            raise _lib.SarsslError("backward() through a forward that ran without gradient tracking (eval() or no_grad())")
        ctx.model.store.reattach_grads()
        sync = getattr(ctx.model, "grad_sync", None)
        ctx.model.engine.backward(ctx.saved, gscale=g_loss.reshape(1).float().contiguous(), on_ready=sync.bucket_ready if sync is not None else None)
        ctx.saved = None
        return None, None, None, None, None, None


class _PlainFn(torch.autograd.Function):
    """MCConformer: encoders + decoder on the un-masked input; the caller's loss supplies the gradient of the prediction."""

    @staticmethod
    def forward(ctx, anchor, model, patches, want):
        _, pred, saved = model._engine().forward(patches, None, None, 0, training=model.training, want_grad=want, plain=True)
        ctx.model, ctx.saved = model, saved
        return pred.float() if pred.dtype != torch.float32 else pred

    @staticmethod
    def backward(ctx, g_pred):
        if ctx.saved is None:
            raise _lib.SarsslError("backward() through a forward that ran without gradient tracking (eval() or no_grad())")
        ctx.model.store.reattach_grads()
        ctx.saved["dpred"] = g_pred.to(ctx.model.compute_dtype).contiguous()
        sync = getattr(ctx.model, "grad_sync", None)
        ctx.model.engine.backward(ctx.saved, gscale=None, on_ready=sync.bucket_ready if sync is not None else None)
        ctx.saved = None
        return None, None, None, None


class _DownstreamFn(torch.autograd.Function):
    """One autograd node for the downstream branch: (pred, pooled embedding); the loss is formed outside (Learner.loss)."""

    @staticmethod
    def forward(ctx, anchor, model, patches, want):
        pred, pooled, saved = model._engine().forward_downstream(patches, model.embed_use4ds, training=model.training, want_grad=want,
                                                                 head=model.downstream_head)
        ctx.model, ctx.saved = model, saved
        if pred is pooled:                       # head '': the embedding is the prediction (model.py:705-706); hand out two tensors
            pred = pooled.clone()
        ctx.mark_non_differentiable(pooled)
        return pred, pooled

    @staticmethod
    def backward(ctx, g_pred, g_pooled):
        if ctx.saved is None:
            raise _lib.SarsslError("backward() through a forward that ran without gradient tracking (eval() or no_grad())")
        ctx.model.store.reattach_grads()
        sync = getattr(ctx.model, "grad_sync", None)
        ctx.model.engine.backward_downstream(ctx.saved, g_pred.float().contiguous(), on_ready=sync.bucket_ready if sync is not None else None)
        ctx.saved = None
        return None, None, None, None


class VisDict(dict):
    """data_vis of model.py:595-599 ('mask', 'pred', 'tar').  'pred' / 'tar' are zero-copy views; the dense 'mask' is only
    materialised when somebody reads it (the reference folds all three every step and uses them once per epoch)."""

    def __init__(self, pred, patches, flag, ch):
        nb, nt, nf = patches.shape[:3]
        super().__init__(pred=pred.view(nb, nt, nf, 2, 2).permute(0, 2, 1, 3, 4), tar=patches.permute(0, 2, 1, 3, 4), mask=None)
        self._flag, self._ch, self._shape = flag, ch, (nb, nt, nf, 2)

    def __getitem__(self, key):
        if key == "mask" and dict.__getitem__(self, "mask") is None:
            nb, nt, nf, nmic = self._shape
            m = [torch.empty(self._shape, device=self._flag.device) for _ in range(3)]
            _lib.check(_lib.lib().sarssl_expand_masks(_lib.ptr(self._flag), _lib.ptr(self._ch), _lib.ptr(m[0]), _lib.ptr(m[1]), _lib.ptr(m[2]), nb, nt,
                                                      nf, nmic, _lib.stream_ptr(self._flag.device)), "sarssl_expand_masks")
            dict.__setitem__(self, "mask", m[0].permute(0, 2, 1, 3))
        return dict.__getitem__(self, key)


class SARSSL(nn.Module):
    """model.py:350-719: pre-training branch, frozen-encoder continuation (pretrain_frozen_encoder) and the downstream branch
    (spec/spat = ['cnn', 'conformer'], in_ver 'separate', decoders ['', 'fc'])."""

    def __init__(self, sig_shape=[256, 256, 2, 2], patch_shape=(256, 1), patch_mode="T", nmasked_patch=128 * 1, pretrain=True, use_cls=False,
                 downstream_token="all", downstream_head="mlp", downstream_embed="spec_spat", downstream_dlabel=1, device="cpu",
                 pretrain_frozen_encoder=False, _nmic_pair=0, _factor=1, _tree_prefix=""):
        super().__init__()
        nf, nt, nreim, nmic = sig_shape
        if use_cls:
            raise _lib.SarsslError("sarssl_b200.SARSSL implements pretrain=True, pretrain_frozen_encoder=True and the downstream branch; use_cls is not built")
        pretrain_frozen_encoder = bool(pretrain_frozen_encoder) and not pretrain            # model.py:463,469: `pretrain` wins
        if not pretrain and not pretrain_frozen_encoder and (downstream_head not in ("mlp", "") or downstream_dlabel < 1 or downstream_token != "all" or
                                                             downstream_embed not in ("spec_spat", "spec", "spat")):
            raise _lib.SarsslError("downstream branch: heads 'mlp' (mlp_head for dlabel 1, joint_head for dlabel > 1) / '' (none), token 'all', "
                                   "embed in {spec_spat, spec, spat}; the crnn_* heads of model.py:508-517 are commented out in the reference too")
        if tuple(patch_shape) != (nf, 1) or patch_mode != "T" or nreim != 2 or nmic != 2:
            raise _lib.SarsslError("only frame patches (patch_shape == (nf, 1), patch_mode 'T') of 2-microphone re/im spectrograms are on the hot path")
        npatch = nt
        if nmasked_patch != npatch // 2:                   # model.py:361-364
            nmasked_patch = npatch // 2
        self.pretrain, self.pretrain_frozen_encoder, self.use_cls = pretrain, pretrain_frozen_encoder, use_cls
        self.sig_shape, self.nmasked_patch, self.in_ver = tuple(sig_shape), nmasked_patch, "separate"
        self.device = torch.device(device if str(device) != "cpu" else ("cuda" if torch.cuda.is_available() else "cpu"))
        self.embed_use4ds = downstream_embed
        dembed_ds = {"spec_spat": 768, "spec": 512, "spat": 256}.get(downstream_embed, 768)
        # what the engine runs after the time-mean pooling: 'mlp', 'joint' (dlabel > 1), '' or an integer (SARSSL_MultiCH.head_mch over that many pairs)
        self.downstream_head = _nmic_pair if _nmic_pair else ("joint" if (downstream_head == "mlp" and downstream_dlabel > 1) else downstream_head)
        self.downstream_dlabel = downstream_dlabel
        self.store = ParamStore(self, nf=nf, device=self.device, pretrain=pretrain, dembed_ds=dembed_ds, frozen=pretrain_frozen_encoder,
                                head=downstream_head, nmic_pair=_nmic_pair, factor=_factor, tree_prefix=_tree_prefix, dlabel=downstream_dlabel)
        # the rest are plain attributes (not sub-modules) so that state_dict() holds exactly the reference's 214 entries
        object.__setattr__(self, "patch_split", PatchSplit(patch_shape=patch_shape, f_first=False))
        object.__setattr__(self, "patch_recover", PatchRecover(output_shape=(nf, nt), patch_shape=patch_shape, f_first=False))
        object.__setattr__(self, "patch_mask", PatchMask(patch_mode=patch_mode, nmasked_patch=nmasked_patch, npatch_shape=[1, npatch], device=self.device))
        self.engine = None
        self.compute_dtype = torch.float32
        self.dropout_p = 0.1
        self.rng_state = None            # None: consume python's global `random` stream like the reference; else np.uint32[625]
        self.grad_sync = None
        self.dp = None                   # (rank, world) once data parallel: masks are drawn for the global batch and sliced
        self._anchor = torch.zeros((), requires_grad=True)
        self._last = None

    # ---- device / dtype plumbing
    def _apply(self, fn, recurse=True):
        probe = fn(torch.zeros(1, device=self.store.device))
        if probe.dtype != torch.float32:
            raise _lib.SarsslError("master parameters stay fp32; use set_compute_dtype(torch.bfloat16) for mixed precision")
        self.store.to(probe.device)
        self.device = self.store.device
        self.patch_mask.device = self.device
        self.engine = None
        return self

    def set_compute_dtype(self, dtype):
        self.compute_dtype = dtype
        self.engine = None

    def set_dropout(self, p):
        self.dropout_p = float(p)
        if self.engine is not None:
            self.engine.dropout_p = float(p)

    def _engine(self):
        if self.engine is None:
            if self.device.type != "cuda":
                raise _lib.SarsslError("SARSSL.forward needs the model on a CUDA device (no CPU path)")
            self.engine = Engine(self.store, self.device, self.compute_dtype, self.dropout_p)
        return self.engine

    # ---- forward
    def forward(self, x, _static_masks=None):
        """pretrain=True : x (nb, 2, nf, nt, 2) -> (loss, diff, {'mask': (nb,nf,nt,2), 'pred': (nb,nf,nt,2,2), 'tar': (nb,nf,nt,2,2)})   model.py:519-601
        pretrain=False: x -> (pred (nb, 1), time-mean embedding (nb, dembed))                                                       model.py:667-719
        pretrain_frozen_encoder=True (with pretrain=False): x -> (loss, loss * 0, data_vis): un-masked channel of the masked frames into the spectral
        encoder, `spec_spat_decoder`, gen_loss_spec(tar_maskch=True)                                                               model.py:603-666,749-774"""
        self._engine()
        patches = as_patch_layout(x)
        nb, nt, nf = patches.shape[:3]
        if (nf, nt) != tuple(self.sig_shape[:2]):
            raise _lib.SarsslError(f"input is {nf} bins x {nt} frames but the model was built for {self.sig_shape[:2]}")
        if not self.pretrain and not self.pretrain_frozen_encoder:
            want = torch.is_grad_enabled() and self.training
            return _DownstreamFn.apply(self._anchor, self, patches, want)
        if _static_masks is not None:            # CUDA-graph replay (graph.py): the masks already sit in static device buffers
            flag, cidx = _static_masks
            pidx = None
        else:
            pidx, cidx, flag = self.patch_mask.draw(nb, nt, 2, self.rng_state, dp=self.dp)
        want = torch.is_grad_enabled() and self.training
        loss, diff = _PretrainFn.apply(self._anchor, self, patches, flag, cidx, want)
        vis = VisDict(self._last["pred"], patches, flag, cidx)
        vis.mask_patch_idx, vis.mask_ch_idx = pidx, cidx
        if self.pretrain_frozen_encoder:
            return loss, loss * 0.0, vis                 # model.py:666
        return loss, diff, vis


class SARSSL_MultiCH(SARSSL):
    """model.py:793-821: the single-pair model (spatial encoder embedding, no head) applied to every microphone pair of an item and an MLP head
    over the concatenated pair embeddings.  Same constructor, `forward(x) -> (pred (nb, factor), embed (nb, nmic_pair * 256))` and state_dict
    keys (`model_sch.*`, `head_mch.*`) as the reference; everything runs on the fused engine (one parameter arena, one optimizer launch)."""

    def __init__(self, sig_shape, nmic_pair, task, device):
        factor = nmic_pair if task == "TDOA" else 1
        super().__init__(sig_shape=sig_shape, pretrain=False, device=device, downstream_token="all", downstream_head="", downstream_embed="spat",
                         downstream_dlabel=1, _nmic_pair=int(nmic_pair), _factor=factor, _tree_prefix="model_sch.")
        self.nmic_pair = int(nmic_pair)

    def forward(self, x):
        """x (nb * nmic_pair, 2, nf, nt, 2): the pairs of an item are consecutive (model.py:816-819)."""
        pred, pooled = super().forward(x)
        return pred, pooled.reshape(-1, self.nmic_pair * pooled.shape[-1])


class MCConformer(SARSSL):
    """model.py:824-912 with its defaults (['cnn', 'conformer'] encoders of width 512 / 256, decoder ['', 'fc'], frame patches): spectrogram in,
    reconstructed spectrogram out - `forward(x (nb, 2, nf, nt, 2)) -> data_pred (nb, nf, nt, 2, 2)`; no masking, no loss.  The state_dict
    is the pre-training model's (spec_encoder.*, spat_encoder.*, decoder.*), so pre-trained checkpoints load directly."""

    def __init__(self, sig_shape=[256, 256, 2, 2], patch_shape=(256, 1), spec_model=["cnn", "conformer"], spat_model=["cnn", "conformer"],
                 dembed={"spec": 512, "spat": 256}, dec_model=["", "fc"], device="cpu"):
        if list(spec_model) != ["cnn", "conformer"] or list(spat_model) != ["cnn", "conformer"] or list(dec_model) != ["", "fc"] or \
                dict(dembed) != {"spec": 512, "spat": 256}:
            raise _lib.SarsslError("MCConformer: only the shipped architecture (CNN + Conformer encoders of width 512 / 256, fc decoder) is on the hot path")
        super().__init__(sig_shape=sig_shape, patch_shape=patch_shape, pretrain=True, device=device)
        self.dembed = dict(dembed)

    def forward(self, x):
        self._engine()
        patches = as_patch_layout(x)
        nb, nt, nf = patches.shape[:3]
        if (nf, nt) != tuple(self.sig_shape[:2]):
            raise _lib.SarsslError(f"input is {nf} bins x {nt} frames but the model was built for {self.sig_shape[:2]}")
        want = torch.is_grad_enabled() and self.training
        pred = _PlainFn.apply(self._anchor, self, patches, want)
        return pred.view(nb, nt, nf, 2, 2).permute(0, 2, 1, 3, 4)                 # PatchRecover (model.py:910) as a view
