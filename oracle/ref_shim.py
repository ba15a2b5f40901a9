"""TEST INFRASTRUCTURE ONLY - import shim for the *real* SAR-SSL reference.

Used only inside the build container (where /root/reference exists) by
`oracle/make_golden.py` to generate the fixtures under tests/golden/ and by
`tests/test_oracle_vs_reference.py` (auto-skipped when the reference is absent,
e.g. on the GPU box).  Nothing in the product package imports this module.

The reference tree is not self-contained: model.py imports
four files that are not in the repository (common.NBC/FNSSL/UNet/CNN, model.py:12-15),
timm (model.py:5), and learner.py pulls torchmetrics / soundfile / matplotlib through
common/utils.py.  None of them is reachable from the default pre-training
architecture, so they are replaced by empty module stubs before the import.
"""
import os
import sys
import types

REF_CANDIDATES = ("/root/reference/code",)


def reference_root():
    for p in REF_CANDIDATES:
        if os.path.isfile(os.path.join(p, "model.py")):
            return p
    return None


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = {}


def load_reference():
    """Return (model_module, learner_module, utils_module_module, common_utils_module)."""
    if _loaded:
        return _loaded["mods"]
    root = reference_root()
    if root is None:
        raise RuntimeError("SAR-SSL reference not present (expected /root/reference/code)")
    import torch

    if "timm" not in sys.modules:
        _stub("timm")
        _stub("timm.models")
        _stub("timm.models.layers", trunc_normal_=torch.nn.init.trunc_normal_)
    sys.path.insert(0, root)
    import common  # noqa: F401  (the real package; its missing submodules are faked below)

    for n, names in (("common.NBC", ["NBC"]), ("common.FNSSL", ["FNblock"]), ("common.UNet", ["UNet"]),
                     ("common.CNN", ["resnet50", "res2net50", "densenet121"])):
        _stub(n, **{k: None for k in names})
    for n in ("soundfile", "matplotlib", "matplotlib.pyplot", "torchmetrics", "torchmetrics.functional",
              "torchmetrics.functional.audio"):
        if n not in sys.modules:
            _stub(n)
    if "torchmetrics.functional.audio.pesq" not in sys.modules:
        _stub("torchmetrics.functional.audio.pesq", perceptual_evaluation_speech_quality=None)
    import model as ref_model
    import learner as ref_learner
    import common.utils_module as ref_ops
    import common.utils as ref_utils

    _loaded["mods"] = (ref_model, ref_learner, ref_ops, ref_utils)
    return _loaded["mods"]
