"""Import the UNMODIFIED reference (/root/reference) on a CPU-only box (TEST INFRASTRUCTURE).

Only used by oracle/gen_golden.py, which runs in the build container where
/root/reference is mounted. Nothing here is imported by the product, by the
`-m gpu` tests, by smoke() or by bench.py (the reference does not exist on the GPU box).

Shims:
  * utils/utils.py:10      `import imgaug.augmenters`        -> stub module
  * utils/config.py:1      `from yacs.config import CfgNode` -> stub class
  * transformer/mixSTE.py:8 `from timm.models.layers import DropPath, ...` -> stubs
  * manopth/manopth/manolayer.py:65 `ready_arguments(pkl)`  -> synthetic MANO dict
  * models/dir.py:490-491  ImageNet weight download          -> weights=None
  * models/loss.py:9,39 / models/dir.py:514 `.cuda()`        -> identity on CPU
"""
import sys
import types

import numpy as np
import scipy.sparse as sp
import torch

REF_ROOT = "/root/reference"


class _R:
    """Mimics chumpy arrays: the reference reads `.r` (manolayer.py:72-85)."""

    def __init__(self, a):
        self.r = a

    def copy(self):
        return self.r.copy()


def _install_stub_modules():
    if "imgaug" not in sys.modules:
        imgaug = types.ModuleType("imgaug")
        aug = types.ModuleType("imgaug.augmenters")
        for n in ("Sequential", "Sometimes", "MotionBlur", "GaussianBlur", "AdditiveGaussianNoise"):
            setattr(aug, n, lambda *a, **k: None)
        imgaug.augmenters = aug
        sys.modules["imgaug"] = imgaug
        sys.modules["imgaug.augmenters"] = aug
    if "yacs" not in sys.modules:
        yacs = types.ModuleType("yacs")
        cfgm = types.ModuleType("yacs.config")

        class CfgNode(dict):
            def __getattr__(self, k):
                return self[k]

            def __setattr__(self, k, v):
                self[k] = v

            def clone(self):
                return self

            def merge_from_file(self, *_a, **_k):
                pass

            def freeze(self):
                pass

        cfgm.CfgNode = CfgNode
        yacs.config = cfgm
        sys.modules["yacs"] = yacs
        sys.modules["yacs.config"] = cfgm
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")

        class DropPath(torch.nn.Identity):
            def __init__(self, *_a, **_k):
                super().__init__()

        layers.DropPath = DropPath
        layers.to_2tuple = lambda x: (x, x)
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        timm.models = models
        models.layers = layers
        sys.modules["timm"] = timm
        sys.modules["timm.models"] = models
        sys.modules["timm.models.layers"] = layers


def _fake_ready_arguments(path, *_a, **_k):
    from oracle.synth import make_mano_arrays

    side = "left" if "LEFT" in str(path).upper() else "right"
    a = make_mano_arrays(side)
    return {
        "hands_components": a["hands_components"],
        "hands_mean": a["hands_mean"],
        "betas": _R(a["betas"]),
        "shapedirs": _R(a["shapedirs"]),
        "posedirs": _R(a["posedirs"]),
        "v_template": _R(a["v_template"]),
        "J_regressor": sp.csc_matrix(a["J_regressor"]),
        "weights": _R(a["weights"]),
        "f": a["f"].astype(np.uint32),
        "kintree_table": a["kintree_table"],
    }


_loaded = {}


def load_reference():
    """Returns a namespace with the reference's classes (DIR, ManoLayer, STE, ...)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    _install_stub_modules()
    for p in (REF_ROOT, REF_ROOT + "/manopth"):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self

    import manopth.manolayer as ref_manolayer

    ref_manolayer.ready_arguments = _fake_ready_arguments
    import torchvision.models as tvm
    import models.dir as ref_dir

    ref_dir.resnet50 = lambda weights=None, **k: tvm.resnet50(weights=None)
    import SemGCN.p_gcn as ref_pgcn
    import SemGCN.utils as ref_gcn_utils
    import transformer.mixSTE as ref_ste
    import models.backbone.hourglass as ref_hg
    import models.backbone.resnet as ref_resnet
    import utils.utils as ref_utils

    _loaded.update(
        dir=ref_dir, manolayer=ref_manolayer, pgcn=ref_pgcn, gcn_utils=ref_gcn_utils, ste=ref_ste,
        hourglass=ref_hg, resnet=ref_resnet, utils=ref_utils,
    )
    return types.SimpleNamespace(**_loaded)
