"""Real-asset path: the released checkpoint and the MANO pickles, without chumpy.

Reference behaviour restated here:
  * `apps/eval.py:107-108`, `train.py:137-149`: checkpoints are `torch.save({'net': state_dict, 'optimizer': ...,
    'schedule': ..., 'last_epoch': ...})`; eval loads `['net']` with `strict=False`. A model saved from
    `nn.DataParallel` carries a `module.` prefix on every key.
  * `manopth/mano/webuser/smpl_handpca_wrapper_HAND_only.py:22-67` (`ready_arguments`): unpickle MANO_{LEFT,RIGHT}.pkl
    with `encoding='latin1'`; add `betas = zeros(shapedirs.shape[-1])`, `pose = zeros(3 * kintree.shape[1])`,
    `trans = zeros(3)` if absent; wrap the array members in chumpy. ManoLayer then reads only `.r` of them
    (`manopth/manopth/manolayer.py:65-108`) and registers ten buffers per hand.
  * `models/dir.py:306-309` (`fix_shape`): when left and right shapedirs[:, 0, :] are (nearly) identical, the left ones
    are negated in place ("Fix shapedirs bug of MANO").

The MANO pickles were written by chumpy: several members are `chumpy.ch.Ch` instances whose pickled state is the
instance `__dict__` with the numeric payload under 'x' (chumpy's `Ch.__getstate__`). chumpy is neither installed nor
installable on current numpy, so the unpickler below resolves every `chumpy.*` class to a passive stand-in that just
keeps that state; `scipy.sparse` (J_regressor is a csc_matrix) is a normal dependency. Nothing here touches the GPU.
"""
import io
import os
import pickle

import numpy as np
import torch

MANO_BUFFER_NAMES = ("th_betas", "th_shapedirs", "th_posedirs", "th_v_template", "th_J_regressor", "th_weights",
                     "th_faces", "th_hands_mean", "th_comps", "th_selected_comps")
# every sub-module of DIR that owns a pair of ManoLayers (models/dir.py:221-224, 315-318 via :64 and :501-509)
MANO_OWNERS = ("init_regressor", "decoder.projecter_4.regressor", "decoder.projecter_3.regressor")
NCOMPS = 45  # models/dir.py:222 (use_pca=True, ncomps=45, flat_hand_mean=False)


class _ChStandIn:
    """Receives the pickled state of any chumpy object; `.r` is chumpy's name for 'the value'."""

    def __init__(self, *a, **k):
        self._args = a

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):  # (dict, slots) protocol-2 form
            merged = dict(state[0] or {})
            merged.update(state[1])
            state = merged
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["x"] = state

    @property
    def r(self):
        d = self.__dict__
        if "x" in d:
            return np.asarray(_value(d["x"]))
        if "a" in d and "b" not in d:  # unary wrapper around another Ch
            return np.asarray(_value(d["a"]))
        raise ValueError("chumpy object in the MANO pickle is not a plain array (no 'x' term); "
                         f"found terms {sorted(k for k in d if not k.startswith('_'))}")


def _value(v):
    """numpy value of a pickle member that may be an ndarray, a chumpy stand-in or a scipy sparse matrix."""
    if isinstance(v, _ChStandIn):
        return v.r
    if hasattr(v, "toarray"):
        return np.asarray(v.toarray())
    return np.asarray(v)


# A MANO pickle only needs ndarrays, scipy sparse matrices and (stand-ins for) chumpy objects: every other global is
# refused, so loading a third-party pickle cannot run arbitrary code.
_ALLOWED_GLOBAL_ROOTS = ("numpy", "scipy.sparse")
_ALLOWED_GLOBALS = {("builtins", n) for n in ("dict", "list", "tuple", "set", "frozenset", "int", "float", "complex",
                                              "bool", "str", "bytes", "bytearray", "slice", "range", "object")}
_ALLOWED_GLOBALS |= {("__builtin__", n) for _, n in list(_ALLOWED_GLOBALS)} | {("copy_reg", "_reconstructor"),
                                                                               ("copyreg", "_reconstructor"),
                                                                               ("collections", "OrderedDict"),
                                                                               ("_codecs", "encode")}  # numpy's py2 bytes


class _ManoUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "chumpy" or module.startswith("chumpy."):
            return _ChStandIn
        if (module, name) in _ALLOWED_GLOBALS or any(module == r or module.startswith(r + ".")
                                                     for r in _ALLOWED_GLOBAL_ROOTS):
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"MANO pickle references {module}.{name}, which is not on the allow-list "
                                     "(numpy, scipy.sparse, chumpy stand-in, plain containers)")


def read_mano_pickle(path_or_bytes):
    """`ready_arguments` without chumpy: returns a dict of plain numpy arrays (float64/int as stored)."""
    if isinstance(path_or_bytes, (bytes, bytearray)):
        f = io.BytesIO(path_or_bytes)
    else:
        f = open(path_or_bytes, "rb")
    with f:
        dd = _ManoUnpickler(f, encoding="latin1").load()
    if not isinstance(dd, dict):
        raise ValueError("a MANO pickle holds a dict")
    need = ("hands_components", "hands_mean", "shapedirs", "posedirs", "v_template", "J_regressor", "weights", "f",
            "kintree_table")
    missing = [k for k in need if k not in dd]
    if missing:
        raise KeyError(f"MANO pickle lacks {missing}")
    out = {}
    for k, v in dd.items():
        if isinstance(v, (str, bytes)):
            out[k] = v
            continue
        try:
            out[k] = _value(v)
        except Exception:  # members the layer never reads (e.g. lazily-evaluated chumpy expressions)
            if k in need:
                raise
    nposeparms = out["kintree_table"].shape[1] * 3                       # wrapper :36
    out.setdefault("trans", np.zeros(3))                                  # :38-39
    out.setdefault("pose", np.zeros(nposeparms))                          # :40-41
    if "betas" not in dd:
        out["betas"] = np.zeros(out["shapedirs"].shape[-1])               # :42-43
    return out


def mano_layer_buffers(dd, ncomps=NCOMPS, flat_hand_mean=False):
    """The ten registered buffers of `ManoLayer.__init__` (manolayer.py:71-101) from a `read_mano_pickle` dict."""
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32)))
    comps = np.asarray(dd["hands_components"])
    mean = np.zeros(comps.shape[1]) if flat_hand_mean else np.asarray(dd["hands_mean"])
    return {
        "th_betas": f32(dd["betas"]).unsqueeze(0),
        "th_shapedirs": f32(dd["shapedirs"]),
        "th_posedirs": f32(dd["posedirs"]),
        "th_v_template": f32(dd["v_template"]).unsqueeze(0),
        "th_J_regressor": f32(dd["J_regressor"]),
        "th_weights": f32(dd["weights"]),
        "th_faces": torch.from_numpy(np.asarray(dd["f"]).astype(np.int32)).long(),
        "th_hands_mean": f32(mean).unsqueeze(0),
        "th_comps": f32(comps),
        "th_selected_comps": f32(comps[:ncomps]),
    }


def fix_shape(left, right):
    """models/dir.py:306-309, on the two buffer dicts; returns True when the flip was applied."""
    if torch.sum(torch.abs(left["th_shapedirs"][:, 0, :] - right["th_shapedirs"][:, 0, :])) < 1:
        left["th_shapedirs"] = left["th_shapedirs"].clone()
        left["th_shapedirs"][:, 0, :] *= -1
        return True
    return False


def mano_state_from_dir(mano_root):
    """All `*.mano_layer_{left,right}.th_*` entries of the DIR state_dict (60 keys), read from
    `<mano_root>/MANO_LEFT.pkl` and `MANO_RIGHT.pkl` the way `DIR.__init__` does (three regressors, each with its own
    pair of layers, each pair passed through fix_shape)."""
    paths = {s: os.path.join(mano_root, f"MANO_{s.upper()}.pkl") for s in ("left", "right")}
    for p in paths.values():
        if not os.path.isfile(p):
            raise FileNotFoundError(p)
    left = mano_layer_buffers(read_mano_pickle(paths["left"]))
    right = mano_layer_buffers(read_mano_pickle(paths["right"]))
    fix_shape(left, right)
    state = {}
    for owner in MANO_OWNERS:
        for side, bufs in (("left", left), ("right", right)):
            for k, v in bufs.items():
                state[f"{owner}.mano_layer_{side}.{k}"] = v
    return state


def has_mano_pickles(mano_root):
    return isinstance(mano_root, (str, os.PathLike)) and all(
        os.path.isfile(os.path.join(mano_root, f"MANO_{s}.pkl")) for s in ("LEFT", "RIGHT"))


def read_checkpoint(path_or_obj, expected_keys=None, trusted=False):
    """`torch.load(path, map_location='cpu')['net']` (apps/eval.py:107) made tolerant of the variants in the wild:
    a bare state_dict, a `module.` (DataParallel) prefix. Returns (state_dict, report) where report lists the
    keys of `expected_keys` that are absent and the checkpoint keys that are not expected.
    The file is read with `weights_only=True` (tensors and plain containers only: nothing in it can execute);
    `trusted=True` is the explicit opt-in to the reference's unrestricted `torch.load` for checkpoints that carry
    arbitrary pickled objects (e.g. an optimizer with custom classes)."""
    obj = path_or_obj
    if isinstance(obj, (str, os.PathLike)):
        try:
            obj = torch.load(obj, map_location="cpu", weights_only=True)
        except pickle.UnpicklingError:
            if not trusted:
                raise
            obj = torch.load(obj, map_location="cpu", weights_only=False)
    if isinstance(obj, dict) and "net" in obj and isinstance(obj["net"], dict):
        obj = obj["net"]
    if not isinstance(obj, dict) or not all(isinstance(k, str) for k in obj):
        raise ValueError("not a DIR checkpoint: expected {'net': state_dict, ...} or a state_dict")
    if obj and all(k.startswith("module.") for k in obj):
        obj = {k[len("module."):]: v for k, v in obj.items()}
    report = {"missing": [], "unexpected": []}
    if expected_keys is not None:
        exp = set(expected_keys)
        report["missing"] = sorted(exp - set(obj))
        report["unexpected"] = sorted(set(obj) - exp)
    return obj, report
