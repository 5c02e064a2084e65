"""Deterministic synthetic weights for the DIR hot path (TEST INFRASTRUCTURE).

No checkpoint or MANO asset is available offline, and a full state_dict (92.7 M
parameters) cannot be committed, so every machine regenerates the SAME weights
from per-tensor seeds (crc32 of the key name) with torch's CPU generator. The
key/shape inventory is tests/golden/state_dict_keys.json, dumped from the
unmodified reference (`DIR(21, ...).state_dict()`, models/dir.py:486-511) by
oracle/gen_golden.py; tests/golden/weight_fingerprint.json pins a few values so
a drift of the generator across torch versions is detected instead of silently
invalidating the committed golden outputs.

The recipe is "calibrated random" (SURVEY.md 8d): He-scaled convs, non-trivial
BN running stats (so BN folding is exercised), damped residual branches (so 50
layers of random weights stay O(1) — important for the bf16 path), and MANO
regression heads scaled so the predicted hands are hand-sized and their 2D
projections land inside (and sometimes outside) the feature map.
"""
import json
import math
import os
import zlib

import torch

from oracle.synthetic_mano import mano_buffers

_HERE = os.path.dirname(os.path.abspath(__file__))
KEYS_JSON = os.path.join(_HERE, "..", "tests", "golden", "state_dict_keys.json")


def load_key_shapes(path: str = KEYS_JSON) -> dict:
    with open(path) as f:
        return json.load(f)


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) + 7919 * seed) & 0x7FFFFFFF)
    return g


def _normal(shape, std, g, mean=0.0):
    return torch.randn(shape, generator=g, dtype=torch.float32) * std + mean


def _uniform(shape, lo, hi, g):
    return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo


def _mano_head(shape, g, w_std):
    w = _normal(shape, w_std, g)
    return w


def _mano_head_bias(g):
    b = _normal((64,), 0.3, g)
    b[61] = 3.0
    b[62:64] = _uniform((2,), -0.3, 0.3, g)
    return b


def make_tensor(name: str, shape, keys, seed: int = 0) -> torch.Tensor:
    leaf = name.split(".")[-1]
    shape = tuple(shape)
    g = _gen(name, seed)
    if leaf == "num_batches_tracked":
        return torch.zeros((), dtype=torch.long)
    if "mano_layer_" in name:
        side = "left" if "mano_layer_left" in name else "right"
        arr = mano_buffers(side)[leaf]
        t = torch.from_numpy(arr.copy())
        return t.long() if leaf == "th_faces" else t.float()
    if name == "seg_loss.weight":
        return torch.tensor([0.1, 0.45, 0.45])
    if leaf == "img_gird":
        # models/dir.py:66-70: (col+.5, row+.5), row-major over (row, col)
        s = int(round(math.sqrt(shape[0])))
        r = torch.arange(s, dtype=torch.float32) + 0.5
        gx, gy = torch.meshgrid(r, r, indexing="ij")
        return torch.stack((gy, gx), dim=-1).reshape(s * s, 2).contiguous()
    if leaf == "running_mean":
        return _normal(shape, 0.1, g)
    if leaf == "running_var":
        return _uniform(shape, 0.5, 1.5, g)
    if leaf == "W":  # PGraphConv per-joint weights (SemGCN/p_graph_conv.py:19)
        a = 0.8 * math.sqrt(3.0 / shape[2])
        return _uniform(shape, -a, a, g)
    if leaf == "e_0":
        return torch.ones(shape)
    if leaf == "e_1":
        return _normal(shape, 0.5, g, mean=1.0)
    if leaf == "spatial_pos_embed":
        return _normal(shape, 0.1, g)
    # regression heads (models/dir.py:243-245, 323-325)
    if name.endswith(("mano_left.weight", "mano_right.weight")):
        gain = 0.4 if name.startswith("init_regressor") else 0.75
        return _normal(shape, gain / math.sqrt(shape[1]), g)
    if name.endswith(("mano_left.bias", "mano_right.bias")):
        return _mano_head_bias(g)
    if name.endswith("offset.weight"):
        gain = 0.4 if name.startswith("init_regressor") else 0.75
        return _normal(shape, gain / math.sqrt(shape[1]), g)
    if name.endswith("offset.bias"):
        return _normal(shape, 0.3, g)
    if leaf == "weight" and len(shape) == 1:  # BN / LN scale
        w = _uniform(shape, 0.8, 1.2, g)
        if name.startswith("backbone") and ".bn3." in name:
            w = w * 0.3  # damp the residual branch
        return w
    if leaf == "bias":
        return _normal(shape, 0.1, g)
    if leaf == "weight":
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        std = math.sqrt(2.0 / fan_in)
        if ".conv3.conv." in name:  # last conv of hourglass Residual
            std *= 0.3
        if ".attn.qkv." in name or ".attn.proj." in name or ".mlp." in name or ".head." in name:
            std = 1.0 / math.sqrt(fan_in)
        return _normal(shape, std, g)
    raise KeyError(f"no synthetic rule for {name} {shape}")


def make_state_dict(seed: int = 0, prefix: str = "", key_shapes: dict = None) -> dict:
    """Full (or prefix-filtered) synthetic state_dict with the reference's 963 keys."""
    ks = key_shapes or load_key_shapes()
    out = {}
    for name, shape in ks.items():
        if prefix and not name.startswith(prefix):
            continue
        out[name] = make_tensor(name, shape, ks, seed)
    return out


def fingerprint(sd: dict) -> dict:
    """A few pinned values + a global checksum; compared against the committed JSON."""
    names = ["backbone.conv1.weight", "decoder.projecter_3.fusion.0.weight",
             "decoder.projecter_4.gcn_left.gconv_layers.2.gconv.W",
             "init_regressor.mano_left.bias",
             "decoder.projecter_3.regressor.mano_layer_left.th_posedirs"]
    fp = {}
    for n in names:
        t = sd[n].double().flatten()
        fp[n] = [float(t[0]), float(t[t.numel() // 2]), float(t[-1]), float(t.sum())]
    return fp
