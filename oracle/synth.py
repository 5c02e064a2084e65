"""Random-init weights and synthetic MANO assets for the DIR architecture.

No checkpoint and no licence-gated MANO pickle is available offline, and a full state_dict
(92.7 M parameters) cannot be shipped, so benchmarks, smoke tests and parity tests all regenerate
the SAME weights from per-tensor seeds (crc32 of the reference's key name) with torch's CPU
generator. The key/shape inventory (state_dict_keys.json) was dumped from the unmodified reference
(`DIR(21, ...).state_dict()`, models/dir.py:486-511) by oracle/gen_golden.py;
tests/golden/weight_fingerprint.json pins a few values so a drift of the generator across torch
versions is detected instead of silently invalidating the committed golden outputs.

Recipe ("calibrated random"): He-scaled convs, non-trivial BN running stats (so BN
folding is exercised), damped residual branches (50 layers of random weights stay O(1)), MANO
regression heads scaled so the predicted hands are hand-sized and their projections land inside
(and sometimes outside) the feature map.

Synthetic MANO: the reference reads (manopth/manopth/manolayer.py:65-108) hands_components (45,45),
hands_mean (45,), betas (10,), shapedirs (778,3,10), posedirs (778,3,135), v_template (778,3),
J_regressor (16,778), weights (778,16), f (1538,3), kintree_table (2,16); we synthesise arrays of the
same shape and role (hand-sized template in metres, convex skinning weights and joint regressor).
"""
import json
import math
import os
import zlib

import numpy as np
import torch


KINTREE_PARENTS = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14]
N_VERTS = 778
N_FACES = 1538


def make_mano_arrays(side: str) -> dict:
    """Return a dict of float32/int numpy arrays with the MANO schema."""
    assert side in ("left", "right")
    seed = 1 if side == "left" else 2
    rng = np.random.RandomState(seed)
    sign = -1.0 if side == "left" else 1.0

    # A crude hand: 16 joint centres (wrist + 5 fingers x 3) in metres, x mirrored per side.
    joints = np.zeros((16, 3), np.float64)
    finger_dirs = np.array([[-0.6, 0.7, 0.3], [-0.25, 1.0, 0.05], [0.0, 1.0, 0.0],
                            [0.25, 0.95, -0.03], [0.5, 0.8, -0.08]])
    finger_dirs /= np.linalg.norm(finger_dirs, axis=1, keepdims=True)
    # MANO joint order: index(1-3), middle(4-6), pinky(7-9), ring(10-12), thumb(13-15)
    order = [1, 2, 4, 3, 0]
    for f in range(5):
        d = finger_dirs[order[f]]
        base = 0.09 if order[f] != 0 else 0.035
        for k in range(3):
            joints[1 + 3 * f + k] = d * (base + 0.028 * k)
    joints[:, 0] *= sign

    # Vertices: each assigned to a "home" joint, scattered around it.
    home = rng.randint(0, 16, size=N_VERTS)
    home[:16] = np.arange(16)  # every joint owns at least one vertex
    v_template = joints[home] + rng.normal(0, 0.008, size=(N_VERTS, 3))

    # Skinning weights: convex, concentrated on home joint and its parent.
    w = rng.uniform(0, 0.05, size=(N_VERTS, 16))
    w[np.arange(N_VERTS), home] += 1.0
    par = np.array([max(p, 0) for p in KINTREE_PARENTS])[home]
    w[np.arange(N_VERTS), par] += rng.uniform(0, 0.6, size=N_VERTS)
    weights = w / w.sum(1, keepdims=True)

    # Joint regressor: convex combination of vertices homed at that joint (+ sparse noise).
    jr = np.zeros((16, N_VERTS))
    for j in range(16):
        idx = np.nonzero(home == j)[0]
        jr[j, idx] = rng.uniform(0.5, 1.0, size=idx.size)
        extra = rng.choice(N_VERTS, 12, replace=False)
        jr[j, extra] += rng.uniform(0, 0.05, size=12)
    jr /= jr.sum(1, keepdims=True)

    shapedirs = rng.normal(0, 0.004, size=(N_VERTS, 3, 10))
    if side == "left":
        # real MANO_LEFT ships shapedirs[:,0,:] un-mirrored ("shapedirs bug");
        # models/dir.py:306-309 flips it iff L~=R. Keep L != R here so that
        # branch is a no-op and both layers are used exactly as constructed.
        pass
    posedirs = rng.normal(0, 0.002, size=(N_VERTS, 3, 135))
    comps = rng.normal(0, 0.35, size=(45, 45))
    hands_mean = rng.normal(0, 0.25, size=(45,))
    betas = np.zeros((10,))
    faces = rng.randint(0, N_VERTS, size=(N_FACES, 3)).astype(np.int64)
    kintree = np.stack([np.array([4294967295] + KINTREE_PARENTS[1:], dtype=np.int64),
                        np.arange(16, dtype=np.int64)])
    return {
        "hands_components": comps.astype(np.float32),
        "hands_mean": hands_mean.astype(np.float32),
        "betas": betas.astype(np.float32),
        "shapedirs": shapedirs.astype(np.float32),
        "posedirs": posedirs.astype(np.float32),
        "v_template": v_template.astype(np.float32),
        "J_regressor": jr.astype(np.float32),
        "weights": weights.astype(np.float32),
        "f": faces,
        "kintree_table": kintree,
    }


def mano_buffers(side: str) -> dict:
    """The registered buffers of manopth ManoLayer (manolayer.py:71-101) as float32 numpy
    arrays, keyed by buffer name (th_*). These are what lives in the reference state_dict."""
    a = make_mano_arrays(side)
    return {
        "th_betas": a["betas"][None, :],
        "th_shapedirs": a["shapedirs"],
        "th_posedirs": a["posedirs"],
        "th_v_template": a["v_template"][None],
        "th_J_regressor": a["J_regressor"],
        "th_weights": a["weights"],
        "th_faces": a["f"],
        "th_hands_mean": a["hands_mean"][None, :],
        "th_comps": a["hands_components"],
        "th_selected_comps": a["hands_components"][:45],
    }


_HERE = os.path.dirname(os.path.abspath(__file__))
KEYS_JSON = os.path.join(os.path.dirname(_HERE), "dir_b200", "state_dict_keys.json")  # shipped with the product module


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
        if name.startswith("backbone") and ".branches." in name and ".bn2." in name:
            w = w * 0.2  # HRNet BasicBlock: damp the residual branch (32 blocks per branch in sequence)
        if name.startswith("backbone") and ".fuse_layers." in name:
            w = w * 0.2  # HRNet fuse layers sum up to four branches per module, eight modules in sequence
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


def hrnet_key_shapes(width: int = 32) -> dict:
    """Key/shape inventory of the HRNet extension (oracle/hrnet_oracle.py; parity unpinned)."""
    from .hrnet_oracle import dir_key_shapes

    return dir_key_shapes(width, load_key_shapes())


def make_state_dict(seed: int = 0, prefix: str = "", key_shapes: dict = None, backbone: str = "resnet50") -> dict:
    """Full (or prefix-filtered) synthetic state_dict with the reference's 963 keys (backbone='hrnet_w32'|'hrnet_w48': the
    HRNet extension's inventory instead)."""
    ks = key_shapes or (load_key_shapes() if backbone == "resnet50" else hrnet_key_shapes(int(backbone.split("_w")[1])))
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
