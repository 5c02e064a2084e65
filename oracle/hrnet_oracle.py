"""Self-authored oracle for the HRNet extension (BASELINE.json configs 3-5).  PARITY UNPINNED.

TEST INFRASTRUCTURE ONLY (same import rules as dir_oracle.py).

The reference contains NO HRNet (`models/backbone/` holds resnet.py and hourglass.py only; `cfg.backbone`
is used in a log-file name). What the reference does define are the knobs an HRNet variant would turn:
`FusionJointInterIterDecoder(inDim=[2048,1024,512,256])` (models/dir.py:390) and `InitRegressor(feat_dim)`
(models/dir.py:219,501). This file therefore restates, functionally and in torch fp32,
  * the HRNet-W{32,48} backbone as published (Sun et al., "Deep High-Resolution Representation Learning", CVPR 2019;
    structure and state_dict names of the authors' `pose_hrnet.py` / `cls_hrnet.py`: stem of two stride-2 3x3 convs,
    `layer1` = 4 Bottlenecks, three stages of HighResolutionModules with 2/3/4 branches of 4 BasicBlocks each,
    nearest-neighbour upsampling in the fuse layers, all four branches returned), and
  * DIR assembled around it exactly as models/dir.py assembles it around ResNet-50: c2..c4 = branches 1..3
    (W32: 64@32x32, 128@16x16, 256@8x8), decoder `inDim=[8w,4w,2w,w]`, `InitRegressor(feat_dim=8w)`; every other module
    is the reference's own (dir_oracle.py).
Nothing pins this to an external implementation: no HRNet code, weights or outputs exist offline. The CUDA path is
tested against THIS file; the claim is internal consistency, not parity with a reference.
"""
import torch
import torch.nn.functional as F

from . import dir_oracle as O

STAGES = ((2, 1, 2), (3, 4, 3), (4, 3, 4))  # (stage index, modules, branches)


def branch_channels(width):
    return [width, 2 * width, 4 * width, 8 * width]


def backbone_key_shapes(width, p="backbone."):
    """state_dict inventory of the backbone (names follow the HRNet authors' modules)."""
    keys = {}

    def conv(name, cout, cin, k):
        keys[p + name + ".weight"] = [cout, cin, k, k]

    def bn(name, c):
        for leaf in ("weight", "bias", "running_mean", "running_var"):
            keys[f"{p}{name}.{leaf}"] = [c]
        keys[f"{p}{name}.num_batches_tracked"] = []

    C = branch_channels(width)
    conv("conv1", 64, 3, 3), bn("bn1", 64), conv("conv2", 64, 64, 3), bn("bn2", 64)
    inpl = 64
    for b in range(4):
        q = f"layer1.{b}."
        conv(q + "conv1", 64, inpl, 1), bn(q + "bn1", 64), conv(q + "conv2", 64, 64, 3), bn(q + "bn2", 64)
        conv(q + "conv3", 256, 64, 1), bn(q + "bn3", 256)
        if b == 0:
            conv(q + "downsample.0", 256, 64, 1), bn(q + "downsample.1", 256)
        inpl = 256
    conv("transition1.0.0", C[0], 256, 3), bn("transition1.0.1", C[0])
    conv("transition1.1.0.0", C[1], 256, 3), bn("transition1.1.0.1", C[1])
    for stage, nmod, nbr in STAGES:
        if stage > 2:
            conv(f"transition{stage - 1}.{nbr - 1}.0.0", C[nbr - 1], C[nbr - 2], 3)
            bn(f"transition{stage - 1}.{nbr - 1}.0.1", C[nbr - 1])
        for m in range(nmod):
            for br in range(nbr):
                for k in range(4):
                    q = f"stage{stage}.{m}.branches.{br}.{k}."
                    conv(q + "conv1", C[br], C[br], 3), bn(q + "bn1", C[br])
                    conv(q + "conv2", C[br], C[br], 3), bn(q + "bn2", C[br])
            for i in range(nbr):
                for j in range(nbr):
                    q = f"stage{stage}.{m}.fuse_layers.{i}.{j}."
                    if j > i:
                        conv(q + "0", C[i], C[j], 1), bn(q + "1", C[i])
                    elif j < i:
                        for k in range(i - j):
                            co = C[i] if k == i - j - 1 else C[j]
                            conv(q + f"{k}.0", co, C[j], 3), bn(q + f"{k}.1", co)
    return keys


def dir_key_shapes(width, resnet_keys):
    """Inventory of DIR with an HRNet backbone: the reference's 963 keys with `backbone.*` replaced and the shapes that
    depend on `inDim` / `feat_dim` (models/dir.py:219-245,390-397) re-derived."""
    C = branch_channels(width)
    feat = C[3]
    out = {k: list(v) for k, v in resnet_keys.items() if not k.startswith("backbone.")}
    out.update(backbone_key_shapes(width))
    for side in ("left", "right"):
        a = f"init_regressor.attention_{side}."
        out[a + "0.weight"], out[a + "0.bias"] = [feat // 2, feat, 3, 3], [feat // 2]
        for leaf in ("weight", "bias", "running_mean", "running_var"):
            out[a + "1." + leaf] = [feat // 2]
        out[a + "3.weight"] = [1, feat // 2, 1, 1]
        out[f"init_regressor.mano_{side}.weight"] = [64, feat]
    out["init_regressor.offset.weight"] = [3, feat]

    def residual(p, cin):  # Residual(cin, 256): models/backbone/hourglass.py:33-53
        for leaf in ("weight", "bias", "running_mean", "running_var"):
            out[f"{p}bn1.{leaf}"] = [cin]
        out[p + "conv1.conv.weight"] = [128, cin, 1, 1]
        out[p + "skip_layer.conv.weight"] = [256, cin, 1, 1]

    residual("decoder.skip_layer4.", C[2])
    residual("decoder.fusion_layer4.", C[3] + 256)
    residual("decoder.skip_layer3.", C[1])
    return out


# --------------------------------------------------------------------------- forward
def _cbr(sd, cp, bp, x, stride=1, pad=1, relu=True):
    y = O.bn2d(sd, bp, F.conv2d(x, sd[cp + "weight"], None, stride=stride, padding=pad))
    return F.relu(y) if relu else y


def _basic_block(sd, q, x):
    y = _cbr(sd, q + "conv1.", q + "bn1.", x)
    y = _cbr(sd, q + "conv2.", q + "bn2.", y, relu=False)
    return F.relu(y + x)


def hrnet(sd, x, width, p="backbone."):
    """-> the four branch maps [w@H/4, 2w@H/8, 4w@H/16, 8w@H/32]."""
    C = branch_channels(width)
    x = _cbr(sd, p + "conv1.", p + "bn1.", x, stride=2)
    x = _cbr(sd, p + "conv2.", p + "bn2.", x, stride=2)
    for b in range(4):
        x = O.bottleneck(sd, f"{p}layer1.{b}.", x, 1)
    xs = [_cbr(sd, p + "transition1.0.0.", p + "transition1.0.1.", x),
          _cbr(sd, p + "transition1.1.0.0.", p + "transition1.1.0.1.", x, stride=2)]
    for stage, nmod, nbr in STAGES:
        if stage > 2:  # the new branch is made from the LAST branch of the previous stage
            t = f"{p}transition{stage - 1}.{nbr - 1}.0."
            xs.append(_cbr(sd, t + "0.", t + "1.", xs[-1], stride=2))
        for m in range(nmod):
            q = f"{p}stage{stage}.{m}."
            for br in range(nbr):
                for k in range(4):
                    xs[br] = _basic_block(sd, f"{q}branches.{br}.{k}.", xs[br])
            fused = []
            for i in range(nbr):
                y = None
                for j in range(nbr):
                    f = f"{q}fuse_layers.{i}.{j}."
                    if j == i:
                        t = xs[j]
                    elif j > i:
                        t = _cbr(sd, f + "0.", f + "1.", xs[j], pad=0, relu=False)
                        t = F.interpolate(t, scale_factor=2 ** (j - i), mode="nearest")
                    else:
                        t = xs[j]
                        for k in range(i - j):
                            t = _cbr(sd, f"{f}{k}.0.", f"{f}{k}.1.", t, stride=2, relu=k != i - j - 1)
                    y = t if y is None else y + t
                fused.append(F.relu(y))
            xs = fused
    assert [t.shape[1] for t in xs] == C
    return xs


def dir_forward(sd, img, width=32):
    """DIR.forward (models/dir.py:513-540, eval branch) with the HRNet backbone in place of ResNet-50."""
    with torch.no_grad():
        feats = hrnet(sd, img, width)
        init_out = O.init_regressor(sd, feats[-1])
        dec = O.decoder(sd, feats, init_out)
        outs = []
        for o in [init_out] + dec["result_list"]:
            d = {k: o[k] for k in O.OUT_KEYS}
            d["pd_rel_joint"] = None
            outs.append(d)
        outs.append({"dense": dec["dense"], "seg": dec["seg"], "proj_feat": dec["proj_feat"]})
        return outs


if __name__ == "__main__":  # python -m oracle.hrnet_oracle  -> regenerates dir_b200/state_dict_keys_hrnet_w{32,48}.json
    import json
    import os

    from .synth import load_key_shapes

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for width in (32, 48):
        path = os.path.join(root, "dir_b200", f"state_dict_keys_hrnet_w{width}.json")
        with open(path, "w") as f:
            json.dump(dir_key_shapes(width, load_key_shapes()), f, indent=0)
        print("wrote", path)
