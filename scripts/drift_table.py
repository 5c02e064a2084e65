"""Per-seam drift table (VERDICT r1 "next" 1a): where do the fp32 and bf16 configurations leave the reference?

Ground truth = the oracle restatement evaluated in float64 on the host (the fp32 oracle itself sits 2-3e-5 away from
it: that is the noise floor of any fp32 implementation, printed as the first rows). For each precision the seams of
SURVEY 8(b-2) are fed with the ORACLE's (fp64->fp32) inputs, so every row isolates one component; the last rows are
the whole forward (errors compound). Also measures what the reference does by itself on this GPU: the same oracle
ops on CUDA in PyTorch eager with TF32 allowed (torch's cudnn default) and disallowed, and under bf16 autocast.

Run on the GPU box:  python scripts/drift_table.py [--batch 4] > gpurun_out/drift_table.txt
Test infrastructure (imports oracle/); not part of the product path.
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import dir_b200  # noqa: E402
from dir_b200 import seams  # noqa: E402
from oracle import dir_oracle as O  # noqa: E402
from oracle.synth import make_state_dict  # noqa: E402

MESH = ("pd_mesh_xyz_left", "pd_mesh_xyz_right")


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def mesh_mm(out, ref):
    d = torch.cat([(out[k].double().cpu() - ref[k].double().cpu()).norm(dim=-1).flatten() for k in MESH]) * 1000
    return float(d.mean()), float(d.max())


def staged_forward(sd, img):
    """oracle forward with every seam input/output kept (same op sequence as O.decoder / models/dir.py:437-483)."""
    with torch.no_grad():
        t = {}
        feats = O.resnet50(sd, img)
        t["c"] = feats
        init_out = O.init_regressor(sd, feats[-1])
        t["init"] = init_out
        p = "decoder."
        _, c2, c3, c4 = feats
        t["skip4"] = O.residual(sd, p + "skip_layer4.", c3)
        t["cat4"] = torch.cat((O.upsample2x(c4), t["skip4"]), 1)
        t["fusion4"] = O.residual(sd, p + "fusion_layer4.", t["cat4"])
        t["res1"], t["f1"] = O.joint2bone(sd, p + "projecter_4.", t["fusion4"], init_out, 16, 1)
        t["enh4_in"] = torch.cat((t["fusion4"], t["f1"]["img_feat"]), 1)
        t["enh4"] = O.residual(sd, p + "enhance_layer4.", t["enh4_in"])
        t["skip3"] = O.residual(sd, p + "skip_layer3.", c2)
        t["cat3"] = torch.cat((O.upsample2x(t["enh4"]), t["skip3"]), 1)
        t["fusion3"] = O.residual(sd, p + "fusion_layer3.", t["cat3"])
        t["res2"], t["f2"] = O.joint2bone(sd, p + "projecter_3.", t["fusion3"], t["res1"], 32, 2)
        return t


def cast(tree, dtype=None, device=None):
    if isinstance(tree, torch.Tensor):
        if tree.is_floating_point():
            return tree.to(dtype=dtype or tree.dtype, device=device or tree.device)
        return tree.to(device=device or tree.device)
    if isinstance(tree, dict):
        return {k: cast(v, dtype, device) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return [cast(v, dtype, device) for v in tree]
    return tree


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    args = ap.parse_args()
    B = args.batch
    torch.set_num_threads(os.cpu_count() or 1)
    sd = make_state_dict(0)
    img = torch.randn(B, 3, 256, 256, generator=torch.Generator().manual_seed(0))
    sd64 = cast(sd, torch.float64)
    T = staged_forward(sd64, img.double())  # ground truth
    T32 = cast(T, torch.float32)
    dev = torch.device("cuda", 0)

    rows = []

    def add(cfg, seam, what, r, mm=None):
        rows.append((cfg, seam, what, r, mm))
        mms = "" if mm is None else f"  mesh drift mean {mm[0]:.4f} mm max {mm[1]:.3f} mm"
        print(f"{cfg:22s} {seam:34s} {what:26s} rel {r:.3e}{mms}", flush=True)

    def whole(cfg, outs_fn):
        for i, key in enumerate(("init", "res1", "res2")):
            o = outs_fn(i)
            worst = max(rel(o[k], T[key][k]) for k in O.OUT_KEYS)
            add(cfg, f"whole forward stage {i}", "worst of 9 tensors", worst, mesh_mm(o, T[key]))

    # ---- the reference's own arithmetic: fp32 on the host, eager on this GPU
    ref32 = O.dir_forward(sd, img)
    whole("oracle fp32 host", lambda i: ref32[i])
    sd_gpu = cast(sd, device=dev)
    for name, tf32, autocast in (("eager cuda fp32", False, False), ("eager cuda tf32 (default)", True, False),
                                 ("eager cuda bf16 autocast", True, True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False  # torch default
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            o = O.dir_forward(sd_gpu, img.to(dev))
        whole(name, lambda i: {k: v.float() for k, v in o[i].items() if v is not None})
    torch.backends.cudnn.allow_tf32 = True
    with torch.autocast("cpu", dtype=torch.bfloat16):
        o = O.dir_forward(sd, img)
    whole("oracle host bf16 autocast", lambda i: {k: v.float() for k, v in o[i].items() if v is not None})

    # ---- our seams, each fed with ground-truth inputs
    for precision in ("fp32", "tf32", "bf16"):
        net = dir_b200.DIR(21, "./misc/mano", precision=precision).to(dev)
        net.load_state_dict(sd, strict=False)
        net.eval()
        cfg = f"dirb200 {precision}"
        cs = seams.backbone(net, img.to(dev))
        for i in range(4):
            add(cfg, "ResNet.forward", f"c{i + 1}", rel(cs[i], T["c"][i]))
        o = seams.init_regressor(net, T32["c"][3].to(dev))
        add(cfg, "InitRegressor.forward", "worst of 9 (+para)",
            max(rel(o[k], T["init"][k]) for k in O.OUT_KEYS + ["pd_mano_para_left", "pd_mano_para_right"]),
            mesh_mm(o, T["init"]))
        for name, xin, want in (("skip_layer4", T32["c"][2], T["skip4"]), ("fusion_layer4", T32["cat4"], T["fusion4"]),
                                ("enhance_layer4", T32["enh4_in"], T["enh4"]), ("skip_layer3", T32["c"][1], T["skip3"]),
                                ("fusion_layer3", T32["cat3"], T["fusion3"])):
            y = seams.residual(net, f"decoder.{name}.", xin.to(dev))
            add(cfg, "Residual.forward", name, rel(y, want))
        for stage, feat, prev, want, wf in ((1, T32["fusion4"], T32["init"], T["res1"], T["f1"]),
                                           (2, T32["fusion3"], T32["res1"], T["res2"], T["f2"])):
            res, feats = seams.joint2bone(net, stage, feat.to(dev), cast(prev, device=dev))
            add(cfg, f"Joint2BoneFeature.forward stage {stage}", "worst of 9 (+para)",
                max(rel(res[k], want[k]) for k in O.OUT_KEYS + ["pd_mano_para_left", "pd_mano_para_right"]),
                mesh_mm(res, want))
            add(cfg, f"Joint2BoneFeature.forward stage {stage}", "img_feat", rel(feats["img_feat"], wf["img_feat"]))
            add(cfg, f"Joint2BoneFeature.forward stage {stage}", "joint_feat",
                max(rel(feats["joint_feat_left"], wf["joint_feat_left"]),
                    rel(feats["joint_feat_right"], wf["joint_feat_right"])))
        outs, _ = net({"img": img.to(dev)}, None, None)
        torch.cuda.synchronize()
        whole(cfg, lambda i: outs[i])
        del net


if __name__ == "__main__":
    main()
