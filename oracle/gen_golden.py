"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run ONLY in the build container (needs /root/reference):   python -m oracle.gen_golden
It (1) dumps the reference's state_dict key/shape inventory, (2) loads the synthetic
weights of oracle/synth.py into the reference modules with strict=True,
(3) executes the reference on seeded inputs at every seam of SURVEY.md 8(b-2) and for the
whole forward, and (4) stores inputs-by-seed + reference outputs as small fixtures.
It also prints the oracle-vs-reference deviation so the restatement is checked at
generation time; tests/test_oracle_golden.py re-checks it against the stored fixtures.
"""
import json
import os
import warnings

import numpy as np
import torch

from oracle import dir_oracle as O
from oracle.ref_shims import load_reference
from oracle.synth import fingerprint, make_state_dict

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
warnings.filterwarnings("ignore")


def rnd(seed, *shape, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def urnd(seed, lo, hi, *shape):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g) * (hi - lo) + lo


def sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def save(name, **arrs):
    np.savez_compressed(os.path.join(GOLD, name), **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v))
                                                     for k, v in arrs.items()})


def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def seam_inputs():
    """Seeded seam inputs, shared with tests (imported from there too)."""
    return {
        "mano_pose": rnd(11, 4, 51, scale=0.5), "mano_beta": rnd(12, 4, 10),
        "ste_x": rnd(13, 2, 42, 128), "gcn_x": rnd(14, 2, 21, 128),
        "i2j_feat": rnd(15, 2, 256, 16, 16), "i2j_uv": urnd(16, -1.3, 1.3, 2, 21, 2),
        "bp_uv16": urnd(17, -0.9, 0.9, 2, 21, 2), "bp_feat": rnd(18, 2, 21, 64),
        "bp_uv32": urnd(19, -1.1, 1.1, 1, 21, 2),
        "res_x": rnd(20, 1, 512, 8, 8), "c4": torch.relu(rnd(21, 2, 2048, 8, 8)),
        "bb_img": rnd(22, 1, 3, 64, 64),
        "j2b_feat": rnd(23, 2, 256, 16, 16), "j2b_xyz_l": rnd(24, 2, 21, 3, scale=0.05),
        "j2b_xyz_r": rnd(25, 2, 21, 3, scale=0.05), "j2b_uv_l": urnd(26, -0.8, 0.8, 2, 21, 2),
        "j2b_uv_r": urnd(27, -0.8, 0.8, 2, 21, 2), "j2b_para_l": rnd(28, 2, 64, scale=0.3),
        "j2b_para_r": rnd(29, 2, 64, scale=0.3), "j2b_off": rnd(30, 2, 3, scale=0.5),
        "img": rnd(0, 2, 3, 256, 256),
    }


def eval_metric_inputs():
    """Seeded inputs of the eval-metric seam (shared with the tests)."""
    B = 3
    gv = {s: rnd(40 + i, B, 778, 3, scale=0.05) + torch.tensor([0.0, 0.0, 0.6]) for i, s in enumerate(("left", "right"))}
    pv = {s: gv[s] * 0.9 + rnd(42 + i, B, 778, 3, scale=0.004) - torch.tensor([0.01, 0.0, 0.55])
          for i, s in enumerate(("left", "right"))}
    cam = torch.tensor([[1400.0, 0, 128.0], [0, 1400.0, 128.0], [0, 0, 1.0]]).repeat(B, 1, 1)
    gv2d = {s: rnd(44 + i, B, 778, 2, scale=30.0) + 128 for i, s in enumerate(("left", "right"))}
    off = rnd(46, B, 3, scale=0.5)
    jreg16 = {s: torch.from_numpy(__import__("oracle.synth", fromlist=["mano_buffers"]).mano_buffers(s)["th_J_regressor"])
              for s in ("left", "right")}
    return {"gt_verts": gv, "pred_verts": pv, "cam": cam, "gt_verts2d": gv2d, "pred_offset": off, "jreg16": jreg16}


def run_reference_eval_lines(E):
    """Execute apps/eval.py's own metric code (class Jr :22-44, xyz2uvd :80-83, loop body :140-241) on E."""
    import textwrap
    import types

    src = open("/root/reference/apps/eval.py").read().replace("\r", "").split("\n")

    def find(prefix, start=0):
        return next(i for i in range(start, len(src)) if src[i].startswith(prefix))

    ns = {"torch": torch, "np": np}
    exec("\n".join(src[find("class Jr"):find("class handDataset")]), ns)  # apps/eval.py:22-44
    i0 = find("def xyz2uvd")
    exec("\n".join(src[i0:i0 + 4]), ns)  # apps/eval.py:80-83
    B = E["cam"].shape[0]
    ns["J_regressor"] = {s: ns["Jr"](E["jreg16"][s], device="cpu") for s in ("left", "right")}
    ns["opt"] = types.SimpleNamespace(root_joint=0, scale=True)
    ns["stage_num"] = 3
    result = [None, None, {"pd_offset": E["pred_offset"], "pd_mesh_xyz_left": E["pred_verts"]["left"],
                           "pd_mesh_xyz_right": E["pred_verts"]["right"]}]
    ns["network"] = lambda inp, a, b: (result, None)
    z = torch.zeros(B, 1)
    ns["data"] = [z, z, z, E["gt_verts"]["left"], z, E["gt_verts"]["right"], z, E["gt_verts2d"]["left"], z,
                  E["gt_verts2d"]["right"], E["cam"]]
    for n in ("joints_loss", "verts_loss", "joints_xyz_list", "joints_xyz_gt_list", "joints_2d_loss", "verts_2d_loss"):
        ns[n] = {"left": [], "right": []}
    ns["root_loss_list"] = []
    ns["idx"] = 0
    b0 = find("        for data in tqdm(dataloader):") + 1
    b1 = find("    joints_loss['left'] = np.concatenate")
    body = textwrap.dedent("\n".join(src[b0:b1]))  # the loop body, apps/eval.py:140-241
    exec(body, ns)
    return {"joint_left": ns["joints_loss"]["left"][0], "joint_right": ns["joints_loss"]["right"][0],
            "vert_left": ns["verts_loss"]["left"][0], "vert_right": ns["verts_loss"]["right"][0],
            "joint2d_left": ns["joints_2d_loss"]["left"][0], "joint2d_right": ns["joints_2d_loss"]["right"][0],
            "vert2d_left": ns["verts_2d_loss"]["left"][0], "vert2d_right": ns["verts_2d_loss"]["right"][0],
            "root": ns["root_loss_list"][0].reshape(-1)}


def main():
    ref = load_reference()
    E = eval_metric_inputs()
    gold = run_reference_eval_lines(E)
    jr = {s: O.eval_jregressor(E["jreg16"][s]) for s in ("left", "right")}
    om = O.eval_metrics(E["pred_verts"], E["pred_offset"], E["gt_verts"], E["gt_verts2d"], E["cam"], jr)
    print("eval metric:", " ".join(f"{k} {rel(om[k], torch.as_tensor(v)):.1e}" for k, v in gold.items()))
    save("eval_metric.npz", **gold)
    net = ref.dir.DIR(21, "./misc/mano")
    net.eval()
    shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
    with open(os.path.join(GOLD, "state_dict_keys.json"), "w") as f:
        json.dump(shapes, f, indent=0)
    sd = make_state_dict(0, key_shapes=shapes)
    print(net.load_state_dict(sd, strict=True))
    with open(os.path.join(GOLD, "weight_fingerprint.json"), "w") as f:
        json.dump(fingerprint(sd), f, indent=1)
    X = seam_inputs()
    p4 = net.decoder.projecter_4
    P4 = "decoder.projecter_4."
    with torch.no_grad():
        # --- MANO layer (both sides)
        for side in ("left", "right"):
            layer = getattr(net.init_regressor, f"mano_layer_{side}")
            v, j = layer(X["mano_pose"], X["mano_beta"])
            ov, oj = O.mano_layer(sd, f"init_regressor.mano_layer_{side}.", X["mano_pose"], X["mano_beta"], side)
            print(f"mano {side}: verts {rel(ov, v):.2e} joints {rel(oj, j):.2e}")
            save(f"mano_{side}.npz", verts=v, joints=j)
        # --- STE
        y = p4.interaction(X["ste_x"].clone())
        print(f"ste: {rel(O.ste(sd, P4 + 'interaction.', X['ste_x']), y):.2e}")
        save("ste.npz", y=y)
        # --- GCN stack
        y = p4.gcn_left(X["gcn_x"])
        print(f"gcn: {rel(O.gcn_stack(sd, P4 + 'gcn_left.', X['gcn_x']), y):.2e}")
        save("gcn.npz", y=y)
        # --- ImgFeature2JointFeature
        y = p4.img2joint_left(X["i2j_feat"], X["i2j_uv"]).reshape(2, -1, 21).permute(0, 2, 1)
        print(f"img2joint: {rel(O.img2joint(sd, P4 + 'img2joint_left.', X['i2j_feat'], X['i2j_uv']), y):.2e}")
        save("img2joint.npz", y=y)
        # --- bone_proj
        y16 = p4.bone_proj(X["bp_uv16"], X["bp_feat"])
        y32 = net.decoder.projecter_3.bone_proj(X["bp_uv32"], X["bp_feat"][:1])
        o16, o32 = O.bone_proj(X["bp_uv16"], X["bp_feat"], 16, 1), O.bone_proj(X["bp_uv32"], X["bp_feat"][:1], 32, 2)
        print(f"bone_proj: {rel(o16, y16):.2e} {rel(o32, y32):.2e} nz {float((y16 != 0).float().mean()):.3f} "
              f"{float((y32 != 0).float().mean()):.3f} mask-eq {bool(((o16 != 0) == (y16 != 0)).all())}")
        save("bone_proj.npz", y16=y16, y32=y32)
        # --- Residual
        y = net.decoder.enhance_layer4(X["res_x"])
        print(f"residual: {rel(O.residual(sd, 'decoder.enhance_layer4.', X['res_x']), y):.2e}")
        save("residual.npz", y=y)
        # --- InitRegressor
        r = net.init_regressor(X["c4"])
        o = O.init_regressor(sd, X["c4"])
        keys = ["pd_offset", "pd_mano_para_left", "pd_mano_para_right", "pd_joint_uv_left", "pd_joint_uv_right",
                "pd_mesh_xyz_left", "pd_mesh_xyz_right", "pd_joint_xyz_left", "pd_joint_xyz_right",
                "pd_mesh_uv_left", "pd_mesh_uv_right"]
        print("init_regressor:", " ".join(f"{rel(o[k], r[k]):.1e}" for k in keys))
        save("init_regressor.npz", **{k: r[k] for k in keys})
        # --- backbone on a 64x64 image
        fs = net.backbone(X["bb_img"])
        of = O.resnet50(sd, X["bb_img"])
        print("resnet50:", " ".join(f"{rel(a, b):.1e}" for a, b in zip(of, fs)))
        save("resnet50.npz", c1=fs[0], c2=fs[1], c3=fs[2], c4=fs[3])
        # --- one refinement stage
        res, feats = p4(X["j2b_feat"], X["j2b_xyz_l"], X["j2b_xyz_r"], X["j2b_uv_l"], X["j2b_uv_r"],
                        X["j2b_para_l"], X["j2b_para_r"], X["j2b_off"].unsqueeze(1))
        prev = {"pd_joint_xyz_left": X["j2b_xyz_l"], "pd_joint_xyz_right": X["j2b_xyz_r"],
                "pd_joint_uv_left": X["j2b_uv_l"], "pd_joint_uv_right": X["j2b_uv_r"],
                "pd_mano_para_left": X["j2b_para_l"], "pd_mano_para_right": X["j2b_para_r"], "pd_offset": X["j2b_off"]}
        ores, ofeats = O.joint2bone(sd, P4, X["j2b_feat"], prev, 16, 1)
        jk = ["pd_offset", "pd_mano_para_left", "pd_mano_para_right", "pd_joint_uv_left", "pd_joint_uv_right",
              "pd_mesh_xyz_left", "pd_mesh_xyz_right", "pd_joint_xyz_left", "pd_joint_xyz_right"]
        print("joint2bone:", " ".join(f"{rel(ores[k], res[k]):.1e}" for k in jk),
              f"img_feat {rel(ofeats['img_feat'], feats['img_feat']):.1e}",
              f"jf {rel(ofeats['joint_feat_left'], feats['joint_feat_left']):.1e}")
        save("joint2bone.npz", img_feat=feats["img_feat"], joint_feat_left=feats["joint_feat_left"],
             joint_feat_right=feats["joint_feat_right"], **{k: res[k] for k in jk})
        # --- whole forward, B=2
        outs, _ = net({"img": X["img"]}, None, None)
        oo = O.dir_forward(sd, X["img"])
        arrs = {}
        for i in range(3):
            for k in O.OUT_KEYS:
                arrs[f"s{i}_{k}"] = outs[i][k]
                print(f"forward stage {i} {k}: {rel(oo[i][k], outs[i][k]):.2e}")
        for k in ("seg", "dense"):
            arrs[k] = outs[3][k]
            print(f"forward {k}: {rel(oo[3][k], outs[3][k]):.2e}")
        pf = outs[3]["proj_feat"]
        print(f"forward proj_feat: {rel(oo[3]['proj_feat'], pf):.2e}")
        arrs["proj_feat_sub"] = pf.flatten()[::61]
        arrs["proj_feat_abs_sum"] = pf.abs().double().sum()
        arrs["proj_feat_nnz"] = (pf != 0).sum()
        save("forward_b2.npz", **arrs)


if __name__ == "__main__":
    main()
