"""GPU parity tests (run on the B200 box with `-m gpu`): every call goes through the C ABI
(dir_b200/libdirb200.so) and is compared with the CPU oracle (oracle/dir_oracle.py) on the same
seeded inputs, with the committed golden outputs of the unmodified reference (tests/golden/), and
— at the benchmark's full batch — through size-independent properties.

Tolerances (relative = max|a-b| / max|b| per tensor):
  fp32 configuration : 1e-4   (north_star's bar). Measured 3e-5..6e-5 on the worst tensor of the whole forward against a
                       float64 evaluation of the oracle; the fp32 oracle itself (torch CPU) sits 2.5e-5 from that truth
                       and PyTorch's fp32 cuDNN forward on the same B200 5.6e-5 (profiles/drift_table_r2.txt): this is
                       the noise floor of fp32 arithmetic through this network, not a property of the kernels.
  bf16 configuration : bf16 feature maps cannot meet 1e-4. The bar is the reference's OWN drift when run in bf16
                       (torch.autocast) on the same weights and inputs, evaluated inside the tests: per stage, our mean
                       per-vertex drift must not exceed it (measured: 2.4 mm vs 3.1-3.5 mm at stage 2).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL32 = 1e-4


def rel(a, b):
    a = a.detach().float().cpu()
    b = torch.as_tensor(b).float().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.fixture(scope="module")
def X():
    from oracle.gen_golden import seam_inputs

    return seam_inputs()


def _make(synth_sd, precision, **kw):
    import dir_b200

    m = dir_b200.DIR(21, "./misc/mano", precision=precision, **kw).cuda()
    m.load_state_dict(synth_sd, strict=False)
    m.eval()
    return m


@pytest.fixture(scope="module")
def m32(synth_sd):
    return _make(synth_sd, "fp32", max_batch=128)


@pytest.fixture(scope="module")
def m16(synth_sd):
    return _make(synth_sd, "bf16", max_batch=128)


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


# ------------------------------------------------------------------------------------------ library / boundary
def test_native_library_is_loaded_and_strict(m32):
    from dir_b200 import capi

    assert os.path.exists(capi.LIB_PATH)
    req = m32.required_keys()
    assert len(req) > 600 and len(set(req)) == len(req)
    sd_keys = set(m32.state_dict().keys())
    assert set(req) <= sd_keys
    # keys the forward never touches
    assert not any(k.startswith("backbone.fc.") or ".STEblocks.0." in k or k.endswith(".e_0") for k in req)


def test_missing_required_key_is_an_error(synth_sd):
    import dir_b200

    sd = {k: v for k, v in synth_sd.items() if k != "decoder.projecter_3.fusion.0.weight"}
    m = dir_b200.DIR(21, "./misc/mano", precision="fp32").cuda()
    m.load_state_dict(sd, strict=False)
    with pytest.raises(KeyError):
        m({"img": torch.zeros(1, 3, 256, 256)}, None, None)


def test_training_branch_and_cpu_are_refused(synth_sd):
    import dir_b200

    m = dir_b200.DIR(21, "./misc/mano", precision="fp32")
    with pytest.raises(dir_b200.DirB200Error):
        m({"img": torch.zeros(1, 3, 256, 256)}, None, None)  # module still on CPU: no fallback
    m.train()
    with pytest.raises(NotImplementedError):
        m({"img": torch.zeros(1, 3, 256, 256)}, None, None)


# ------------------------------------------------------------------------------------------ seams, fp32
@pytest.mark.parametrize("side", ["left", "right"])
def test_mano_layer_vs_golden(m32, golden_dir, X, side):
    from dir_b200 import seams

    g = load(golden_dir, f"mano_{side}.npz")
    B = X["mano_pose"].shape[0]
    para = torch.zeros(B, 2, 64)
    hand = 0 if side == "left" else 1
    para[:, hand, :51] = X["mano_pose"]
    para[:, hand, 51:61] = X["mano_beta"]
    para[:, :, 61] = 1.0
    out = seams.mano(m32, 0, para.cuda())
    assert rel(out[f"pd_mesh_xyz_{side}"], g["verts"]) < TOL32
    assert rel(out[f"pd_joint_xyz_{side}"], g["joints"]) < TOL32
    assert float(out[f"pd_joint_xyz_{side}"][:, 0].abs().max()) == 0.0  # centred on the wrist
    # scale 1, zero translation: uv == xy
    assert rel(out[f"pd_joint_uv_{side}"], g["joints"][..., :2]) < TOL32


def test_mano_all_six_copies_and_projection(m32, synth_sd):
    from dir_b200 import seams
    from oracle import dir_oracle as O

    gen = torch.Generator().manual_seed(5)
    para = torch.randn(7, 2, 64, generator=gen) * 0.4
    para[:, :, 61] = 2.5
    prefixes = ["init_regressor.", "decoder.projecter_4.regressor.", "decoder.projecter_3.regressor."]
    for which, p in enumerate(prefixes):
        out = seams.mano(m32, which, para.cuda())
        for hand, side in enumerate(("left", "right")):
            v, j = O.mano_layer(synth_sd, f"{p}mano_layer_{side}.", para[:, hand, :51], para[:, hand, 51:61], side)
            assert rel(out[f"pd_mesh_xyz_{side}"], v) < TOL32
            assert rel(out[f"pd_joint_xyz_{side}"], j) < TOL32
            assert rel(out[f"pd_joint_uv_{side}"], O.projection_xy(para[:, hand, 61:64], j)) < TOL32


def test_backbone_vs_golden_and_oracle(m32, synth_sd, golden_dir, X):
    from dir_b200 import seams
    from oracle import dir_oracle as O

    g = load(golden_dir, "resnet50.npz")
    feats = seams.backbone(m32, X["bb_img"].cuda())
    for f, k in zip(feats, ("c1", "c2", "c3", "c4")):
        assert rel(f, g[k]) < TOL32, k
    img = X["img"][:1]
    want = O.resnet50(synth_sd, img)
    got = seams.backbone(m32, img.cuda())
    for i, (a, b) in enumerate(zip(got, want)):
        assert rel(a, b) < TOL32, i


@pytest.mark.parametrize("name,cin,S", [("decoder.skip_layer4.", 1024, 16), ("decoder.fusion_layer4.", 2304, 16),
                                        ("decoder.enhance_layer4.", 512, 8), ("decoder.skip_layer3.", 512, 32),
                                        ("decoder.fusion_layer3.", 512, 32), ("decoder.enhance_layer3.", 512, 32)])
def test_residual_blocks(m32, synth_sd, golden_dir, X, name, cin, S):
    from dir_b200 import seams
    from oracle import dir_oracle as O

    if name == "decoder.enhance_layer4.":
        x = X["res_x"]
        assert rel(seams.residual(m32, name, x.cuda()), load(golden_dir, "residual.npz")["y"]) < TOL32
    else:
        x = torch.randn(3, cin, S, S, generator=torch.Generator().manual_seed(cin + S))
    assert rel(seams.residual(m32, name, x.cuda()), O.residual(synth_sd, name, x)) < TOL32


def test_init_regressor(m32, synth_sd, golden_dir, X):
    from dir_b200 import seams

    g = load(golden_dir, "init_regressor.npz")
    out = seams.init_regressor(m32, X["c4"].cuda())
    for k in g.files:
        if k in out:
            assert rel(out[k], g[k]) < TOL32, k


def _prev(X):
    return {"pd_joint_xyz_left": X["j2b_xyz_l"], "pd_joint_xyz_right": X["j2b_xyz_r"],
            "pd_joint_uv_left": X["j2b_uv_l"], "pd_joint_uv_right": X["j2b_uv_r"],
            "pd_mano_para_left": X["j2b_para_l"], "pd_mano_para_right": X["j2b_para_r"], "pd_offset": X["j2b_off"]}


def test_joint2bone_stage1_vs_golden(m32, golden_dir, X):
    from dir_b200 import seams

    g = load(golden_dir, "joint2bone.npz")
    res, feats = seams.joint2bone(m32, 1, X["j2b_feat"].cuda(), {k: v.cuda() for k, v in _prev(X).items()})
    for k in g.files:
        got = feats[k] if k in feats else res[k]
        assert rel(got, g[k]) < TOL32, k


@pytest.mark.parametrize("stage,S,uv_range", [(1, 16, 0.8), (2, 32, 0.8), (2, 32, 3.0), (1, 16, 0.0)])
def test_joint2bone_vs_oracle(m32, synth_sd, stage, S, uv_range):
    """uv_range 3.0 exercises grid_sample zero padding / off-map bones; 0.0 collapses every bone (a==b -> zeros)."""
    from dir_b200 import seams
    from oracle import dir_oracle as O

    gen = torch.Generator().manual_seed(100 + stage + int(uv_range * 10))
    B = 3
    prev = {"pd_joint_xyz_left": torch.randn(B, 21, 3, generator=gen) * 0.05,
            "pd_joint_xyz_right": torch.randn(B, 21, 3, generator=gen) * 0.05,
            "pd_joint_uv_left": (torch.rand(B, 21, 2, generator=gen) * 2 - 1) * uv_range,
            "pd_joint_uv_right": (torch.rand(B, 21, 2, generator=gen) * 2 - 1) * uv_range,
            "pd_mano_para_left": torch.randn(B, 64, generator=gen) * 0.3,
            "pd_mano_para_right": torch.randn(B, 64, generator=gen) * 0.3,
            "pd_offset": torch.randn(B, 3, generator=gen) * 0.5}
    feat = torch.randn(B, 256, S, S, generator=gen)
    p = "decoder.projecter_4." if stage == 1 else "decoder.projecter_3."
    want, wfeats = O.joint2bone(synth_sd, p, feat, prev, S, 1 if stage == 1 else 2)
    res, feats = seams.joint2bone(m32, stage, feat.cuda(), {k: v.cuda() for k, v in prev.items()}, want_vis=True)
    for k in ("pd_offset", "pd_mano_para_left", "pd_mano_para_right", "pd_joint_uv_left", "pd_joint_uv_right",
              "pd_mesh_xyz_left", "pd_mesh_xyz_right", "pd_joint_xyz_left", "pd_joint_xyz_right"):
        assert rel(res[k], want[k]) < TOL32, k
    assert rel(feats["joint_feat_left"], wfeats["joint_feat_left"]) < TOL32
    assert rel(feats["joint_feat_right"], wfeats["joint_feat_right"]) < TOL32
    # rasterisation is discontinuous in uv: compare away from flipped boundary pixels
    wv, gv = wfeats["vis_img_feat"], feats["vis_img_feat"].cpu()
    flipped = ((wv != 0) != (gv != 0)).float().mean()
    assert float(flipped) < 2e-3
    same = (wv != 0) == (gv != 0)
    assert float(((wv - gv).abs() * same).max()) < TOL32 * float(wv.abs().max() + 1e-9) + 1e-6
    assert float(((wfeats["img_feat"] - feats["img_feat"].cpu()).abs() > 1e-3 * wfeats["img_feat"].abs().max()).float().mean()) < 5e-3


def test_bone_proj_vs_golden(m32, golden_dir, X):
    from dir_b200 import seams

    g = load(golden_dir, "bone_proj.npz")
    y16 = seams.bone_proj(m32, X["bp_uv16"].cuda(), X["bp_feat"].cuda(), 16, 1.0).cpu()
    y32 = seams.bone_proj(m32, X["bp_uv32"].cuda(), X["bp_feat"][:1].cuda(), 32, 2.0).cpu()
    for got, want in ((y16, torch.as_tensor(g["y16"])), (y32, torch.as_tensor(g["y32"]))):
        assert bool(((got != 0) == (want != 0)).all())  # identical capsule masks
        assert rel(got, want) < 1e-5


def test_bone_proj_degenerate(m32):
    from dir_b200 import seams

    y = seams.bone_proj(m32, torch.zeros(2, 21, 2).cuda(), torch.ones(2, 21, 64).cuda(), 16, 1.0)
    assert float(y.abs().sum()) == 0.0 and not bool(torch.isnan(y).any())


# ------------------------------------------------------------------------------------------ whole forward
def _check_forward(outs, g, tol):
    from oracle import dir_oracle as O

    worst = 0.0
    for i in range(3):
        for k in O.OUT_KEYS:
            r = rel(outs[i][k], g[f"s{i}_{k}"])
            worst = max(worst, r)
            assert r < tol, (i, k, r)
    return worst


def test_forward_fp32_vs_golden(m32, golden_dir, X):
    g = load(golden_dir, "forward_b2.npz")
    outs, loss = m32({"img": X["img"]}, None, None)  # CPU tensor in, like apps/eval.py would after .cuda()
    assert loss == {} and len(outs) == 4 and outs[0]["pd_rel_joint"] is None
    worst = _check_forward(outs, g, TOL32)
    assert rel(outs[3]["seg"], g["seg"]) < TOL32 and rel(outs[3]["dense"], g["dense"]) < TOL32
    pf = outs[3]["proj_feat"]
    assert tuple(pf.shape) == (2, 1280, 32, 32)
    assert abs(int((pf != 0).sum()) - int(g["proj_feat_nnz"])) <= 64 * 8
    assert abs(float(pf.abs().double().sum()) / float(g["proj_feat_abs_sum"]) - 1) < 1e-3
    print(f"fp32 whole-forward worst relative error vs reference: {worst:.2e}")


def test_forward_mpjpe_delta_fp32(m32, golden_dir, X):
    """|MPJPE/MPVPE(new) - (ref)| < 0.01 mm for any fixed GT (apps/eval.py:151-193 metric core: wrist-aligned L2, mm)."""
    g = load(golden_dir, "forward_b2.npz")
    outs, _ = m32({"img": X["img"]}, None, None)
    gen = torch.Generator().manual_seed(3)
    for side in ("left", "right"):
        ref_v = torch.as_tensor(g[f"s2_pd_mesh_xyz_{side}"])
        ref_j = torch.as_tensor(g[f"s2_pd_joint_xyz_{side}"])
        gt_v = ref_v + torch.randn(ref_v.shape, generator=gen) * 0.01
        gt_j = ref_j + torch.randn(ref_j.shape, generator=gen) * 0.01
        new_v, new_j = outs[2][f"pd_mesh_xyz_{side}"].cpu(), outs[2][f"pd_joint_xyz_{side}"].cpu()
        mpvpe = lambda v: float((v - gt_v).norm(dim=-1).mean() * 1000)
        mpjpe = lambda j: float((j - gt_j).norm(dim=-1).mean() * 1000)
        assert abs(mpvpe(new_v) - mpvpe(ref_v)) < 0.01 and abs(mpjpe(new_j) - mpjpe(ref_j)) < 0.01
        assert float((new_v - ref_v).norm(dim=-1).max() * 1000) < 0.01  # stricter: per-vertex, mm


def test_forward_fp32_odd_batch_vs_oracle(m32, synth_sd):
    from oracle import dir_oracle as O

    img = torch.randn(5, 3, 256, 256, generator=torch.Generator().manual_seed(77))
    want = O.dir_forward(synth_sd, img)
    outs, _ = m32({"img": img.cuda()}, None, None)
    for i in range(3):
        for k in O.OUT_KEYS:
            assert rel(outs[i][k], want[i][k]) < TOL32, (i, k)
    assert rel(outs[3]["seg"], want[3]["seg"]) < TOL32


def test_forward_bf16_vs_golden(m16, golden_dir, X):
    """bf16 feature maps: report the drift; bound chosen from measurement (precision)."""
    g = load(golden_dir, "forward_b2.npz")
    outs, _ = m16({"img": X["img"].cuda()}, None, None)
    worst = _check_forward(outs, g, 0.15)
    ref_v = torch.as_tensor(g["s2_pd_mesh_xyz_left"])
    d_mm = float((outs[2]["pd_mesh_xyz_left"].cpu() - ref_v).norm(dim=-1).mean() * 1000)
    print(f"bf16 whole-forward worst relative error {worst:.2e}; mean per-vertex drift {d_mm:.3f} mm")
    assert d_mm < 5.0


def _mesh_drift_mm(outs, want, i):
    d = torch.cat([(outs[i][k].float().cpu() - want[i][k].float()).norm(dim=-1).flatten()
                   for k in ("pd_mesh_xyz_left", "pd_mesh_xyz_right")]) * 1000
    return float(d.mean()), float(d.max())


def test_forward_fp32_b32_vs_float64_oracle(m32, synth_sd):
    """BASELINE.json configs[1] at its full size (B=32, ResNet-50, 3 stages, fp32): every output tensor of every image
    against the oracle evaluated in float64 on the host; the fp32 oracle's own distance to that truth is printed beside it.
    Stages 0 and 1 are continuous functions of the input: 1e-4 for all 32 images. Stage 2 sits behind the one
    discontinuity of the forward, the capsule mask of stage 1's bone_proj (`hypot(h,c) < distance`, models/dir.py:164):
    an image in which some pixel lies on a capsule boundary to within the fp32 noise of uv (~1e-5 px) flips that pixel
    in ANY fp32 implementation (the torch CPU oracle included). Such images are identified from the float64 truth by
    their boundary margin, must be rare, and are bounded at 5e-2; every other image meets 1e-4."""
    from oracle import dir_oracle as O

    B = 32
    img = torch.randn(B, 3, 256, 256, generator=torch.Generator().manual_seed(3232))
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in synth_sd.items()}
    truth = O.dir_forward(sd64, img.double())
    host32 = O.dir_forward(synth_sd, img)
    outs, _ = m32({"img": img}, None, None)

    def per_image(o, i):
        e = torch.zeros(B, dtype=torch.float64)
        for k in O.OUT_KEYS:
            t = truth[i][k].double()
            d = (o[i][k].detach().double().cpu() - t).abs().flatten(1).max(1).values
            e = torch.maximum(e, d / t.abs().max())
        return e

    for i in (0, 1):
        ours, floor = per_image(outs, i), per_image(host32, i)
        print(f"B=32 fp32 stage {i}: worst image {float(ours.max()):.2e} (torch CPU fp32 oracle {float(floor.max()):.2e})")
        assert float(ours.max()) < TOL32
    # stage 2: margins of stage 1's capsule test, from the truth
    margin = torch.minimum(O.bone_capsule_margin(truth[1]["pd_joint_uv_left"], 16, 1.0),
                           O.bone_capsule_margin(truth[1]["pd_joint_uv_right"], 16, 1.0))
    ours, floor = per_image(outs, 2), per_image(host32, 2)
    on_boundary = margin < 1e-4  # pixels: uv carries ~1e-5 of fp32 noise here (3e-5 relative on the worst tensor), x S/2
    print(f"B=32 fp32 stage 2: worst image off the boundary {float(ours[~on_boundary].max()):.2e} "
          f"(torch CPU fp32 oracle {float(floor[~on_boundary].max()):.2e}); {int(on_boundary.sum())} image(s) within 1e-4 px of a "
          f"capsule boundary: ours {[f'{float(v):.1e}' for v in ours[on_boundary]]}, "
          f"oracle fp32 {[f'{float(v):.1e}' for v in floor[on_boundary]]}, margins {[f'{float(v):.1e}' for v in margin[on_boundary]]}")
    assert float(ours[~on_boundary].max()) < TOL32
    assert int(on_boundary.sum()) <= 4 and (not bool(on_boundary.any()) or float(ours[on_boundary].max()) < 5e-2)
    mm = torch.cat([(outs[2][k].detach().double().cpu() - truth[2][k]).norm(dim=-1)[~on_boundary].flatten()
                    for k in ("pd_mesh_xyz_left", "pd_mesh_xyz_right")]) * 1000
    print(f"B=32 fp32 stage 2 per-vertex drift: mean {float(mm.mean()):.5f} mm max {float(mm.max()):.5f} mm")
    assert float(mm.max()) < 0.01  # north_star's MPVPE budget, applied per vertex


def test_forward_bf16_b128_vs_oracle_and_reference_autocast(m16, synth_sd):
    """The benchmarked configuration at its full batch (B=128, bf16) against the fp32 oracle, with the reference's own
    bf16 behaviour (the same op sequence under torch.autocast(bfloat16) on the host) as the yardstick: per stage our
    mean per-vertex drift must not exceed the reference-autocast drift on the same 128 images."""
    from oracle import dir_oracle as O

    img = torch.randn(128, 3, 256, 256, generator=torch.Generator().manual_seed(128128))
    want = O.dir_forward(synth_sd, img)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        auto = O.dir_forward(synth_sd, img)
    auto = [{k: (v.float() if v is not None else None) for k, v in d.items()} for d in auto[:3]]
    outs, _ = m16({"img": img}, None, None)
    for i in range(3):
        ours, ref = _mesh_drift_mm(outs, want, i), _mesh_drift_mm(auto, want, i)
        worst = max(rel(outs[i][k], want[i][k]) for k in O.OUT_KEYS)
        print(f"B=128 bf16 stage {i}: per-vertex drift mean {ours[0]:.3f} mm max {ours[1]:.2f} mm "
              f"(reference under bf16 autocast: mean {ref[0]:.3f} mm max {ref[1]:.2f} mm); worst relative {worst:.3f}")
        assert ours[0] <= ref[0], (i, ours, ref)
    assert _mesh_drift_mm(outs, want, 2)[0] < 4.0


@pytest.mark.parametrize("which", ["m32", "m16"])
def test_full_batch_properties(which, request):
    """B=128 (BASELINE.json's batch): no oracle run; size-independent properties instead.
    (1) per-image independence: an image's outputs do not depend on its batch or position — bit-exact;
    (2) MANO invariants: wrist joint is the origin; uv == s*xy + t;  (3) everything finite."""
    m = request.getfixturevalue(which)
    gen = torch.Generator().manual_seed(1234)
    img = torch.randn(128, 3, 256, 256, generator=gen).cuda()
    big = m.run_raw(img)
    rec = big["record"]
    assert bool(torch.isfinite(rec).all())
    pick = [5, 77, 127]
    small = m.run_raw(img[pick])
    assert torch.equal(small["record"], rec[pick])
    assert torch.equal(small["seg"], big["seg"][pick])
    outs = m.unpack_record(rec)
    for o in outs[:3]:
        for side in ("left", "right"):
            j, uv, pr = o[f"pd_joint_xyz_{side}"], o[f"pd_joint_uv_{side}"], o[f"pd_proj_{side}"]
            assert float(j[:, 0].abs().max()) == 0.0
            want = pr[:, None, 0:1] * j[..., :2] + pr[:, None, 1:3]
            assert float((uv - want).abs().max()) < 1e-5 * (1 + float(want.abs().max()))


def test_chunking_over_max_batch(synth_sd):
    m = _make(synth_sd, "fp32", max_batch=4, aux_outputs=False)
    img = torch.randn(9, 3, 256, 256, generator=torch.Generator().manual_seed(9)).cuda()
    a = m.run_raw(img)["record"]
    b = torch.cat([m.run_raw(img[i:i + 3])["record"] for i in range(0, 9, 3)])
    assert torch.equal(a, b)
    outs, _ = m({"img": img}, None, None)
    assert outs[3]["seg"] is None


def test_cuda_graph_replay_matches_eager(synth_sd, X):
    m = _make(synth_sd, "fp32", max_batch=8, use_cuda_graph=True)
    e = _make(synth_sd, "fp32", max_batch=8)
    img = X["img"].cuda()
    want = e.run_raw(img)
    for _ in range(3):  # both graph sets, then the first one again
        got = m.run_raw(img)
        assert torch.equal(got["record"], want["record"]) and torch.equal(got["proj_feat"], want["proj_feat"])
    # outputs are the graph's own double-buffered tensors: call i stays intact across call i+1
    first = m.run_raw(img)
    keep = first["record"].clone()
    other = m.run_raw(torch.flip(img, dims=[0]))
    assert torch.equal(first["record"], keep) and not torch.equal(other["record"], keep)
    assert torch.equal(other["record"], want["record"].flip(0))


def test_factored_fusion_equals_dense_path(synth_sd, X, monkeypatch):
    """The factored bone_proj->conv3x3 (fusion.cu) against the dense path (bone raster + 2560-channel conv),
    fp32, whole forward: same math, different summation order."""
    fact = _make(synth_sd, "fp32", max_batch=4)
    monkeypatch.setenv("DIRB200_DENSE_FUSION", "1")
    dense = _make(synth_sd, "fp32", max_batch=4)
    dense._ensure_handle()  # handle reads the env var at creation
    monkeypatch.delenv("DIRB200_DENSE_FUSION")
    img = X["img"].cuda()
    a, b = fact.run_raw(img), dense.run_raw(img)
    assert rel(a["record"], b["record"]) < 2e-5
    assert rel(a["seg"], b["seg"]) < 2e-5 and rel(a["dense"], b["dense"]) < 2e-5
    # proj_feat comes from the same kernel in both; its inputs (stage-2 uv) differ by ~1e-7, which moves values by
    # ~1e-6 and may flip a few capsule-boundary pixels
    pa, pb = a["proj_feat"], b["proj_feat"]
    same = (pa != 0) == (pb != 0)
    assert float((~same).float().mean()) < 1e-4
    assert float(((pa - pb).abs() * same).max()) < 1e-4 * float(pb.abs().max())


def _ref_preprocess(frames_u8):
    """apps/eval.py:56-61 / dataset/interhand.py:223-225 with torch ops: BGR->RGB, /255, HWC->CHW, Normalize."""
    rgb = frames_u8.flip(-1).float() / 255
    x = rgb.permute(0, 3, 1, 2)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    return (x - mean) / std


def test_u8_input_pipeline(m32, m16):
    """Next-row N1: raw uint8 BGR frames in, preprocessing on the device."""
    from dir_b200 import seams

    frames = torch.randint(0, 256, (3, 256, 256, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(8))
    want = _ref_preprocess(frames)
    got = seams.preprocess_u8(m32, frames.cuda())
    assert float((got.cpu() - want).abs().max()) < 2e-6
    # whole forward: uint8 frames == the reference-preprocessed fp32 image
    a = m32.run_raw(frames.cuda())["record"]
    b = m32.run_raw(want.cuda())["record"]
    assert rel(a, b) < 1e-5
    outs, _ = m16({"img": frames}, None, None)  # host uint8 tensor through the public forward
    c = m16.run_raw(want.cuda())["record"]
    assert rel(outs[2]["pd_mesh_xyz_left"], m16.unpack_record(c)[2]["pd_mesh_xyz_left"]) < 2e-2


def test_eval_metric_kernel_vs_golden(m32, golden_dir):
    """Next-row N2: device-side MPJPE/MPVPE/2-D/root errors vs the reference's own metric lines (apps/eval.py:151-241)."""
    from dir_b200 import capi, seams
    from oracle.gen_golden import eval_metric_inputs

    E = eval_metric_inputs()
    g = load(golden_dir, "eval_metric.npz")
    B = E["cam"].shape[0]
    rec = torch.zeros(B, capi.RECORD_FLOATS)
    s2 = 2 * capi.STAGE_FLOATS
    rec[:, s2 + capi.OFF["mesh_l"]:s2 + capi.OFF["mesh_l"] + 2334] = E["pred_verts"]["left"].reshape(B, -1)
    rec[:, s2 + capi.OFF["mesh_r"]:s2 + capi.OFF["mesh_r"] + 2334] = E["pred_verts"]["right"].reshape(B, -1)
    rec[:, s2 + capi.OFF["offset"]:s2 + capi.OFF["offset"] + 3] = E["pred_offset"]
    gv = torch.stack((E["gt_verts"]["left"], E["gt_verts"]["right"]), 1)
    g2 = torch.stack((E["gt_verts2d"]["left"], E["gt_verts2d"]["right"]), 1)
    jr = torch.stack([seams.eval_jregressor(E["jreg16"][s]) for s in ("left", "right")])
    out = seams.eval_metrics(m32, rec.cuda(), gv.cuda(), g2.cuda(), E["cam"].cuda(), jr.cuda())
    for k in g.files:  # errors are differences of aligned coordinates: fp32 cancellation, ~5e-7 m absolute
        assert rel(out[k], g[k]) < 1e-4, k
    mpjpe = float(out["joint_left"].mean() * 1000)
    assert abs(mpjpe - float(g["joint_left"].mean() * 1000)) < 0.01  # mm


def test_pair_fused_convs_match_separate_launches(synth_sd, X, monkeypatch):
    """bf16 path: conv3 + (downsample | skip) as ONE K-concatenated tcgen05 GEMM vs two launches with a residual add.
    Same math up to bf16 rounding of the scale-folded weights and of the (no longer materialised) branch output."""
    fused = _make(synth_sd, "bf16", max_batch=4)
    monkeypatch.setenv("DIRB200_NO_PAIR_FUSION", "1")
    sep = _make(synth_sd, "bf16", max_batch=4)
    sep._ensure_handle()
    monkeypatch.delenv("DIRB200_NO_PAIR_FUSION")
    from dir_b200 import seams

    img = X["img"][:1]
    fa, fb = seams.backbone(fused, img.cuda()), seams.backbone(sep, img.cuda())
    for a, b in zip(fa, fb):
        assert float((a - b).abs().mean() / b.abs().mean()) < 1e-2
    ra, rb = fused.run_raw(X["img"].cuda())["record"], sep.run_raw(X["img"].cuda())["record"]
    assert rel(ra, rb) < 5e-2
    n_fused = fused._handle.lib.dirb200_forward_launches(fused._handle.h, 2)
    n_sep = sep._handle.lib.dirb200_forward_launches(sep._handle.h, 2)
    # 4 downsample + 6 skip convs disappear; the pair also carries layer1.0 -> layer1.1.conv1 back to back (conv_b2b.cu)
    # and lets enhance_layer{4,3} read their two inputs in place (no concat launch)
    assert n_fused == n_sep - 13, (n_fused, n_sep)


@pytest.mark.parametrize("stage,S,B", [(1, 16, 3), (2, 32, 8)])
def test_ste_tcgen05_vs_cuda_core_kernel(synth_sd, monkeypatch, stage, S, B):
    """bf16 configuration: mixSTE on tcgen05 (ste_tc.cu, bf16 operands / fp32 accumulate, residual stream in TMEM)
    against the fp32 CUDA-core ste_kernel (DIRB200_STE_SIMT=1) through the joint2bone seam, same bf16 feature map.
    Odd and even batches (two images per CTA: the last CTA of an odd batch has an empty slot). The joint features
    (Linear over the STE tokens) and the refined MANO outputs must agree to bf16-operand accuracy; measured printed."""
    from dir_b200 import seams

    tc = _make(synth_sd, "bf16", max_batch=8)
    monkeypatch.setenv("DIRB200_STE_SIMT", "1")
    simt = _make(synth_sd, "bf16", max_batch=8)
    simt._ensure_handle()  # handle reads the env var at creation
    monkeypatch.delenv("DIRB200_STE_SIMT")
    gen = torch.Generator().manual_seed(500 + stage)
    prev = {"pd_joint_xyz_left": torch.randn(B, 21, 3, generator=gen) * 0.05,
            "pd_joint_xyz_right": torch.randn(B, 21, 3, generator=gen) * 0.05,
            "pd_joint_uv_left": (torch.rand(B, 21, 2, generator=gen) * 2 - 1) * 0.8,
            "pd_joint_uv_right": (torch.rand(B, 21, 2, generator=gen) * 2 - 1) * 0.8,
            "pd_mano_para_left": torch.randn(B, 64, generator=gen) * 0.3,
            "pd_mano_para_right": torch.randn(B, 64, generator=gen) * 0.3,
            "pd_offset": torch.randn(B, 3, generator=gen) * 0.5}
    feat = torch.randn(B, 256, S, S, generator=gen).cuda()
    prev = {k: v.cuda() for k, v in prev.items()}
    ra, fa = seams.joint2bone(tc, stage, feat, prev)
    rb, fb = seams.joint2bone(simt, stage, feat, prev)
    worst = {}
    for k in ("joint_feat_left", "joint_feat_right"):
        worst[k] = rel(fa[k], fb[k])
    for k in ("pd_mano_para_left", "pd_mano_para_right", "pd_mesh_xyz_left", "pd_mesh_xyz_right", "pd_joint_uv_left"):
        worst[k] = rel(ra[k], rb[k])
    print("tcgen05 STE vs fp32 STE:", {k: f"{v:.2e}" for k, v in worst.items()})
    assert all(bool(torch.isfinite(v).all()) for v in ra.values() if v is not None)
    assert max(worst.values()) < 3e-2
    # per-image independence across the two slots of a CTA: image i alone == image i inside the batch
    one, _ = seams.joint2bone(tc, stage, feat[1:2], {k: v[1:2] for k, v in prev.items()})
    assert torch.equal(one["pd_mesh_xyz_left"], ra["pd_mesh_xyz_left"][1:2])


@pytest.mark.parametrize("stage,S,B", [(1, 16, 3), (2, 32, 70)])
def test_bone_coef_tf32_tensor_core_vs_cuda_core(synth_sd, monkeypatch, stage, S, B):
    """bf16 configuration: bone_coef as tcgen05 kind::tf32 GEMMs against the fp32 CUDA-core kernel
    (DIRB200_COEF_SIMT=1). Only the fused stage feature map depends on it; tf32 operands (10-bit mantissa) under a
    bf16 output: most pixels identical, the rest one bf16 ulp apart. B=70 -> 140 rows = one full + one ragged M tile."""
    from dir_b200 import seams

    tc = _make(synth_sd, "bf16", max_batch=128)
    monkeypatch.setenv("DIRB200_COEF_SIMT", "1")
    simt = _make(synth_sd, "bf16", max_batch=128)
    simt._ensure_handle()
    monkeypatch.delenv("DIRB200_COEF_SIMT")
    gen = torch.Generator().manual_seed(700 + stage)
    prev = {"pd_joint_xyz_left": torch.randn(B, 21, 3, generator=gen) * 0.05,
            "pd_joint_xyz_right": torch.randn(B, 21, 3, generator=gen) * 0.05,
            "pd_joint_uv_left": (torch.rand(B, 21, 2, generator=gen) * 2 - 1) * 0.8,
            "pd_joint_uv_right": (torch.rand(B, 21, 2, generator=gen) * 2 - 1) * 0.8,
            "pd_mano_para_left": torch.randn(B, 64, generator=gen) * 0.3,
            "pd_mano_para_right": torch.randn(B, 64, generator=gen) * 0.3,
            "pd_offset": torch.randn(B, 3, generator=gen) * 0.5}
    feat = torch.randn(B, 256, S, S, generator=gen).cuda()
    prev = {k: v.cuda() for k, v in prev.items()}
    ra, fa = seams.joint2bone(tc, stage, feat, prev)
    rb, fb = seams.joint2bone(simt, stage, feat, prev)
    assert torch.equal(ra["pd_mesh_xyz_left"], rb["pd_mesh_xyz_left"])  # upstream of bone_coef: untouched
    a, b = fa["img_feat"], fb["img_feat"]
    assert bool(torch.isfinite(a).all()) and float(b.abs().max()) > 0
    err = float((a - b).abs().max() / b.abs().max())
    frac = float(((a - b).abs() > 0).float().mean())
    print(f"tf32 bone_coef vs fp32: max rel err {err:.2e}, {100 * frac:.2f}% of outputs differ")
    assert err < 1.5e-2  # a couple of bf16 ulps of the largest value


@pytest.mark.parametrize("B", [1, 3])
def test_fused_stem_pool_vs_split_kernels(synth_sd, monkeypatch, B):
    """bf16 configuration: conv1+bn1+ReLU+maxpool as ONE kernel (stem_pool.cu: no-swizzle overlapping-window UMMA operand,
    conv rows pooled out of a shared-memory ring) against the TMA implicit-GEMM stem + separate max-pool kernel
    (DIRB200_STEM_SPLIT=1). Same bf16 operands and fp32 accumulation: the backbone outputs agree to bf16 rounding."""
    from dir_b200 import seams

    fused = _make(synth_sd, "bf16", max_batch=4)
    monkeypatch.setenv("DIRB200_STEM_SPLIT", "1")
    split = _make(synth_sd, "bf16", max_batch=4)
    split._ensure_handle()
    monkeypatch.delenv("DIRB200_STEM_SPLIT")
    img = torch.randn(B, 3, 256, 256, generator=torch.Generator().manual_seed(40 + B)).cuda()
    fa, fb = seams.backbone(fused, img), seams.backbone(split, img)
    for name, a, b in zip(("c1", "c2", "c3", "c4"), fa, fb):
        assert bool(torch.isfinite(a).all())
        e = float((a - b).abs().max() / b.abs().max())
        print(f"fused stem vs split, {name}: max rel err {e:.2e}")
        assert e < 2e-2, name
    ra, rb = fused.run_raw(img)["record"], split.run_raw(img)["record"]
    assert rel(ra, rb) < 5e-2
    n_f = fused._handle.lib.dirb200_forward_launches(fused._handle.h, B)
    n_s = split._handle.lib.dirb200_forward_launches(split._handle.h, B)
    assert n_f == n_s - 1  # the max-pool launch is gone


@pytest.mark.parametrize("stage,S,B", [(1, 16, 3), (2, 32, 130)])
def test_gcn_tf32_tensor_core_vs_cuda_core(synth_sd, monkeypatch, stage, S, B):
    """bf16 configuration: the SemGCN layer products as tcgen05 kind::tf32 GEMMs (gcn_tc.cu) against the fp32
    CUDA-core kernel (DIRB200_GCN_SIMT=1), through the joint2bone seam; both sides use the fp32 mixSTE so that only
    the GCN differs. B=130: one full 128-image tile plus a ragged one."""
    from dir_b200 import seams

    monkeypatch.setenv("DIRB200_STE_SIMT", "1")
    tc = _make(synth_sd, "bf16", max_batch=130)
    tc._ensure_handle()
    monkeypatch.setenv("DIRB200_GCN_SIMT", "1")
    simt = _make(synth_sd, "bf16", max_batch=130)
    simt._ensure_handle()
    monkeypatch.delenv("DIRB200_GCN_SIMT")
    monkeypatch.delenv("DIRB200_STE_SIMT")
    gen = torch.Generator().manual_seed(900 + stage)
    prev = {"pd_joint_xyz_left": torch.randn(B, 21, 3, generator=gen) * 0.05,
            "pd_joint_xyz_right": torch.randn(B, 21, 3, generator=gen) * 0.05,
            "pd_joint_uv_left": (torch.rand(B, 21, 2, generator=gen) * 2 - 1) * 0.8,
            "pd_joint_uv_right": (torch.rand(B, 21, 2, generator=gen) * 2 - 1) * 0.8,
            "pd_mano_para_left": torch.randn(B, 64, generator=gen) * 0.3,
            "pd_mano_para_right": torch.randn(B, 64, generator=gen) * 0.3,
            "pd_offset": torch.randn(B, 3, generator=gen) * 0.5}
    feat = torch.randn(B, 256, S, S, generator=gen).cuda()
    prev = {k: v.cuda() for k, v in prev.items()}
    ra, fa = seams.joint2bone(tc, stage, feat, prev)
    rb, fb = seams.joint2bone(simt, stage, feat, prev)
    worst = {k: rel(fa[k], fb[k]) for k in ("joint_feat_left", "joint_feat_right")}
    worst.update({k: rel(ra[k], rb[k]) for k in ("pd_mano_para_left", "pd_mesh_xyz_left", "pd_mesh_xyz_right")})
    print("tf32 SemGCN vs fp32:", {k: f"{v:.2e}" for k, v in worst.items()})
    assert all(bool(torch.isfinite(v).all()) for v in ra.values() if v is not None)
    assert max(worst["joint_feat_left"], worst["joint_feat_right"], worst["pd_mano_para_left"]) < 5e-3
    assert max(worst.values()) < 3e-2  # the synthetic MANO heads amplify parameter noise ~5x into the mesh


@pytest.mark.parametrize("stage,S,B,uv_range", [(1, 16, 3, 0.8), (2, 32, 5, 0.8), (2, 32, 4, 3.0), (1, 16, 2, 0.0)])
def test_bone_fusion_tensor_core_vs_cuda_core(synth_sd, monkeypatch, stage, S, B, uv_range):
    """bf16 configuration: the sparse bone accumulate as a per-row-block tcgen05 GEMM (A = capsule weights written in
    place, B = coefficient vectors, both bf16) against the fp32 CUDA-core accumulate (DIRB200_FUSION_SIMT=1).
    uv_range 3.0: bones partly / fully off the map; 0.0: every bone collapsed (empty masks -> bias only)."""
    from dir_b200 import seams

    tc = _make(synth_sd, "bf16", max_batch=8)
    monkeypatch.setenv("DIRB200_FUSION_SIMT", "1")
    simt = _make(synth_sd, "bf16", max_batch=8)
    simt._ensure_handle()
    monkeypatch.delenv("DIRB200_FUSION_SIMT")
    gen = torch.Generator().manual_seed(1100 + stage + int(uv_range * 10))
    prev = {"pd_joint_xyz_left": torch.randn(B, 21, 3, generator=gen) * 0.05,
            "pd_joint_xyz_right": torch.randn(B, 21, 3, generator=gen) * 0.05,
            "pd_joint_uv_left": (torch.rand(B, 21, 2, generator=gen) * 2 - 1) * uv_range,
            "pd_joint_uv_right": (torch.rand(B, 21, 2, generator=gen) * 2 - 1) * uv_range,
            "pd_mano_para_left": torch.randn(B, 64, generator=gen) * 0.3,
            "pd_mano_para_right": torch.randn(B, 64, generator=gen) * 0.3,
            "pd_offset": torch.randn(B, 3, generator=gen) * 0.5}
    feat = torch.randn(B, 256, S, S, generator=gen).cuda()
    prev = {k: v.cuda() for k, v in prev.items()}
    ra, fa = seams.joint2bone(tc, stage, feat, prev)
    rb, fb = seams.joint2bone(simt, stage, feat, prev)
    assert torch.equal(ra["pd_mesh_xyz_left"], rb["pd_mesh_xyz_left"])  # upstream of the fusion: untouched
    a, b = fa["img_feat"], fb["img_feat"]
    assert bool(torch.isfinite(a).all()) and float(b.abs().max()) > 0
    err = float((a - b).abs().max() / b.abs().max())
    mean = float((a - b).abs().mean() / b.abs().mean())
    print(f"tcgen05 bone fusion vs fp32 accumulate: max rel err {err:.2e}, mean rel err {mean:.2e}")
    assert err < 3e-2 and mean < 5e-3
    one, fone = seams.joint2bone(tc, stage, feat[1:2], {k: v[1:2] for k, v in prev.items()})
    assert torch.equal(fone["img_feat"], a[1:2])  # per-image independence, bit-exact


def test_one_refine_iteration_config(synth_sd):
    """BASELINE.json configs[0]: B=1, ResNet-50, "1 refine iter" = init regression + projecter_4 (truncation of the
    reference forward after models/dir.py:456). refine_stages=1 must reproduce stages 0 and 1 of the full
    forward and skip everything behind them."""
    from oracle import dir_oracle as O

    one = _make(synth_sd, "fp32", max_batch=4, aux_outputs=False, refine_stages=1)
    full = _make(synth_sd, "fp32", max_batch=4, aux_outputs=False)
    img = torch.randn(1, 3, 256, 256, generator=torch.Generator().manual_seed(11))
    want = O.dir_forward(synth_sd, img)
    outs, _ = one({"img": img}, None, None)
    assert len(outs) == 3 and outs[2]["seg"] is None  # two stage dicts + the (empty) aux dict
    for i in range(2):
        for k in O.OUT_KEYS:
            assert rel(outs[i][k], want[i][k]) < TOL32, (i, k)
    a, b = one.run_raw(img.cuda()), full.run_raw(img.cuda())
    n = 2 * 4887
    assert torch.equal(a["record"][:, :n], b["record"][:, :n])  # same kernels, same bits
    assert float(a["record"][:, n:].abs().max()) == 0.0
    h1, h2 = one._handle, full._handle
    assert h1.lib.dirb200_forward_launches(h1.h, 1) < h2.lib.dirb200_forward_launches(h2.h, 1) - 20
    import dir_b200

    with pytest.raises(ValueError):
        dir_b200.DIR(21, "./misc/mano", refine_stages=5)  # the reference has no third refinement stage


def test_forward_tf32_matches_what_torch_does_by_default_on_this_gpu(synth_sd):
    """precision='tf32': fp32 activations, plain TF32 tensor-core convs — the arithmetic PyTorch gives the REFERENCE on
    this GPU out of the box (torch.backends.cudnn.allow_tf32 defaults to True). Yardstick evaluated here: the oracle's
    op sequence in PyTorch eager on the same GPU with that default; our drift from the fp32 truth must not exceed 1.5x
    its drift (both are ~0.5 mm: the tf32 truncation noise of ~60 conv layers)."""
    from oracle import dir_oracle as O

    m = _make(synth_sd, "tf32", max_batch=16)
    img = torch.randn(16, 3, 256, 256, generator=torch.Generator().manual_seed(1616))
    want = O.dir_forward(synth_sd, img)
    sd_gpu = {k: v.cuda() for k, v in synth_sd.items()}
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = True
    try:
        eager = O.dir_forward(sd_gpu, img.cuda())
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    outs, _ = m({"img": img}, None, None)
    for i in range(3):
        ours, ref = _mesh_drift_mm(outs, want, i), _mesh_drift_mm(eager, want, i)
        print(f"tf32 stage {i}: per-vertex drift mean {ours[0]:.3f} mm max {ours[1]:.2f} mm "
              f"(PyTorch eager with TF32 convs on this GPU: mean {ref[0]:.3f} mm max {ref[1]:.2f} mm)")
        assert ours[0] <= 1.5 * ref[0] + 0.02, (i, ours, ref)
