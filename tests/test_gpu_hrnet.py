"""GPU: the HRNet-W32 extension (BASELINE.json configs 3-5) against its SELF-AUTHORED oracle
(oracle/hrnet_oracle.py). PARITY UNPINNED: the reference contains no HRNet, so these tests establish that the CUDA
path computes what our own PyTorch restatement of the published architecture computes — nothing more. The modules
around the backbone (InitRegressor, decoder, refinement stages, MANO) are the reference's and are pinned elsewhere."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.fixture(scope="module")
def sd():
    from oracle.synth import make_state_dict

    return make_state_dict(0, backbone="hrnet_w32")


def _make(sd, precision, **kw):
    import dir_b200

    m = dir_b200.DIR(21, "./misc/mano", precision=precision, backbone="hrnet_w32", max_batch=8, **kw).cuda()
    m.load_state_dict(sd, strict=False)
    m.eval()
    return m


@pytest.fixture(scope="module")
def m32(sd):
    return _make(sd, "fp32")


@pytest.fixture(scope="module")
def m16(sd):
    return _make(sd, "bf16")


def test_hrnet_backbone_fp32_vs_oracle(m32, sd):
    from dir_b200 import seams
    from oracle import hrnet_oracle as H

    img = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(5))
    want = H.hrnet(sd, img, 32)
    got = seams.backbone(m32, img.cuda())
    assert tuple(got[0].shape) == (2, 64, 64, 64)  # branch 0 lives in 64 channels, the upper 32 exactly zero
    assert float(got[0][:, 32:].abs().max()) == 0.0
    assert rel(got[0][:, :32], want[0]) < 1e-4
    for i in (1, 2, 3):
        assert got[i].shape == want[i].shape and rel(got[i], want[i]) < 1e-4, i


def test_hrnet_backbone_bf16_vs_oracle(m16, sd):
    from dir_b200 import seams
    from oracle import hrnet_oracle as H

    img = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(6))
    want = H.hrnet(sd, img, 32)
    got = seams.backbone(m16, img.cuda())
    assert float(got[0][:, 32:].abs().max()) == 0.0
    for i, (a, b) in enumerate(zip([got[0][:, :32]] + got[1:], want)):
        err = rel(a, b)
        mean = float((a.cpu() - b).abs().mean() / b.abs().mean())
        print(f"hrnet bf16 c{i + 1}: max-rel {err:.3e} mean-rel {mean:.3e}")
        assert err < 8e-2 and mean < 3e-2, (i, err, mean)  # ~150 chained bf16 convs (ResNet-50 test: 6e-2 / 2e-2 for 53)


def test_hrnet_forward_fp32_vs_oracle(m32, sd):
    from oracle import dir_oracle as O
    from oracle import hrnet_oracle as H

    img = torch.randn(3, 3, 256, 256, generator=torch.Generator().manual_seed(7))
    want = H.dir_forward(sd, img, 32)
    outs, loss = m32({"img": img}, None, None)
    assert loss == {} and len(outs) == 4
    worst = 0.0
    for i in range(3):
        for k in O.OUT_KEYS:
            worst = max(worst, rel(outs[i][k], want[i][k]))
    print(f"hrnet_w32 fp32 whole forward: worst relative error vs the self-authored oracle {worst:.2e}")
    assert worst < 1e-4
    assert rel(outs[3]["seg"], want[3]["seg"]) < 1e-4 and rel(outs[3]["dense"], want[3]["dense"]) < 1e-4


def test_hrnet_forward_bf16_vs_oracle_and_autocast(m16, sd):
    from oracle import hrnet_oracle as H

    img = torch.randn(8, 3, 256, 256, generator=torch.Generator().manual_seed(8))
    want = H.dir_forward(sd, img, 32)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        auto = H.dir_forward(sd, img, 32)
    outs, _ = m16({"img": img}, None, None)

    def drift(o, i):
        d = torch.cat([(o[i][k].float().cpu() - want[i][k]).norm(dim=-1).flatten()
                       for k in ("pd_mesh_xyz_left", "pd_mesh_xyz_right")]) * 1000
        return float(d.mean())

    for i in range(3):
        ours, ref = drift(outs, i), drift(auto, i)
        print(f"hrnet_w32 bf16 stage {i}: mean per-vertex drift {ours:.3f} mm (same oracle under bf16 autocast: {ref:.3f} mm)")
        assert ours <= 1.25 * ref + 0.05, (i, ours, ref)
    assert bool(torch.isfinite(outs[2]["pd_mesh_xyz_left"]).all())


def test_hrnet_u8_frames_and_batch_independence(m16):
    frames = torch.randint(0, 256, (5, 256, 256, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(9)).cuda()
    big = m16.run_raw(frames)["record"]
    small = m16.run_raw(frames[[1, 4]])["record"]
    assert torch.equal(small, big[[1, 4]]) and bool(torch.isfinite(big).all())


# ---------------------------------------------------------------------------------------------- HRNet-W48 (configs[4])
@pytest.fixture(scope="module")
def sd48():
    from oracle.synth import make_state_dict

    return make_state_dict(0, backbone="hrnet_w48")


def _make48(sd48, precision):
    import dir_b200

    m = dir_b200.DIR(21, "./misc/mano", precision=precision, backbone="hrnet_w48", max_batch=8).cuda()
    m.load_state_dict(sd48, strict=False)
    m.eval()
    return m


def test_hrnet_w48_fp32_backbone_and_forward_vs_oracle(sd48):
    """Widths 48 / 96 / 192 / 384: the 48- and 96-channel branches are stored zero-padded in 64 and 128 channels (padding
    exactly zero everywhere), the decoder's skip_layer3 consumes the padded c2 with zero-padded weights and pre-activation,
    InitRegressor runs with feat_dim 384. Against the width-generic self-authored oracle (parity unpinned)."""
    from dir_b200 import seams
    from oracle import dir_oracle as O
    from oracle import hrnet_oracle as H

    m = _make48(sd48, "fp32")
    img = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(11))
    want = H.hrnet(sd48, img, 48)
    got = seams.backbone(m, img.cuda())
    assert [tuple(t.shape[1:]) for t in got] == [(64, 64, 64), (128, 32, 32), (192, 16, 16), (384, 8, 8)]
    assert float(got[0][:, 48:].abs().max()) == 0.0 and float(got[1][:, 96:].abs().max()) == 0.0
    for i, n in enumerate((48, 96, 192, 384)):
        assert rel(got[i][:, :n], want[i]) < 1e-4, i
    wf = H.dir_forward(sd48, img, 48)
    outs, _ = m({"img": img}, None, None)
    worst = max(rel(outs[i][k], wf[i][k]) for i in range(3) for k in O.OUT_KEYS)
    print(f"hrnet_w48 fp32 whole forward: worst relative error vs the self-authored oracle {worst:.2e}")
    assert worst < 1e-4
    assert rel(outs[3]["seg"], wf[3]["seg"]) < 1e-4 and rel(outs[3]["dense"], wf[3]["dense"]) < 1e-4


def test_hrnet_w48_bf16_vs_oracle_and_autocast(sd48):
    from oracle import hrnet_oracle as H

    m = _make48(sd48, "bf16")
    img = torch.randn(4, 3, 256, 256, generator=torch.Generator().manual_seed(12))
    want = H.dir_forward(sd48, img, 48)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        auto = H.dir_forward(sd48, img, 48)
    outs, _ = m({"img": img}, None, None)

    def drift(o, i):
        d = torch.cat([(o[i][k].float().cpu() - want[i][k]).norm(dim=-1).flatten()
                       for k in ("pd_mesh_xyz_left", "pd_mesh_xyz_right")]) * 1000
        return float(d.mean())

    for i in range(3):
        ours, ref = drift(outs, i), drift(auto, i)
        print(f"hrnet_w48 bf16 stage {i}: mean per-vertex drift {ours:.3f} mm (same oracle under bf16 autocast: {ref:.3f} mm)")
        assert ours <= 1.25 * ref + 0.05, (i, ours, ref)
    assert bool(torch.isfinite(outs[2]["pd_mesh_xyz_left"]).all())
