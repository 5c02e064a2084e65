"""GPU: every distinct conv shape of the path, one layer at a time, through dirb200_conv_layer.

fp32 handle  -> error-compensated 3xTF32 tcgen05 kernel with chunked round-to-nearest accumulation (conv_tf32.cu; asserted
                via used_tensor_cores), compared with a FLOAT64 conv2d: the bound 3e-6 of the tensor's max is what an
                fp32 library conv achieves (torch's own CPU fp32 conv2d is checked against the same truth beside it);
                the CUDA-core kernel of round 1 (DIRB200_FP32_SIMT=1) stays as the A/B at 1e-4.
tf32 handle  -> the same kernel with one MMA per k-step (plain TF32, PyTorch's cuDNN default): operands truncated to
                10 mantissa bits -> 2e-3 of max.
bf16 handle  -> tcgen05/TMA kernel (asserted via used_tensor_cores), compared with an fp32 conv2d on the
                bf16-ROUNDED inputs and weights, i.e. the exact arithmetic the tensor cores do (products of bf16
                are exact in fp32, accumulation fp32); the only remaining differences are accumulation order and
                the final bf16 rounding of the output (2^-9 relative), hence tol 6e-3 of the tensor's max.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BN_EPS = 1e-5

# (weight key, B, H, W, with_residual)
CASES = [
    ("backbone.layer1.0.conv1.weight", 2, 64, 64, False),       # 1x1 64->64   (BN=64 tile)
    ("backbone.layer1.0.conv2.weight", 2, 64, 64, False),       # 3x3 64->64
    ("backbone.layer1.0.conv3.weight", 2, 64, 64, True),        # 1x1 64->256 + residual + relu
    ("backbone.layer1.0.downsample.0.weight", 1, 64, 64, False),
    ("backbone.layer2.0.conv2.weight", 2, 64, 64, False),       # 3x3 stride 2 (TMA elementStrides)
    ("backbone.layer2.0.downsample.0.weight", 2, 64, 64, False),  # 1x1 stride 2
    ("backbone.layer3.0.conv2.weight", 3, 32, 32, False),       # stride 2 -> 16x16
    ("backbone.layer3.0.conv2.weight", 5, 8, 8, False),         # tiny map 4x4: tile spans 8 images, ragged M
    ("backbone.layer4.0.conv2.weight", 3, 16, 16, False),       # stride 2 -> 8x8, odd batch (partial tile)
    ("backbone.layer4.2.conv3.weight", 3, 8, 8, True),          # 1x1 512->2048 + residual
    ("backbone.layer4.2.conv1.weight", 1, 2, 2, False),         # 2x2 map (64x64 image case)
    ("init_regressor.attention_left.0.weight", 2, 8, 8, False),  # 3x3 2048->2x1024, K=18432
    ("decoder.fusion_layer4.conv1.conv.weight", 2, 16, 16, False),  # 1x1 2304->128
    ("decoder.fusion_layer4.conv2.conv.weight", 2, 16, 16, False),  # 3x3 128->128
    ("decoder.fusion_layer4.conv3.conv.weight", 2, 16, 16, True),   # 1x1 + residual, no relu
    ("decoder.fusion_layer4.skip_layer.conv.weight", 1, 16, 16, False),
    ("decoder.projecter_4.fusion.0.weight", 2, 16, 16, False),  # 3x3 2560->256
    ("decoder.projecter_3.fusion.0.weight", 1, 32, 32, False),  # the largest op, K=23040
    ("decoder.projecter_3.fusion.3.weight", 2, 32, 32, False),
    ("decoder.conv_final.0.weight", 2, 32, 32, False),
    ("decoder.seg.0.weight", 2, 32, 32, False),                 # packed seg|dense
]


def epilogue_spec(key):
    """-> list of (weight_key, bias_key|None, bn_prefix|None) per packed part, relu, stride, pad."""
    p = key[: -len("weight")]
    parts = key.split(".")
    if key.startswith("backbone.layer"):
        blk = ".".join(parts[:3]) + "."
        if parts[3] == "downsample":
            stride = 1 if parts[1] == "layer1" else 2
            return [(key, None, blk + "downsample.1.")], False, stride, 0
        n = parts[3][-1]
        stride = 2 if (n == "2" and parts[2] == "0" and parts[1] != "layer1") else 1
        return [(key, None, f"{blk}bn{n}.")], True, stride, 1 if n == "2" else 0
    if ".conv1.conv." in key:
        return [(key, p + "bias", key.split("conv1.conv.")[0] + "bn2.")], True, 1, 0
    if ".conv2.conv." in key:
        return [(key, p + "bias", key.split("conv2.conv.")[0] + "bn3.")], True, 1, 1
    if ".conv3.conv." in key or ".skip_layer.conv." in key:
        return [(key, p + "bias", None)], False, 1, 0
    if key.startswith("init_regressor.attention_left.0"):
        return [(key, p + "bias", "init_regressor.attention_left.1."),
                (key.replace("left", "right"), p.replace("left", "right") + "bias",
                 "init_regressor.attention_right.1.")], True, 1, 1
    if key.endswith("fusion.0.weight"):
        return [(key, p + "bias", key.replace("fusion.0.weight", "fusion.1."))], True, 1, 1
    if key.endswith("fusion.3.weight"):
        return [(key, p + "bias", None)], False, 1, 0
    if key == "decoder.conv_final.0.weight":
        return [(key, None, "decoder.conv_final.1.")], True, 1, 1
    if key == "decoder.seg.0.weight":
        return [(key, "decoder.seg.0.bias", "decoder.seg.1."),
                ("decoder.dense.0.weight", "decoder.dense.0.bias", "decoder.dense.1.")], True, 1, 1
    raise KeyError(key)


def expected(sd, key, x, res, round_bf16, dtype=torch.float32):
    spec, relu, stride, pad = epilogue_spec(key)
    src = sd

    class _Cast(dict):  # the needed tensors converted on access
        def __getitem__(self, k):
            return src[k].to(dtype)

    sd = _Cast()
    x = x.to(dtype)
    res = None if res is None else res.to(dtype)
    rb = (lambda t: t.bfloat16().to(dtype)) if round_bf16 else (lambda t: t)
    outs = []
    for wk, bk, bn in spec:
        y = F.conv2d(rb(x), rb(sd[wk]), None, stride=stride, padding=pad)
        scale = torch.ones(y.shape[1], dtype=dtype)
        shift = torch.zeros(y.shape[1], dtype=dtype)
        if bn:
            scale = sd[bn + "weight"] / torch.sqrt(sd[bn + "running_var"] + BN_EPS)
            shift = sd[bn + "bias"] - sd[bn + "running_mean"] * scale
        if bk:
            shift = shift + sd[bk] * scale
        outs.append(y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    y = torch.cat(outs, 1)
    if res is not None:
        y = y + rb(res)
    return F.relu(y) if relu else y


def _model(synth_sd, precision):
    import dir_b200

    m = dir_b200.DIR(21, "./misc/mano", precision=precision, max_batch=8).cuda()
    m.load_state_dict(synth_sd, strict=False)
    return m


@pytest.fixture(scope="module")
def m32(synth_sd):
    return _model(synth_sd, "fp32")


@pytest.fixture(scope="module")
def m16(synth_sd):
    return _model(synth_sd, "bf16")


@pytest.fixture(scope="module")
def mtf32(synth_sd):
    return _model(synth_sd, "tf32")


def _inputs(synth_sd, key, B, H, W, with_res):
    spec, relu, stride, pad = epilogue_spec(key)
    w = synth_sd[key]
    g = torch.Generator().manual_seed(H * 131 + W * 7 + B + w.shape[0])
    x = torch.relu(torch.randn(B, w.shape[1], H, W, generator=g)) * 1.5
    kh = w.shape[2]
    Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kh) // stride + 1
    cout = sum(synth_sd[s[0]].shape[0] for s in spec)
    res = torch.randn(B, cout, Ho, Wo, generator=g) if with_res else None
    return x, res


@pytest.mark.parametrize("key,B,H,W,with_res", CASES)
def test_conv_fp32_3xtf32_tcgen05(m32, synth_sd, key, B, H, W, with_res):
    from dir_b200 import seams

    x, res = _inputs(synth_sd, key, B, H, W, with_res)
    truth = expected(synth_sd, key, x, res, round_bf16=False, dtype=torch.float64)
    host32 = expected(synth_sd, key, x, res, round_bf16=False)
    got, used = seams.conv_layer(m32, key, x.cuda(), None if res is None else res.cuda())
    dense_fusion = key.endswith("fusion.0.weight")  # only exists as a dense conv for the A/B of the factored form
    assert used == (0 if dense_fusion else 1), "layer did not run on the tcgen05 tf32 kernel"
    err = float((got.cpu().double() - truth).abs().max() / truth.abs().max())
    ref = float((host32.double() - truth).abs().max() / truth.abs().max())
    print(f"{key}: dirb200 fp32 {err:.2e} of max vs float64; torch CPU fp32 conv2d {ref:.2e}")
    assert err < (1e-4 if dense_fusion else 3e-6), err


@pytest.mark.parametrize("key,B,H,W,with_res", CASES)
def test_conv_tf32_tcgen05(mtf32, synth_sd, key, B, H, W, with_res):
    from dir_b200 import seams

    x, res = _inputs(synth_sd, key, B, H, W, with_res)
    truth = expected(synth_sd, key, x, res, round_bf16=False, dtype=torch.float64)
    got, used = seams.conv_layer(mtf32, key, x.cuda(), None if res is None else res.cuda())
    err = float((got.cpu().double() - truth).abs().max() / truth.abs().max())
    print(f"{key}: dirb200 tf32 {err:.2e} of max vs float64")
    assert err < (1e-4 if key.endswith("fusion.0.weight") else 2e-3), err


@pytest.mark.parametrize("key,B,H,W,with_res", CASES)
def test_conv_bf16_tcgen05(m16, synth_sd, key, B, H, W, with_res):
    from dir_b200 import seams

    x, res = _inputs(synth_sd, key, B, H, W, with_res)
    want = expected(synth_sd, key, x, res, round_bf16=True)
    got, used = seams.conv_layer(m16, key, x.cuda(), None if res is None else res.cuda())
    torch.cuda.synchronize()
    assert used == 1, "layer did not run on the tcgen05 kernel"
    err = float((got.cpu() - want).abs().max() / want.abs().max())
    assert err < 6e-3, err


def test_backbone_bf16_tensor_core_stem(m16, synth_sd):
    """256x256 input: the 7x7 stem runs on the tcgen05 stem variant (overlapping-stride TMA view); feature
    pyramid vs the fp32 oracle within bf16 drift."""
    from dir_b200 import seams
    from oracle import dir_oracle as O

    img = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(42))
    want = O.resnet50(synth_sd, img)
    got = seams.backbone(m16, img.cuda())
    for i, (a, b) in enumerate(zip(got, want)):
        err = float((a.cpu() - b).abs().max() / b.abs().max())
        mean = float((a.cpu() - b).abs().mean() / b.abs().mean())
        print(f"c{i + 1}: max-rel {err:.3e} mean-rel {mean:.3e}")
        assert err < 6e-2 and mean < 2e-2, (i, err, mean)


@pytest.mark.parametrize("B,H,W,with_res", [(2, 64, 64, False), (3, 32, 32, True), (5, 16, 16, True), (1, 16, 16, False)])
def test_halo_conv_vs_per_tap_kernel_and_reference(synth_sd, monkeypatch, B, H, W, with_res):
    """conv_halo.cu (one input fetch per tile, chunk-major no-swizzle A operand, three column-masked copies for the
    horizontal padding) against the per-tap TMA kernel it replaces (DIRB200_NO_HALO=1) and against the exact bf16-operand
    model: 64 -> 64 channels, 3x3, maps of 64 / 32 / 16 pixels (2 / 4 / 8 image rows per tile), with and without the
    residual + ReLU epilogue HRNet's BasicBlocks use. Same products, same fp32 accumulation: the two kernels may differ
    by summation order only."""
    from dir_b200 import seams

    key = "backbone.layer1.0.conv2.weight"
    halo = _model(synth_sd, "bf16")
    monkeypatch.setenv("DIRB200_NO_HALO", "1")
    taps = _model(synth_sd, "bf16")
    taps._ensure_handle()  # the handle reads the switch at creation
    monkeypatch.delenv("DIRB200_NO_HALO")
    g = torch.Generator().manual_seed(B * 1000 + H)
    x = torch.relu(torch.randn(B, 64, H, W, generator=g)) * 1.5
    res = torch.randn(B, 64, H, W, generator=g) if with_res else None
    want = expected(synth_sd, key, x, res, round_bf16=True)
    a, ua = seams.conv_layer(halo, key, x.cuda(), None if res is None else res.cuda())
    b, ub = seams.conv_layer(taps, key, x.cuda(), None if res is None else res.cuda())
    assert ua == 1 and ub == 1
    scale = float(want.abs().max())
    assert float((a.cpu() - want).abs().max()) / scale < 6e-3
    assert float((a - b).abs().max()) / scale < 4e-3  # one bf16 ulp of the output at most, from the summation order


RESIDUALS = [("decoder.skip_layer4.", 1024, 16), ("decoder.fusion_layer4.", 2304, 16), ("decoder.enhance_layer4.", 512, 16),
             ("decoder.skip_layer3.", 512, 32), ("decoder.fusion_layer3.", 512, 32), ("decoder.enhance_layer3.", 512, 32)]


@pytest.mark.parametrize("B", [1, 3, 4])
def test_residual_preactivation_folded_into_conv1_is_bit_identical(synth_sd, monkeypatch, B):
    """hourglass.py:60-61 (bn1 + ReLU of a Residual) applied by conv1 to its own A tiles in shared memory (PRE variant of
    conv_tc_kernel: transform warps between the TMA arrival and the MMA; 2-CTA pairs for even tile counts, single CTAs for
    B = 1 and 3 at 16x16) against the separate pre-activation pass it replaces (DIRB200_NO_PREACT_FOLD=1). Same bf16
    input, same fp32 fma + ReLU, same rounding, same MMA order: the block outputs must be bit-identical; and both must
    sit at the bf16-operand distance from the fp32 oracle."""
    from dir_b200 import seams
    from oracle import dir_oracle as O

    fold = _model(synth_sd, "bf16")
    monkeypatch.setenv("DIRB200_NO_PREACT_FOLD", "1")
    sep = _model(synth_sd, "bf16")
    sep._ensure_handle()  # the handle reads the switch at creation
    monkeypatch.delenv("DIRB200_NO_PREACT_FOLD")
    for name, cin, S in RESIDUALS:
        g = torch.Generator().manual_seed(B * 100 + cin + S)
        x = torch.randn(B, cin, S, S, generator=g)
        a = seams.residual(fold, name, x.cuda())
        b = seams.residual(sep, name, x.cuda())
        assert torch.equal(a, b), (name, float((a - b).abs().max()))
        with torch.no_grad():
            want = O.residual(synth_sd, name, x)
        assert float((a.cpu() - want).abs().max() / want.abs().max()) < 3e-2, name


def test_whole_forward_with_and_without_preact_fold_bit_identical(synth_sd, monkeypatch):
    """The forward with every Residual pre-activation folded (4 of the 6 concat_preact launches gone, the other 2 reduced
    to upsample + concat; both sources of enhance_layer{4,3}'s virtual concat pre-activated in flight) equals the forward
    with the separate pass, bit for bit, at an even and an odd batch."""
    fold = _model(synth_sd, "bf16")
    monkeypatch.setenv("DIRB200_NO_PREACT_FOLD", "1")
    sep = _model(synth_sd, "bf16")
    sep._ensure_handle()
    monkeypatch.delenv("DIRB200_NO_PREACT_FOLD")
    for B in (4, 5):
        img = torch.randn(B, 3, 256, 256, generator=torch.Generator().manual_seed(B)).cuda()
        oa = {k: v.clone() for k, v in fold.run_raw(img).items()}
        ob = sep.run_raw(img)
        torch.cuda.synchronize()
        for k in ("record", "mano_para", "seg", "dense", "proj_feat"):
            assert torch.equal(oa[k], ob[k]), (B, k, float((oa[k] - ob[k]).abs().max()))
        for _ in range(25):  # the transform warps race nobody: every repetition reproduces the same bits
            oc = fold.run_raw(img)
            assert torch.equal(oa["record"], oc["record"])
        n_fold = fold._handle.lib.dirb200_forward_launches(fold._handle.h, B)
        n_sep = sep._handle.lib.dirb200_forward_launches(sep._handle.h, B)
        assert n_fold == n_sep - 4, (n_fold, n_sep)


@pytest.mark.parametrize("B", [1, 2, 5])
def test_bottleneck_tail_and_next_conv1_back_to_back_bit_identical(synth_sd, monkeypatch, B):
    """conv_b2b.cu: conv3 (+identity, or the conv3+downsample pair of block 0) of a layer1 bottleneck and conv1 of the
    next block (layer1.1, layer1.2, layer2.0) as back-to-back GEMMs, the 256-channel block output handed from the
    first GEMM's store staging to the second GEMM as its A operand (models/backbone/resnet.py:120-140). Against the
    separate launches (DIRB200_NO_B2B=1): same bf16 operands, same fp32 accumulation order over K, same epilogue
    arithmetic, so c1..c4 must be bit-identical, and three launches disappear."""
    from dir_b200 import seams

    fused = _model(synth_sd, "bf16")
    monkeypatch.setenv("DIRB200_NO_B2B", "1")
    sep = _model(synth_sd, "bf16")
    sep._ensure_handle()
    monkeypatch.delenv("DIRB200_NO_B2B")
    img = torch.randn(B, 3, 256, 256, generator=torch.Generator().manual_seed(40 + B)).cuda()
    a = seams.backbone(fused, img)
    b = seams.backbone(sep, img)
    for i, (u, v) in enumerate(zip(a, b)):
        assert torch.equal(u, v), (f"c{i + 1}", float((u - v).abs().max()), float(v.abs().max()))
    for _ in range(10):  # no race between the store staging, the second GEMM and the next tile's epilogue
        c = seams.backbone(fused, img)
        assert all(torch.equal(u, v) for u, v in zip(a, c))
    oa = fused.run_raw(img)["record"].clone()
    ob = sep.run_raw(img)["record"]
    assert torch.equal(oa, ob)
    n_f = fused._handle.lib.dirb200_forward_launches(fused._handle.h, B)
    n_s = sep._handle.lib.dirb200_forward_launches(sep._handle.h, B)
    assert n_f == n_s - 3, (n_f, n_s)
