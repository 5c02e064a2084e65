"""Per-seam entry points of the C ABI, mirroring the reference sub-modules' call
signatures so the parity tests read like calls into the reference:

    ResNet.forward, Residual.forward, InitRegressor.forward, ManoLayer.forward (+projection),
    Joint2BoneFeature.forward, Joint2BoneFeature.bone_proj

All tensors are fp32 CUDA tensors in the reference's own layouts (NCHW feature maps).
"""
import ctypes as C

import torch

from . import capi


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _prep(model, B):
    model._ensure_handle()
    if not model._packed:
        model._pack()
    ws = model._workspace_for(max(B, 1))
    return model._handle, ws


def _f32(t):
    return t.detach().to(dtype=torch.float32).contiguous()


def stage_dict(rec, para=None):
    """(B,4887) stage slice -> dict with the reference's result keys."""
    out = {}
    for key, off, shp in __import__("dir_b200.module", fromlist=["STAGE_KEYS"]).STAGE_KEYS:
        a = capi.OFF[off]
        n = 1
        for s in shp:
            n *= s
        out[key] = rec[:, a:a + n].reshape(rec.shape[0], *shp)
    if para is not None:
        out["pd_mano_para_left"] = para[:, 0]
        out["pd_mano_para_right"] = para[:, 1]
    return out


def pack_prev(prev, device):
    """Build the (B,4887) record slice + (B,2,64) para a stage consumes from a dict with the reference's keys
    (pd_joint_xyz_*, pd_joint_uv_*, pd_mano_para_*, pd_offset) — models/dir.py:446-454."""
    B = prev["pd_offset"].shape[0]
    rec = torch.zeros(B, capi.STAGE_FLOATS, device=device)
    rec[:, capi.OFF["joint_l"]:capi.OFF["joint_l"] + 63] = prev["pd_joint_xyz_left"].reshape(B, -1)
    rec[:, capi.OFF["joint_r"]:capi.OFF["joint_r"] + 63] = prev["pd_joint_xyz_right"].reshape(B, -1)
    rec[:, capi.OFF["uv_l"]:capi.OFF["uv_l"] + 42] = prev["pd_joint_uv_left"].reshape(B, -1)
    rec[:, capi.OFF["uv_r"]:capi.OFF["uv_r"] + 42] = prev["pd_joint_uv_right"].reshape(B, -1)
    rec[:, capi.OFF["offset"]:capi.OFF["offset"] + 3] = prev["pd_offset"].reshape(B, 3)
    para = torch.stack((prev["pd_mano_para_left"], prev["pd_mano_para_right"]), 1).to(device).contiguous()
    return rec, para


def backbone(model, img):
    img = _f32(img)
    B, _, H, W = img.shape
    h, ws = _prep(model, B)
    dev = img.device
    # packed channel counts: ResNet-50, or HRNet-W32 / -W48 (branch widths that are not multiples of 64 are stored zero-
    # padded: W32's 32-channel branch in 64, W48's 48 in 64 and 96 in 128)
    ch = {"hrnet_w32": (64, 64, 128, 256), "hrnet_w48": (64, 128, 192, 384)}.get(
        getattr(model, "backbone_name", "resnet50"), (256, 512, 1024, 2048))
    c1 = torch.empty(B, ch[0], H // 4, W // 4, device=dev)
    c2 = torch.empty(B, ch[1], H // 8, W // 8, device=dev)
    c3 = torch.empty(B, ch[2], H // 16, W // 16, device=dev)
    c4 = torch.empty(B, ch[3], H // 32, W // 32, device=dev)
    h.check(h.lib.dirb200_backbone(h.h, _ptr(img), B, H, W, _ptr(c1), _ptr(c2), _ptr(c3), _ptr(c4), _ptr(ws),
                                   ws.numel(), _stream()), "backbone")
    return [c1, c2, c3, c4]


def residual(model, name, x):
    x = _f32(x)
    B, Cin, H, W = x.shape
    h, ws = _prep(model, B)
    cout = model.state_dict()[name + "conv3.conv.weight"].shape[0]
    y = torch.empty(B, cout, H, W, device=x.device)
    h.check(h.lib.dirb200_residual(h.h, name.encode(), _ptr(x), B, Cin, H, W, _ptr(y), _ptr(ws), ws.numel(),
                                   _stream()), "residual")
    return y


def init_regressor(model, c4):
    c4 = _f32(c4)
    B = c4.shape[0]
    h, ws = _prep(model, B)
    rec = torch.zeros(B, capi.STAGE_FLOATS, device=c4.device)
    para = torch.zeros(B, 2, 64, device=c4.device)
    h.check(h.lib.dirb200_init_regressor(h.h, _ptr(c4), B, _ptr(rec), _ptr(para), _ptr(ws), ws.numel(), _stream()),
            "init_regressor")
    return stage_dict(rec, para)


def mano(model, which, para):
    """para (B,2,64) -> stage dict (verts/joints/uv for both hands)."""
    para = _f32(para)
    B = para.shape[0]
    h, _ = _prep(model, B)
    rec = torch.zeros(B, capi.STAGE_FLOATS, device=para.device)
    h.check(h.lib.dirb200_mano(h.h, which, _ptr(para), B, _ptr(rec), _stream()), "mano")
    return stage_dict(rec)


def joint2bone(model, stage, img_feat, prev, want_vis=False):
    img_feat = _f32(img_feat)
    B, _, S, _ = img_feat.shape
    h, ws = _prep(model, B)
    dev = img_feat.device
    prev_rec, prev_para = pack_prev(prev, dev)
    rec = torch.zeros(B, capi.STAGE_FLOATS, device=dev)
    para = torch.zeros(B, 2, 64, device=dev)
    out_feat = torch.empty(B, 256, S, S, device=dev)
    jf = torch.empty(B, 2, 21, 64, device=dev)
    vis = torch.empty(B, 1280, S, S, device=dev) if want_vis else None
    h.check(h.lib.dirb200_joint2bone(h.h, stage, _ptr(img_feat), _ptr(prev_rec), _ptr(prev_para), B, _ptr(rec),
                                     _ptr(para), _ptr(out_feat), _ptr(jf), _ptr(vis), _ptr(ws), ws.numel(),
                                     _stream()), "joint2bone")
    feats = {"img_feat": out_feat, "joint_feat_left": jf[:, 0], "joint_feat_right": jf[:, 1], "vis_img_feat": vis}
    return stage_dict(rec, para), feats


def img2joint(model, stage, img_feat, uv_left, uv_right):
    """ImgFeature2JointFeature.forward of both hands (models/dir.py:197-200) -> two (B,21,128) tensors."""
    img_feat, uv_left, uv_right = _f32(img_feat), _f32(uv_left), _f32(uv_right)
    B = img_feat.shape[0]
    h, ws = _prep(model, B)
    out = [torch.empty(B, 21, 128, device=img_feat.device) for _ in range(2)]
    h.check(h.lib.dirb200_img2joint(h.h, stage, _ptr(img_feat), _ptr(uv_left), _ptr(uv_right), B, _ptr(out[0]),
                                    _ptr(out[1]), _ptr(ws), ws.numel(), _stream()), "img2joint")
    return out


def gcn(model, stage, x_left, x_right):
    """ResSimplePGCN.forward of gcn_left / gcn_right (SemGCN/p_gcn.py:63-73): (B,21,128) -> (B,21,128) per hand."""
    x_left, x_right = _f32(x_left), _f32(x_right)
    B = x_left.shape[0]
    h, ws = _prep(model, B)
    out = [torch.empty(B, 21, 128, device=x_left.device) for _ in range(2)]
    h.check(h.lib.dirb200_gcn(h.h, stage, _ptr(x_left), _ptr(x_right), B, _ptr(out[0]), _ptr(out[1]), _ptr(ws),
                              ws.numel(), _stream()), "gcn")
    return out


def ste(model, stage, x):
    """STE.forward (transformer/mixSTE.py:194-205): (B,42,128) -> (B,42,64)."""
    x = _f32(x)
    B = x.shape[0]
    h, _ = _prep(model, B)
    y = torch.empty(B, 42, 64, device=x.device)
    h.check(h.lib.dirb200_ste(h.h, stage, _ptr(x), B, _ptr(y), _stream()), "ste")
    return y


def regressor_offset(model, stage, feat_left, feat_right, para_left, para_right, offset):
    """RegressorOffset.forward (models/dir.py:339-381) -> stage dict incl. pd_mano_para_*."""
    args = [_f32(t) for t in (feat_left, feat_right, para_left, para_right, offset.reshape(offset.shape[0], 3))]
    B = args[0].shape[0]
    h, ws = _prep(model, B)
    rec = torch.zeros(B, capi.STAGE_FLOATS, device=args[0].device)
    para = torch.zeros(B, 2, 64, device=args[0].device)
    h.check(h.lib.dirb200_regressor_offset(h.h, stage, *[_ptr(a) for a in args], B, _ptr(rec), _ptr(para), _ptr(ws),
                                           ws.numel(), _stream()), "regressor_offset")
    return stage_dict(rec, para)


def bone_fusion(model, stage, uv_left, uv_right, feat_left, feat_right):
    """bone_proj x2 -> cat -> fusion (models/dir.py:118-122): uv (B,21,2), joint features (B,21,64) -> (B,256,S,S)."""
    args = [_f32(t) for t in (uv_left, uv_right, feat_left, feat_right)]
    B = args[0].shape[0]
    h, ws = _prep(model, B)
    S = 16 if stage == 1 else 32
    out = torch.empty(B, 256, S, S, device=args[0].device)
    h.check(h.lib.dirb200_bone_fusion(h.h, stage, *[_ptr(a) for a in args], B, _ptr(out), _ptr(ws), ws.numel(),
                                      _stream()), "bone_fusion")
    return out


def bone_proj(model, uv, feat, size, distance):
    uv, feat = _f32(uv), _f32(feat)
    B = uv.shape[0]
    h, _ = _prep(model, B)
    out = torch.empty(B, 1280, size, size, device=uv.device)
    h.check(h.lib.dirb200_bone_proj(h.h, _ptr(uv), _ptr(feat), B, size, float(distance), _ptr(out), _stream()),
            "bone_proj")
    return out


def conv_layer(model, weight_key, x, res=None):
    """One conv with its fused epilogue, addressed by the reference state_dict key of its weight.
    Returns (y NCHW fp32, used_tensor_cores)."""
    x = _f32(x)
    B, _, H, W = x.shape
    h, ws = _prep(model, B)
    w = model.state_dict()[weight_key]
    cout, _, kh, kw = w.shape
    parts = weight_key.split(".")
    stride, pad = 1, (kh - 1) // 2
    if weight_key == "backbone.conv1.weight":
        stride = 2
    elif parts[0] == "backbone" and parts[2] == "0" and parts[1] != "layer1" and (parts[3] in ("conv2", "downsample")):
        stride = 2  # resnet.py:111,227 (v1.5: stride on the 3x3 and on the downsample 1x1)
    Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    if weight_key.startswith(("init_regressor.attention_left.0", "decoder.seg.0")):
        cout *= 2  # packed together with its right/dense twin
    y = torch.empty(B, cout, Ho, Wo, device=x.device)
    if res is not None:
        res = _f32(res)
    used = C.c_int(0)
    h.check(h.lib.dirb200_conv_layer(h.h, weight_key.encode(), _ptr(x), _ptr(res), B, H, W, _ptr(y), C.byref(used),
                                     _ptr(ws), ws.numel(), _stream()), "conv_layer")
    return y, used.value


def preprocess_u8(model, frames):
    """apps/eval.py:56-61 on the device: (B,H,W,3) uint8 BGR -> (B,3,H,W) fp32 normalised RGB."""
    frames = frames.contiguous()
    B, H, W, _ = frames.shape
    h, _ = _prep(model, B)
    out = torch.empty(B, 3, H, W, device=frames.device)
    h.check(h.lib.dirb200_preprocess_u8(h.h, _ptr(frames), B, H, W, _ptr(out), _stream()), "preprocess_u8")
    return out


def eval_jregressor(jreg16):
    """class Jr of apps/eval.py:22-44: (16,778) MANO joint regressor -> (21,778) incl. the 5 tip vertices, reordered."""
    tips = torch.zeros(5, 778, device=jreg16.device)
    for i, v in enumerate([745, 317, 444, 556, 673]):
        tips[i, v] = 1.0
    order = [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]
    return torch.cat([jreg16.float(), tips], 0)[order].contiguous()


def eval_metrics(model, record, gt_verts, gt_verts2d, cam, jreg21, scale=True):
    """apps/eval.py:151-241 on the device. record (B,14661); gt_verts (B,2,778,3); gt_verts2d (B,2,778,2);
    cam (B,3,3); jreg21 (2,21,778). Returns dict of per-sample error tensors (metres / pixels)."""
    B = record.shape[0]
    h, _ = _prep(model, B)
    dev = record.device
    args = [_f32(t) for t in (record, gt_verts, gt_verts2d, cam, jreg21)]
    je, j2 = torch.empty(B, 2, 21, device=dev), torch.empty(B, 2, 21, device=dev)
    ve, v2 = torch.empty(B, 2, 778, device=dev), torch.empty(B, 2, 778, device=dev)
    re = torch.empty(B, device=dev)
    h.check(h.lib.dirb200_eval_metrics(h.h, *[_ptr(a) for a in args], B, int(bool(scale)), _ptr(je), _ptr(ve), _ptr(j2),
                                       _ptr(v2), _ptr(re), _stream()), "eval_metrics")
    return {"joint_left": je[:, 0], "joint_right": je[:, 1], "vert_left": ve[:, 0], "vert_right": ve[:, 1],
            "joint2d_left": j2[:, 0], "joint2d_right": j2[:, 1], "vert2d_left": v2[:, 0], "vert2d_right": v2[:, 1],
            "root": re}
