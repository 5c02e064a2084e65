"""CPU oracle: a functional fp32 restatement of DIR's eval-mode forward.

TEST INFRASTRUCTURE ONLY. Importers allowed: tests/, __graft_entry__.smoke(), and
bench.py's cpu_baseline / `--impl reference` legs. The product (dir_b200/) never
imports this file; it fails loudly when its CUDA library is missing.

Parity pinning: the reference ships no golden vectors for this path,
so this restatement is pinned against outputs of the UNMODIFIED reference executed
in the build container (oracle/gen_golden.py -> tests/golden/*.npz); see
tests/test_oracle_golden.py. Floating-point path => torch fp32 ops are used for
the dense contractions (conv2d / matmul), everything else is spelled out.

Every function cites the reference lines (relative to /root/reference) it restates.
State is a flat dict `sd` with the reference's state_dict keys.
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BONE_PARENT = [0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11, 0, 13, 14, 15, 0, 17, 18, 19]  # models/dir.py:25
BONE_CHILD = list(range(1, 21))  # models/dir.py:26
SKELETON_EDGES = [[0, 1], [1, 2], [2, 3], [3, 4], [0, 5], [5, 6], [6, 7], [7, 8], [0, 9], [9, 10], [10, 11],
                  [11, 12], [0, 13], [13, 14], [14, 15], [15, 16], [0, 17], [17, 18], [18, 19], [19, 20]]
FK_PARENTS = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14]
JOINT_REORDER = [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]  # manolayer.py:259


# --------------------------------------------------------------------------- basic blocks
def bn_affine(sd, p):
    """Eval-mode BatchNorm as y = x*scale + shift (torch.nn.BatchNorm*, eps 1e-5)."""
    scale = sd[p + "weight"] / torch.sqrt(sd[p + "running_var"] + BN_EPS)
    shift = sd[p + "bias"] - sd[p + "running_mean"] * scale
    return scale, shift


def bn2d(sd, p, x):
    s, b = bn_affine(sd, p)
    return x * s.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def bn1d(sd, p, x):  # x: (B, C, L)
    s, b = bn_affine(sd, p)
    return x * s.view(1, -1, 1) + b.view(1, -1, 1)


def conv(sd, p, x, stride=1, pad=0):
    return F.conv2d(x, sd[p + "weight"], sd.get(p + "bias"), stride=stride, padding=pad)


# --------------------------------------------------------------------------- backbone
def bottleneck(sd, p, x, stride):
    """models/backbone/resnet.py:120-140 (v1.5: stride on the 3x3)."""
    out = F.relu(bn2d(sd, p + "bn1.", conv(sd, p + "conv1.", x)))
    out = F.relu(bn2d(sd, p + "bn2.", conv(sd, p + "conv2.", out, stride=stride, pad=1)))
    out = bn2d(sd, p + "bn3.", conv(sd, p + "conv3.", out))
    if (p + "downsample.0.weight") in sd:
        x = bn2d(sd, p + "downsample.1.", conv(sd, p + "downsample.0.", x, stride=stride))
    return F.relu(out + x)


def resnet50(sd, x, p="backbone."):
    """models/backbone/resnet.py:243-255 -> [c1, c2, c3, c4]."""
    x = F.relu(bn2d(sd, p + "bn1.", conv(sd, p + "conv1.", x, stride=2, pad=3)))
    x = F.max_pool2d(x, 3, 2, 1)
    feats = []
    for li, nblocks in enumerate([3, 4, 6, 3], start=1):
        for b in range(nblocks):
            stride = 2 if (b == 0 and li > 1) else 1
            x = bottleneck(sd, f"{p}layer{li}.{b}.", x, stride)
        feats.append(x)
    return feats


def residual(sd, p, x):
    """models/backbone/hourglass.py:55-70 (pre-activation bottleneck, biased convs)."""
    if (p + "skip_layer.conv.weight") in sd and sd[p + "skip_layer.conv.weight"].shape[0] != x.shape[1]:
        res = conv(sd, p + "skip_layer.conv.", x)
    else:
        res = x
    out = conv(sd, p + "conv1.conv.", F.relu(bn2d(sd, p + "bn1.", x)))
    out = conv(sd, p + "conv2.conv.", F.relu(bn2d(sd, p + "bn2.", out)), pad=1)
    out = conv(sd, p + "conv3.conv.", F.relu(bn2d(sd, p + "bn3.", out)))
    return out + res


def upsample2x(x):
    """nn.Upsample(scale_factor=2, mode='bilinear') => align_corners=False (models/dir.py:392,398).
    out[i] samples src coordinate (i+0.5)/2-0.5, clamped at 0 from below, neighbours clamped at the edge."""
    B, C, H, W = x.shape

    def idx(n):
        o = torch.arange(2 * n, dtype=x.dtype, device=x.device)
        s = torch.clamp((o + 0.5) * 0.5 - 0.5, min=0.0)
        i0 = s.floor().long()
        i1 = torch.clamp(i0 + 1, max=n - 1)
        w1 = s - i0.to(x.dtype)
        return i0, i1, w1

    y0, y1, wy = idx(H)
    x0, x1, wx = idx(W)
    top = x[:, :, y0, :]
    bot = x[:, :, y1, :]
    rows = top * (1 - wy).view(1, 1, -1, 1) + bot * wy.view(1, 1, -1, 1)
    return rows[:, :, :, x0] * (1 - wx).view(1, 1, 1, -1) + rows[:, :, :, x1] * wx.view(1, 1, 1, -1)


# --------------------------------------------------------------------------- MANO (manopth)
def _normalize(v):
    """manopth/manopth/rot6d.py:54-60 (norm clamped at 1e-8)."""
    return v / torch.clamp(torch.sqrt((v * v).sum(1, keepdim=True)), min=1e-8)


def rot6d_robust(p6):
    """manopth/manopth/rot6d.py:26-51 -> (B,3,3) with columns x,y,z."""
    x = _normalize(p6[:, 0:3])
    y = _normalize(p6[:, 3:6])
    mid = _normalize(x + y)
    orth = _normalize(x - y)
    x = _normalize(mid + orth)
    y = _normalize(mid - orth)
    z = _normalize(torch.cross(x, y, dim=1))
    return torch.stack((x, y, z), dim=2)


def rodrigues(aa):
    """manopth/manopth/rodrigues_layer.py:15-54: axis-angle (N,3) -> (N,3,3) via quaternion."""
    theta = torch.sqrt(((aa + 1e-8) ** 2).sum(1, keepdim=True))
    axis = aa / theta
    half = theta * 0.5
    q = torch.cat([torch.cos(half), torch.sin(half) * axis], dim=1)
    q = q / torch.sqrt((q * q).sum(1, keepdim=True))
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    R = torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                     2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                     2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1)
    return R.view(-1, 3, 3)


def mano_layer(sd, p, pose, betas, side, center_idx=0):
    """manopth/manopth/manolayer.py:110-270 with root_rot_mode='6D', use_pca, ncomps=45,
    flat_hand_mean=False, robust_rot=True (models/dir.py:221-224).
    pose (B,51) = [6D root | 45 PCA coeffs], betas (B,10) -> verts (B,778,3), joints (B,21,3) metres."""
    B = pose.shape[0]
    comps, mean = sd[p + "th_selected_comps"], sd[p + "th_hands_mean"]
    shapedirs, posedirs = sd[p + "th_shapedirs"], sd[p + "th_posedirs"]
    v_template, jreg, skin_w = sd[p + "th_v_template"], sd[p + "th_J_regressor"], sd[p + "th_weights"]

    aa = mean + pose[:, 6:51] @ comps  # :124-133
    R = rodrigues(aa.reshape(-1, 3)).view(B, 15, 3, 3)  # tensutils.py:6-12
    pose_map = (R - torch.eye(3, dtype=pose.dtype, device=pose.device)).reshape(B, 135)
    R_root = rot6d_robust(pose[:, :6])

    v_shaped = torch.einsum("vck,bk->bvc", shapedirs, betas) + v_template  # :173-176
    J = torch.einsum("jv,bvc->bjc", jreg, v_shaped)  # :177
    v_posed = v_shaped + torch.einsum("vck,bk->bvc", posedirs, pose_map)  # :180-181

    # forward kinematics, 3 levels (:186-227), expressed per joint with FK_PARENTS
    G_R = [None] * 16
    G_t = [None] * 16
    G_R[0], G_t[0] = R_root, J[:, 0]
    for j in range(1, 16):
        par = FK_PARENTS[j]
        G_R[j] = G_R[par] @ R[:, j - 1]
        G_t[j] = (G_R[par] @ (J[:, j] - J[:, par]).unsqueeze(-1)).squeeze(-1) + G_t[par]
    G_R = torch.stack(G_R, 1)
    G_t = torch.stack(G_t, 1)  # (B,16,3) = posed joints
    # A_j = G_j - [0 | G_j[:3,:3] J_j]  (:229-231)
    A_t = G_t - (G_R @ J.unsqueeze(-1)).squeeze(-1)
    # LBS (:233-244)
    T_R = torch.einsum("vj,bjmn->bvmn", skin_w, G_R)
    T_t = torch.einsum("vj,bjm->bvm", skin_w, A_t)
    verts = (T_R @ v_posed.unsqueeze(-1)).squeeze(-1) + T_t
    tips_idx = [745, 317, 444, 556, 673] if side == "right" else [745, 317, 445, 556, 673]  # :249-252
    joints = torch.cat([G_t, verts[:, tips_idx]], 1)[:, JOINT_REORDER]
    if center_idx is not None:  # :261-265
        c = joints[:, center_idx:center_idx + 1]
        joints = joints - c
        verts = verts - c
    return verts, joints


def projection_xy(para, xyz):
    """utils/utils.py:47-63 with scale=para[:,0], trans=para[:,1:]."""
    return para[:, 0].view(-1, 1, 1) * xyz[..., :2] + para[:, 1:3].unsqueeze(1)


# --------------------------------------------------------------------------- init regressor
def init_regressor(sd, c4, p="init_regressor."):
    """models/dir.py:260-305."""
    out = {}
    pooled = {}
    for side in ("left", "right"):
        a = conv(sd, f"{p}attention_{side}.0.", c4, pad=1)
        a = F.relu(bn2d(sd, f"{p}attention_{side}.1.", a))
        a = torch.sigmoid(conv(sd, f"{p}attention_{side}.3.", a))
        pooled[side] = (c4 * a).sum(-1).sum(-1) / (a.sum(-1).sum(-1) + 1e-8)
    out["pd_offset"] = F.linear(c4.mean(-1).mean(-1), sd[p + "offset.weight"], sd[p + "offset.bias"])
    for side in ("left", "right"):
        para = F.linear(pooled[side], sd[f"{p}mano_{side}.weight"], sd[f"{p}mano_{side}.bias"])
        _finish_hand(sd, f"{p}mano_layer_{side}.", para, side, out)
    return out


def _finish_hand(sd, mano_prefix, para, side, out):
    pose, beta, proj = para[:, :51], para[:, 51:61], para[:, 61:64]  # split [51,10,3] dir.py:272
    verts, joints = mano_layer(sd, mano_prefix, pose, beta, side)
    out[f"pd_mano_para_{side}"] = para
    out[f"pd_proj_{side}"] = proj
    out[f"pd_mesh_xyz_{side}"] = verts
    out[f"pd_joint_xyz_{side}"] = joints
    out[f"pd_joint_uv_{side}"] = projection_xy(proj, joints)
    out[f"pd_mesh_uv_{side}"] = projection_xy(proj, verts)


# --------------------------------------------------------------------------- joint space
def grid_sample_joints(feat, uv):
    """F.grid_sample(feat, uv[:,None]) with defaults bilinear / zeros / align_corners=False
    (models/dir.py:198). feat (B,C,H,W), uv (B,J,2) with x=col, y=row in [-1,1] -> (B,C,J)."""
    B, C, H, W = feat.shape
    x = ((uv[..., 0] + 1) * W - 1) * 0.5
    y = ((uv[..., 1] + 1) * H - 1) * 0.5
    x0, y0 = torch.floor(x), torch.floor(y)
    out = torch.zeros(B, C, uv.shape[1], dtype=feat.dtype, device=feat.device)
    bidx = torch.arange(B, device=feat.device).view(B, 1).expand(B, uv.shape[1])
    for dy in (0, 1):
        for dx in (0, 1):
            xi, yi = x0 + dx, y0 + dy
            w = (1 - (x - xi).abs()) * (1 - (y - yi).abs())
            ok = (xi >= 0) & (xi <= W - 1) & (yi >= 0) & (yi <= H - 1)
            xi_c, yi_c = xi.clamp(0, W - 1).long(), yi.clamp(0, H - 1).long()
            v = feat[bidx, :, yi_c, xi_c]  # (B,J,C)
            out = out + (v * (w * ok).unsqueeze(-1)).permute(0, 2, 1)
    return out


def pointwise_mlp(sd, p, x):
    """Conv1d(k=1) -> BN1d -> ReLU -> Conv1d(k=1) on (B, Cin, L) (models/dir.py:31-56,180-185)."""
    h = F.conv1d(x, sd[p + "0.weight"], sd[p + "0.bias"])
    h = F.relu(bn1d(sd, p + "1.", h))
    return F.conv1d(h, sd[p + "3.weight"], sd[p + "3.bias"])


def img2joint(sd, p, feat, uv):
    """models/dir.py:197-200 + the reshape/permute at :94 -> (B,21,128)."""
    return pointwise_mlp(sd, p + "filters.", grid_sample_joints(feat, uv)).permute(0, 2, 1)


def gcn_adjacency_logits_index():
    """Row-major nonzero positions of the symmetric skeleton adjacency without self loops
    (SemGCN/utils.py:27-43 with eye=False; SemGCN/p_graph_conv.py:26-29)."""
    adj = torch.zeros(21, 21)
    for a, b in SKELETON_EDGES:
        adj[a, b] = 1
        adj[b, a] = 1
    return adj.nonzero()  # (40,2) row-major


def gcn_softmax_adjacency(e1):
    """A_1 = softmax over each row of the 40 learned edge logits, -9e15 elsewhere
    (SemGCN/p_graph_conv.py:43-50). A_0 = softmax(diag logits) == I exactly."""
    nz = gcn_adjacency_logits_index()
    A = torch.full((21, 21), -9e15, dtype=e1.dtype, device=e1.device)
    A[nz[:, 0], nz[:, 1]] = e1.flatten()
    return torch.softmax(A, dim=1)


def pgraph_conv(sd, p, x):
    """SemGCN/p_graph_conv.py:39-60. x (B,21,128)."""
    W = sd[p + "W"]
    h0 = torch.einsum("bjc,jcd->bjd", x, W[0])
    h1 = torch.einsum("bjc,jcd->bjd", x, W[1])
    A1 = gcn_softmax_adjacency(sd[p + "e_1"])
    return h0 + torch.einsum("ij,bjd->bid", A1, h1) + sd[p + "bias"].view(1, 1, -1)


def gcn_stack(sd, p, x):
    """SemGCN/p_gcn.py:20-27,63-73: 4 x (PGraphConv -> BN1d -> ReLU), no residual."""
    for l in range(4):
        q = f"{p}gconv_layers.{l}."
        x = pgraph_conv(sd, q + "gconv.", x)
        s, b = bn_affine(sd, q + "bn.")
        x = F.relu(x * s + b)
    return x


def layer_norm(x, w, b, eps):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu_erf(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def ste(sd, p, x):
    """transformer/mixSTE.py:194-205: blocks 1..3 only, shared spatial_norm after every block,
    block LNs eps 1e-6 (:177), head LN eps 1e-5 (:190). x (B,42,128) -> (B,42,64)."""
    B, N, C = x.shape
    H, D = 4, C // 4
    x = x + sd[p + "spatial_pos_embed"]
    for i in (1, 2, 3):
        q = f"{p}STEblocks.{i}."
        h = layer_norm(x, sd[q + "norm1.weight"], sd[q + "norm1.bias"], 1e-6)
        qkv = F.linear(h, sd[q + "attn.qkv.weight"], sd[q + "attn.qkv.bias"]).view(B, N, 3, H, D)
        qh, kh, vh = (qkv[:, :, j].permute(0, 2, 1, 3) for j in range(3))  # (B,H,N,D)
        att = torch.softmax((qh @ kh.transpose(-1, -2)) * (D ** -0.5), dim=-1)
        o = (att @ vh).permute(0, 2, 1, 3).reshape(B, N, C)
        x = x + F.linear(o, sd[q + "attn.proj.weight"], sd[q + "attn.proj.bias"])
        h = layer_norm(x, sd[q + "norm2.weight"], sd[q + "norm2.bias"], 1e-6)
        h = gelu_erf(F.linear(h, sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"]))
        x = x + F.linear(h, sd[q + "mlp.fc2.weight"], sd[q + "mlp.fc2.bias"])
        x = layer_norm(x, sd[p + "spatial_norm.weight"], sd[p + "spatial_norm.bias"], 1e-6)
    h = layer_norm(x, sd[p + "head.0.weight"], sd[p + "head.0.bias"], 1e-5)
    return F.linear(h, sd[p + "head.1.weight"], sd[p + "head.1.bias"])


def regressor_offset(sd, p, feat_l, feat_r, para_l, para_r, offset):
    """models/dir.py:339-381. feat (B,21,64), para (B,64), offset (B,3)."""
    B = feat_l.shape[0]
    fl, fr = feat_l.reshape(B, -1), feat_r.reshape(B, -1)
    out = {"pd_offset": F.linear(torch.cat((fl, fr, offset), -1), sd[p + "offset.weight"], sd[p + "offset.bias"])}
    for side, f, prev in (("left", fl, para_l), ("right", fr, para_r)):
        para = F.linear(torch.cat((f, prev), -1), sd[f"{p}mano_{side}.weight"], sd[f"{p}mano_{side}.bias"])
        _finish_hand(sd, f"{p}mano_layer_{side}.", para, side, out)
    return out


def bone_proj(uv, feat, S, distance):
    """models/dir.py:132-174: rasterise the 20 bone capsules.
    uv (B,21,2) in [-1,1], feat (B,21,C) -> (B, 20*C, S, S), channel = bone*C + c."""
    B, J, C = feat.shape
    p = (uv + 1) / 2 * S
    a = p[:, BONE_PARENT].unsqueeze(1)  # (B,1,20,2)
    b = p[:, BONE_CHILD].unsqueeze(1)
    r = torch.arange(S, dtype=uv.dtype, device=uv.device) + 0.5
    gy, gx = torch.meshgrid(r, r, indexing="ij")  # pixel (row, col) -> P = (col+.5, row+.5)
    zero = torch.zeros((), dtype=uv.dtype, device=uv.device)
    P = torch.stack((gx, gy), -1).reshape(1, S * S, 1, 2)
    dba = b - a
    d = dba / torch.hypot(dba[..., 0], dba[..., 1]).unsqueeze(-1)  # NaN when a == b
    s = ((a - P) * d).sum(-1)
    t = ((P - b) * d).sum(-1)
    h = torch.maximum(torch.maximum(s, t), zero)
    dpa = P - a
    c = dpa[..., 0] * d[..., 1] - dpa[..., 1] * d[..., 0]
    mask = torch.hypot(h, c) < distance  # NaN -> False
    da = torch.sqrt(((P - a + 1e-6) ** 2).sum(-1))  # F.pairwise_distance eps on the difference
    db = torch.sqrt(((P - b + 1e-6) ** 2).sum(-1))
    wa = 1 - da / (da + db)
    wb = 1 - db / (da + db)
    fa = feat[:, BONE_PARENT].unsqueeze(1)  # (B,1,20,C)
    fb = feat[:, BONE_CHILD].unsqueeze(1)
    img = fa * wa.unsqueeze(-1) + fb * wb.unsqueeze(-1)
    img = torch.where(mask.unsqueeze(-1), img, zero)
    return img.reshape(B, S, S, 20 * C).permute(0, 3, 1, 2)


def bone_capsule_margin(uv, S, distance):
    """Distance of the closest (pixel, bone) pair to the capsule boundary of bone_proj, per image: min |hypot(h,c) - distance|
    (same arithmetic as bone_proj above). The mask `hypot(h,c) < distance` (models/dir.py:164) is the one discontinuity
    of the forward: an image whose margin is below the arithmetic's noise may legitimately flip a pixel."""
    p = (uv + 1) / 2 * S
    a = p[:, BONE_PARENT].unsqueeze(1)
    b = p[:, BONE_CHILD].unsqueeze(1)
    r = torch.arange(S, dtype=uv.dtype, device=uv.device) + 0.5
    gy, gx = torch.meshgrid(r, r, indexing="ij")
    P = torch.stack((gx, gy), -1).reshape(1, S * S, 1, 2)
    dba = b - a
    d = dba / torch.hypot(dba[..., 0], dba[..., 1]).unsqueeze(-1)
    s = ((a - P) * d).sum(-1)
    t = ((P - b) * d).sum(-1)
    h = torch.maximum(torch.maximum(s, t), torch.zeros((), dtype=uv.dtype, device=uv.device))
    dpa = P - a
    c = dpa[..., 0] * d[..., 1] - dpa[..., 1] * d[..., 0]
    m = (torch.hypot(h, c) - distance).abs()
    m = torch.where(torch.isnan(m), torch.full_like(m, float("inf")), m)
    return m.flatten(1).min(1).values


def joint2bone(sd, p, img_feat, prev, S, distance):
    """models/dir.py:86-130. `prev` holds pd_joint_xyz_*, pd_joint_uv_*, pd_mano_para_*, pd_offset."""
    offset = prev["pd_offset"].unsqueeze(1)  # (B,1,3)
    joint_feat = []
    for side, sgn in (("left", -1.0), ("right", 1.0)):
        xyz = prev[f"pd_joint_xyz_{side}"] / 0.15
        f = img2joint(sd, f"{p}img2joint_{side}.", img_feat, prev[f"pd_joint_uv_{side}"])
        f = f + pointwise_mlp(sd, f"{p}pos_emb_{side}.", xyz.permute(0, 2, 1)).permute(0, 2, 1)
        f = gcn_stack(sd, f"{p}gcn_{side}.", f)
        g = pointwise_mlp(sd, f"{p}global_pos_emb.", (xyz + sgn * offset / 2).permute(0, 2, 1)).permute(0, 2, 1)
        joint_feat.append(f + g)
    tok = ste(sd, p + "interaction.", torch.cat(joint_feat, 1))
    feat_l, feat_r = tok[:, :21], tok[:, 21:]
    res = regressor_offset(sd, p + "regressor.", feat_l, feat_r, prev["pd_mano_para_left"],
                           prev["pd_mano_para_right"], prev["pd_offset"])
    jf_l = pointwise_mlp(sd, p + "proj_feat_emb.", feat_l.permute(0, 2, 1)).permute(0, 2, 1)
    jf_r = pointwise_mlp(sd, p + "proj_feat_emb.", feat_r.permute(0, 2, 1)).permute(0, 2, 1)
    bl = bone_proj(res["pd_joint_uv_left"], jf_l, S, distance)
    br = bone_proj(res["pd_joint_uv_right"], jf_r, S, distance)
    x = conv(sd, p + "fusion.0.", torch.cat((bl, br), 1), pad=1)
    x = F.relu(bn2d(sd, p + "fusion.1.", x))
    x = conv(sd, p + "fusion.3.", x)
    feats = {"img_feat": x, "joint_feat_left": jf_l, "joint_feat_right": jf_r, "vis_img_feat": bl + br}
    return res, feats


def conv_bn_relu_conv(sd, p, x):
    """conv3x3 -> BN -> ReLU -> conv1x1 heads (models/dir.py:404-420)."""
    h = F.relu(bn2d(sd, p + "1.", conv(sd, p + "0.", x, pad=1)))
    return conv(sd, p + "3.", h)


def decoder(sd, feats, init_out, p="decoder."):
    """models/dir.py:437-483."""
    _, c2, c3, c4 = feats
    fusion = residual(sd, p + "fusion_layer4.", torch.cat((upsample2x(c4), residual(sd, p + "skip_layer4.", c3)), 1))
    res1, f1 = joint2bone(sd, p + "projecter_4.", fusion, init_out, 16, 1)
    enh = residual(sd, p + "enhance_layer4.", torch.cat((fusion, f1["img_feat"]), 1))
    fusion = residual(sd, p + "fusion_layer3.", torch.cat((upsample2x(enh), residual(sd, p + "skip_layer3.", c2)), 1))
    res2, f2 = joint2bone(sd, p + "projecter_3.", fusion, res1, 32, 2)
    enh = residual(sd, p + "enhance_layer3.", torch.cat((fusion, f2["img_feat"]), 1))
    feat = conv_bn_relu_conv(sd, p + "conv_final.", enh)
    return {"result_list": [res1, res2], "seg": conv_bn_relu_conv(sd, p + "seg.", feat),
            "dense": conv_bn_relu_conv(sd, p + "dense.", feat), "proj_feat": f2["vis_img_feat"],
            "_feats": [f1, f2]}


OUT_KEYS = ["pd_joint_uv_left", "pd_joint_uv_right", "pd_mesh_xyz_left", "pd_mesh_xyz_right",
            "pd_joint_xyz_left", "pd_joint_xyz_right", "pd_proj_left", "pd_proj_right", "pd_offset"]


def dir_forward(sd, img):
    """models/dir.py:513-540 (eval branch). img (B,3,256,256) fp32 -> list of 4 dicts."""
    with torch.no_grad():
        feats = resnet50(sd, img)
        init_out = init_regressor(sd, feats[-1])
        dec = decoder(sd, feats, init_out)
        outs = []
        for o in [init_out] + dec["result_list"]:
            d = {k: o[k] for k in OUT_KEYS}
            d["pd_rel_joint"] = None
            outs.append(d)
        outs.append({"dense": dec["dense"], "seg": dec["seg"], "proj_feat": dec["proj_feat"]})
        return outs


# --------------------------------------------------------------------------- eval metric (next row N2)
def eval_jregressor(jreg16):
    """apps/eval.py:27-41 (class Jr): 16-joint MANO regressor -> 21 joints (5 tip vertices appended, reordered).
    Note: the reference uses tip vertex 444 for BOTH hands here (unlike manopth's 445 for the left hand)."""
    tips = torch.zeros(5, jreg16.shape[1], dtype=jreg16.dtype, device=jreg16.device)
    for i, v in enumerate([745, 317, 444, 556, 673]):
        tips[i, v] = 1.0
    return torch.cat([jreg16, tips], 0)[JOINT_REORDER].contiguous()


def xyz2uvd(xyz, cam):
    """apps/eval.py:80-83."""
    p = xyz @ cam.permute(0, 2, 1)
    return p[:, :, :2] / p[:, :, 2:]


def eval_metrics(pred_verts, pred_offset, gt_verts, gt_verts2d, cam, jreg21, scale=True):
    """apps/eval.py:151-241 for root_joint=0. pred_verts/gt_verts: dict side->(B,778,3); gt_verts2d: side->(B,778,2);
    jreg21: side->(21,778). Returns per-sample error arrays like the lists the reference accumulates."""
    out = {}
    roots = {}
    for side in ("left", "right"):
        J = jreg21[side]
        j_gt = J @ gt_verts[side]
        j2d_gt = xyz2uvd(j_gt, cam)
        root_gt = j_gt[:, 0:1].clone()
        roots[side] = root_gt
        len_gt = torch.linalg.norm(j_gt[:, 9] - j_gt[:, 0], dim=-1)
        j_gt = j_gt - root_gt
        v_gt = gt_verts[side] - root_gt
        j_ori = J @ pred_verts[side]
        root_p = j_ori[:, 0:1].clone()
        len_p = torch.linalg.norm(j_ori[:, 9] - j_ori[:, 0], dim=-1)
        sc = (len_gt / len_p).unsqueeze(-1).unsqueeze(-1) if scale else 1
        j_p = (j_ori - root_p) * sc
        v_p = (pred_verts[side] - root_p) * sc
        out[f"joint_{side}"] = torch.linalg.norm(j_p - j_gt, dim=-1)
        out[f"vert_{side}"] = torch.linalg.norm(v_p - v_gt, dim=-1)
        out[f"vert2d_{side}"] = torch.linalg.norm(xyz2uvd(v_p + root_gt, cam) - gt_verts2d[side], dim=-1)
        out[f"joint2d_{side}"] = torch.linalg.norm(xyz2uvd(j_p + root_gt, cam) - j2d_gt, dim=-1)
    gt_offset = roots["right"] - roots["left"]
    out["root"] = torch.linalg.norm(gt_offset - pred_offset.unsqueeze(1) * 0.15, dim=-1).reshape(-1)
    return out
