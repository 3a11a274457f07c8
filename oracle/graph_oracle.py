"""oracle/graph_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

torch-CPU restatement (fp32 by default, fp64 on request) of the TensorFlow-1.12 graph blocks on
EfficientLO-Net's hot path.  PARITY UNPINNED: TensorFlow is not installable here and the reference
ships no golden vectors for these blocks, so this file follows the
reference source line by line with the documented TF semantics (gather_nd row-major, scatter_nd
accumulates, to_int32 truncates, inference batch-norm with eps 1e-3, softmax(dim=2) over K,
boolean_mask then softmax over axis 0).  The two custom ops inside the blocks go through
oracle/index_oracle.port, which IS pinned to the reference's own kernels.

Follows (file:line relative to /root/reference):
  utils/tf_util.py:120-185, 52-115, 512-531     conv2d / conv1d / batch norm
  utils/pointnet_util.py:33-149                  cost_volume
  utils/pointnet_util.py:153-175                 flow_predictor
  utils/pointnet_util.py:179-250                 down_conv
  utils/pointnet_util.py:254-316                 up_conv
  model_util.py:17-69                            quaternion algebra
  model_util.py:181-292                          ProjectPC2SphericalRing
  model_util.py:296-316, 319-343, 346-445        get_selected_idx, softmax_valid, PreProcess
  pwclo_model.py:30-433, 437-481                 get_model, get_loss

Parameters are a flat dict under the reference's variable names (see efficientlo-net_b200/params.py
for the naming; the oracle only reads the dict it is given).
"""
import math

import numpy as np
import torch

from . import index_oracle as io

BN_EPS = 1e-3


# ----------------------------------------------------------------------------------------------
# primitives
_TRAIN = []          # stack of dict(decay=..., moving={}) while a training-mode forward is being restated


class training:
    """with training(bn_decay) as moving: ...   conv2d then normalises with the statistics of the tensor
    it is given (biased variance, tf_util.py:527 / FusedBatchNorm) and records the moving averages it
    would assign (Bessel-corrected variance, moving -= (moving - batch) * (1 - decay)) in `moving`;
    P itself is left untouched.  Dropout is NOT restated (random); compare with dropout disabled."""

    def __init__(self, bn_decay=None):
        self.state = dict(decay=0.9 if bn_decay is None else float(bn_decay), moving={})

    def __enter__(self):
        _TRAIN.append(self.state)
        return self.state["moving"]

    def __exit__(self, *a):
        _TRAIN.pop()


def conv2d(x, P, scope, relu=True, bn=True):
    """1x1 conv + bias + batch-norm + ReLU on the last axis (tf_util.py:120-185); inference statistics
    unless inside a `training` block."""
    y = x @ P[scope + "/weights"].to(x.dtype) + P[scope + "/biases"].to(x.dtype)
    if bn and _TRAIN:
        st = _TRAIN[-1]
        flat = y.reshape(-1, y.shape[-1])
        n = flat.shape[0]
        mean = flat.mean(0)
        var = ((flat - mean) ** 2).mean(0)
        mv = st["moving"]
        mm0 = mv.get(scope + "/bn/moving_mean", P[scope + "/bn/moving_mean"]).to(x.dtype)
        mv0 = mv.get(scope + "/bn/moving_variance", P[scope + "/bn/moving_variance"]).to(x.dtype)
        mv[scope + "/bn/moving_mean"] = (mm0 - (mm0 - mean) * (1 - st["decay"])).detach()
        mv[scope + "/bn/moving_variance"] = (mv0 - (mv0 - var * (n / max(n - 1, 1))) * (1 - st["decay"])).detach()
        y = (y - mean) / torch.sqrt(var + BN_EPS) * P[scope + "/bn/gamma"].to(x.dtype) + P[scope + "/bn/beta"].to(x.dtype)
        return torch.relu(y) if relu else y
    if bn:
        mean = P[scope + "/bn/moving_mean"].to(x.dtype)
        var = P[scope + "/bn/moving_variance"].to(x.dtype)
        y = (y - mean) * torch.rsqrt(var + BN_EPS) * P[scope + "/bn/gamma"].to(x.dtype) + P[scope + "/bn/beta"].to(x.dtype)
    return torch.relu(y) if relu else y


def conv1d(x, P, scope):
    """conv1d kernel 1, no BN, no activation (tf_util.py:52-115 as called at pwclo_model.py:197)."""
    return x @ P[scope + "/weights"].to(x.dtype) + P[scope + "/biases"].to(x.dtype)


def gather_nd(t, idx):
    """tf.gather_nd with (..., 3) [b, h, w] indices into a (B, H, W, C) tensor."""
    idx = idx.long()
    return t[idx[..., 0], idx[..., 1], idx[..., 2]]


def get_hw_idx(B, H, W):
    hh = torch.arange(H, dtype=torch.int32).view(1, H, 1, 1).expand(B, H, W, 1)
    ww = torch.arange(W, dtype=torch.int32).view(1, 1, W, 1).expand(B, H, W, 1)
    return torch.cat([hh, ww], -1).reshape(B, -1, 2).contiguous()


def get_selected_idx(B, stride_h, stride_w, out_h, out_w):
    hh = torch.arange(0, out_h * stride_h, stride_h, dtype=torch.int32).view(1, -1, 1, 1).expand(B, out_h, out_w, 1)
    ww = torch.arange(0, out_w * stride_w, stride_w, dtype=torch.int32).view(1, 1, -1, 1).expand(B, out_h, out_w, 1)
    bb = torch.arange(B, dtype=torch.int32).view(-1, 1, 1, 1).expand(B, out_h, out_w, 1)
    return torch.cat([bb, hh, ww], -1).contiguous()


def fused_conv(mode, xyz1, xyz2, idx_n2, random_hw, kH, kW, K, distance, stride_h, stride_w, flag_copy=0):
    """The custom op through the pinned C restatement; returns (idx int64 (B,n,K,3), mask (B,n,K,1))."""
    B, H, W, _ = xyz1.shape
    n = idx_n2.shape[1]
    sel, _, _, mask = io.port(mode, xyz1.detach().float().numpy(), xyz2.detach().float().numpy(), idx_n2.numpy(),
                              np.asarray(random_hw, dtype=np.int32), H, W, n, kH, kW, K, flag_copy,
                              float(distance), stride_h, stride_w, nthreads=8)
    return torch.from_numpy(sel).long(), torch.from_numpy(mask).to(xyz1.dtype)


# ----------------------------------------------------------------------------------------------
# blocks
def down_conv(xyz_proj, points_proj, selected_idx, K_sample, kernel_size, distance, mlp_scopes, P, random_hw,
              debug=None):
    """Set-conv (pointnet_util.py:179-250).  Returns ((B, n, C_out), new_xyz_proj (B, oh, ow, 3))."""
    B = xyz_proj.shape[0]
    idx_n2 = selected_idx.reshape(B, -1, 3)[:, :, 1:].contiguous()
    idx, mask = fused_conv("random", xyz_proj, xyz_proj, idx_n2, random_hw, kernel_size[0], kernel_size[1],
                           K_sample, distance, 1, 1)
    if debug is not None:
        debug["idx"], debug["mask"] = idx, mask
    new_xyz_group = gather_nd(xyz_proj, idx) * mask
    new_points_group = gather_nd(points_proj, idx) * mask
    new_xyz_proj = gather_nd(xyz_proj, selected_idx)
    new_xyz = new_xyz_proj.reshape(B, -1, 1, 3)
    x = torch.cat([new_xyz_group - new_xyz, new_points_group], -1)
    for s in mlp_scopes:
        x = conv2d(x, P, s)
    x = x * mask
    return x.max(dim=2).values, new_xyz_proj


def up_conv(xyz1_proj, xyz2_proj, feat1_proj, feat2_proj, kernel_size, stride_h, stride_w, nsample, distance,
            scope, P, random_hw, debug=None):
    """Set-upconv (pointnet_util.py:254-316): dense queries xyz1 against the sparse grid xyz2."""
    B, H, W, _ = xyz1_proj.shape
    xyz1 = xyz1_proj.reshape(B, H * W, 1, 3)
    points1 = feat1_proj.reshape(B, H * W, -1)
    idx, mask = fused_conv("random", xyz1_proj, xyz2_proj, get_hw_idx(B, H, W), random_hw, kernel_size[0],
                           kernel_size[1], nsample, distance, stride_h, stride_w)
    if debug is not None:
        debug["idx"], debug["mask"] = idx, mask
    g_xyz = gather_nd(xyz2_proj, idx) * mask
    g_feat = gather_nd(feat2_proj, idx) * mask
    x = torch.cat([g_xyz - xyz1, g_feat], -1)
    for s in ("up_1_0", "up_1_1"):
        x = conv2d(x, P, scope + "/" + s)
    x = (x * mask).max(dim=2).values
    x = torch.cat([x, points1], -1)
    for s in ("up_2_0", "up_2_1"):
        x = conv2d(x, P, scope + "/" + s)
    return x


def cost_volume(warped_xyz1_proj, xyz2_proj, points1_proj, points2_proj, kernel_size1, kernel_size2, nsample,
                nsample_q, distance, scope, P, random_hw_q, random_hw_p, debug=None):
    """Two-stage attentive cost volume (pointnet_util.py:33-149).  Returns (B, H*W, 64)."""
    B, H, W, _ = warped_xyz1_proj.shape
    dt = warped_xyz1_proj.dtype
    xyz1 = warped_xyz1_proj.reshape(B, H * W, 1, 3)
    points1 = points1_proj.reshape(B, H * W, 1, -1)
    idx_hw = get_hw_idx(B, H, W)
    # stage 1: point-to-patch, select-K in frame 2; `distance` is ignored here (hard-coded 1000, :51)
    idx, mask = fused_conv("select", warped_xyz1_proj, xyz2_proj, idx_hw, random_hw_q, kernel_size2[0],
                           kernel_size2[1], nsample_q, 1000.0, 1, 1)
    qi_xyz = gather_nd(xyz2_proj, idx) * mask
    qi_pts = gather_nd(points2_proj, idx) * mask
    pi_xyz = xyz1.expand(-1, -1, nsample_q, -1)
    pi_pts = points1.expand(-1, -1, nsample_q, -1)
    diff = qi_xyz - pi_xyz
    euc = torch.sqrt((diff * diff).sum(-1, keepdim=True) + 1e-20)
    xyz10 = torch.cat([pi_xyz, qi_xyz, diff, euc], -1)
    feat = torch.cat([xyz10, pi_pts, qi_pts], -1)
    for s in ("CV_0", "CV_1", "CV_2"):
        feat = conv2d(feat, P, scope + "/" + s)
    enc = conv2d(xyz10, P, scope + "/CV_xyz")
    w = torch.cat([enc, feat], -1)
    for s in ("sum_CV_0", "sum_CV_1"):
        w = conv2d(w, P, scope + "/" + s)
    w = torch.where(mask == 1.0, w, torch.full_like(w, -1e10))
    w = torch.softmax(w, dim=2)
    stage1 = (w * feat).sum(2)                                     # (B, N, 64)
    stage1_proj = stage1.reshape(B, H, W, -1)
    # stage 2: patch-to-patch, random-K among frame-1 neighbours
    idx2, mask2 = fused_conv("random", warped_xyz1_proj, warped_xyz1_proj, idx_hw, random_hw_p, kernel_size1[0],
                             kernel_size1[1], nsample, distance, 1, 1)
    if debug is not None:
        debug.update(idx_q=idx, mask_q=mask, idx_p=idx2, mask_p=mask2, stage1=stage1)
    pc_pts = gather_nd(stage1_proj, idx2) * mask2
    pc_xyz = gather_nd(warped_xyz1_proj, idx2) * mask2
    p_xyz = xyz1.expand(-1, -1, nsample, -1)
    p_pts = points1.expand(-1, -1, nsample, -1)
    d2 = pc_xyz - p_xyz
    e2 = torch.sqrt((d2 * d2).sum(-1, keepdim=True) + 1e-20)
    enc2 = conv2d(torch.cat([p_xyz, pc_xyz, d2, e2], -1), P, scope + "/sum_xyz_encoding")
    w2 = torch.cat([enc2, p_pts, pc_pts], -1)
    for s in ("sum_cost_volume_0", "sum_cost_volume_1"):
        w2 = conv2d(w2, P, scope + "/" + s)
    w2 = torch.where(mask2 == 1.0, w2, torch.full_like(w2, -1e10))
    w2 = torch.softmax(w2, dim=2)
    return (w2 * pc_pts).sum(2).to(dt)


def flow_predictor(points_f1, upsampled_feat, cost_vol, scope, P):
    """Shared MLP on concatenated per-point features (pointnet_util.py:153-175)."""
    parts = [points_f1] + ([upsampled_feat] if upsampled_feat is not None else []) + \
            ([cost_vol] if cost_vol is not None else [])
    x = torch.cat(parts, -1)
    for s in ("conv_predictor0", "conv_predictor1"):
        x = conv2d(x, P, scope + "/" + s)
    return x


def softmax_valid(feature_bnc, weight_bnc, mask_valid):
    """Masked softmax over the point axis per channel, weighted sum (model_util.py:319-343)."""
    out = []
    for b in range(feature_bnc.shape[0]):
        m = mask_valid[b]
        f, w = feature_bnc[b][m], weight_bnc[b][m]
        out.append((f * torch.softmax(w, dim=0)).sum(0, keepdim=True))
    return torch.stack(out, 0)                                      # (B, 1, C)


# ----------------------------------------------------------------------------------------------
# quaternions (w, x, y, z), model_util.py:17-69
def hamilton(a, b):
    a0, a1, a2, a3 = a.unbind(-1)
    b0, b1, b2, b3 = b.unbind(-1)
    return torch.stack([a0 * b0 - a1 * b1 - a2 * b2 - a3 * b3,
                        a0 * b1 + a1 * b0 + a2 * b3 - a3 * b2,
                        a0 * b2 - a1 * b3 + a2 * b0 + a3 * b1,
                        a0 * b3 + a1 * b2 - a2 * b1 + a3 * b0], -1)


def mul_q_point(q_a, q_b):
    return hamilton(q_a.reshape(q_a.shape[0], 1, 4), q_b)


def mul_point_q(q_a, q_b):
    return hamilton(q_a, q_b.reshape(q_b.shape[0], 1, 4))


def inv_q(q):
    q = q.reshape(q.shape[0], 4)
    q2 = (q * q).sum(-1, keepdim=True) + 1e-10
    return torch.cat([q[:, :1], -q[:, 1:]], -1) / q2


def normalize_q(q):
    return q / (torch.sqrt((q * q).sum(-1, keepdim=True) + 1e-10) + 1e-10)


def warp(xyz, q, t):
    """(q (x) [0,p] (x) q^-1)[1:] + t, zero points stay zero (pwclo_model.py:213-227)."""
    B = xyz.shape[0]
    valid = (~(xyz == 0).all(-1, keepdim=True)).to(xyz.dtype)
    pq = torch.cat([torch.zeros_like(xyz[..., :1]), xyz], -1)
    w = mul_point_q(mul_q_point(q.reshape(B, 1, 4), pq), inv_q(q.reshape(B, 1, 4)))
    return (w[..., 1:] + t.reshape(B, 1, 3)) * valid


# ----------------------------------------------------------------------------------------------
def project_constants(H_input, W_input):
    """float64 python arithmetic, then fp32 tf.constant (model_util.py:189-210)."""
    d2r = math.pi / 180
    az = (360.0 / W_input) * d2r
    down, up = -24.8 * d2r, 2.0 * d2r
    vres = (up - down) / (H_input - 1)
    voff = -down / vres
    return np.float32(np.pi), np.float32(az), np.float32(vres), np.float32(voff)


def ProjectPC2SphericalRing(PC, Feature, H_input, W_input):
    """Spherical binning with min-range winner per cell, ties summed (model_util.py:181-292).
    Emulates the GPU placement of the reference graph for zero points: asin(0/0) = NaN casts to 0
    (SURVEY.md Appendix A.5), which x86 would cast to INT_MIN."""
    B, N, _ = PC.shape
    dt = PC.dtype
    PI, AZ, VRES, VOFF = (torch.tensor(float(c), dtype=torch.float32).to(dt) for c in project_constants(H_input, W_input))
    out_pc, out_ft = [], []
    for b in range(B):
        pc = PC[b, :, :3]
        x, y, z = pc[:, 0], pc[:, 1], pc[:, 2]
        r = torch.sqrt((x * x + y * y) + z * z)
        col = ((PI - torch.atan2(y, x)) / AZ)
        col = torch.where(torch.isfinite(col), col, torch.zeros_like(col)).to(torch.int32)     # trunc toward zero
        tmp = torch.asin(z / r) / VRES + VOFF
        tmp = torch.where(torch.isfinite(tmp), tmp, torch.zeros_like(tmp)).to(torch.int32)     # NaN -> 0 (GPU cast)
        row = (H_input - tmp).clamp(0, H_input - 1)
        col = col.clamp(0, W_input - 1)
        cell = (row * W_input + col).long()
        min_r = torch.full((H_input * W_input,), float("inf"), dtype=dt)
        min_r = min_r.scatter_reduce(0, cell, r, reduce="amin", include_self=True)
        same = (r == min_r[cell]).to(dt).unsqueeze(-1)
        grid = torch.zeros(H_input * W_input, 3, dtype=dt).index_add_(0, cell, pc * same)
        out_pc.append(grid.reshape(H_input, W_input, 3))
        if Feature is not None:
            C = Feature.shape[-1]
            fg = torch.zeros(H_input * W_input, C, dtype=dt).index_add_(0, cell, Feature[b] * same)
            out_ft.append(fg.reshape(H_input, W_input, C))
    pc_final = torch.stack(out_pc, 0)
    return pc_final, (torch.stack(out_ft, 0) if Feature is not None else pc_final)


def mat2euler(M):
    cy = torch.sqrt(M[2, 2] * M[2, 2] + M[1, 2] * M[1, 2])
    return torch.atan2(-M[0, 1], M[0, 0]), torch.atan2(M[0, 2], cy), torch.atan2(-M[1, 2], M[2, 2])


def euler2quat(z, y, x):
    z, y, x = z / 2.0, y / 2.0, x / 2.0
    cz, sz, cy, sy, cx, sx = torch.cos(z), torch.sin(z), torch.cos(y), torch.sin(y), torch.cos(x), torch.sin(x)
    return torch.stack([cx * cy * cz - sx * sy * sz, cx * sy * sz + cy * cz * sx,
                        cx * cz * sy - sx * cy * sz, cx * cy * sz + sx * cz * sy])


def PreProcess(PC_f1, PC_f2, T_gt, T_trans, T_trans_inv, aug_frame):
    """35 m crop, optional rigid augmentation of one frame, GT pose as (q, t) (model_util.py:346-445)."""
    B = PC_f1.shape[0]
    dt = PC_f1.dtype
    outs1, outs2, qs, ts = [], [], [], []
    for i in range(B):
        f1, f2 = PC_f1[i], PC_f2[i]
        m1 = (~(f1 == 0).all(-1, keepdim=True)).to(dt)
        m2 = (~(f2 == 0).all(-1, keepdim=True)).to(dt)
        c1 = torch.cat([f1, torch.ones_like(f1[:, :1])], -1)
        c2 = torch.cat([f2, torch.ones_like(f2[:, :1])], -1)
        r1 = torch.sqrt(c1[:, 0] * c1[:, 0] + c1[:, 1] * c1[:, 1])
        c1 = torch.where((r1 > 35).unsqueeze(-1), torch.zeros_like(c1), c1)
        r2 = torch.sqrt(c2[:, 0] * c2[:, 0] + c2[:, 1] * c2[:, 1])
        c2 = torch.where((r2 > 35).unsqueeze(-1), torch.zeros_like(c2), c2)
        Tg, Tt, Ti = T_gt[i].to(dt), T_trans[i].to(dt), T_trans_inv[i].to(dt)
        if aug_frame[i] == 2:
            a1, a2 = c1[:, :3], (Tt @ c2.T).T[:, :3]
            Tg = Tt @ Tg
        else:
            a1, a2 = (Tt @ c1.T).T[:, :3], c2[:, :3]
            Tg = Tg @ Ti
        outs1.append(a1 * m1)
        outs2.append(a2 * m2)
        z, y, x = mat2euler(Tg[:3, :3])
        qs.append(euler2quat(z, y, x))
        ts.append(Tg[:3, 3:])
    return torch.stack(outs1), torch.stack(outs2), torch.stack(qs), torch.stack(ts)


# ----------------------------------------------------------------------------------------------
def pyramid_shapes(H_input, W_input):
    sh, sw = [1, 1, 4, 2, 2, 1], [1, 1, 8, 2, 2, 2]
    oh, ow = [math.ceil(H_input / sh[0])], [math.ceil(W_input / sw[0])]
    for i in range(1, 6):
        oh.append(math.ceil(oh[-1] / sh[i]))
        ow.append(math.ceil(ow[-1] / sw[i]))
    return sh, sw, oh, ow


def pose_head(feat, P, lvl):
    """conv1d 64->256 (no act), [dropout is identity at inference], q (normalised) and t heads."""
    big = conv1d(feat, P, "l%d_big" % lvl)
    suffix = "coarse" if lvl == 3 else "det"
    q = normalize_q(conv1d(big, P, "l%d_q_%s" % (lvl, suffix)))
    t = conv1d(big, P, "l%d_t_%s" % (lvl, suffix))
    return q, t


def get_model(point_cloud, H_input, W_input, T_gt, T_trans, T_trans_inv, P, perms, aug_frame=None,
              dtype=torch.float32, keep=None):
    """Inference forward of the whole network (pwclo_model.py:30-433).  `keep`, if a dict, receives
    named intermediates for layer-wise parity tests."""
    B = point_cloud.shape[0]
    N = point_cloud.shape[1] // 2
    pc = point_cloud.to(dtype)
    if aug_frame is None:
        aug_frame = [2] * B
    sh, sw, oh, ow = pyramid_shapes(H_input, W_input)
    K = keep if keep is not None else {}

    f1, f2, q_gt, t_gt = PreProcess(pc[:, :N, 0:3], pc[:, N:, 0:3], T_gt, T_trans, T_trans_inv, aug_frame)
    xyz_f1, _ = ProjectPC2SphericalRing(f1, None, H_input, W_input)
    xyz_f2, _ = ProjectPC2SphericalRing(f2, None, H_input, W_input)
    K["xyz_f1_proj"], K["xyz_f2_proj"] = xyz_f1, xyz_f2
    pts_in = torch.zeros(B, H_input, W_input, 3, dtype=dtype)

    pre2_idx = get_selected_idx(B, sh[1], sw[1], oh[1], ow[1])
    pre2_f1, pre2_f2 = gather_nd(xyz_f1, pre2_idx), gather_nd(xyz_f2, pre2_idx)
    sel = [get_selected_idx(B, sh[i + 2], sw[i + 2], oh[i + 2], ow[i + 2]) for i in range(4)]   # l0..l3

    # siamese feature pyramid (pwclo_model.py:117-165)
    cfg = [(32, (9, 15), 0.5), (32, (7, 11), 3.0), (16, (5, 9), 6.0), (16, (5, 9), 12.0)]
    xyz = {"f1": [None] * 4, "f2": [None] * 4}
    pts = {"f1": [None] * 4, "f2": [None] * 4}
    for f, x0 in (("f1", pre2_f1), ("f2", pre2_f2)):
        cur_xyz, cur_pts = x0, pts_in
        for l in range(4):
            scopes = ["sa1/layer%d/conv%d" % (l, j) for j in range(3)]
            feat, new_xyz = down_conv(cur_xyz, cur_pts, sel[l], cfg[l][0], cfg[l][1], cfg[l][2], scopes, P,
                                      perms["sa1/layer%d/%s" % (l, f)])
            xyz[f][l], pts[f][l] = new_xyz, feat
            cur_xyz, cur_pts = new_xyz, feat.reshape(B, oh[l + 2], ow[l + 2], -1)
            K["l%d_points_%s" % (l, f)] = feat

    def proj(l, t):
        return t.reshape(B, oh[l + 2], ow[l + 2], -1)

    # initial cost volume at level 2 and its set-conv to level 3 (pwclo_model.py:170-178)
    l2_new = cost_volume(xyz["f1"][2], xyz["f2"][2], proj(2, pts["f1"][2]), proj(2, pts["f2"][2]), (3, 5), (5, 35),
                         4, 32, 4.0, "flow_embedding_l2_origin", P, perms["flow_embedding_l2_origin/q"],
                         perms["flow_embedding_l2_origin/p"])
    K["l2_points_f1_new"] = l2_new
    l3_cv, _ = down_conv(xyz["f1"][2], proj(2, l2_new), sel[3], 16, (5, 9), 12.0,
                         ["new_layer3/conv%d" % j for j in range(3)], P, perms["new_layer3"])
    K["l3_points_f1_cost_volume"] = l3_cv
    l3_w = flow_predictor(pts["f1"][3], None, l3_cv, "l3_costvolume_predict_ww", P)
    l3_xyz = xyz["f1"][3].reshape(B, -1, 3)
    l3_valid = ~(l3_xyz == 0).all(-1)
    l3_feat = softmax_valid(l3_cv, l3_w, l3_valid)
    q, t = pose_head(l3_feat, P, 3)
    q, t = q.squeeze(1), t.squeeze(1)
    qs, ts = {3: q}, {3: t}
    K["l3_q"], K["l3_t"] = q, t

    up_xyz = xyz["f1"][3]                  # level 2 up-samples from the UN-warped level-3 grid (:247)
    up_w, up_pred = proj(3, l3_w), proj(3, l3_cv)
    cv_dis = {2: 4.0, 1: 2.0, 0: 1.0}
    up_dis = {2: 9.0, 1: 6.0, 0: 3.0}
    kq = {2: (5, 15), 1: (7, 25), 0: (11, 41)}
    for lvl in (2, 1, 0):
        h, w_ = oh[lvl + 2], ow[lvl + 2]
        q_c, t_c = q.reshape(B, 1, 4), t.reshape(B, 1, 3)
        xyz_l = xyz["f1"][lvl].reshape(B, -1, 3)
        warped = warp(xyz_l, q_c, t_c)
        K["l%d_flow_warp" % lvl] = warped
        xyz_wp, pts_wp = ProjectPC2SphericalRing(warped, pts["f1"][lvl], h, w_)
        K["l%d_xyz_warp_proj" % lvl], K["l%d_points_warp_proj" % lvl] = xyz_wp, pts_wp
        valid_w = ~(xyz_wp.reshape(B, -1, 3) == 0).all(-1)
        pts_w = pts_wp.reshape(B, h * w_, -1)
        cv = cost_volume(xyz_wp, xyz["f2"][lvl], pts_wp, proj(lvl, pts["f2"][lvl]), (3, 5), kq[lvl], 4, 6,
                         cv_dis[lvl], "flow_embedding_l%d" % lvl, P, perms["flow_embedding_l%d/q" % lvl],
                         perms["flow_embedding_l%d/p" % lvl])
        K["l%d_cost_volume" % lvl] = cv
        s_h, s_w = sh[lvl + 3], sw[lvl + 3]
        w_up = up_conv(xyz_wp, up_xyz, pts_wp, up_w, (7, 15), s_h, s_w, 8, up_dis[lvl],
                       "up_sa_layer_layer_l%dw" % lvl, P, perms["up_sa_layer_layer_l%dw" % lvl])
        p_up = up_conv(xyz_wp, up_xyz, pts_wp, up_pred, (7, 15), s_h, s_w, 8, up_dis[lvl],
                       "up_sa_layer_layer_l%dcostvolume" % lvl, P, perms["up_sa_layer_layer_l%dcostvolume" % lvl])
        K["l%d_w_up" % lvl], K["l%d_p_up" % lvl] = w_up, p_up
        pred = flow_predictor(pts_w, p_up, cv, "l%d_costvolume_predict" % lvl, P)
        wgt = flow_predictor(pts_w, w_up, cv, "l%d_w_predict" % lvl, P)
        K["l%d_predict" % lvl], K["l%d_w" % lvl] = pred, wgt
        feat = softmax_valid(pred, wgt, valid_w)
        K["l%d_pooled" % lvl] = feat
        q_det, t_det = pose_head(feat, P, lvl)
        tq = torch.cat([torch.zeros(B, 1, 1, dtype=dtype), t_c], -1)
        tq = mul_point_q(mul_q_point(q_det, tq), inv_q(q_det))[..., 1:]
        q = mul_point_q(q_det, q_c).squeeze(1)
        t = (tq + t_det).squeeze(1)
        qs[lvl], ts[lvl] = q, t
        K["l%d_q" % lvl], K["l%d_t" % lvl] = q, t
        up_xyz, up_w, up_pred = xyz_wp, wgt.reshape(B, h, w_, -1), pred.reshape(B, h, w_, -1)

    l0_xyz_f1 = xyz["f1"][0].reshape(B, -1, 3)
    return (normalize_q(qs[0]), ts[0], normalize_q(qs[1]), ts[1], normalize_q(qs[2]), ts[2],
            normalize_q(qs[3]), ts[3], l0_xyz_f1, q_gt, t_gt)


def get_loss(l0_q, l0_t, l1_q, l1_t, l2_q, l2_t, l3_q, l3_t, q_gt, t_gt, w_x, w_q):
    """pwclo_model.py:437-481"""
    t_gt = t_gt.squeeze(-1)
    total = 0.0
    for wgt, q, t in ((0.2, l0_q, l0_t), (0.4, l1_q, l1_t), (0.8, l2_q, l2_t), (1.6, l3_q, l3_t)):
        qn = normalize_q(q)
        lq = torch.sqrt(((q_gt - qn) ** 2).sum(-1, keepdim=True) + 1e-10).mean()
        lx = torch.sqrt((t - t_gt) * (t - t_gt) + 1e-10).mean()
        total = total + wgt * (lx * torch.exp(-w_x) + w_x + lq * torch.exp(-w_q) + w_q)
    return total
