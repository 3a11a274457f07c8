"""Geometry helpers of the PWCLO network on B200: quaternion algebra, spherical re-projection, masked
attention pooling, pre-processing.

Host-side mirror of the reference's model_util.py (mul_q_point :17, mul_point_q :39, inv_q :61,
ProjectPC2SphericalRing :181, get_selected_idx :296, softmax_valid :319, PreProcess :346).  The heavy
functions launch fused sm_100a kernels (csrc/project_pose.cu); the quaternion helpers are plain torch
expressions kept for API completeness -- inside the model the warp is fused into the re-projection
kernel (project_points(mode=2)) and the pose composition into the pose-head kernel.
"""
import math

import numpy as np
import torch

from . import _lib
from . import store as _store
from .pointnet_util import SelectedIdx


# ---- quaternions (w, x, y, z) ------------------------------------------------------------------
def _hamilton(a, b):
    a0, a1, a2, a3 = a.unbind(-1)
    b0, b1, b2, b3 = b.unbind(-1)
    return torch.stack([a0 * b0 - a1 * b1 - a2 * b2 - a3 * b3, a0 * b1 + a1 * b0 + a2 * b3 - a3 * b2,
                        a0 * b2 - a1 * b3 + a2 * b0 + a3 * b1, a0 * b3 + a1 * b2 - a2 * b1 + a3 * b0], -1)


def mul_q_point(q_a, q_b, batch_size):
    """q_a (B,1,4) (x) q_b (B,N,4) (model_util.py:17-36)."""
    return _hamilton(q_a.reshape(batch_size, 1, 4), q_b)


def mul_point_q(q_a, q_b, batch_size):
    """q_a (B,N,4) (x) q_b (B,1,4) (model_util.py:39-58)."""
    return _hamilton(q_a, q_b.reshape(batch_size, 1, 4))


def inv_q(q, batch_size):
    """conj(q) / (|q|^2 + 1e-10), (B,1,4) -> (B,4) (model_util.py:61-69)."""
    q = q.reshape(batch_size, 4)
    return torch.cat([q[:, :1], -q[:, 1:]], -1) / ((q * q).sum(-1, keepdim=True) + 1e-10)


# ---- strided index grids -------------------------------------------------------------------------
def get_selected_idx(array, stride_h, stride_w, out_h, out_w):
    """model_util.py:296-316.  Returns a symbolic SelectedIdx (call .tensor() for the (B,oh,ow,3) grid)."""
    return SelectedIdx(array.shape[0], stride_h, stride_w, out_h, out_w, array.device)


def xyz_pyramid(xyz_in, out_hs, out_ws, strides_h, strides_w):
    """The four strided xyz grids of pwclo_model.py:88-114 in one launch.  strides are cumulative with
    respect to xyz_in (S,H,W,3).  Returns a list of (S, out_h[l], out_w[l], 3) tensors."""
    import ctypes
    _lib.require_cuda("xyz_pyramid", xyz_in)
    S, H, W, _ = xyz_in.shape
    dev = xyz_in.device
    xyz_in = xyz_in.contiguous().float()
    outs = [torch.empty((S, h, w, 3), dtype=torch.float32, device=dev) for h, w in zip(out_hs, out_ws)]
    arr = lambda v: (ctypes.c_int * 4)(*[int(x) for x in v])
    ptrs = (ctypes.c_void_p * 4)(*[o.data_ptr() for o in outs])
    with torch.cuda.device(dev):
        rc = _lib.lib().elo_pyramid_xyz(S, H, W, arr(out_hs), arr(out_ws), arr(strides_h), arr(strides_w),
                                        xyz_in.data_ptr(), ptrs, _lib.stream_ptr(dev))
    _lib.check(rc, "elo_pyramid_xyz")
    return outs


# ---- spherical projection --------------------------------------------------------------------------
def projection_constants(H_input, W_input):
    """The fp32 constants of model_util.py:189-210 (float64 python arithmetic, then float32)."""
    d2r = math.pi / 180
    az = (360.0 / W_input) * d2r
    down, up = -24.8 * d2r, 2.0 * d2r
    vres = (up - down) / (H_input - 1)
    voff = -down / vres
    return float(np.float32(np.pi)), float(np.float32(az)), float(np.float32(vres)), float(np.float32(voff))


def _scratch(name, shape, dtype, device, fill=None):
    """Persistent scratch from the current store if there is one, else a fresh tensor."""
    stack = _store.current_stack()
    if stack:
        return stack[-1].scratch(name, shape, dtype, fill=fill)
    t = torch.empty(shape, dtype=dtype, device=device)
    if fill is not None:
        t.fill_(fill)
    return t


def project_points(PC, Feature, H_input, W_input, mode=0, T=None, q=None, t=None, inner_batch=0,
                   outer_stride=0, batch_size=None, want_points=False, want_cells=False, T_apply=None):
    """elo_project: optional PreProcess (mode 1) / pose warp (mode 2) fused with the spherical projection.
    PC may be any view whose last dimension is contiguous (e.g. point_cloud[:, :N, 0:3] of the
    (B, 2N, 6) input): it is read in place through its strides.  Returns (xyz (B,H,W,3),
    feat (B,H,W,C) or None, transformed points (B,N,3) or None)."""
    _lib.require_cuda("ProjectPC2SphericalRing", PC, Feature, T, q, t)
    if PC.dtype != torch.float32 or PC.stride(-1) != 1:
        PC = PC.float().contiguous()
    B = PC.shape[0] if batch_size is None else batch_size
    N = PC.shape[1]
    dev = PC.device
    # epoch-tagged cell minima + (epoch, done-counter): persistent per (store scratch name space, image shape)
    cellmin = _scratch("cellmin64", (B, H_input, W_input), torch.int64, dev, fill=-1)
    state = _scratch("project_state_%dx%dx%d" % (B, H_input, W_input), (2,), torch.int32, dev, fill=0)
    out_xyz = torch.empty((B, H_input, W_input, 3), dtype=torch.float32, device=dev)
    out_feat = out_pts = None
    d = _lib.ProjectDesc()
    d.batch_size, d.num_points, d.H, d.W, d.mode = B, N, H_input, W_input, mode
    d.points, d.point_stride, d.batch_stride = PC.data_ptr(), PC.stride(1), PC.stride(0)
    d.inner_batch, d.outer_stride = inner_batch, outer_stride
    if Feature is not None:
        Feature = Feature.contiguous().float()
        d.C = Feature.shape[-1]
        out_feat = torch.empty((B, H_input, W_input, d.C), dtype=torch.float32, device=dev)
        d.feat, d.out_feat = Feature.data_ptr(), out_feat.data_ptr()
    keep = [x.contiguous().float() if x is not None else None for x in (T, q, t)]
    d.T, d.q, d.t = (_lib.ptr(x) for x in keep)
    if T_apply is not None:
        T_apply = T_apply.to(device=dev, dtype=torch.int32).contiguous()
        d.T_apply = T_apply.data_ptr()
    d.pi, d.az_res, d.v_res, d.v_off = projection_constants(H_input, W_input)
    d.cellmin, d.state, d.out_xyz = cellmin.data_ptr(), state.data_ptr(), out_xyz.data_ptr()
    if want_points:
        out_pts = torch.empty((B, N, 3), dtype=torch.float32, device=dev)
        d.out_points = out_pts.data_ptr()
    keys = _scratch("project_keys_%dx%d" % (B, N), (B, N, 2), torch.int32, dev)      # binning pass -> scatter pass
    d.point_keys = keys.data_ptr()
    out_cell = None
    if want_cells:
        out_cell = torch.empty((B, N), dtype=torch.int32, device=dev)
        d.out_cell = out_cell.data_ptr()
    _lib.call("elo_project", d, dev)
    if want_cells:
        return out_xyz, out_feat, out_pts, out_cell
    return out_xyz, out_feat, out_pts


def ProjectPC2SphericalRing(PC, Feature, H_input, W_input):
    """Project (B,N,3) points (and optional (B,N,C) features) onto the H x W spherical ring image; per
    cell the nearest point wins, exact ties accumulate (model_util.py:181-292).
    Returns (PC_project (B,H,W,3), Feature_project (B,H,W,C)); like the reference, the second value is
    the xyz image again when Feature is None."""
    xyz, feat, _ = project_points(PC, Feature, H_input, W_input, mode=0)
    return xyz, (feat if Feature is not None else xyz)


# ---- masked attention pooling ---------------------------------------------------------------------
def pose_head_call(feature_bnc, weight_bnc, xyz_bn3, heads=None, coarse=None, want_pooled=False):
    """elo_pose_head.  heads = (w_big, b_big, w_q, b_q, w_t, b_t) device tensors or None (pool only);
    coarse = (q (B,4), t (B,3)) or None.  Returns dict(q, t, q_norm, pooled)."""
    _lib.require_cuda("softmax_valid", feature_bnc, weight_bnc, xyz_bn3)
    B, N, C = feature_bnc.shape
    if C != 64:
        raise NotImplementedError("softmax_valid kernel is built for 64 channels")
    dev = feature_bnc.device
    f, w, x = feature_bnc.contiguous().float(), weight_bnc.contiguous().float(), xyz_bn3.contiguous().float()
    G = (N + 31) // 32
    partial = _scratch("pose_partial", (B, G, 192), torch.float32, dev)
    counter = _scratch("pose_counter", (B,), torch.int32, dev, fill=0)
    out = {}
    d = _lib.PoseHeadDesc()
    d.batch_size, d.num_points, d.num_slices = B, N, G
    d.feature, d.weight, d.xyz = f.data_ptr(), w.data_ptr(), x.data_ptr()
    d.partial, d.counter = partial.data_ptr(), counter.data_ptr()
    keep = []
    if want_pooled or heads is None:
        out["pooled"] = torch.empty((B, 1, 64), dtype=torch.float32, device=dev)
        d.pooled_out = out["pooled"].data_ptr()
    if heads is not None:
        d.w_big, d.b_big, d.w_q, d.b_q, d.w_t, d.b_t = (h.data_ptr() for h in heads)
        out["q"] = torch.empty((B, 4), dtype=torch.float32, device=dev)
        out["t"] = torch.empty((B, 3), dtype=torch.float32, device=dev)
        out["q_norm"] = torch.empty((B, 4), dtype=torch.float32, device=dev)
        d.q_out, d.t_out, d.q_norm_out = out["q"].data_ptr(), out["t"].data_ptr(), out["q_norm"].data_ptr()
        if coarse is not None:
            keep = [coarse[0].contiguous().float(), coarse[1].contiguous().float()]
            d.has_coarse, d.q_coarse, d.t_coarse = 1, keep[0].data_ptr(), keep[1].data_ptr()
    _lib.call("elo_pose_head", d, dev)
    return out


def softmax_valid(feature_bnc, weight_bnc, mask_valid):
    """Per sample and channel: softmax of weight over the VALID points, weighted sum of feature
    (model_util.py:319-343).  feature, weight (B,N,64); mask_valid (B,N) bool.  Returns (B,1,64)."""
    m = mask_valid.to(feature_bnc.dtype)[..., None].expand(-1, -1, 3)
    return pose_head_call(feature_bnc, weight_bnc, m)["pooled"]


# ---- pre-processing ---------------------------------------------------------------------------------
def gt_pose(T_gt, T_trans, T_trans_inv, aug_frame=None):
    """(q_gt (B,4), t_gt (B,3,1)) of model_util.py:386-426 for (B,4,4) device matrices."""
    _lib.require_cuda("PreProcess", T_gt, T_trans, T_trans_inv)
    B = T_gt.shape[0]
    dev = T_gt.device
    Tg, Tt, Ti = (x.contiguous().float() for x in (T_gt, T_trans, T_trans_inv))
    af = None if aug_frame is None else torch.as_tensor(aug_frame).to(device=dev, dtype=torch.int32).contiguous()
    q = torch.empty((B, 4), dtype=torch.float32, device=dev)
    t = torch.empty((B, 3, 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().elo_gt_pose(B, Tg.data_ptr(), Tt.data_ptr(), Ti.data_ptr(), _lib.ptr(af), q.data_ptr(),
                                    t.data_ptr(), _lib.stream_ptr(dev))
    _lib.check(rc, "elo_gt_pose")
    return q, t


def aug_setup(T_trans, aug_frame, B, device):
    """Augmentation of model_util.py:386-417 for samples stacked frame-major (s = f*B + b): the matrices
    (2B,4,4) and the flags (2B,) int32 saying which samples are multiplied -- only the frame named by
    aug_frame[b] is; the other one is passed through untouched (no identity matmul: that would turn a
    -0.0 coordinate into +0.0).  aug_frame None = frame 2 everywhere, built without a host copy."""
    T = T_trans.to(device).float()
    if aug_frame is None:
        apply = torch.cat([torch.zeros(B, dtype=torch.int32, device=device), torch.ones(B, dtype=torch.int32, device=device)])
    else:
        af = torch.as_tensor(aug_frame).to(device)
        apply = torch.cat([af == 1, af == 2]).to(torch.int32)
    return torch.cat([T, T], 0).contiguous(), apply


def PreProcess(PC_f1, PC_f2, T_gt, T_trans, T_trans_inv, aug_frame):
    """35 m crop, rigid augmentation of one frame, ground truth as (q, t) (model_util.py:346-445).
    Returns (PC_f1_aft_aug (B,N,3), PC_f2_aft_aug (B,N,3), q_gt (B,4), t_gt (B,3,1)).  Inside get_model
    this work is fused into the projection kernel instead."""
    B = PC_f1.shape[0]
    dev = PC_f1.device
    outs = []
    T2, apply = aug_setup(T_trans, aug_frame, B, dev)
    for frame, pc in ((0, PC_f1), (1, PC_f2)):
        _, _, pts = project_points(pc, None, 2, 8, mode=1, T=T2[:B], want_points=True,
                                   T_apply=apply[frame * B:(frame + 1) * B])
        outs.append(pts)
    q_gt, t_gt = gt_pose(T_gt, T_trans, T_trans_inv, aug_frame)
    return outs[0], outs[1], q_gt, t_gt
