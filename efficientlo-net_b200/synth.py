"""Synthetic KITTI-shaped LiDAR scans (there is no KITTI data in this environment).

Generator specified in SURVEY.md section 8(d): a ground plane + a smooth closed wall seen by an
HDL-64-like sensor whose beams sit at the centres of the H x W cells of the reference's spherical
projection (model_util.py:189-200, 234-242), so every point re-projects to its own cell; 1 % range
noise; 10 % of the cells dropped to (0,0,0); frame 2 is the same scene under a small rigid motion
of KITTI magnitude (ground_truth_pose/kitti_T_diff).  Points are flattened row-major and zero-padded
to NUM_POINTS like kitti_dataset.py:38-103 does.
"""
import math

import torch

VERTICAL_DOWN_DEG = -24.8   # model_util.py:192
VERTICAL_UP_DEG = 2.0       # model_util.py:193


def sensor_angles(H, W):
    """Elevation per row and azimuth per column of the cell centres (float64)."""
    d2r = math.pi / 180.0
    daz = 2.0 * math.pi / W
    dv = (VERTICAL_UP_DEG - VERTICAL_DOWN_DEG) * d2r / (H - 1)
    off = -VERTICAL_DOWN_DEG * d2r / dv
    i = torch.arange(H, dtype=torch.float64)
    j = torch.arange(W, dtype=torch.float64)
    beta = ((H - i) - off + 0.5) * dv
    alpha = math.pi - (j + 0.5) * daz
    return beta, alpha


def synth_scan(H=64, W=1800, seed=0, dropout=0.10):
    """One range image as an (H, W, 3) float32 tensor of xyz (empty cells are exactly zero)."""
    g = torch.Generator().manual_seed(seed)
    beta, alpha = sensor_angles(H, W)
    a = torch.rand(4, generator=g, dtype=torch.float64) * 0.25
    phi = torch.rand(4, generator=g, dtype=torch.float64) * 2.0 * math.pi
    k = torch.arange(1, 5, dtype=torch.float64)
    r_wall = 12.0 + 8.0 * (a[None, :] * torch.sin(k[None, :] * alpha[:, None] + phi[None, :])).sum(1)  # (W,)
    # a few azimuth bands open onto far structure (> 35 m) so the crop of model_util.py:380-383 is exercised
    r_wall = r_wall + 30.0 * (torch.sin(3.0 * alpha + phi[0]) > 0.95).double()
    r_ground = torch.where(beta < 0, 1.73 / torch.sin(-beta).clamp_min(1e-9), torch.full_like(beta, 1e9))
    r = torch.minimum(r_ground[:, None], r_wall[None, :])
    r = r * (1.0 + 0.01 * torch.randn(H, W, generator=g, dtype=torch.float64))
    r = r.clamp(2.0, 80.0)
    cb, sb = torch.cos(beta)[:, None], torch.sin(beta)[:, None]
    xyz = torch.stack([r * cb * torch.cos(alpha)[None, :], r * cb * torch.sin(alpha)[None, :],
                       r * sb.expand(H, W)], dim=-1)
    keep = torch.rand(H, W, generator=g) >= dropout
    xyz = xyz * keep[..., None]
    return xyz.to(torch.float32)


def synth_motion(seed):
    """A KITTI-magnitude rigid motion as a 4x4 float64 matrix (yaw, forward x, small y/z)."""
    g = torch.Generator().manual_seed(1000 + seed)
    yaw = float(torch.randn(1, generator=g, dtype=torch.float64) * 0.01)
    tx = float(torch.rand(1, generator=g, dtype=torch.float64) + 0.5)
    ty, tz = (torch.randn(2, generator=g, dtype=torch.float64) * 0.02).tolist()
    T = torch.eye(4, dtype=torch.float64)
    c, s = math.cos(yaw), math.sin(yaw)
    T[0, 0], T[0, 1], T[1, 0], T[1, 1] = c, -s, s, c
    T[0, 3], T[1, 3], T[2, 3] = tx, ty, tz
    return T


def synth_pair(H=64, W=1800, seed=0, num_points=150000):
    """One frame pair in the reference's input layout: pointclouds (2*num_points, 6) float32 with
    frame 1 in rows [0, N) and frame 2 in rows [N, 2N) (pwclo_model.py:54-55; only xyz is read),
    and T_gt (4, 4) float32 with p2 = T_gt p1."""
    if H * W > num_points:
        raise ValueError("num_points must be at least H*W")
    f1 = synth_scan(H, W, seed).reshape(-1, 3)
    T = synth_motion(seed)
    nz = (f1 != 0).any(dim=1, keepdim=True)
    f2 = (f1.double() @ T[:3, :3].T + T[:3, 3]).float() * nz
    pc = torch.zeros(2 * num_points, 6, dtype=torch.float32)
    pc[: H * W, :3] = f1
    pc[num_points: num_points + H * W, :3] = f2
    return pc, T.float()


def synth_batch(B, H=64, W=1800, num_points=150000, seed0=0):
    pcs, Ts = zip(*[synth_pair(H, W, seed0 + s, num_points) for s in range(B)])
    return torch.stack(pcs), torch.stack(Ts)


def hw_index(B, H, W, device="cpu"):
    """(B, H*W, 2) int32 [h, w] of every cell, row-major -- utils/pointnet_util.py:23-30 (get_hw_idx)."""
    hh = torch.arange(H, dtype=torch.int32, device=device)[:, None].expand(H, W)
    ww = torch.arange(W, dtype=torch.int32, device=device)[None, :].expand(H, W)
    return torch.stack([hh, ww], dim=-1).reshape(1, H * W, 2).expand(B, -1, -1).contiguous()
