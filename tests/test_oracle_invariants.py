"""CPU: independent checks of the graph restatement (oracle/graph_oracle.py).

The graph blocks have no TensorFlow oracle here ("parity unpinned"): the torch restatement is the
yardstick of every floating-point GPU test.  These tests pin the restatement itself from the outside -- against
independent implementations (scipy's rotations, numpy's float64 trigonometry, torch's own batch norm, the reference's
quat2mat as pinned by tests/golden/kitti_golden.npz) and against algebraic properties the reference's formulas must
have (model_util.py / pointnet_util.py / pwclo_model.py lines cited per test).  Everything in float64 unless the
property is about fp32 behaviour."""
import importlib
import math

import numpy as np
import pytest
import torch
from scipy.spatial.transform import Rotation

from oracle import graph_oracle as go

D = torch.float64


def unit_q(n, seed):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(n, 4, generator=g, dtype=D)
    return q / q.norm(dim=-1, keepdim=True)


def scipy_rot(q_wxyz):
    q = q_wxyz.numpy()
    return Rotation.from_quat(np.stack([q[:, 1], q[:, 2], q[:, 3], q[:, 0]], -1))      # scipy wants (x, y, z, w)


# ---- quaternion warp (model_util.py:17-69, pwclo_model.py:213-227) ---------------------------------------------------
def test_warp_is_the_rotation_scipy_computes():
    q = unit_q(5, 0)
    g = torch.Generator().manual_seed(1)
    xyz = torch.randn(5, 40, 3, generator=g, dtype=D) * 20
    xyz[:, ::7] = 0.0                                                 # empty points stay empty, whatever t is
    t = torch.randn(5, 3, generator=g, dtype=D)
    got = go.warp(xyz, q, t)
    want = torch.from_numpy(np.stack([scipy_rot(q)[b].apply(xyz[b].numpy()) for b in range(5)])) + t[:, None]
    want[:, ::7] = 0.0
    # inv_q divides by |q|^2 + 1e-10 (model_util.py:66): a relative 1e-10 on coordinates of ~60
    assert torch.allclose(got, want, rtol=0, atol=1e-7)
    # a rotation preserves lengths (t = 0)
    assert torch.allclose(go.warp(xyz, q, torch.zeros_like(t)).norm(dim=-1), xyz.norm(dim=-1), atol=1e-7)


def test_quaternion_inverse_and_kitti_quat2mat_agree():
    kitti = importlib.import_module("efficientlo-net_b200.kitti")
    q = unit_q(6, 2) * torch.tensor([[1.0], [2.0], [0.5], [1.0], [3.0], [1.0]], dtype=D)     # non-unit allowed
    one = go.hamilton(q, go.inv_q(q))
    assert torch.allclose(one, torch.tensor([1.0, 0, 0, 0], dtype=D).expand(6, 4), atol=1e-8)
    # the reference's quat2mat (main.py:401-434; kitti.quat2mat is pinned to it by the golden file) gives the matrix of
    # the same rotation the warp applies
    p = torch.randn(6, 1, 3, generator=torch.Generator().manual_seed(3), dtype=D)
    for b in range(6):
        R = kitti.quat2mat(q[b].numpy())
        qn = (q[b] / q[b].norm())[None]
        assert np.allclose(go.warp(p[b:b + 1], qn, torch.zeros(1, 3, dtype=D))[0, 0].numpy(), R @ p[b, 0].numpy(), atol=1e-8)


def test_pose_composition_is_the_product_of_the_two_motions():
    """pwclo_model.py:264-280: q = q_det (x) q_coarse, t = q_det t_coarse q_det^-1 + t_det -- warping by the coarse pose
    and then by the refinement equals warping once by the composed pose."""
    qc, qd = unit_q(4, 4), unit_q(4, 5)
    g = torch.Generator().manual_seed(6)
    tc, td = torch.randn(4, 3, generator=g, dtype=D), torch.randn(4, 3, generator=g, dtype=D)
    p = torch.randn(4, 30, 3, generator=g, dtype=D) * 10
    q = go.mul_point_q(qd.reshape(4, 1, 4), qc).squeeze(1)
    tq = torch.cat([torch.zeros(4, 1, 1, dtype=D), tc.reshape(4, 1, 3)], -1)
    t = go.mul_point_q(go.mul_q_point(qd, tq), go.inv_q(qd))[..., 1:].squeeze(1) + td
    twice = go.warp(go.warp(p, qc, tc), qd, td)
    assert torch.allclose(go.warp(p, q, t), twice, atol=1e-7)
    both = scipy_rot(qd) * scipy_rot(qc)
    assert np.allclose(np.abs((scipy_rot(q).inv() * both).magnitude()), 0, atol=1e-9)


# ---- PreProcess ground truth (model_util.py:72-177, 386-426) --------------------------------------------------------
def test_ground_truth_quaternion_is_the_rotation_of_T_gt():
    rots = Rotation.random(8, random_state=7)
    eye = torch.eye(4, dtype=D).expand(8, 4, 4).contiguous()
    T = eye.clone()
    T[:, :3, :3] = torch.from_numpy(rots.as_matrix())
    T[:, :3, 3] = torch.randn(8, 3, generator=torch.Generator().manual_seed(8), dtype=D)
    pc = torch.zeros(8, 4, 3, dtype=D)
    _, _, q_gt, t_gt = go.PreProcess(pc, pc, T, eye, eye, [2] * 8)
    assert torch.allclose(t_gt.squeeze(-1), T[:, :3, 3])
    # mat2euler -> euler2quat is a matrix -> quaternion conversion: same rotation as scipy's, up to the sign of q
    assert np.allclose((scipy_rot(q_gt).inv() * rots).magnitude(), 0, atol=1e-7)
    # augmentation bookkeeping: frame 2 augmented -> T_trans @ T_gt, frame 1 -> T_gt @ T_trans^-1
    A = Rotation.random(8, random_state=9)
    Tt = eye.clone()
    Tt[:, :3, :3] = torch.from_numpy(A.as_matrix())
    Tt[:, :3, 3] = 0.3
    Ti = torch.linalg.inv(Tt)
    _, _, q2, t2 = go.PreProcess(pc, pc, T, Tt, Ti, [2] * 8)
    _, _, q1, t1 = go.PreProcess(pc, pc, T, Tt, Ti, [1] * 8)
    assert torch.allclose(t2.squeeze(-1), (Tt @ T)[:, :3, 3]) and torch.allclose(t1.squeeze(-1), (T @ Ti)[:, :3, 3])
    assert np.allclose((scipy_rot(q2).inv() * (A * rots)).magnitude(), 0, atol=1e-7)
    assert np.allclose((scipy_rot(q1).inv() * (rots * A.inv())).magnitude(), 0, atol=1e-7)


def test_preprocess_crops_at_35_m_and_moves_only_the_augmented_frame():
    pc = torch.tensor([[[10.0, 0, 1], [34.9, 0, 0], [0, 35.1, 2], [30, 30, 0], [0, 0, 0]]], dtype=D)
    eye = torch.eye(4, dtype=D)[None]
    Tt = eye.clone()
    Tt[0, :3, 3] = torch.tensor([1.0, 2.0, 3.0], dtype=D)
    f1, f2, _, _ = go.PreProcess(pc, pc, eye, Tt, torch.linalg.inv(Tt), [2])
    keep = torch.tensor([1.0, 1, 0, 0, 0], dtype=D)[:, None]                    # rows 2, 3 are beyond 35 m in the xy plane
    assert torch.equal(f1[0], pc[0] * keep)                                        # frame 1 untouched apart from the crop
    want2 = (pc[0] + Tt[0, :3, 3]) * keep                                          # cropped points stay (0,0,0): the mask
    assert torch.allclose(f2[0], want2)                                            # is the ORIGINAL emptiness (:419-420)


# ---- spherical projection (model_util.py:181-292) --------------------------------------------------------------------
def numpy_bins(pc, H, W):
    """The binning formulas of model_util.py:234-245 in numpy float64, written independently of the restatement."""
    d2r = math.pi / 180
    az, vres = (360.0 / W) * d2r, (2.0 * d2r + 24.8 * d2r) / (H - 1)
    voff = 24.8 * d2r / vres
    x, y, z = pc[:, 0], pc[:, 1], pc[:, 2]
    r = np.sqrt(x * x + y * y + z * z)
    col = np.trunc((math.pi - np.arctan2(y, x)) / az).astype(np.int64)
    row = H - np.trunc(np.arcsin(z / r) / vres + voff).astype(np.int64)
    return np.clip(row, 0, H - 1), np.clip(col, 0, W - 1), r


def test_projection_keeps_the_nearest_point_of_every_cell_and_sums_exact_ties():
    H, W = 16, 90
    g = np.random.default_rng(10)
    n = 4000
    rng = g.uniform(3, 40, n)
    az, el = g.uniform(-math.pi, math.pi, n), g.uniform(-24 * math.pi / 180, 1.5 * math.pi / 180, n)
    pc = np.stack([rng * np.cos(el) * np.cos(az), rng * np.cos(el) * np.sin(az), rng * np.sin(el)], -1)
    pc = np.concatenate([pc, pc[:50]], 0)                      # 50 exact duplicates: ties accumulate (scatter_nd, :271)
    feat = g.standard_normal((pc.shape[0], 5))
    img, fimg = go.ProjectPC2SphericalRing(torch.from_numpy(pc)[None], torch.from_numpy(feat)[None], H, W)
    row, col, r = numpy_bins(pc, H, W)
    want = np.zeros((H, W, 3))
    wantf = np.zeros((H, W, 5))
    for cell in set(zip(row.tolist(), col.tolist())):
        members = np.nonzero((row == cell[0]) & (col == cell[1]))[0]
        winners = members[r[members] == r[members].min()]
        want[cell] = pc[winners].sum(0)
        wantf[cell] = feat[winners].sum(0)
    assert np.allclose(img[0].numpy(), want, atol=1e-9) and np.allclose(fimg[0].numpy(), wantf, atol=1e-9)
    assert (np.abs(want).sum(-1) > 0).sum() > 1000


def test_projection_of_a_range_image_returns_the_range_image():
    """Synthetic scans put every point at the centre of its own cell (synth.py, the inverse of model_util.py:189-242):
    projecting the flattened image gives the image back -- except the cells the empty pixels land in, which are forced to
    zero."""
    synth = importlib.import_module("efficientlo-net_b200.synth")
    H, W = 64, 1800
    scan = synth.synth_scan(H, W, seed=5).double()
    img, _ = go.ProjectPC2SphericalRing(scan.reshape(1, -1, 3), None, H, W)
    diff = (img[0] != scan).any(-1)
    cells = {tuple(c) for c in diff.nonzero().tolist()}
    # empty pixels are (+-0, +-0, 0): atan2 sends them to azimuth 0, pi or -pi, i.e. the middle or either end of the
    # bottom row, and their range 0 wins there (signed zeros survive the reference's `* mask`, model_util.py:419)
    assert cells <= {(H - 1, 0), (H - 1, W // 2), (H - 1, W - 1)} and (H - 1, W // 2) in cells, cells
    assert bool((scan.reshape(-1, 3) == 0).all(-1).any())      # the scan does contain empty pixels


# ---- masked attention pooling (model_util.py:319-343) ---------------------------------------------------------------
def test_softmax_valid_is_a_convex_combination_of_the_valid_points_only():
    g = torch.Generator().manual_seed(11)
    f, w = torch.randn(2, 50, 8, generator=g, dtype=D), torch.randn(2, 50, 8, generator=g, dtype=D) * 3
    valid = torch.rand(2, 50, generator=g) > 0.3
    out = go.softmax_valid(f, w, valid)
    for b in range(2):
        fv, wv = f[b][valid[b]].numpy(), w[b][valid[b]].numpy()
        e = np.exp(wv - wv.max(0))
        assert np.allclose(out[b, 0].numpy(), (fv * e / e.sum(0)).sum(0), atol=1e-12)
        assert bool((out[b, 0] <= f[b][valid[b]].max(0).values + 1e-12).all())
        assert bool((out[b, 0] >= f[b][valid[b]].min(0).values - 1e-12).all())
    # logits are defined up to a constant per channel; invalid points have no say
    assert torch.allclose(go.softmax_valid(f, w + 7.0, valid), out, atol=1e-12)
    f2, w2 = f.clone(), w.clone()
    f2[~valid] = 1e6
    w2[~valid] = 50.0
    assert torch.equal(go.softmax_valid(f2, w2, valid), out)


# ---- the MLP primitive (tf_util.py:120-185, 512-531) ----------------------------------------------------------------
def layer_params(cin, cout, seed):
    g = torch.Generator().manual_seed(seed)
    return {"s/weights": torch.randn(cin, cout, generator=g, dtype=D), "s/biases": torch.randn(cout, generator=g, dtype=D),
            "s/bn/gamma": torch.rand(cout, generator=g, dtype=D) + 0.5, "s/bn/beta": torch.randn(cout, generator=g, dtype=D),
            "s/bn/moving_mean": torch.randn(cout, generator=g, dtype=D),
            "s/bn/moving_variance": torch.rand(cout, generator=g, dtype=D) + 0.1}


def test_inference_layer_is_torch_batch_norm_and_the_products_bn_fold():
    params = importlib.import_module("efficientlo-net_b200.params")
    P = layer_params(6, 5, 12)
    x = torch.randn(3, 7, 4, 6, generator=torch.Generator().manual_seed(13), dtype=D)
    got = go.conv2d(x, P, "s")
    z = (x @ P["s/weights"] + P["s/biases"]).reshape(-1, 5)
    want = torch.relu(torch.nn.functional.batch_norm(z, P["s/bn/moving_mean"], P["s/bn/moving_variance"], P["s/bn/gamma"],
                                                     P["s/bn/beta"], training=False, eps=1e-3)).reshape(3, 7, 4, 5)
    assert torch.allclose(got, want, atol=1e-12)
    # the folding the CUDA path's weights go through (params.fold_bn) is the same function of the same six tensors
    Wf, bf = params.fold_bn(P, "s")
    assert torch.allclose(torch.relu(x @ Wf.to(D) + bf.to(D)), got, atol=1e-12)


# ---- set-conv (pointnet_util.py:179-250) ------------------------------------------------------------------------------
def test_set_conv_does_not_depend_on_the_scan_order_when_every_candidate_fits():
    """Max over the K neighbours of (mlp * mask): with K >= window cells the neighbour SET is every in-range pixel of the
    window whatever the scan order is, and a maximum does not care about order."""
    params = importlib.import_module("efficientlo-net_b200.params")
    synth = importlib.import_module("efficientlo-net_b200.synth")
    P = {k: v.double() for k, v in params.init_params(0).items()}
    H, W = 12, 40
    xyz = synth.synth_scan(H, W, seed=2).double()[None]
    pts = torch.zeros(1, H, W, 3, dtype=D)
    sel = go.get_selected_idx(1, 2, 4, 6, 10)
    scopes = ["sa1/layer0/conv%d" % j for j in range(3)]
    kH, kW = 3, 5
    outs = []
    for seed in (0, 1, 2):
        perm = np.random.default_rng(seed).permutation(kH * kW).astype(np.int32)
        out, new_xyz = go.down_conv(xyz, pts, sel, kH * kW, (kH, kW), 6.0, scopes, P, perm)
        outs.append(out)
    assert torch.allclose(outs[0], outs[1], atol=1e-12) and torch.allclose(outs[0], outs[2], atol=1e-12)
    assert torch.equal(new_xyz, xyz[:, ::2, ::4][:, :6, :10])
    assert float(outs[0].abs().max()) > 0


# ---- loss (pwclo_model.py:437-481) -----------------------------------------------------------------------------------
def test_loss_at_the_ground_truth_and_its_level_weights():
    q = unit_q(3, 14)
    t = torch.randn(3, 3, generator=torch.Generator().manual_seed(15), dtype=D)
    wx, wq = torch.tensor(0.3, dtype=D), torch.tensor(-2.5, dtype=D)
    eps = 1e-5                                              # sqrt(1e-10) under both square roots (:445-446)
    perfect = go.get_loss(q, t, q, t, q, t, q, t, q, t[..., None], wx, wq)
    level = eps * math.exp(-0.3) + 0.3 + eps * math.exp(2.5) - 2.5
    assert abs(float(perfect) - 3.0 * level) < 1e-9          # 1.6 + 0.8 + 0.4 + 0.2
    # an error at level 3 weighs eight times an equal error at level 0
    off = t + 0.5
    l3 = go.get_loss(q, t, q, t, q, t, q, off, q, t[..., None], wx, wq) - perfect
    l0 = go.get_loss(q, off, q, t, q, t, q, t, q, t[..., None], wx, wq) - perfect
    assert abs(float(l3) / float(l0) - 8.0) < 1e-6
    # the loss normalises q itself (:444): scaling a predicted quaternion changes nothing
    assert abs(float(go.get_loss(q * 3.0, t, q, t, q, t, q, t, q, t[..., None], wx, wq)) - float(perfect)) < 1e-6
