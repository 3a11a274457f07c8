"""CPU: the host-side front end (KITTI reader, augmentation, batch assembly), the trajectory accumulation and the
KITTI relative-pose metric against outputs of the reference's own code (tests/golden/kitti_golden.npz, written by
tests/golden/make_kitti_golden.py from /root/reference)."""
import importlib
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(HERE, "golden", "kitti_golden.npz"))


@pytest.fixture(scope="module")
def kitti():
    return importlib.import_module("efficientlo-net_b200.kitti")


@pytest.fixture()
def tree(tmp_path, G):
    """The fake KITTI tree the goldens were made from, rebuilt byte for byte."""
    for seq in ("00", "04"):
        d = tmp_path / "dataset" / seq / "velodyne"
        d.mkdir(parents=True)
        (tmp_path / "dataset" / seq / "calib.txt").write_bytes(G["calib_" + seq].tobytes())
        for i in range(3):
            G["bin_%s_%d" % (seq, i)].tofile(str(d / ("%06d.bin" % i)))
    pose = tmp_path / "poses"
    pose.mkdir()
    np.save(str(pose / "00_diff.npy"), G["diff_00"])
    np.save(str(pose / "04_diff.npy"), G["diff_04"])
    return str(tmp_path / "dataset"), str(pose)


def test_dataset_items_match_the_reference_reader(kitti, G, tree):
    root, pose = tree
    ds = kitti.OdometryDataset(root=root, NUM_POINTS=128, pose_dir=pose)
    assert len(ds) == 43552
    for k, idx in enumerate(G["ds_idx"]):
        pos2, pos1, n2, n1, T_gt = ds[int(idx)]
        assert np.array_equal(pos2, G["ds_pos2_%d" % k]) and np.array_equal(pos1, G["ds_pos1_%d" % k])
        assert [n2, n1] == G["ds_n_%d" % k].tolist()
        assert np.array_equal(T_gt, G["ds_T_%d" % k])
    # first frame of a sequence is paired with itself; otherwise (frame i, frame i-1), frame i FIRST
    assert ds.locate(0) == (0, 0, 0) and ds.locate(2) == (0, 1, 2) and ds.locate(4541) == (1, 0, 0)
    with pytest.raises(IndexError):
        ds.locate(43552)


def test_batch_assembly(kitti, G, tree):
    root, pose = tree
    ds = kitti.OdometryDataset(root=root, NUM_POINTS=128, pose_dir=pose)
    idxs = [int(i) for i in G["ds_idx"]]
    data, T_gt, T_trans, T_inv = kitti.get_batch(ds, idxs, 1, 4, training=0, NUM_POINTS=128)
    assert data.shape == (3, 256, 6) and T_gt.shape == (3, 4, 4)
    # main.py:320-323 unpacks dataset[...] as (pc1, pc2, ...): the reader's FIRST value (the later scan) fills rows [0,N)
    assert np.array_equal(data[0, :128, :3], G["ds_pos2_1"]) and np.array_equal(data[0, 128:, :3], G["ds_pos1_1"])
    assert np.array_equal(data[..., 3:], np.zeros_like(data[..., 3:]))
    assert np.array_equal(T_gt[2], G["ds_T_3"])
    assert np.array_equal(T_trans, np.tile(np.eye(4), (3, 1, 1))) and np.array_equal(T_inv, T_trans)
    rng = np.random.RandomState(5)
    _, _, T_trans, T_inv = kitti.get_batch(ds, idxs, 0, 2, training=1, NUM_POINTS=128, rng=rng)
    assert np.allclose(T_trans @ T_inv, np.tile(np.eye(4), (2, 1, 1)), atol=1e-12)
    assert not np.allclose(T_trans[0], np.eye(4))


def test_augmentation_draws_like_the_reference(kitti, G):
    np.random.seed(1234)
    got = np.stack([kitti.DataAugmentation() for _ in range(5)])
    assert np.array_equal(got, G["aug"])
    assert np.all(np.abs(got[:, :3, 3]) <= [1.0, 0.2, 0.15])


def test_quat2mat_and_trajectory(kitti, G):
    for q, M in zip(G["quat_in"], G["quat_mat"]):
        assert np.array_equal(kitti.quat2mat(q), M)
    # chaining identity motions in any calibration frame stays at the origin
    Tr = np.eye(4)
    Tr[:3, :3] = kitti.quat2mat([0.5, 0.5, -0.5, 0.5])
    Tr[:3, 3] = [0.1, -0.2, 0.3]
    tj = kitti.Trajectory(Tr)
    tj.extend([[1, 0, 0, 0]] * 3, [[0, 0, 0]] * 3)
    assert np.allclose(tj.rows(), np.tile(np.eye(4)[:3].reshape(12), (3, 1)))
    # a pure forward motion in the velodyne frame (x) is a forward motion in the camera frame (z) for KITTI's Tr
    R_c2l = np.array([[0, -1, 0, 0], [0, 0, -1, 0], [1, 0, 0, 0], [0, 0, 0, 1.0]])
    tj = kitti.Trajectory(R_c2l)
    tj.append([1, 0, 0, 0], [1.0, 0, 0])
    tj.append([1, 0, 0, 0], [1.0, 0, 0])
    assert np.allclose(tj.rows()[-1].reshape(3, 4)[:, 3], [0, 0, 2.0])


def test_kitti_metric_matches_the_reference(kitti, G, tmp_path):
    gt = kitti.poses_from_rows(G["metric_gt_rows"])
    pred = kitti.poses_from_rows(G["metric_pred_rows"])
    assert np.array_equal(np.array(kitti.trajectory_distances(gt)), G["metric_dist"])
    err = kitti.calc_sequence_errors(gt, pred)
    assert np.array_equal(np.array(err), G["metric_err"])
    assert np.array_equal(np.array(kitti.compute_overall_err(err)), G["metric_overall"])
    # through files, like the reference's evaluation script
    f = tmp_path / "04.txt"
    np.savetxt(str(f), G["metric_gt_rows"], fmt="%.12e")
    t_rel, r_rel = kitti.evaluate_sequence(str(f), G["metric_pred_rows"])
    assert abs(t_rel - G["metric_overall"][0] * 100) < 1e-6
    assert abs(r_rel - G["metric_overall"][1] / np.pi * 180 * 100) < 1e-6
    # a perfect trajectory has zero error
    e0 = kitti.calc_sequence_errors(gt, gt)
    assert max(e[2] for e in e0) < 1e-12 and len(e0) == len(err)


def test_metric_of_a_truncated_run_is_nan_not_a_crash(kitti, G):
    """A trajectory with no completed 100 m segment (eval_kitti --max_frames) has no metric: NaN, no ZeroDivisionError."""
    gt = kitti.poses_from_rows(G["metric_gt_rows"][:5])
    err = kitti.calc_sequence_errors(gt, gt)
    assert err == []
    t, r = kitti.compute_overall_err(err)
    assert np.isnan(t) and np.isnan(r)


def test_packed_batch_is_the_reference_batch_without_padding(kitti, G, tree):
    """get_batch_packed: the xyz rows that hold points, as float32 -- equal to get_batch's rows, which are zero beyond."""
    import torch
    root, pose = tree
    ds = kitti.OdometryDataset(root=root, NUM_POINTS=128, pose_dir=pose)
    idxs = [int(i) for i in G["ds_idx"]]
    data, T_gt, T_trans, T_inv = kitti.get_batch(ds, idxs, 1, 4, training=0, NUM_POINTS=128)
    f1, f2, T_gt_p, T_trans_p, T_inv_p = kitti.get_batch_packed(ds, idxs, 1, 4, training=0)
    assert f1.dtype == np.float32 and f1.shape[0] == 3 and f1.shape[2] == 3 and f1.shape[1] <= 128
    n1, n2 = f1.shape[1], f2.shape[1]
    assert np.array_equal(data[:, :n1, :3], f1) and not data[:, n1:128].any()
    assert np.array_equal(data[:, 128:128 + n2, :3], f2) and not data[:, 128 + n2:].any()
    assert np.array_equal(T_gt, T_gt_p) and np.array_equal(T_trans, T_trans_p) and np.array_equal(T_inv, T_inv_p)
    # in place into preallocated (pinned in production) torch staging buffers; stale rows of an earlier batch are cleared
    out = (torch.full((4, 128, 3), 7.0), torch.full((4, 128, 3), 7.0))
    g1, g2, _, _, _ = kitti.get_batch_packed(ds, idxs, 1, 4, training=0, out=out)
    assert np.array_equal(g1.numpy(), f1) and np.array_equal(g2.numpy(), f2)
    assert g1.data_ptr() == out[0].data_ptr()
    # the augmentation draws are the reference's (same rng stream as get_batch)
    a = kitti.get_batch(ds, idxs, 0, 2, training=1, NUM_POINTS=128, rng=np.random.RandomState(5))
    b = kitti.get_batch_packed(ds, idxs, 0, 2, training=1, rng=np.random.RandomState(5))
    assert np.array_equal(a[2], b[3]) and np.array_equal(a[3], b[4])
