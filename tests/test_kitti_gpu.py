"""GPU: front end -> hot path -> trajectory on a synthetic KITTI-format tree (three 64x1800 scans written as
velodyne .bin files): the streamed sequence run equals frame-by-frame synchronous inference, chained the same way."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_run_sequence_on_a_synthetic_kitti_tree(elo, cuda, tmp_path):
    kitti = elo.kitti
    H, W, N = 64, 1800, 150000
    seq_dir = tmp_path / "dataset" / "04" / "velodyne"
    seq_dir.mkdir(parents=True)
    tr = np.array([0, -1, 0, 0.01, 0, 0, -1, -0.05, 1, 0, 0, -0.3], dtype=np.float64)
    (tmp_path / "dataset" / "04" / "calib.txt").write_text("Tr: " + " ".join("%.12e" % v for v in tr) + "\n")
    diffs = []
    for i in range(3):
        pc, T = elo.synth.synth_pair(H, W, seed=i, num_points=N)
        pts = pc[:N, :3].numpy()
        pts = pts[np.any(pts != 0, axis=1)]
        np.concatenate([pts, np.ones((len(pts), 1), np.float32)], 1).astype(np.float32).tofile(str(seq_dir / ("%06d.bin" % i)))
        diffs.append(T.numpy()[:3].reshape(12))
    pose_dir = tmp_path / "poses"
    pose_dir.mkdir()
    np.save(str(pose_dir / "04_diff.npy"), np.stack(diffs))
    ds = kitti.OdometryDataset(root=str(tmp_path / "dataset"), NUM_POINTS=N, pose_dir=str(pose_dir))
    ds.len_list = [0, 0, 0, 0, 0, 3] + [3] * 17            # sequence 04 holds the three scans
    store = elo.ParamStore(elo.params.init_params(0), cuda)
    perms = elo.params.make_perms(0)
    traj = kitti.run_sequence(ds, 4, store, batch_size=2, perms=perms)
    rows = traj.rows()
    assert rows.shape == (3, 12) and np.isfinite(rows).all()
    # the same three pairs, one by one through the synchronous engine, chained by hand
    eng = elo.PWCLOEngine(1, H, W, N, params=store, perms=perms, device=cuda).capture()
    Tr, Tr_inv = kitti.calib_Tr(str(tmp_path / "dataset" / "04" / "calib.txt"))
    want = kitti.Trajectory(Tr, Tr_inv)
    for i in range(3):
        data, _, _, _ = kitti.get_batch(ds, np.arange(3), i, i + 1, NUM_POINTS=N)
        q, t = eng.infer(torch.from_numpy(data).float().pin_memory())
        want.append(q[0].numpy().astype(np.float64), t[0].numpy().astype(np.float64))
    assert np.allclose(rows, want.rows(), rtol=0, atol=1e-5)
    # first item: a scan paired with itself
    pos2, pos1, n2, n1, T_gt = ds[0]
    assert n1 == n2 and np.array_equal(pos1, pos2)
    out = tmp_path / "04_pred.txt"
    traj.save(str(out))
    assert np.allclose(np.loadtxt(str(out)), rows, atol=1e-8)
