"""Writes tests/golden/kitti_golden.npz from the REFERENCE's own code (run here, where /root/reference exists):

  * kitti_dataset.OdometryDataset.__getitem__ on a tiny fake KITTI tree (two sequences, three scans each;
    calib + poses: the reference's ground_truth_pose/kitti_T_diff/*.npy);
  * main.DataAugmentation and main.quat2mat (main.py imports TensorFlow, so the two functions are executed from
    their source text);
  * kitti_evaluation.kittiOdomEval.{loadPoses, trajectoryDistances, calcSequenceErrors, computeOverallErr} on the
    reference's ground-truth trajectory of sequence 04 against a perturbed copy.

matplotlib and the reference's missing `tools` package are stubbed: only the plotting code uses them.
    python tests/golden/make_kitti_golden.py
"""
import ast
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def stub(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


for n in ("matplotlib", "matplotlib.pyplot", "matplotlib.backends", "matplotlib.backends.backend_pdf", "tools",
          "tools.transformations", "tools.pose_evaluation_utils"):
    stub(n)
sys.modules["matplotlib"].use = lambda *a, **k: None
sys.modules["matplotlib.pyplot"].switch_backend = lambda *a, **k: None
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["matplotlib"].backends = sys.modules["matplotlib.backends"]
sys.modules["matplotlib.backends"].backend_pdf = sys.modules["matplotlib.backends.backend_pdf"]
sys.modules["tools.pose_evaluation_utils"].quat_pose_to_mat = None
sys.path.insert(0, REF)
os.chdir(REF)                                   # the reference opens ground_truth_pose/... relative to the cwd
import kitti_dataset  # noqa: E402
import kitti_evaluation  # noqa: E402


def functions_from_source(path, names):
    tree = ast.parse(open(path).read())
    ns = {"np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return [ns[n] for n in names]


def fake_kitti_tree(root, rng):
    """Two sequences with three small scans each; returns the calib Tr rows used."""
    trs = {}
    for seq in ("00", "04"):
        d = os.path.join(root, seq, "velodyne")
        os.makedirs(d)
        tr = np.array([4.276802385584e-04, -9.999672484946e-01, -8.084491683471e-03, -1.198459927713e-02,
                       -7.210626507497e-03, 8.081198471645e-03, -9.999413164504e-01, -5.403984729748e-02,
                       9.999738645903e-01, 4.859485810390e-04, -7.206933692422e-03, -2.921968648686e-01])
        tr = tr + (0.001 if seq == "04" else 0.0)
        trs[seq] = tr
        with open(os.path.join(root, seq, "calib.txt"), "w") as f:
            f.write("P0: 7.070912000000e+02 0.000000000000e+00 6.018873000000e+02 0.0\n")
            f.write("Tr: " + " ".join("%.12e" % v for v in tr) + "\n")
        for i in range(3):
            n = 50 + 7 * i + (3 if seq == "04" else 0)
            pts = (rng.standard_normal((n, 4)) * 10).astype(np.float32)
            pts.tofile(os.path.join(d, "%06d.bin" % i))
    return trs


def main():
    out = {}
    rng = np.random.default_rng(0)
    with tempfile.TemporaryDirectory() as root:
        fake_kitti_tree(root, rng)
        ds = kitti_dataset.OdometryDataset(root=root, NUM_POINTS=128)
        # items: first frame of seq 00, second and third frame of seq 00, frames of seq 04
        base04 = ds.len_list[4]
        idxs = [0, 1, 2, base04, base04 + 1, base04 + 2]
        out["ds_idx"] = np.array(idxs)
        for k, i in enumerate(idxs):
            pos2, pos1, n2, n1, T_gt = ds[i]
            out["ds_pos2_%d" % k], out["ds_pos1_%d" % k] = pos2, pos1
            out["ds_n_%d" % k] = np.array([n2, n1])
            out["ds_T_%d" % k] = T_gt
        # the tree itself, so the test can rebuild it byte for byte
        for seq in ("00", "04"):
            out["calib_" + seq] = np.frombuffer(open(os.path.join(root, seq, "calib.txt"), "rb").read(), dtype=np.uint8)
            for i in range(3):
                out["bin_%s_%d" % (seq, i)] = np.fromfile(os.path.join(root, seq, "velodyne", "%06d.bin" % i), dtype=np.float32)
        out["diff_00"] = np.load(os.path.join(REF, "ground_truth_pose/kitti_T_diff/00_diff.npy"))[:4]
        out["diff_04"] = np.load(os.path.join(REF, "ground_truth_pose/kitti_T_diff/04_diff.npy"))[:4]

    DataAugmentation, quat2mat = functions_from_source(os.path.join(REF, "main.py"), ["DataAugmentation", "quat2mat"])
    np.random.seed(1234)
    out["aug"] = np.stack([DataAugmentation() for _ in range(5)])
    qs = rng.standard_normal((6, 4))
    qs[5] = 1e-6                                   # the near-zero branch
    out["quat_in"] = qs
    out["quat_mat"] = np.stack([quat2mat(q) for q in qs])

    # metric: reference GT trajectory of sequence 04 vs a perturbed relative-pose chain
    ev = kitti_evaluation.kittiOdomEval.__new__(kitti_evaluation.kittiOdomEval)
    ev.lengths = [100, 200, 300, 400, 500, 600, 700, 800]
    ev.num_lengths = 8
    gt_file = os.path.join(REF, "ground_truth_pose/04.txt")
    poses_gt = ev.loadPoses(gt_file, toCameraCoord=False)
    diffs = np.load(os.path.join(REF, "ground_truth_pose/kitti_T_diff/04_diff.npy"))
    T = np.eye(4)
    rows = []
    for d in diffs:
        D = np.eye(4)
        D[:3, :] = d.reshape(3, 4)
        D[:3, 3] *= 1.0 + 0.02 * rng.standard_normal()          # 2 % scale noise on every step
        a = 0.001 * rng.standard_normal()
        Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
        D[:3, :3] = Rz @ D[:3, :3]
        T = T @ D
        rows.append(T[:3, :].reshape(12).copy())
    rows = np.stack(rows)
    pred_file = os.path.join(tempfile.gettempdir(), "elo_golden_04_pred.txt")
    np.savetxt(pred_file, rows, fmt="%.08f")
    poses_res = ev.loadPoses(pred_file, toCameraCoord=False)
    os.remove(pred_file)
    err = ev.calcSequenceErrors(poses_gt, poses_res)
    out["metric_gt_rows"] = np.stack([poses_gt[k][:3, :].reshape(12) for k in sorted(poses_gt)])
    out["metric_pred_rows"] = np.stack([poses_res[k][:3, :].reshape(12) for k in sorted(poses_res)])
    out["metric_dist"] = np.array(ev.trajectoryDistances(poses_gt))
    out["metric_err"] = np.array(err)
    out["metric_overall"] = np.array(ev.computeOverallErr(err))
    np.savez_compressed(os.path.join(HERE, "kitti_golden.npz"), **out)
    print("wrote kitti_golden.npz:", len(out), "arrays;", len(err), "segments; overall", out["metric_overall"])


if __name__ == "__main__":
    main()
