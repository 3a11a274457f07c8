"""KITTI odometry evaluation through the B200 path (the reference's `main.py --mode test` + kitti_evaluation.py):

    python tools/eval_kitti.py --data_root /data/kitti/dataset --checkpoint /path/to/pretrained_model \
        --gt_dir ground_truth_pose --seqs 7 8 9 10 --out result/

For every sequence: read scans (kitti.OdometryDataset), stream frame pairs through PWCLOPipeline, chain the poses,
write <seq>_pred.txt and print t_rel [%] / r_rel [deg/100 m] in the reference's output format."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import elo_b200 as elo  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data_root", required=True, help="KITTI odometry `dataset` directory (NN/velodyne, NN/calib.txt)")
    ap.add_argument("--checkpoint", required=True, help="directory of the TF checkpoint, or an .npz of named tensors")
    ap.add_argument("--pose_dir", default="ground_truth_pose/kitti_T_diff")
    ap.add_argument("--gt_dir", default="ground_truth_pose")
    ap.add_argument("--seqs", nargs="+", type=int, default=[7, 8, 9, 10])
    ap.add_argument("--batch_size", type=int, default=1)
    ap.add_argument("--max_frames", type=int, default=None)
    ap.add_argument("--out", default="result")
    a = ap.parse_args()
    if a.checkpoint.endswith(".npz"):
        import numpy as np
        import torch
        P = {k: torch.from_numpy(v) for k, v in np.load(a.checkpoint).items()}
    else:
        P = elo.tf_checkpoint.load_reference_checkpoint(a.checkpoint)
    store = elo.ParamStore(P, "cuda:0")
    ds = elo.kitti.OdometryDataset(root=a.data_root, pose_dir=a.pose_dir)
    os.makedirs(a.out, exist_ok=True)
    for seq in a.seqs:
        traj = elo.kitti.run_sequence(ds, seq, store, batch_size=a.batch_size, max_frames=a.max_frames)
        pred = os.path.join(a.out, "%02d_pred.txt" % seq)
        traj.save(pred)
        gt = os.path.join(a.gt_dir, "%02d.txt" % seq)
        if os.path.exists(gt):
            t_rel, r_rel = elo.kitti.evaluate_sequence(gt, traj.rows())
            print("seq%02d Average_t_error %.2f Average_r_error %.2f" % (seq, t_rel, r_rel))
        else:
            print("seq%02d: %d poses written to %s (no ground truth)" % (seq, len(traj.rows()), pred))


if __name__ == "__main__":
    main()
