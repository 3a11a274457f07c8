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


def load_params(checkpoint):
    """Flat parameter dict of params.py from an .npz of named tensors or a TF-1.x checkpoint (directory or prefix)."""
    if checkpoint.endswith(".npz"):
        import numpy as np
        import torch
        return {k: torch.from_numpy(v) for k, v in np.load(checkpoint).items()}
    P, _step = elo.tf_checkpoint.load_reference_checkpoint(elo.tf_checkpoint.resolve_prefix(checkpoint))
    return P


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--data_root", required=True, help="KITTI odometry `dataset` directory (NN/velodyne, NN/calib.txt)")
    ap.add_argument("--checkpoint", required=True,
                    help="TF checkpoint: its directory (resolved through the `checkpoint` file or *.ckpt.index) or the "
                         "prefix `.../pretrained_model.ckpt`; or an .npz of named tensors")
    ap.add_argument("--pose_dir", default="ground_truth_pose/kitti_T_diff")
    ap.add_argument("--gt_dir", default="ground_truth_pose")
    ap.add_argument("--seqs", nargs="+", type=int, default=[7, 8, 9, 10])
    ap.add_argument("--batch_size", type=int, default=1)
    ap.add_argument("--max_frames", type=int, default=None)
    ap.add_argument("--out", default="result")
    ap.add_argument("--check_only", action="store_true", help="load the checkpoint, print its size and exit (no GPU)")
    a = ap.parse_args(argv)
    P = load_params(a.checkpoint)
    if a.check_only:
        print("%d tensors, %d values" % (len(P), sum(v.numel() for v in P.values())))
        return
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
            if t_rel != t_rel:        # NaN: no 100 m segment was completed (a short --max_frames run)
                print("seq%02d: trajectory too short for the KITTI metric (no completed 100 m segment)" % seq)
            else:
                print("seq%02d Average_t_error %.2f Average_r_error %.2f" % (seq, t_rel, r_rel))
        else:
            print("seq%02d: %d poses written to %s (no ground truth)" % (seq, len(traj.rows()), pred))


if __name__ == "__main__":
    main()
