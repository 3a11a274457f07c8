"""The steps either side of the hot path, on the host (numpy): KITTI odometry front end, trajectory
accumulation and the KITTI relative-pose metric.

Mirrors, by name and argument order:
  kitti_dataset.OdometryDataset            kitti_dataset.py:21-103   (note the (pc2, pc1) return order and
                                                                    T_gt = Tr^-1 . T_diff . Tr)
  main.DataAugmentation / get_batch        main.py:259-341
  main.quat2mat + pose chaining            main.py:401-434, 538-565
  kittiOdomEval.loadPoses / trajectoryDistances / rotationError / translationError /
  lastFrameFromSegmentLength / calcSequenceErrors / computeOverallErr      kitti_evaluation.py:74-195
The plotting half of kitti_evaluation.py (matplotlib, the missing `tools/` package) is not part of the path.
Pinned against outputs of the reference's own code: tests/golden/kitti_golden.npz (made by
tests/golden/make_kitti_golden.py, which imports the reference modules).
"""
import os

import numpy as np

LEN_LIST = [0, 4541, 5642, 10303, 11104, 11375, 14136, 15237, 16338, 20409, 22000, 23201, 24122, 25183, 28464,
            29095, 30996, 32727, 33218, 35019, 40000, 40831, 43552]
FILE_MAP = ["%02d" % i for i in range(22)]


def read_calib_file(path):
    """KITTI calib.txt -> {key: float array, or the raw string when it is not numeric} (kitti_dataset.py:107-125)."""
    numeric = set("0123456789.e+- ")
    out = {}
    with open(path, "r") as f:
        for line in f:
            if ":" not in line:
                continue
            key, value = (part.strip() if i else part for i, part in enumerate(line.split(":", 1)))
            out[key] = value
            if set(value) <= numeric:
                try:
                    out[key] = np.array([float(tok) for tok in value.split(" ")])
                except ValueError:
                    pass                                  # e.g. doubled spaces: keep the string, like the reference
    return out


def calib_Tr(path):
    """4x4 velodyne->camera matrix Tr of a calib.txt and its inverse."""
    Tr = np.vstack((read_calib_file(path)["Tr"].reshape(3, 4), np.array([0, 0, 0, 1.0])))
    return Tr, np.linalg.inv(Tr)


class OdometryDataset:
    """Frame pairs of the KITTI odometry benchmark (kitti_dataset.py:21-103).  Item `index` counts frames over
    the concatenated sequences; it pairs frame i (returned FIRST, as pos2) with frame i-1 (pos1), the first
    frame of a sequence with itself.  Point clouds are zero-padded to NUM_POINTS rows; T_gt is the relative
    pose in the velodyne frame.  `pose_dir` holds the reference's <seq>_diff.npy relative camera poses."""

    def __init__(self, root="/tmp/data_odometry_velodyne/dataset", NUM_POINTS=150000, H_input=64, W_input=1800,
                 pose_dir="ground_truth_pose/kitti_T_diff"):
        self.num_points = NUM_POINTS
        self.datapath = root
        self.pose_dir = pose_dir
        self.len_list = LEN_LIST
        self.file_map = FILE_MAP
        self._calib, self._pose = {}, {}

    def locate(self, index):
        """(sequence, frame of pc1, frame of pc2) of item `index`."""
        for seq_idx, seq_num in enumerate(self.len_list):
            if index < seq_num:
                cur_idx_pc2 = index - self.len_list[seq_idx - 1]
                return seq_idx - 1, (0 if cur_idx_pc2 == 0 else cur_idx_pc2 - 1), cur_idx_pc2
        raise IndexError(index)

    def __getitem__(self, index):
        point2, point1, T_gt = self.frames(index)
        n1, n2 = point1.shape[0], point2.shape[0]
        pos1 = np.zeros((self.num_points, 3))
        pos2 = np.zeros((self.num_points, 3))
        pos1[:n1, :3] = point1
        pos2[:n2, :3] = point2
        return pos2, pos1, n2, n1, T_gt

    def frames(self, index):
        """The item without its zero padding, in __getitem__'s order (the LATER scan first): (pos2 (n2, 3) float32,
        pos1 (n1, 3) float32, T_gt) -- what the packed upload path (get_batch_packed, PWCLOEngine(packed=True)) sends
        to the GPU; the padding to NUM_POINTS rows of kitti_dataset.py:76-80 then happens on the device."""
        cur_seq, i1, i2 = self.locate(index)
        seq_dir = os.path.join(self.datapath, self.file_map[cur_seq])
        if cur_seq not in self._calib:
            self._calib[cur_seq] = calib_Tr(os.path.join(seq_dir, "calib.txt"))
        Tr, Tr_inv = self._calib[cur_seq]
        point1 = np.fromfile(os.path.join(seq_dir, "velodyne", "%06d.bin" % i1), dtype=np.float32).reshape(-1, 4)[:, :3]
        point2 = np.fromfile(os.path.join(seq_dir, "velodyne", "%06d.bin" % i2), dtype=np.float32).reshape(-1, 4)[:, :3]
        if cur_seq > 10:
            T_diff = np.ones((1, 12))                      # test sequences have no ground truth (:84-85)
        else:
            if cur_seq not in self._pose:
                self._pose[cur_seq] = np.load(os.path.join(self.pose_dir, self.file_map[cur_seq] + "_diff.npy"))
            T_diff = self._pose[cur_seq][i2:i2 + 1, :]
        T_diff = np.concatenate([T_diff.reshape(3, 4), np.array([[0.0, 0.0, 0.0, 1.0]])], axis=0)
        T_gt = np.matmul(np.matmul(Tr_inv, T_diff), Tr)
        return point2, point1, T_gt

    def __len__(self):
        return self.len_list[-1]


def _rot(axis, angle):
    c, s_ = np.cos(angle), np.sin(angle)
    i, j = [(1, 2), (2, 0), (0, 1)][axis]          # x: rows/cols (1,2); y: (2,0); z: (0,1)
    R = np.eye(3)
    R[i, i], R[i, j], R[j, i], R[j, j] = c, -s_, s_, c
    return R


def DataAugmentation(rng=np.random):
    """Random rigid augmentation matrix of main.py:259-297.  The random draws happen in the reference's order
    (three angles, the unit 'scale', three offsets) so that a seeded numpy state gives the same matrix."""
    def draw(sigma, lim):
        return np.clip(sigma * rng.randn(), -lim, lim).astype(np.float32)

    ax, ay, az = (draw(s_, l) * np.pi / 4.0 for s_, l in ((0.01, 0.02), (0.01, 0.02), (0.05, 0.1)))
    scale = np.diag(rng.uniform(1.00, 1.00, 3).astype(np.float32))
    shift = [draw(0.5, 1.0), draw(0.1, 0.2), draw(0.05, 0.15)]
    T = np.eye(4)
    T[:3, :3] = _rot(0, ax).dot(_rot(1, ay)).dot(_rot(2, az)).dot(scale.T)
    T[:3, 3] = shift
    return T


def get_batch(dataset, idxs, start_idx, end_idx, training=0, NUM_POINTS=150000, rng=np.random):
    """(batch_data (b, 2N, 6), T_gt, T_trans, T_trans_inv) of main.py:301-341: frame pc1 in rows [0,N),
    pc2 in [N,2N); identity augmentation unless training."""
    bsize = end_idx - start_idx
    batch_data = np.zeros((bsize, NUM_POINTS * 2, 6))
    batch_T_gt = np.zeros((bsize, 4, 4))
    batch_T_trans = np.tile(np.expand_dims(np.eye(4), axis=0), [bsize, 1, 1])
    batch_T_trans_inv = batch_T_trans.copy()
    for i in range(bsize):
        pc1, pc2, n1, n2, T_gt = dataset[idxs[i + start_idx]]
        batch_data[i, :NUM_POINTS, :3] = pc1
        batch_data[i, NUM_POINTS:, :3] = pc2
        batch_T_gt[i] = T_gt
        if training != 0:
            T_trans = DataAugmentation(rng)
            batch_T_trans[i] = T_trans
            batch_T_trans_inv[i] = np.linalg.inv(T_trans)
    return batch_data, batch_T_gt, batch_T_trans, batch_T_trans_inv


def get_batch_packed(dataset, idxs, start_idx, end_idx, training=0, rng=np.random, out=None):
    """The same batch as get_batch in the packed upload format: (xyz_f1 (b, n1, 3) float32, xyz_f2 (b, n2, 3) float32,
    T_gt, T_trans, T_trans_inv) where n1 / n2 are the largest point counts of the batch (shorter samples are
    zero-padded to them, nothing is padded to NUM_POINTS and channels 3:6 -- always zero, main.py:327 -- are not
    materialised).  get_batch(...)[0][:, :n1, :3] == xyz_f1 and rows [n1, NUM_POINTS) of it are zero.
    `out`: optional pair of preallocated (ideally pinned) float32 arrays / tensors (b_max, NUM_POINTS, 3) to fill in
    place; the returned frames are views of their first n rows."""
    bsize = end_idx - start_idx
    items = [dataset.frames(idxs[i + start_idx]) for i in range(bsize)]
    n1 = max(it[0].shape[0] for it in items)
    n2 = max(it[1].shape[0] for it in items)
    if max(n1, n2) > dataset.num_points:
        raise ValueError("a scan has %d points, more than NUM_POINTS = %d" % (max(n1, n2), dataset.num_points))
    if out is None:
        f1, f2 = np.zeros((bsize, n1, 3), np.float32), np.zeros((bsize, n2, 3), np.float32)
    else:
        f1, f2 = out[0][:bsize, :n1], out[1][:bsize, :n2]
    batch_T_gt = np.zeros((bsize, 4, 4))
    batch_T_trans = np.tile(np.expand_dims(np.eye(4), axis=0), [bsize, 1, 1])
    batch_T_trans_inv = batch_T_trans.copy()
    for i, (pc1, pc2, T_gt) in enumerate(items):
        for dst, src in ((f1, pc1), (f2, pc2)):
            dst[i, :src.shape[0]] = _like(dst, src)
            dst[i, src.shape[0]:] = 0
        batch_T_gt[i] = T_gt
        if training != 0:
            T_trans = DataAugmentation(rng)
            batch_T_trans[i] = T_trans
            batch_T_trans_inv[i] = np.linalg.inv(T_trans)
    return f1, f2, batch_T_gt, batch_T_trans, batch_T_trans_inv


def _like(dst, src):
    """src as something dst[...] = accepts (dst may be a numpy array or a torch tensor)."""
    if isinstance(dst, np.ndarray):
        return src
    import torch
    return torch.from_numpy(np.ascontiguousarray(src))


# ---- trajectory ---------------------------------------------------------------------------------------
def quat2mat(q):
    """Rotation matrix of a (w, x, y, z) quaternion, non-unit allowed (main.py:401-434)."""
    w, x, y, z = q
    Nq = w * w + x * x + y * y + z * z
    if Nq < 1e-8:
        return np.eye(3)
    s = 2.0 / Nq
    X, Y, Z = x * s, y * s, z * s
    wX, wY, wZ = w * X, w * Y, w * Z
    xX, xY, xZ = x * X, x * Y, x * Z
    yY, yZ, zZ = y * Y, y * Z, z * Z
    return np.array([[1.0 - (yY + zZ), xY - wZ, xZ + wY],
                     [xY + wZ, 1.0 - (xX + zZ), yZ - wX],
                     [xZ - wY, yZ + wX, 1.0 - (xX + yY)]])


class Trajectory:
    """Accumulates per-pair network outputs (q, t) into absolute camera-frame poses, main.py:538-565:
    TT = Tr . [R(q) | t] . Tr^-1, T_k = T_{k-1} . TT; rows() is what the reference writes to <seq>_pred.txt."""

    def __init__(self, Tr, Tr_inv=None):
        self.Tr = np.asarray(Tr, dtype=np.float64)
        self.Tr_inv = np.linalg.inv(self.Tr) if Tr_inv is None else np.asarray(Tr_inv, dtype=np.float64)
        self.T_final = None
        self._rows = []

    def append(self, q, t):
        TT = np.concatenate([np.concatenate([quat2mat(np.reshape(q, [4])), np.reshape(t, [3, 1])], axis=-1),
                             np.array([[0.0, 0.0, 0.0, 1.0]])], axis=0)
        TT = np.matmul(np.matmul(self.Tr, TT), self.Tr_inv)
        self.T_final = TT if self.T_final is None else np.matmul(self.T_final, TT)
        self._rows.append(self.T_final[:3, :].reshape(12))
        return self.T_final

    def extend(self, qs, ts):
        for q, t in zip(np.asarray(qs), np.asarray(ts)):
            self.append(q, t)

    def rows(self):
        return np.stack(self._rows, 0) if self._rows else np.zeros((0, 12))

    def save(self, path):
        np.savetxt(path, self.rows(), fmt="%.08f")


# ---- KITTI relative-pose metric ---------------------------------------------------------------------
LENGTHS = [100, 200, 300, 400, 500, 600, 700, 800]


def load_poses(file_name):
    """{frame: 4x4} from a KITTI pose file, 12 numbers per line or 13 with a leading index
    (kitti_evaluation.py:74-101, toCameraCoord=False)."""
    poses = {}
    with open(file_name, "r") as f:
        for cnt, line in enumerate(f.readlines()):
            v = [float(i) for i in line.split()]
            with_idx = int(len(v) == 13)
            P = np.eye(4)
            P[:3, :] = np.asarray(v[with_idx:with_idx + 12]).reshape(3, 4)
            poses[v[0] if with_idx else cnt] = P
    return poses


def poses_from_rows(rows):
    out = {}
    for i, r in enumerate(np.asarray(rows)):
        P = np.eye(4)
        P[:3, :] = r.reshape(3, 4)
        out[i] = P
    return out


def trajectory_distances(poses):
    dist = [0]
    keys = sorted(poses.keys())
    for i in range(len(keys) - 1):
        P1, P2 = poses[keys[i]], poses[keys[i + 1]]
        dx, dy, dz = P1[0, 3] - P2[0, 3], P1[1, 3] - P2[1, 3], P1[2, 3] - P2[2, 3]
        dist.append(dist[i] + np.sqrt(dx ** 2 + dy ** 2 + dz ** 2))
    return dist


def rotation_error(pose_error):
    d = 0.5 * (pose_error[0, 0] + pose_error[1, 1] + pose_error[2, 2] - 1.0)
    return np.arccos(max(min(d, 1.0), -1.0))


def translation_error(pose_error):
    return np.sqrt(pose_error[0, 3] ** 2 + pose_error[1, 3] ** 2 + pose_error[2, 3] ** 2)


def last_frame_from_segment_length(dist, first_frame, len_):
    for i in range(first_frame, len(dist), 1):
        if dist[i] > (dist[first_frame] + len_):
            return i
    return -1


def calc_sequence_errors(poses_gt, poses_result, step_size=10):
    """[first_frame, r_err/len, t_err/len, len, speed] for every start frame (every 10th) and every segment
    length 100..800 m that fits (kitti_evaluation.py:141-176)."""
    err = []
    dist = trajectory_distances(poses_gt)
    for first_frame in range(0, len(poses_gt), step_size):
        for len_ in LENGTHS:
            last_frame = last_frame_from_segment_length(dist, first_frame, len_)
            if last_frame == -1 or last_frame not in poses_result or first_frame not in poses_result:
                continue
            pose_delta_gt = np.dot(np.linalg.inv(poses_gt[first_frame]), poses_gt[last_frame])
            pose_delta_result = np.dot(np.linalg.inv(poses_result[first_frame]), poses_result[last_frame])
            pose_error = np.dot(np.linalg.inv(pose_delta_result), pose_delta_gt)
            num_frames = last_frame - first_frame + 1.0
            err.append([first_frame, rotation_error(pose_error) / len_, translation_error(pose_error) / len_, len_,
                        len_ / (0.1 * num_frames)])
    return err


def compute_overall_err(seq_err):
    """(mean t_err, mean r_err) over the segments; the reference prints t*100 [%] and r/pi*180*100 [deg/100 m]
    (kitti_evaluation.py:185-195, 626)."""
    if not seq_err:               # no completed 100 m segment (a truncated run): the metric is undefined
        return float("nan"), float("nan")
    t_err = sum(e[2] for e in seq_err)
    r_err = sum(e[1] for e in seq_err)
    return t_err / len(seq_err), r_err / len(seq_err)


def evaluate_sequence(gt_file, pred_rows):
    """t_rel [%] and r_rel [deg/100 m] of a predicted trajectory (rows as Trajectory.rows())."""
    err = calc_sequence_errors(load_poses(gt_file), poses_from_rows(pred_rows))
    t, r = compute_overall_err(err)
    return t * 100, r / np.pi * 180 * 100


# ---- sequence evaluation: front end -> hot path -> trajectory (main.py:459-582) -----------------------------
def run_sequence(dataset, seq, params, batch_size=1, H_input=64, W_input=1800, device="cuda:0", max_frames=None,
                 perms=None):
    """Predict the trajectory of KITTI sequence `seq`: every frame of the sequence paired with its predecessor
    (the first with itself), batches streamed through PWCLOPipeline from pinned host memory, finest-level (q, t)
    chained in the camera frame.  `params`: a ParamStore or flat parameter dict.  Returns a Trajectory."""
    import torch

    from .engine import PWCLOPipeline

    start, end = dataset.len_list[seq], dataset.len_list[seq + 1]
    if max_frames is not None:
        end = min(end, start + max_frames)
    idxs = np.arange(start, end)
    Tr, Tr_inv = calib_Tr(os.path.join(dataset.datapath, dataset.file_map[seq], "calib.txt"))
    N = dataset.num_points
    # packed upload: only the xyz rows that hold points travel (pinned fp32 staging buffers, filled in place); the
    # zero padding to NUM_POINTS and the (unused) channels 3:6 of main.py:327 exist only on the device
    pipe = PWCLOPipeline(batch_size, H_input, W_input, N, params=params, perms=perms, device=device, packed=True)
    hosts = [(torch.zeros(batch_size, N, 3).pin_memory(), torch.zeros(batch_size, N, 3).pin_memory())
             for _ in range(len(pipe.engines) + 1)]
    sizes = []

    def batches():
        for k, b0 in enumerate(range(0, len(idxs), batch_size)):
            b1 = min(len(idxs), b0 + batch_size)
            buf = hosts[k % len(hosts)]
            f1, f2, T_gt, _, _ = get_batch_packed(dataset, idxs, b0, b1, training=0, out=buf)
            sizes.append(b1 - b0)
            # a short last batch keeps the previous batch's samples in its tail rows, like main.py:509-515
            yield (buf[0][:, :f1.shape[1]], buf[1][:, :f2.shape[1]]), None

    traj = Trajectory(Tr, Tr_inv)
    for k, (q, t) in enumerate(pipe.run(batches())):
        n = sizes[k]
        traj.extend(q[:n].numpy().astype(np.float64), t[:n].numpy().astype(np.float64))
    return traj
