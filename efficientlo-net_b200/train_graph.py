"""Training-mode PWCLO graph (is_training=True): batch-statistics batch norm, dropout, autograd.

What the reference does in training (config 4 of BASELINE.json): the same graph as inference with
``tf.contrib.layers.batch_norm`` normalising by the statistics of the current tensor (all B*N*K
positions of ONE call, utils/tf_util.py:527 -- the two frames of the siamese pyramid are separate
calls), moving averages updated in place (``updates_collections=None``), dropout 0.5 on the 256-wide
pose feature (pwclo_model.py:199,266,342,410), and gradients through every gather / scatter / softmax.

Split of the work here:
  * everything DISCRETE runs in the sm_100a kernels of the inference path and carries no gradient,
    exactly as in the reference where the custom ops register no gradient (fused_conv_*_k.py:31-32):
    PreProcess + spherical binning of the inputs (elo_project), the strided xyz pyramid, all
    projection-aware neighbour searches (elo_multi_search), and the binning of the warped points
    (elo_project reporting each point's cell and winner flag);
  * everything DIFFERENTIABLE is composed from torch tensor ops on the GPU so autograd provides the
    backward pass: gathers by the neighbour tables, the shared MLPs with batch-stat BN, masked
    soft-max pooling, the quaternion warp, the scatter of warped points / features into their cells
    (index_add == tf.scatter_nd's accumulate), the pose heads and the loss.

The fused inference kernels cannot be used for the dense part in this mode: batch-norm statistics
couple all rows of a layer, so a layer cannot be folded into its neighbours.
"""
import math

import torch

from . import model_util as mu
from . import pointnet_util as pu
from .params import make_perms

BN_EPS = 1e-3

# main.py:44-53,105-108
BASE_LEARNING_RATE = 0.001
DECAY_STEP = 200000
DECAY_RATE = 0.7
BN_INIT_DECAY = 0.5
BN_DECAY_DECAY_RATE = 0.5
BN_DECAY_CLIP = 0.99


def get_learning_rate(batch, batch_size, base=BASE_LEARNING_RATE, decay_step=DECAY_STEP, decay_rate=DECAY_RATE):
    """Staircase exponential decay clipped at 1e-5 (main.py:120-128)."""
    return max(base * decay_rate ** ((batch * batch_size) // decay_step), 0.00001)


def get_bn_decay(batch, batch_size, decay_step=DECAY_STEP):
    """main.py:130-138"""
    return min(BN_DECAY_CLIP, 1 - BN_INIT_DECAY * BN_DECAY_DECAY_RATE ** ((batch * batch_size) // decay_step))


class TrainableParams:
    """Flat parameter set under the reference's variable names, resident on one GPU.

    weights / biases / gamma / beta and the two loss weights w_x, w_q (main.py:151-152) are leaf
    tensors with requires_grad; moving_mean / moving_variance are plain tensors updated in place by
    the forward pass."""

    def __init__(self, flat, device="cuda", w_x=0.0, w_q=-2.5):
        self.device = torch.device(device)
        self.t = {}
        for name, v in flat.items():
            x = torch.as_tensor(v).detach().to(self.device, torch.float32).clone()
            if not name.endswith(("moving_mean", "moving_variance")):
                x.requires_grad_(True)
            self.t[name] = x
        for name, v in (("w_x", w_x), ("w_q", w_q)):
            if name not in self.t:
                self.t[name] = torch.tensor(float(v), device=self.device, requires_grad=True)

    def __getitem__(self, name):
        return self.t[name]

    def named_parameters(self):
        return [(n, x) for n, x in self.t.items() if x.requires_grad]

    def parameters(self):
        return [x for _, x in self.named_parameters()]

    def export(self):
        """Flat dict of CPU tensors in the layout ParamStore / the inference engine take."""
        return {n: x.detach().cpu().clone() for n, x in self.t.items()}


class _Linear(torch.autograd.Function):
    """y = x @ W + b on (rows, Cin) with a weight gradient that fills the GPU.  rows is ~10^5..10^6 and Cin, Cout
    <= 192, so dW = x^T @ dy is a tiny output with an enormous reduction dimension: cuBLAS runs it on a handful of
    CTAs (the largest single item of the eager step's GPU time).  Here the rows are cut into S slabs, one batched
    GEMM gives S partial (Cin, Cout) products and their sum is dW -- the same fp32 arithmetic, S-fold parallelism."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return torch.addmm(b, x, w)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dy @ w.t() if ctx.needs_input_grad[0] else None
        rows = x.shape[0]
        S = max(1, min(256, rows // 4096))
        r0 = (rows // S) * S
        dw = torch.bmm(x[:r0].view(S, r0 // S, -1).transpose(1, 2), dy[:r0].view(S, r0 // S, -1)).sum(0)
        if r0 < rows:
            dw = dw + x[r0:].t() @ dy[r0:]
        return dx, dw, dy.sum(0)


class _Net:
    def __init__(self, p, bn_decay, dropout, generator):
        self.p, self.dropout, self.generator = p, dropout, generator
        self.decay = 0.9 if bn_decay is None else float(bn_decay)          # tf_util.py:526

    def conv(self, x, scope, relu=True):
        """1x1 conv + bias + batch-stat BN + ReLU on the last axis (tf_util.py:120-185, 512-531).
        Three kernels forward, three backward: a GEMM with the bias folded in, ONE fused batch-norm kernel (batch
        mean / biased variance over all rows, normalise, scale and shift, moving averages updated in place with the
        Bessel-corrected variance -- exactly TF's FusedBatchNorm with updates_collections=None; torch's momentum is
        1 - decay), ReLU."""
        p = self.p
        w = p[scope + "/weights"]
        flat = _Linear.apply(x.reshape(-1, x.shape[-1]), w, p[scope + "/biases"])
        flat = torch.nn.functional.batch_norm(flat, p[scope + "/bn/moving_mean"], p[scope + "/bn/moving_variance"],
                                              p[scope + "/bn/gamma"], p[scope + "/bn/beta"], training=True,
                                              momentum=1.0 - self.decay, eps=BN_EPS)
        if relu:
            flat = torch.relu_(flat)
        return flat.view(*x.shape[:-1], w.shape[-1])

    def linear(self, x, scope):
        return x @ self.p[scope + "/weights"] + self.p[scope + "/biases"]

    def drop(self, x):
        if self.dropout <= 0:
            return x
        keep = (torch.rand(x.shape, device=x.device, generator=self.generator) >= self.dropout).to(x.dtype)
        return x * keep / (1 - self.dropout)


def _gather(grid, nbr):
    """grid (B, cells, C), nbr (B, n, K) int32 with -1 = masked -> (B, n, K, C) zeros where masked."""
    B, cells, C = grid.shape
    mask = nbr >= 0
    lin = nbr.clamp(min=0).long() + (torch.arange(B, device=grid.device) * cells).view(B, 1, 1)
    # index_select, not advanced indexing: its backward is index_add_ (atomic adds), where the backward of
    # x[idx] is index_put_(accumulate=True), which sorts the ~10^6 indices of every gather
    g = grid.reshape(B * cells, C).index_select(0, lin.reshape(-1)).view(*nbr.shape, C)
    return g * mask.unsqueeze(-1), mask.unsqueeze(-1)


def _flat(t):
    return t.reshape(t.shape[0], -1, t.shape[-1])


def set_conv(net, xyz_grid, feat_grid, centres, nbr, scopes):
    """pointnet_util.py:179-250 given the neighbour table.  xyz_grid (B,H,W,3), feat_grid (B,H,W,C) or
    None (zero features, 3 channels), centres (B,n,3)."""
    g_xyz, mask = _gather(_flat(xyz_grid), nbr)
    if feat_grid is None:
        g_feat = torch.zeros_like(g_xyz)
    else:
        g_feat, _ = _gather(_flat(feat_grid), nbr)
    x = torch.cat([g_xyz - centres.unsqueeze(2), g_feat], -1)
    for s in scopes:
        x = net.conv(x, s)
    return (x * mask).max(dim=2).values


def up_conv(net, xyz1_grid, xyz2_grid, feat1, feat2_grid, nbr, scope):
    """pointnet_util.py:254-316 given the neighbour table; feat1 (B, H*W, C1)."""
    g_xyz, mask = _gather(_flat(xyz2_grid), nbr)
    g_feat, _ = _gather(_flat(feat2_grid), nbr)
    x = torch.cat([g_xyz - _flat(xyz1_grid).unsqueeze(2), g_feat], -1)
    for s in ("up_1_0", "up_1_1"):
        x = net.conv(x, scope + "/" + s)
    x = torch.cat([(x * mask).max(dim=2).values, feat1], -1)
    for s in ("up_2_0", "up_2_1"):
        x = net.conv(x, scope + "/" + s)
    return x


def cost_volume(net, xyz1_grid, xyz2_grid, pts1, pts2_grid, nbr_q, nbr_p, scope):
    """pointnet_util.py:33-149 given both neighbour tables; pts1 (B, H*W, C)."""
    xyz1 = _flat(xyz1_grid)
    kq, kp = nbr_q.shape[-1], nbr_p.shape[-1]
    qi_xyz, mask = _gather(_flat(xyz2_grid), nbr_q)
    qi_pts, _ = _gather(_flat(pts2_grid), nbr_q)
    pi_xyz = xyz1.unsqueeze(2).expand(-1, -1, kq, -1)
    pi_pts = pts1.unsqueeze(2).expand(-1, -1, kq, -1)
    diff = qi_xyz - pi_xyz
    euc = torch.sqrt((diff * diff).sum(-1, keepdim=True) + 1e-20)
    xyz10 = torch.cat([pi_xyz, qi_xyz, diff, euc], -1)
    feat = torch.cat([xyz10, pi_pts, qi_pts], -1)
    for s in ("CV_0", "CV_1", "CV_2"):
        feat = net.conv(feat, scope + "/" + s)
    w = torch.cat([net.conv(xyz10, scope + "/CV_xyz"), feat], -1)
    for s in ("sum_CV_0", "sum_CV_1"):
        w = net.conv(w, scope + "/" + s)
    w = torch.softmax(torch.where(mask, w, torch.full_like(w, -1e10)), dim=2)
    stage1 = (w * feat).sum(2)
    pc_pts, mask2 = _gather(stage1, nbr_p)
    pc_xyz, _ = _gather(xyz1, nbr_p)
    p_xyz = xyz1.unsqueeze(2).expand(-1, -1, kp, -1)
    d2 = pc_xyz - p_xyz
    e2 = torch.sqrt((d2 * d2).sum(-1, keepdim=True) + 1e-20)
    enc = net.conv(torch.cat([p_xyz, pc_xyz, d2, e2], -1), scope + "/sum_xyz_encoding")
    w2 = torch.cat([enc, pts1.unsqueeze(2).expand(-1, -1, kp, -1), pc_pts], -1)
    for s in ("sum_cost_volume_0", "sum_cost_volume_1"):
        w2 = net.conv(w2, scope + "/" + s)
    w2 = torch.softmax(torch.where(mask2, w2, torch.full_like(w2, -1e10)), dim=2)
    return (w2 * pc_pts).sum(2)


def flow_predictor(net, parts, scope):
    x = torch.cat([t for t in parts if t is not None], -1)
    for s in ("conv_predictor0", "conv_predictor1"):
        x = net.conv(x, scope + "/" + s)
    return x


def softmax_valid(feature, weight, valid):
    """model_util.py:319-343 without the boolean_mask: invalid rows get weight 0."""
    w = torch.where(valid.unsqueeze(-1), weight, torch.full_like(weight, -float("inf")))
    m = torch.where(valid.any(1).view(-1, 1, 1), w.max(dim=1, keepdim=True).values, torch.zeros_like(w[:, :1]))
    e = torch.exp(w - m)
    return (feature * e).sum(1, keepdim=True) / e.sum(1, keepdim=True).clamp_min(1e-38)


def normalize_q(q):
    return q / (torch.sqrt((q * q).sum(-1, keepdim=True) + 1e-10) + 1e-10)


def warp(xyz, q, t):
    """pwclo_model.py:213-227: rotate by q, translate by t, zero points stay zero."""
    B = xyz.shape[0]
    valid = (~(xyz == 0).all(-1, keepdim=True)).to(xyz.dtype)
    pq = torch.cat([torch.zeros_like(xyz[..., :1]), xyz], -1)
    r = mu.mul_point_q(mu.mul_q_point(q.reshape(B, 1, 4), pq, B), mu.inv_q(q.reshape(B, 1, 4), B), B)
    return (r[..., 1:] + t.reshape(B, 1, 3)) * valid


def reproject(warped, feats, h, w):
    """ProjectPC2SphericalRing (model_util.py:181-292) of the warped points with gradients: the kernel
    decides cell and winner of every point, index_add accumulates winners like tf.scatter_nd."""
    B, n, _ = warped.shape
    _, _, _, cell = mu.project_points(warped.detach(), None, h, w, mode=0, want_cells=True)
    win = (cell >= 0)
    lin = (cell.clamp(min=0).long() + (torch.arange(B, device=warped.device) * (h * w)).view(B, 1)).reshape(-1)
    wf = win.reshape(-1, 1).to(warped.dtype)
    xyz = torch.zeros((B * h * w, 3), device=warped.device).index_add(0, lin, warped.reshape(-1, 3) * wf)
    C = feats.shape[-1]
    ft = torch.zeros((B * h * w, C), device=warped.device).index_add(0, lin, feats.reshape(-1, C) * wf)
    return xyz.view(B, h, w, 3), ft.view(B, h, w, C)


def pose_head(net, feat, lvl):
    suffix = "coarse" if lvl == 3 else "det"
    big = net.drop(net.linear(feat, "l%d_big" % lvl))
    return normalize_q(net.linear(big, "l%d_q_%s" % (lvl, suffix))), net.linear(big, "l%d_t_%s" % (lvl, suffix))


def get_model(point_cloud, H_input, W_input, T_gt, T_trans, T_trans_inv, params, bn_decay=None, perms=None,
              aug_frame=None, dropout=0.5, generator=None, keep=None):
    """Training-mode forward (pwclo_model.py:30-433 with is_training=True).  ``params`` is a
    TrainableParams; returns the reference's 11-tuple with autograd history."""
    from . import pwclo_model as pm
    if not isinstance(params, TrainableParams):
        raise TypeError("training needs a TrainableParams (leaf tensors with gradients), got %r" % type(params))
    net = _Net(params, bn_decay, dropout, generator)
    if perms is None:
        perms = make_perms(int(torch.randint(0, 2 ** 31 - 1, (1,))))
    B, N = point_cloud.shape[0], point_cloud.shape[1] // 2
    dev = point_cloud.device
    if point_cloud.dtype != torch.float32 or not point_cloud.is_contiguous():
        point_cloud = point_cloud.float().contiguous()
    oh, ow = pm.pyramid_shapes(H_input, W_input)
    K = keep if keep is not None else {}
    eye = torch.eye(4, device=dev).expand(B, 4, 4).contiguous()
    T_gt = eye if T_gt is None else T_gt.to(dev)
    T_trans = eye if T_trans is None else T_trans.to(dev)
    T_trans_inv = eye if T_trans_inv is None else T_trans_inv.to(dev)

    with torch.no_grad():
        q_gt, t_gt = mu.gt_pose(T_gt, T_trans, T_trans_inv, aug_frame)
        T_aug, T_apply = mu.aug_setup(T_trans, aug_frame, B, dev)
        xyz_in, _, _ = mu.project_points(point_cloud[:, :N, 0:3], None, H_input, W_input, mode=1, T=T_aug, T_apply=T_apply,
                                         inner_batch=B, outer_stride=N * point_cloud.stride(1), batch_size=2 * B)
        csh, csw, ch, cw = [], [], pm.STRIDE_H[1], pm.STRIDE_W[1]
        for l in range(4):
            ch, cw = ch * pm.STRIDE_H[l + 2], cw * pm.STRIDE_W[l + 2]
            csh.append(ch)
            csw.append(cw)
        xyz = mu.xyz_pyramid(xyz_in, oh[2:], ow[2:], csh, csw)
        grids = [xyz_in] + xyz[:3]
        specs = []
        for l in range(4):
            K_l, ks = pm.DOWN_CFG[l]
            qsh = pm.STRIDE_H[l + 2] * (pm.STRIDE_H[1] if l == 0 else 1)
            qsw = pm.STRIDE_W[l + 2] * (pm.STRIDE_W[1] if l == 0 else 1)
            for f, half in (("f1", slice(0, B)), ("f2", slice(B, 2 * B))):
                g_ = grids[l][half]
                specs.append(pu.search_spec(False, g_, g_, (oh[l + 2], ow[l + 2], qsh, qsw), ks, K_l,
                                            pm.DOWN_CONV_DIS[l], 1, 1, perms["sa1/layer%d/%s" % (l, f)]))
        x2f1, x2f2 = xyz[2][:B], xyz[2][B:]
        all2 = (oh[4], ow[4], 1, 1)
        specs.append(pu.search_spec(True, x2f1, x2f2, all2, (5, 35), 32, 1000.0, 1, 1, perms["flow_embedding_l2_origin/q"]))
        specs.append(pu.search_spec(False, x2f1, x2f1, all2, (3, 5), 4, pm.COST_VOLUME_DIS[2], 1, 1,
                                    perms["flow_embedding_l2_origin/p"]))
        specs.append(pu.search_spec(False, x2f1, x2f1, (oh[5], ow[5], pm.STRIDE_H[5], pm.STRIDE_W[5]), (5, 9), 16,
                                    pm.DOWN_CONV_DIS[3], 1, 1, perms["new_layer3"]))
        tables = pu.multi_search(specs)

    # siamese feature pyramid (:117-165): one call -- and one set of batch statistics -- per frame
    pts = {0: [], 1: []}
    for f, half in ((0, slice(0, B)), (1, slice(B, 2 * B))):
        src_xyz, src_pts = xyz_in[half], None
        for l in range(4):
            scopes = ["sa1/layer%d/conv%d" % (l, j) for j in range(3)]
            feat = set_conv(net, src_xyz, src_pts, _flat(xyz[l][half]), tables[2 * l + f], scopes)
            pts[f].append(feat)
            src_xyz, src_pts = xyz[l][half], feat.view(B, oh[l + 2], ow[l + 2], -1)
            K["l%d_points_f%d" % (l, f + 1)] = feat

    def f1(t):
        return t[:B]

    def f2(t):
        return t[B:]

    def grid(l, t):
        return t.reshape(B, oh[l + 2], ow[l + 2], -1)

    l2_new = cost_volume(net, f1(xyz[2]), f2(xyz[2]), pts[0][2], grid(2, pts[1][2]), tables[8], tables[9],
                         "flow_embedding_l2_origin")
    l3_cv = set_conv(net, f1(xyz[2]), grid(2, l2_new), _flat(f1(xyz[3])), tables[10],
                     ["new_layer3/conv%d" % j for j in range(3)])
    l3_w = flow_predictor(net, [pts[0][3], None, l3_cv], "l3_costvolume_predict_ww")
    l3_valid = ~(_flat(f1(xyz[3])) == 0).all(-1)
    q, t = pose_head(net, softmax_valid(l3_cv, l3_w, l3_valid), 3)
    q, t = q.squeeze(1), t.squeeze(1)
    qs, ts = {3: q}, {3: t}
    K.update(l2_points_f1_new=l2_new, l3_points_f1_cost_volume=l3_cv, l3_w=l3_w, l3_q=q, l3_t=t)

    up_xyz, up_w, up_pred = f1(xyz[3]), grid(3, l3_w), grid(3, l3_cv)
    for lvl in (2, 1, 0):
        h, w_ = oh[lvl + 2], ow[lvl + 2]
        warped = warp(_flat(f1(xyz[lvl])), q, t)
        xyz_wp, pts_wp = reproject(warped, pts[0][lvl], h, w_)
        xyz_wpd = xyz_wp.detach()
        names = ["up_sa_layer_layer_l%dw" % lvl, "up_sa_layer_layer_l%dcostvolume" % lvl]
        allq = (h, w_, 1, 1)
        s_h, s_w = pm.STRIDE_H[lvl + 3], pm.STRIDE_W[lvl + 3]
        with torch.no_grad():
            up_xyzd = up_xyz.detach()
            nq_, np_, nu0, nu1 = pu.multi_search([
                pu.search_spec(True, xyz_wpd, f2(xyz[lvl]), allq, pm.CV_KERNEL_Q[lvl], 6, 1000.0, 1, 1,
                               perms["flow_embedding_l%d/q" % lvl]),
                pu.search_spec(False, xyz_wpd, xyz_wpd, allq, (3, 5), 4, pm.COST_VOLUME_DIS[lvl], 1, 1,
                               perms["flow_embedding_l%d/p" % lvl]),
                pu.search_spec(False, xyz_wpd, up_xyzd, allq, (7, 15), 8, pm.UP_CONV_DIS[lvl], s_h, s_w, perms[names[0]]),
                pu.search_spec(False, xyz_wpd, up_xyzd, allq, (7, 15), 8, pm.UP_CONV_DIS[lvl], s_h, s_w, perms[names[1]])])
        pts_w = _flat(pts_wp)
        cv = cost_volume(net, xyz_wp, f2(xyz[lvl]), pts_w, grid(lvl, pts[1][lvl]), nq_, np_, "flow_embedding_l%d" % lvl)
        w_up = up_conv(net, xyz_wp, up_xyz, pts_w, up_w, nu0, names[0])
        p_up = up_conv(net, xyz_wp, up_xyz, pts_w, up_pred, nu1, names[1])
        pred = flow_predictor(net, [pts_w, p_up, cv], "l%d_costvolume_predict" % lvl)
        wgt = flow_predictor(net, [pts_w, w_up, cv], "l%d_w_predict" % lvl)
        valid = ~(_flat(xyz_wp) == 0).all(-1)
        q_det, t_det = pose_head(net, softmax_valid(pred, wgt, valid), lvl)
        q_c, t_c = q.reshape(B, 1, 4), t.reshape(B, 1, 3)
        tq = torch.cat([torch.zeros_like(t_c[..., :1]), t_c], -1)
        tq = mu.mul_point_q(mu.mul_q_point(q_det, tq, B), mu.inv_q(q_det, B), B)[..., 1:]
        q = mu.mul_point_q(q_det, q_c, B).squeeze(1)
        t = (tq + t_det).squeeze(1)
        qs[lvl], ts[lvl] = q, t
        K.update({"l%d_flow_warp" % lvl: warped, "l%d_xyz_warp_proj" % lvl: xyz_wp, "l%d_points_warp_proj" % lvl: pts_wp,
                  "l%d_cost_volume" % lvl: cv, "l%d_w_up" % lvl: w_up, "l%d_p_up" % lvl: p_up, "l%d_predict" % lvl: pred,
                  "l%d_w" % lvl: wgt, "l%d_q" % lvl: q, "l%d_t" % lvl: t})
        up_xyz, up_w, up_pred = xyz_wp, wgt.view(B, h, w_, -1), pred.view(B, h, w_, -1)

    return (normalize_q(qs[0]), ts[0], normalize_q(qs[1]), ts[1], normalize_q(qs[2]), ts[2], normalize_q(qs[3]), ts[3],
            _flat(f1(xyz[0])), q_gt, t_gt)


class Trainer:
    """One optimisation step of main.py:344-397: forward, loss, Adam with the staircase learning-rate
    decay, batch-norm decay schedule, global step.  With ``process_group`` set, the gradients of all
    ranks are averaged with one flat all-reduce per step (data parallel over frame pairs)."""

    def __init__(self, params, batch_size, H_input=64, W_input=1800, optimizer="adam", momentum=0.9,
                 base_lr=BASE_LEARNING_RATE, decay_step=DECAY_STEP, decay_rate=DECAY_RATE, dropout=0.5,
                 process_group=None, seed=0, use_graph=False):
        self.p, self.batch_size, self.H, self.W = params, batch_size, H_input, W_input
        self.base_lr, self.decay_step, self.decay_rate, self.dropout = base_lr, decay_step, decay_rate, dropout
        if optimizer == "adam":
            # tf.train.AdamOptimizer defaults: beta1 0.9, beta2 0.999, epsilon 1e-8
            self.opt = torch.optim.Adam(params.parameters(), lr=base_lr, betas=(0.9, 0.999), eps=1e-8)
        else:
            self.opt = torch.optim.SGD(params.parameters(), lr=base_lr, momentum=momentum)
        self.batch = 0
        self.group = process_group
        self.gen = torch.Generator(device=params.device).manual_seed(seed)
        # whole-step CUDA graph (use_graph=True): forward, loss, backward, gradient all-reduce and the optimizer update
        # are ~6000 small launches whose issue time, not their run time, bounds the eager step
        self.use_graph = bool(use_graph)
        self._graph = None
        if self.use_graph:
            if optimizer != "adam":
                raise ValueError("use_graph needs the capturable Adam")
            self.opt = torch.optim.Adam(params.parameters(), lr=torch.tensor(base_lr, device=params.device),
                                        betas=(0.9, 0.999), eps=1e-8, capturable=True)

    # ---- captured step ------------------------------------------------------------------------------------------
    def _capture(self, point_cloud, T_gt, T_trans, T_trans_inv, perms, bn_decay):
        from .pwclo_model import get_loss
        dev = self.p.device
        st = self._static = dict(pc=point_cloud.detach().to(dev, torch.float32).clone(), T=T_gt.detach().to(dev).float().clone())
        eye = torch.eye(4, device=dev).expand(point_cloud.shape[0], 4, 4).contiguous()
        st["Tt"] = (eye if T_trans is None else T_trans.to(dev).float()).clone()
        st["Ti"] = (eye if T_trans_inv is None else T_trans_inv.to(dev).float()).clone()
        st["perms"] = {k: torch.as_tensor(v).to(dev, torch.int32).clone() for k, v in perms.items()}
        self._graph_decay = bn_decay

        def body():
            out = get_model(st["pc"], self.H, self.W, st["T"], st["Tt"], st["Ti"], self.p, bn_decay=bn_decay,
                            perms=st["perms"], aug_frame=None, dropout=self.dropout, generator=self.gen)
            loss = get_loss(*out[:8], out[9], out[10], self.p["w_x"], self.p["w_q"])
            loss.backward()
            if self.group is not None:
                all_reduce_gradients(self.p.parameters(), self.group)
            self.opt.step()
            return loss.detach()

        # Warm-up on a side stream (allocator pools, lazily created Adam slots, NCCL communicator) must not count as
        # training: parameters, moving averages and optimizer slots are put back afterwards, IN PLACE -- the captured
        # graph refers to these very tensors.
        snap = {n: t.detach().clone() for n, t in self.p.t.items()}
        slots = {id(q): {k: v.detach().clone() for k, v in self.opt.state[q].items() if torch.is_tensor(v)}
                 for q in self.p.parameters() if q in self.opt.state}
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                self.opt.zero_grad(set_to_none=True)
                body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        with torch.no_grad():
            for n, t in self.p.t.items():
                t.copy_(snap[n])
            for q in self.p.parameters():
                for k, v in self.opt.state[q].items():
                    if torch.is_tensor(v):
                        old = slots.get(id(q), {}).get(k)
                        v.copy_(old) if old is not None else v.zero_()
        self.opt.zero_grad(set_to_none=True)
        g = torch.cuda.CUDAGraph()
        g.register_generator_state(self.gen)
        with torch.cuda.graph(g):
            self._static_loss = body()
        self._graph = g

    def _graph_step(self, point_cloud, T_gt, T_trans, T_trans_inv, perms):
        lr = get_learning_rate(self.batch, self.batch_size, self.base_lr, self.decay_step, self.decay_rate)
        bn_decay = get_bn_decay(self.batch, self.batch_size, self.decay_step)
        if perms is None:
            perms = make_perms(self.batch)
        if self._graph is None or bn_decay != self._graph_decay:        # the decay is baked into the captured BN kernels
            self._graph = None
            for g in self.opt.param_groups:
                g["lr"].fill_(lr)
            self._capture(point_cloud, T_gt, T_trans, T_trans_inv, perms, bn_decay)
        st = self._static
        st["pc"].copy_(point_cloud, non_blocking=True)
        st["T"].copy_(T_gt, non_blocking=True)
        if T_trans is not None:
            st["Tt"].copy_(T_trans, non_blocking=True)
            st["Ti"].copy_(T_trans_inv, non_blocking=True)
        for k, v in perms.items():
            if v is not st["perms"][k]:
                st["perms"][k].copy_(torch.as_tensor(v), non_blocking=True)
        for g in self.opt.param_groups:
            g["lr"].fill_(lr)
        self._graph.replay()
        self.batch += 1
        return self._static_loss

    def step(self, point_cloud, T_gt, T_trans=None, T_trans_inv=None, perms=None, aug_frame=None):
        from .pwclo_model import get_loss
        if self.use_graph and aug_frame is None:
            return self._graph_step(point_cloud, T_gt, T_trans, T_trans_inv, perms)
        lr = get_learning_rate(self.batch, self.batch_size, self.base_lr, self.decay_step, self.decay_rate)
        for g in self.opt.param_groups:
            if torch.is_tensor(g["lr"]):
                g["lr"].fill_(lr)
            else:
                g["lr"] = lr
        out = get_model(point_cloud, self.H, self.W, T_gt, T_trans, T_trans_inv, self.p,
                        bn_decay=get_bn_decay(self.batch, self.batch_size, self.decay_step), perms=perms,
                        aug_frame=aug_frame, dropout=self.dropout, generator=self.gen)
        loss = get_loss(*out[:8], out[9], out[10], self.p["w_x"], self.p["w_q"])
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        if self.group is not None:
            all_reduce_gradients(self.p.parameters(), self.group)
        self.opt.step()
        self.batch += 1
        return loss.detach()


def all_reduce_gradients(params, group=None):
    """Average the gradients over the ranks with ONE flat all-reduce (NCCL over NVLink on the GPU box,
    gloo in the CPU tests)."""
    import torch.distributed as dist
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for p, g in zip(params, grads):
        n = g.numel()
        p.grad = flat[off:off + n].view_as(g).clone()
        off += n
