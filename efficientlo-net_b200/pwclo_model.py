"""PWCLO-Net forward on B200 -- mirror of the reference's pwclo_model.py (placeholder_inputs :19,
get_model :30-433, get_loss :437-481).

get_model takes the reference's arguments and returns its 11-tuple.  The wiring follows the reference
block for block (the comments cite its lines); what differs is the execution: ~40 fused sm_100a kernel
launches per forward instead of several thousand TensorFlow ops, both frames of the siamese pyramid and
both up-convs / predictors of a level sharing launches, the pose warp fused into the re-projection, and
the pose composition into the attention-pooling kernel.  Nothing here synchronises with the host, so a
whole forward can be captured in a CUDA graph (engine.py).
"""
import math
import os
import threading

import torch

from . import _lib
from . import model_util as mu
from . import pointnet_util as pu
from .params import make_perms
from .store import ParamStore, current_store, use_store

# hyper-parameters, literals of pwclo_model.py:38-43 and of the call sites
DOWN_CONV_DIS = [0.5, 3.0, 6.0, 12.0]
UP_CONV_DIS = [3.0, 6.0, 9.0]
COST_VOLUME_DIS = [1.0, 2.0, 4.0]
STRIDE_H = [1, 1, 4, 2, 2, 1]
STRIDE_W = [1, 1, 8, 2, 2, 2]
DOWN_CFG = [(32, (9, 15)), (32, (7, 11)), (16, (5, 9)), (16, (5, 9))]     # (K_sample, kernel_size) layer0..3
CV_KERNEL_Q = {2: (5, 15), 1: (7, 25), 0: (11, 41)}


def placeholder_inputs(batch_size, NUM_POINTS, device="cuda"):
    """Zero tensors with the shapes of the reference's placeholders (pwclo_model.py:19-27)."""
    eye = torch.eye(4, device=device).expand(batch_size, 4, 4).contiguous()
    return (torch.zeros(batch_size, NUM_POINTS * 2, 6, device=device), eye.clone(), eye.clone(), eye.clone())


def pyramid_shapes(H_input, W_input):
    """out_h_list / out_w_list of pwclo_model.py:45-50."""
    oh, ow = [math.ceil(H_input / STRIDE_H[0])], [math.ceil(W_input / STRIDE_W[0])]
    for i in range(1, 6):
        oh.append(math.ceil(oh[-1] / STRIDE_H[i]))
        ow.append(math.ceil(ow[-1] / STRIDE_W[i]))
    return oh, ow


# Branches of a level on two streams (fork / join): 1 = always, 0 = never, unset = under the latency tile policy only
# (with several forwards in flight the SMs are busy anyway and the extra graph edges cost more than they give).
_FORK = os.environ.get("ELO_FORK", "")
_SIDE = {}


def _fork_branches(dev):
    """(current stream, side stream) when the level's independent chains should run side by side, else None."""
    if dev.type != "cuda" or _FORK == "0":
        return None
    if _FORK != "1" and _lib.get_tile_policy() != 0:
        return None
    key = (dev.index, threading.get_ident())
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return torch.cuda.current_stream(dev), _SIDE[key]


def _heads(store, lvl):
    suffix = "coarse" if lvl == 3 else "det"
    names = ["l%d_big" % lvl, "l%d_q_%s" % (lvl, suffix), "l%d_t_%s" % (lvl, suffix)]
    out = []
    for n in names:
        out += [store.tensor(n + "/weights"), store.tensor(n + "/biases")]
    return tuple(out)


class RowBand:
    """Row-band partition of ONE frame pair over the ranks of a process group (BASELINE.json north_star, SURVEY.md
    section 8(e)).  Rank r owns image rows [r0, r1) of every banded level.  The heavy blocks of a banded level --
    the level's four neighbour searches, both stages of the cost volume, both set-upconvs and the predictors, and
    layer 0 of the feature pyramid -- run only for the queries of the owned rows (`query_begin / query_end` of the
    C ABI); what they read (the level's grids, the coarser level's embeddings) is held in full by every rank, so
    the window of a query never needs rows that are not there and no halo has to be fetched.  ONE all-gather per
    banded block chain then makes the level's result complete on every rank again: per level the rows of
    (embedding, mask logits), once for layer 0 the rows of both frames' features.  Small levels are computed
    by every rank (they are a handful of single-wave kernels; exchanging them would cost more than computing them).
    """

    def __init__(self, rank, world, group=None, min_rows_per_rank=2, min_points=4096, skip=()):
        self.rank, self.world, self.group, self.min_rows = int(rank), int(world), group, int(min_rows_per_rank)
        # An exchange costs ~25 us (measured, NCCL all-gather inside the captured graph, 2 GPUs); a level with fewer
        # points than this is a few single-wave kernels that banding does not shorten by that much: at 128x2048 layer 0
        # and level 0 (8192 points each) are banded, at 64x1800 (3600 points) nothing is.
        self.min_points = int(min_points)
        self.skip = set(skip)         # tags ("layer0", "l0", "l1", "l2") that are NOT banded
        self.exchanges = 0

    def rows(self, h, tag=None, w=None):
        """(r0, r1) of an h-row (w-column) level for this rank, or None when the level is not banded: rows do not
        divide evenly over the ranks, bands would be thinner than min_rows_per_rank, the level has fewer than
        min_points points, or its tag is in `skip`."""
        if self.world <= 1 or h % self.world or h // self.world < self.min_rows or tag in self.skip:
            return None
        if w is not None and h * w < self.min_points:
            return None
        per = h // self.world
        return self.rank * per, (self.rank + 1) * per

    def gather_rows(self, tensors, h, w):
        """tensors: list of (S, h*w, C) tensors whose rows [r0*w, r1*w) this rank has just computed.  After the call
        every rank holds all rows of all of them: one all-gather (NCCL over NVLink) of the rank's slabs."""
        import torch.distributed as dist
        per = h // self.world
        r0, r1 = self.rank * per, (self.rank + 1) * per
        n = per * w
        send = torch.cat([t[:, r0 * w:r1 * w].reshape(-1) for t in tensors])
        recv = torch.empty((self.world, send.numel()), dtype=send.dtype, device=send.device)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        off = 0
        for t in tensors:
            S, _, C = t.shape
            part = recv[:, off:off + S * n * C].view(self.world, S, n, C)
            t.view(S, self.world, n, C).copy_(part.permute(1, 0, 2, 3))
            off += S * n * C
        self.exchanges += 1


def get_model(point_cloud, H_input, W_input, T_gt, T_trans, T_trans_inv, is_training, bn_decay=None,
              params=None, perms=None, aug_frame=None, keep=None, band=None):
    """Whole network (pwclo_model.py:30-433).

    point_cloud (B, 2*N, 6) fp32 on the GPU, frame 1 in rows [0,N), frame 2 in [N,2N), xyz in channels
    0:3; T_gt / T_trans / T_trans_inv (B,4,4).  ``params``: a ParamStore or a flat parameter dict
    (params.py); ``perms``: scan orders per call site (params.make_perms) -- the reference redraws them
    every run; ``aug_frame``: which frame each sample augments (the reference draws it with numpy at
    graph-build time, :59; default 2).  ``keep``: optional dict that receives named intermediates.
    ``band``: a RowBand -- this rank computes only its row band of the banded levels (B must be 1) and the ranks
    exchange the bands; every rank returns the same poses as a single-GPU call.
    Returns (l0_q, l0_t, l1_q, l1_t, l2_q, l2_t, l3_q, l3_t, l0_xyz_f1, q_gt, t_gt)."""
    if pu._is_training(is_training):
        # batch-statistics batch norm + autograd: the differentiable composition in train_graph.py
        from . import train_graph
        return train_graph.get_model(point_cloud, H_input, W_input, T_gt, T_trans, T_trans_inv, params,
                                     bn_decay=bn_decay, perms=perms, aug_frame=aug_frame, keep=keep)
    store = params if isinstance(params, ParamStore) else (ParamStore(params, point_cloud.device) if params is not None
                                                           else current_store())
    if perms is None:
        perms = make_perms(int(torch.randint(0, 2 ** 31 - 1, (1,))))
    B = point_cloud.shape[0]
    N = point_cloud.shape[1] // 2
    dev = point_cloud.device
    if point_cloud.dtype != torch.float32 or not point_cloud.is_contiguous():
        point_cloud = point_cloud.float().contiguous()
    oh, ow = pyramid_shapes(H_input, W_input)
    K = keep if keep is not None else {}
    want = keep is not None
    if band is not None and (B != 1 or want):
        raise ValueError("row bands shard ONE frame pair (B = 1) and keep no intermediates")

    def band_rows(h, tag, w=None):
        return band.rows(h, tag, w) if band is not None else None

    with use_store(store):
        # ---- PreProcess (:61) + ProjectPC2SphericalRing x2 (:63-64), both frames in one pass.
        # Samples are stacked frame-major: s = f*B + b reads point_cloud[b, f*N:(f+1)*N, 0:3] in place.
        eye = store.eye(B)
        T_gt = T_gt.to(dev) if T_gt is not None else eye
        # Independent chains run side by side on a second stream under the latency tile policy (_fork_branches): the
        # ground-truth pose (needed only as an output), pyramid layer 3 beside the initial cost volume, and in every
        # refinement level the set-upconvs beside the cost volume.
        fork = _fork_branches(dev)
        gt_args = (T_gt, eye if T_trans is None else T_trans.to(dev), eye if T_trans_inv is None else T_trans_inv.to(dev),
                   aug_frame)
        gt_done = None
        if fork is not None:
            main, side = fork
            started = torch.cuda.Event()
            started.record(main)
            with torch.cuda.stream(side):
                side.wait_event(started)
                q_gt, t_gt = mu.gt_pose(*gt_args)
                gt_done = torch.cuda.Event()
                gt_done.record(side)
        else:
            q_gt, t_gt = mu.gt_pose(*gt_args)
        if T_trans is None and aug_frame is None:
            T_aug, T_apply = store.default_aug(B)               # constants: no per-forward tensor ops
        else:
            T_aug, T_apply = mu.aug_setup(eye if T_trans is None else T_trans, aug_frame, B, dev)
        stride_pt = point_cloud.stride(1)
        xyz_in, _, _ = mu.project_points(point_cloud[:, :N, 0:3], None, H_input, W_input, mode=1, T=T_aug,
                                         T_apply=T_apply, inner_batch=B, outer_stride=N * stride_pt, batch_size=2 * B)
        if want:
            K["xyz_f1_proj"], K["xyz_f2_proj"] = xyz_in[:B], xyz_in[B:]

        # ---- strided xyz pyramid (:88-114): pure slicing of the range image
        csh, csw, ch, cw = [], [], STRIDE_H[1], STRIDE_W[1]
        for l in range(4):
            ch, cw = ch * STRIDE_H[l + 2], cw * STRIDE_W[l + 2]
            csh.append(ch)
            csw.append(cw)
        xyz = mu.xyz_pyramid(xyz_in, oh[2:], ow[2:], csh, csw)                  # 4 x (2B, oh, ow, 3)

        # ---- every neighbour search that only needs the un-warped xyz pyramid, in ONE launch: the 8
        # set-convs of the siamese pyramid, the initial cost volume's two, and new_layer3's.
        sels = [pu.SelectedIdx(B, STRIDE_H[l + 2] * (STRIDE_H[1] if l == 0 else 1),
                               STRIDE_W[l + 2] * (STRIDE_W[1] if l == 0 else 1), oh[l + 2], ow[l + 2], dev)
                for l in range(4)]
        sel3 = pu.SelectedIdx(B, STRIDE_H[5], STRIDE_W[5], oh[5], ow[5], dev)
        grids = [xyz_in] + xyz[:3]                                              # grid searched by layer l
        specs, nbr_pyr = [], []
        for l in range(4):
            K_l, ks = DOWN_CFG[l]
            table = torch.empty((2 * B, oh[l + 2] * ow[l + 2], K_l), dtype=torch.int32, device=dev)
            nbr_pyr.append(table)
            for f, half in (("f1", slice(0, B)), ("f2", slice(B, 2 * B))):
                g_ = grids[l][half]
                rb = band_rows(oh[l + 2], "layer0", ow[l + 2]) if l == 0 else None
                specs.append(pu.search_spec(False, g_, g_, (sels[l].out_h, sels[l].out_w, sels[l].stride_h,
                                                            sels[l].stride_w), ks, K_l, DOWN_CONV_DIS[l], 1, 1,
                                            perms["sa1/layer%d/%s" % (l, f)], out=table[half],
                                            qrange=None if rb is None else (rb[0] * ow[l + 2], rb[1] * ow[l + 2])))
        x2f1, x2f2 = xyz[2][:B], xyz[2][B:]
        all2 = (oh[4], ow[4], 1, 1)
        specs.append(pu.search_spec(True, x2f1, x2f2, all2, (5, 35), 32, 1000.0, 1, 1, perms["flow_embedding_l2_origin/q"]))
        specs.append(pu.search_spec(False, x2f1, x2f1, all2, (3, 5), 4, COST_VOLUME_DIS[2], 1, 1,
                                    perms["flow_embedding_l2_origin/p"]))
        specs.append(pu.search_spec(False, x2f1, x2f1, (sel3.out_h, sel3.out_w, sel3.stride_h, sel3.stride_w), (5, 9),
                                    16, DOWN_CONV_DIS[3], 1, 1, perms["new_layer3"]))
        tables = pu.multi_search(specs)
        nbr_l2o_q, nbr_l2o_p, nbr_new3 = tables[8], tables[9], tables[10]

        # ---- siamese feature pyramid (:117-165): frames share weights, each call its own scan order
        pts = [None] * 4
        src_xyz, src_pts, src_c = xyz_in, None, 3
        for l in range(4):
            _lib.PROFILE_TAG[0] = "sa%d" % l
            sel = sels[l]
            scopes = ["sa1/layer%d/conv%d" % (l, j) for j in range(3)]
            rb = band_rows(oh[l + 2], "layer0", ow[l + 2]) if l == 0 else None        # layer 0 (the 64x1800 / 128x2048 image) is banded
            qr = None if rb is None else (rb[0] * ow[l + 2], rb[1] * ow[l + 2])
            sa3_done = None
            if l == 3 and fork is not None:         # layer 3 beside the initial cost volume, joined before level 3's mask
                fed = torch.cuda.Event()
                fed.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(fed)
                    feat = pu.set_conv(src_xyz, src_pts, sel, DOWN_CFG[l][0], DOWN_CFG[l][1], DOWN_CONV_DIS[l], scopes,
                                       store, [perms["sa1/layer%d/f1" % l], perms["sa1/layer%d/f2" % l]],
                                       feat_channels=src_c, set_batch_offsets=(0, B), nbr=nbr_pyr[l], qrange=qr)
                    sa3_done = torch.cuda.Event()
                    sa3_done.record(side)
            else:
                feat = pu.set_conv(src_xyz, src_pts, sel, DOWN_CFG[l][0], DOWN_CFG[l][1], DOWN_CONV_DIS[l], scopes, store,
                                   [perms["sa1/layer%d/f1" % l], perms["sa1/layer%d/f2" % l]], feat_channels=src_c,
                                   set_batch_offsets=(0, B), nbr=nbr_pyr[l], qrange=qr)
            if rb is not None:
                band.gather_rows([feat], oh[l + 2], ow[l + 2])
            pts[l] = feat                                                       # (2B, n_l, C_l)
            src_xyz, src_pts, src_c = xyz[l], feat.view(2 * B, oh[l + 2], ow[l + 2], -1), feat.shape[-1]
            if want:
                K["l%d_points_f1" % l], K["l%d_points_f2" % l] = feat[:B], feat[B:]

        def f1(t):
            return t[:B]

        def f2(t):
            return t[B:]

        def grid(l, t):
            return t.reshape(t.shape[0], oh[l + 2], ow[l + 2], -1)

        # ---- initial cost volume on level 2 and its set-conv to level 3 (:170-178)
        _lib.PROFILE_TAG[0] = "l2o"
        l2_new = pu.cost_volume(f1(xyz[2]), f2(xyz[2]), grid(2, f1(pts[2])), grid(2, f2(pts[2])), [3, 5], [5, 35],
                                4, 32, COST_VOLUME_DIS[2], [128, 64, 64], [128, 64], False, bn_decay,
                                "flow_embedding_l2_origin", random_hw_q=perms["flow_embedding_l2_origin/q"],
                                random_hw_p=perms["flow_embedding_l2_origin/p"], nbr_q=nbr_l2o_q, nbr_p=nbr_l2o_p)
        _lib.PROFILE_TAG[0] = "l3"
        l3_cv = pu.set_conv(f1(xyz[2]), grid(2, l2_new), sel3, 16, (5, 9), DOWN_CONV_DIS[3],
                            ["new_layer3/conv%d" % j for j in range(3)], store, [perms["new_layer3"]], nbr=nbr_new3)
        if sa3_done is not None:
            main.wait_event(sa3_done)
        # ---- level 3: embedding mask, attention pooling, coarse pose (:181-208)
        l3_w = pu.flow_predictor(f1(pts[3]), None, l3_cv, [128, 64], False, bn_decay, "l3_costvolume_predict_ww")
        l3_xyz = f1(xyz[3]).reshape(B, -1, 3)
        pose = mu.pose_head_call(l3_cv, l3_w, l3_xyz, heads=_heads(store, 3), want_pooled=want)
        q, t = pose["q"], pose["t"]
        q_norm, t_lvl = {3: pose["q_norm"]}, {3: t}
        if want:
            K.update(l2_points_f1_new=l2_new, l3_points_f1_cost_volume=l3_cv, l3_w=l3_w, l3_q=q, l3_t=t,
                     l3_pooled=pose["pooled"])

        up_xyz = f1(xyz[3])                 # level 2 up-samples from the UN-warped level-3 grid (:247)
        up_w, up_pred = grid(3, l3_w), grid(3, l3_cv)
        for lvl in (2, 1, 0):
            _lib.PROFILE_TAG[0] = "l%d" % lvl
            h, w_ = oh[lvl + 2], ow[lvl + 2]
            C = pts[lvl].shape[-1]
            # warp with the coarse pose and re-project (:213-237): one fused pass
            xyz_wp, pts_wp, warped = mu.project_points(f1(xyz[lvl]).reshape(B, -1, 3), f1(pts[lvl]), h, w_, mode=2,
                                                       q=q, t=t, want_points=want)
            # the level's four neighbour searches (cost volume q / p, the two up-convs) in one launch
            names = ["up_sa_layer_layer_l%dw" % lvl, "up_sa_layer_layer_l%dcostvolume" % lvl]
            allq = (h, w_, 1, 1)
            s_h, s_w = STRIDE_H[lvl + 3], STRIDE_W[lvl + 3]
            # row band of this rank (None: the level is computed whole).  Stage 2 of the cost volume looks one row up
            # and down (3x5 window), so stage 1 and its search also cover those two rows; nothing is exchanged for it.
            rb = band_rows(h, "l%d" % lvl, w_)
            qr = None if rb is None else (rb[0] * w_, rb[1] * w_)
            qr1 = None if rb is None else (max(rb[0] - 1, 0) * w_, min(rb[1] + 1, h) * w_)
            nq_, np_, nu0, nu1 = pu.multi_search([
                pu.search_spec(True, xyz_wp, f2(xyz[lvl]), allq, CV_KERNEL_Q[lvl], 6, 1000.0, 1, 1,
                               perms["flow_embedding_l%d/q" % lvl], qrange=qr1),
                pu.search_spec(False, xyz_wp, xyz_wp, allq, (3, 5), 4, COST_VOLUME_DIS[lvl], 1, 1,
                               perms["flow_embedding_l%d/p" % lvl], qrange=qr),
                pu.search_spec(False, xyz_wp, up_xyz, allq, (7, 15), 8, UP_CONV_DIS[lvl], s_h, s_w, perms[names[0]],
                               qrange=qr),
                pu.search_spec(False, xyz_wp, up_xyz, allq, (7, 15), 8, UP_CONV_DIS[lvl], s_h, s_w, perms[names[1]],
                               qrange=qr)])
            # The level's two chains after the searches -- cost volume (stage 1 -> stage 2) and the first half of the two
            # set-upconvs -- do not depend on each other: the set-upconvs go to a side stream (a fork / join that a
            # captured forward keeps as two branches of the graph), the predictors wait for both.
            if fork is not None:
                searched = torch.cuda.Event()
                searched.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(searched)
                    ups = pu.up_conv_group(xyz_wp, up_xyz, [up_w, up_pred], (7, 15), s_h, s_w, 8,
                                           UP_CONV_DIS[lvl], [["%s/up_1_%d" % (n, j) for j in range(2)] for n in names],
                                           store, [perms[n] for n in names], nbrs=[nu0, nu1], qrange=qr)
                    joined = torch.cuda.Event()
                    joined.record(side)
            # cost volume between the warped frame 1 and frame 2 (:242-244)
            cv = pu.cost_volume(xyz_wp, f2(xyz[lvl]), pts_wp, grid(lvl, f2(pts[lvl])), [3, 5], CV_KERNEL_Q[lvl], 4, 6,
                                COST_VOLUME_DIS[lvl], [128, 64, 64], [128, 64], False, bn_decay,
                                "flow_embedding_l%d" % lvl, random_hw_q=perms["flow_embedding_l%d/q" % lvl],
                                random_hw_p=perms["flow_embedding_l%d/p" % lvl], nbr_q=nq_, nbr_p=np_,
                                qrange1=qr1, qrange2=qr)
            # the two set-upconvs of the level (:247-251) share one launch for their first half ...
            if fork is not None:
                main.wait_event(joined)
            else:
                ups = pu.up_conv_group(xyz_wp, up_xyz, [up_w, up_pred], (7, 15), s_h, s_w, 8,
                                       UP_CONV_DIS[lvl], [["%s/up_1_%d" % (n, j) for j in range(2)] for n in names],
                                       store, [perms[n] for n in names], nbrs=[nu0, nu1], qrange=qr)
            # ... and one launch for their second half chained into the two predictors (:253-254)
            rows = B * h * w_
            pts_w = pts_wp.reshape(rows, C)
            cv_r = cv.reshape(rows, 64)
            pred_names = ["l%d_w_predict" % lvl, "l%d_costvolume_predict" % lvl]
            streams = [store.stream(["%s/up_2_0" % n, "%s/up_2_1" % n, "%s/conv_predictor0" % p,
                                     "%s/conv_predictor1" % p]) for n, p in zip(names, pred_names)]
            phases = [dict(sources=[[u.reshape(rows, 64) for u in ups], [pts_w]], channels=[64, C], couts=[128, 64]),
                      dict(sources=[[pts_w], None, [cv_r]], channels=[C, 64, 64], couts=[128, 64])]
            outs, up2 = pu.row_mlp(rows, phases, streams, 64, dev, want_phase0=want, phase0_channels=64, rrange=qr)
            wgt, pred = outs[0].view(B, h * w_, 64), outs[1].view(B, h * w_, 64)
            if rb is not None:          # the level's one exchange: every rank gets all rows of (mask logits, embedding)
                band.gather_rows([wgt, pred], h, w_)
            # attention pooling over the valid re-projected pixels + pose refinement (:262-280)
            pose = mu.pose_head_call(pred, wgt, xyz_wp.reshape(B, -1, 3), heads=_heads(store, lvl), coarse=(q, t),
                                     want_pooled=want)
            if want:
                K.update({"l%d_flow_warp" % lvl: warped, "l%d_xyz_warp_proj" % lvl: xyz_wp,
                          "l%d_points_warp_proj" % lvl: pts_wp, "l%d_cost_volume" % lvl: cv,
                          "l%d_w_up" % lvl: up2[0].view(B, h * w_, 64), "l%d_p_up" % lvl: up2[1].view(B, h * w_, 64),
                          "l%d_predict" % lvl: pred, "l%d_w" % lvl: wgt, "l%d_pooled" % lvl: pose["pooled"],
                          "l%d_q" % lvl: pose["q"], "l%d_t" % lvl: pose["t"]})
            q, t = pose["q"], pose["t"]
            q_norm[lvl], t_lvl[lvl] = pose["q_norm"], t
            up_xyz, up_w, up_pred = xyz_wp, wgt.view(B, h, w_, 64), pred.view(B, h, w_, 64)

    if gt_done is not None:
        main.wait_event(gt_done)
    _lib.PROFILE_TAG[0] = ""
    l0_xyz_f1 = f1(xyz[0]).reshape(B, -1, 3)
    return (q_norm[0], t_lvl[0], q_norm[1], t_lvl[1], q_norm[2], t_lvl[2], q_norm[3], t_lvl[3], l0_xyz_f1,
            q_gt, t_gt)


def get_loss(l0_q, l0_t, l1_q, l1_t, l2_q, l2_t, l3_q, l3_t, q_gt, t_gt, w_x, w_q):
    """Multi-scale pose loss with learnable uncertainty weights (pwclo_model.py:437-481)."""
    t_gt = t_gt.squeeze(-1)

    def level(q, t):
        qn = q / (torch.sqrt((q * q).sum(-1, keepdim=True) + 1e-10) + 1e-10)
        lq = torch.sqrt(((q_gt - qn) * (q_gt - qn)).sum(-1, keepdim=True) + 1e-10).mean()
        lx = torch.sqrt((t - t_gt) * (t - t_gt) + 1e-10).mean()
        return lx * torch.exp(-w_x) + w_x + lq * torch.exp(-w_q) + w_q

    return 1.6 * level(l3_q, l3_t) + 0.8 * level(l2_q, l2_t) + 0.4 * level(l1_q, l1_t) + 0.2 * level(l0_q, l0_t)
