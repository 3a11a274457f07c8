"""Building blocks of the PWCLO network on B200: set-conv, set-upconv, attentive cost volume, predictor.

Host-side mirror of the reference's utils/pointnet_util.py -- same function names, argument order and
meaning (cost_volume :33, flow_predictor :153, down_conv :179, up_conv :254, get_hw_idx :23) -- so that
pwclo_model.py wires up the same way.  Each block is one or two fused sm_100a kernels behind the C ABI
(include/elo_b200.h) instead of the reference's chain of TensorFlow ops.  Differences a caller sees:

  * weights come from the current ParamStore (store.py) under ``scope`` -- TensorFlow's implicit
    variable store made explicit;
  * the scan-order permutation each block draws with tf.random_shuffle (:45,104,193,270) can be passed
    in (``random_hw=``) so results are reproducible; if omitted it is drawn from torch's global RNG;
  * these fused blocks fold batch norm into the weights, i.e. serve ``is_training=False``; with
    is_training=True they raise NotImplementedError rather than silently using moving statistics --
    the training-mode graph (batch statistics + autograd) is train_graph.py, reached through
    get_model(..., is_training=True).
"""
import torch

from . import _lib
from .store import current_store, scoped


# ----------------------------------------------------------------------------------------------
def get_hw_idx(B, H, W, device="cuda"):
    """(B, H*W, 2) int32 [h, w] of every pixel, row-major (utils/pointnet_util.py:23-30)."""
    hh = torch.arange(H, dtype=torch.int32, device=device)[:, None].expand(H, W)
    ww = torch.arange(W, dtype=torch.int32, device=device)[None, :].expand(H, W)
    return torch.stack([hh, ww], dim=-1).reshape(1, H * W, 2).expand(B, -1, -1).contiguous()


class SelectedIdx:
    """Strided centre grid of model_util.get_selected_idx (model_util.py:296-316) kept symbolic: the
    kernels derive (b, h, w) from the query id instead of reading a (B, oh, ow, 3) index tensor."""

    def __init__(self, batch, stride_h, stride_w, out_h, out_w, device):
        self.batch, self.stride_h, self.stride_w, self.out_h, self.out_w = batch, stride_h, stride_w, out_h, out_w
        self.device = device

    def tensor(self):
        hh = torch.arange(0, self.out_h * self.stride_h, self.stride_h, dtype=torch.int32, device=self.device)
        ww = torch.arange(0, self.out_w * self.stride_w, self.stride_w, dtype=torch.int32, device=self.device)
        bb = torch.arange(self.batch, dtype=torch.int32, device=self.device)
        B, oh, ow = self.batch, self.out_h, self.out_w
        return torch.stack([bb.view(B, 1, 1).expand(B, oh, ow), hh.view(1, oh, 1).expand(B, oh, ow),
                            ww.view(1, 1, ow).expand(B, oh, ow)], dim=-1).contiguous()

    @staticmethod
    def from_tensor(idx):
        """Recover the strides from an explicit (B, oh, ow, 3) index tensor (host round trip)."""
        t = idx.detach().cpu()
        B, oh, ow, _ = t.shape
        sh = int(t[0, 1, 0, 1] - t[0, 0, 0, 1]) if oh > 1 else 1
        sw = int(t[0, 0, 1, 2] - t[0, 0, 0, 2]) if ow > 1 else 1
        sel = SelectedIdx(B, max(sh, 1), max(sw, 1), oh, ow, idx.device)
        if not torch.equal(sel.tensor().cpu(), t.to(torch.int32)):
            raise ValueError("selected_idx is not a regular strided grid; only get_selected_idx grids are supported")
        return sel


def _perm(random_hw, kt, device):
    if random_hw is None:
        random_hw = torch.randperm(kt)          # tf.random_shuffle(tf.range(kt))
    random_hw = torch.as_tensor(random_hw).to(device=device, dtype=torch.int32).contiguous()
    if random_hw.numel() != kt:
        raise ValueError("FusedConv expects (kernel_size_h * kernel_size_w) random_hw shape.")
    return random_hw


def _window(kernel_size, K, distance, stride_h, stride_w, small_h, small_w, random_hw):
    w = _lib.Window()
    w.kernel_size_H, w.kernel_size_W, w.K = int(kernel_size[0]), int(kernel_size[1]), int(K)
    w.distance = float(distance)
    w.stride_h, w.stride_w, w.small_h, w.small_w = int(stride_h), int(stride_w), int(small_h), int(small_w)
    w.random_hw = random_hw.data_ptr()
    return w


import os as _os
_SMALL_ONLY = _os.environ.get("ELO_SETCONV_SMALL_ONLY", "0") == "1"      # A/B: keep narrow chains on set_conv_small


def _is_training(is_training):
    return is_training is True or (isinstance(is_training, torch.Tensor) and bool(is_training))


def _check_training(is_training):
    if _is_training(is_training):
        raise NotImplementedError("the fused kernels behind this block fold batch norm into the weights and so "
                                  "only serve is_training=False; training-mode batch norm (batch statistics over "
                                  "B*N*K rows, utils/tf_util.py:527) is provided for the whole graph by "
                                  "get_model(..., is_training=True, params=TrainableParams) / train_graph.py")


def _f32(t):
    return t.contiguous().float()


# ----------------------------------------------------------------------------------------------
def search_spec(select, xyz1, xyz2, queries, kernel_size, K, distance, stride_h, stride_w, random_hw, out=None,
                qrange=None):
    """One entry for multi_search.  queries = (out_h, out_w, q_stride_h, q_stride_w) inside xyz1's image;
    `out` an optional pre-allocated (B, out_h*out_w, K) int32 view to write into; `qrange` = (begin, end): only
    these linear queries are searched (row bands), the other rows of the table are left untouched."""
    return dict(select=select, xyz1=xyz1, xyz2=xyz2, queries=queries, kernel_size=kernel_size, K=K,
                distance=distance, stride_h=stride_h, stride_w=stride_w, random_hw=random_hw, out=out, qrange=qrange)


def multi_search(specs):
    """Run up to 16 independent projection-aware neighbour searches in ONE launch (elo_multi_search).
    Returns one (B, n, K) int32 table per spec: linear cell of the searched grid, -1 = masked slot."""
    n = len(specs)
    arr = (_lib.SearchDesc * n)()
    outs, keep = [], []
    dev = specs[0]["xyz1"].device
    for i, sp in enumerate(specs):
        xyz1, xyz2 = _f32(sp["xyz1"]), _f32(sp["xyz2"])
        _lib.require_cuda("multi_search", xyz1, xyz2)
        B, H, W, _ = xyz1.shape
        oh, ow, qsh, qsw = sp["queries"]
        kt = sp["kernel_size"][0] * sp["kernel_size"][1]
        perm = _perm(sp["random_hw"], kt, dev)
        out = sp["out"] if sp["out"] is not None else torch.empty((B, oh * ow, sp["K"]), dtype=torch.int32, device=dev)
        if not out.is_contiguous() or out.dtype != torch.int32:
            raise ValueError("multi_search: out must be a contiguous int32 tensor")
        d = arr[i]
        d.select, d.batch_size = int(bool(sp["select"])), B
        d.queries = _lib.Queries(H, W, oh, ow, qsh, qsw)
        d.window = _window(sp["kernel_size"], sp["K"], sp["distance"], sp["stride_h"], sp["stride_w"],
                           xyz2.shape[1], xyz2.shape[2], perm)
        d.xyz1, d.xyz2, d.out_nbr = xyz1.data_ptr(), xyz2.data_ptr(), out.data_ptr()
        if sp.get("qrange") is not None:
            d.query_begin, d.query_end = int(sp["qrange"][0]), int(sp["qrange"][1])
        outs.append(out)
        keep += [xyz1, xyz2, perm]
    with torch.cuda.device(dev):
        rc = _lib.lib().elo_multi_search(arr, n, _lib.stream_ptr(dev))
    _lib.check(rc, "elo_multi_search")
    return outs


def set_conv(xyz_proj, points_proj, sel, K_sample, kernel_size, distance, layer_scopes, store, random_hws,
             feat_channels=None, set_batch_offsets=(0,), debug=None, nbr=None, qrange=None):
    """Set-conv kernel launch shared by down_conv and the batched pyramid of pwclo_model.

    xyz_proj (Bt, H, W, 3), points_proj (Bt, H, W, C) or None (zero features); ``sel`` a SelectedIdx
    whose .batch is the number of samples PER parameter set; ``random_hws`` one scan order per set;
    ``set_batch_offsets`` the first sample of each set inside the Bt stacked samples."""
    _lib.require_cuda("set_conv", xyz_proj, points_proj)
    Bt, H, W, _ = xyz_proj.shape
    dev = xyz_proj.device
    xyz_proj = _f32(xyz_proj)
    C = feat_channels if points_proj is None else points_proj.shape[-1]
    if points_proj is not None:
        points_proj = _f32(points_proj)
    widths = store.widths(layer_scopes)
    if store.cin(layer_scopes[0]) != 3 + C:
        raise ValueError("%s expects %d input channels, got 3 + %d" % (layer_scopes[0], store.cin(layer_scopes[0]), C))
    n = sel.out_h * sel.out_w
    nsets = len(random_hws)
    kt = kernel_size[0] * kernel_size[1]
    perms = [_perm(r, kt, dev) for r in random_hws]
    out = torch.empty((Bt, n, widths[-1]), dtype=torch.float32, device=dev)
    dbg = torch.full((Bt, n, K_sample), -2, dtype=torch.int32, device=dev) if debug is not None else None

    d = _lib.GroupMlpDesc()
    d.batch_size = sel.batch
    d.queries = _lib.Queries(H, W, sel.out_h, sel.out_w, sel.stride_h, sel.stride_w)
    d.nsets = nsets
    for s in range(nsets):
        d.set_batch_offset[s] = int(set_batch_offsets[s])
        d.window[s] = _window(kernel_size, K_sample, distance, 1, 1, H, W, perms[s])
    d.feat_channels = C
    d.num_layers = len(layer_scopes)
    for i, wd in enumerate(widths):
        d.cout[i] = wd
    d.xyz1 = d.xyz2 = xyz_proj.data_ptr()
    # the GEMM engines take 64- / 128-wide layers.  A chain that ENDS that wide but has narrower hidden layers (pyramid
    # layer 2: 35 -> 32 -> 32 -> 64) runs there too, its hidden layers widened with zero columns: 57 tensor-core tiles
    # instead of 58 CTAs that each push 4192 serial FMAs per row through registers (34 -> 13 us at B = 1).
    native = all(wd in (64, 128) for wd in widths)
    big = native or (widths[-1] in (64, 128) and C % 4 == 0 and all(wd <= 64 for wd in widths[:-1]) and not _SMALL_ONLY)
    pad = 64 if big and not native else None
    if pad:
        for i in range(len(widths) - 1):
            d.cout[i] = max(widths[i], pad) if widths[i] not in (64, 128) else widths[i]
    if big and points_proj is None:
        points_proj = torch.zeros((Bt, H, W, C), dtype=torch.float32, device=dev)
    weights = store.stream(layer_scopes, pad_hidden=pad) if big else store.plain(layer_scopes)
    for s in range(2):
        d.feat2[s] = _lib.ptr(points_proj)
        d.weights[s] = weights.data_ptr()
        d.out[s] = out.data_ptr()
        d.dbg_nbr[s] = _lib.ptr(dbg)
        d.nbr[s] = _lib.ptr(nbr)
    if qrange is not None:          # only these queries of every set (row bands); other output rows are not written
        d.query_begin, d.query_end = int(qrange[0]), int(qrange[1])
    _lib.call("elo_group_mlp_max" if big else "elo_set_conv_small", d, dev)
    if debug is not None:
        debug["nbr"] = dbg
    return out


def down_conv(xyz_proj, points_proj, selected_idx, K_sample, kernel_size, distance, mlp, mlp2, flag_add,
              is_training, bn_decay, scope, bn=True, pooling='max', knn=False, use_xyz=True, use_nchw=False,
              random_hw=None, params=None, debug=None):
    """Set-conv (utils/pointnet_util.py:179-250): random-K neighbours of each strided centre in its own
    range image, [xyz_diff, feat] -> mlp -> * mask -> max over K.
    Returns (new_points (B, n, mlp[-1]), new_xyz_proj (B, out_h, out_w, 3))."""
    _check_training(is_training)
    if mlp2 is not None or pooling != 'max' or not bn or use_nchw:
        raise NotImplementedError("down_conv: only mlp2=None, pooling='max', bn=True, NHWC (what pwclo_model uses)")
    store = current_store(params)
    sel = selected_idx if isinstance(selected_idx, SelectedIdx) else SelectedIdx.from_tensor(selected_idx)
    scopes = [scoped("%s/conv%d" % (scope, i)) for i in range(len(mlp))]
    if store.widths(scopes) != list(mlp):
        raise ValueError("%s: mlp %s does not match the stored weights %s" % (scope, list(mlp), store.widths(scopes)))
    out = set_conv(xyz_proj, points_proj, sel, K_sample, kernel_size, distance, scopes, store, [random_hw],
                   debug=debug)
    new_xyz_proj = xyz_proj[:, ::sel.stride_h, ::sel.stride_w][:, :sel.out_h, :sel.out_w].contiguous()
    return out, new_xyz_proj


# ----------------------------------------------------------------------------------------------
def up_conv_group(xyz1_proj, xyz2_proj, feat2_projs, kernel_size, stride_h, stride_w, nsample, distance,
                  scopes_per_set, store, random_hws, debug=None, nbrs=None, qrange=None):
    """First half of set-upconv for one or two parameter sets in one launch: random-K (with stride)
    neighbours of every dense pixel in the sparse grid, [xyz_diff, feat2] -> up_1_* -> * mask -> max."""
    _lib.require_cuda("up_conv", xyz1_proj, xyz2_proj, *feat2_projs)
    B, H, W, _ = xyz1_proj.shape
    h2, w2 = xyz2_proj.shape[1], xyz2_proj.shape[2]
    dev = xyz1_proj.device
    xyz1_proj, xyz2_proj = _f32(xyz1_proj), _f32(xyz2_proj)
    feats = [_f32(f) for f in feat2_projs]
    nsets = len(feats)
    widths = store.widths(scopes_per_set[0])
    kt = kernel_size[0] * kernel_size[1]
    perms = [_perm(r, kt, dev) for r in random_hws]
    outs = [torch.empty((B, H * W, widths[-1]), dtype=torch.float32, device=dev) for _ in range(nsets)]
    dbgs = [torch.full((B, H * W, nsample), -2, dtype=torch.int32, device=dev) if debug is not None else None
            for _ in range(nsets)]
    d = _lib.GroupMlpDesc()
    d.batch_size = B
    d.queries = _lib.Queries(H, W, H, W, 1, 1)
    d.nsets = nsets
    d.feat_channels = feats[0].shape[-1]
    d.num_layers = len(widths)
    for i, wd in enumerate(widths):
        d.cout[i] = wd
    d.xyz1, d.xyz2 = xyz1_proj.data_ptr(), xyz2_proj.data_ptr()
    streams = [store.stream(sc) for sc in scopes_per_set]
    for s in range(2):
        u = min(s, nsets - 1)
        d.set_batch_offset[s] = 0
        d.window[s] = _window(kernel_size, nsample, distance, stride_h, stride_w, h2, w2, perms[u])
        d.feat2[s] = feats[u].data_ptr()
        d.weights[s] = streams[u].data_ptr()
        d.out[s] = outs[u].data_ptr()
        d.dbg_nbr[s] = _lib.ptr(dbgs[u])
        d.nbr[s] = _lib.ptr(nbrs[u]) if nbrs is not None else None
    if qrange is not None:
        d.query_begin, d.query_end = int(qrange[0]), int(qrange[1])
    _lib.call("elo_group_mlp_max", d, dev)
    if debug is not None:
        debug["nbr"] = dbgs
    return outs


def row_mlp(rows, phases, weights_per_set, out_channels, device, want_phase0=False, phase0_channels=0, rrange=None,
            outs=None):
    """Launch elo_row_mlp.  phases: list of dicts(sources=[per-set list of tensors or None for
    'previous phase'], channels=[...], couts=[...]).  Returns (outs per set, phase-0 outs per set).
    `rrange` = (begin, end): only these rows are computed (rows are independent; the rest of the outputs is left
    untouched); `outs`: optional pre-allocated (rows, out_channels) tensors per set."""
    nsets = len(weights_per_set)
    r0, r1 = (0, int(rows)) if rrange is None else (int(rrange[0]), int(rrange[1]))
    d = _lib.RowMlpDesc()
    d.rows = r1 - r0
    d.nsets = nsets
    d.num_phases = len(phases)
    keep = []
    for ph, spec in enumerate(phases):
        f = d.phase[ph]
        f.num_sources = len(spec["channels"])
        f.num_layers = len(spec["couts"])
        for i, c in enumerate(spec["channels"]):
            f.channels[i] = int(c)
            src = spec["sources"][i]
            f.from_previous[i] = 1 if src is None else 0
            for s in range(2):
                if src is None:
                    f.src[s][i] = None
                else:
                    t = _f32(src[min(s, len(src) - 1)])
                    keep.append(t)
                    f.src[s][i] = t.data_ptr() + r0 * int(c) * 4
        for i, c in enumerate(spec["couts"]):
            f.cout[i] = int(c)
    if outs is None:
        outs = [torch.empty((rows, out_channels), dtype=torch.float32, device=device) for _ in range(nsets)]
    p0 = [torch.empty((rows, phase0_channels), dtype=torch.float32, device=device) if want_phase0 else None
          for _ in range(nsets)]
    for s in range(2):
        u = min(s, nsets - 1)
        d.weights[s] = weights_per_set[u].data_ptr()
        d.out[s] = outs[u].data_ptr() + r0 * out_channels * 4
        d.out_phase0[s] = (p0[u].data_ptr() + r0 * phase0_channels * 4) if p0[u] is not None else None
    if d.rows > 0:
        _lib.call("elo_row_mlp", d, device)
    return outs, p0


def up_conv(xyz1_proj, xyz2_proj, feat1_proj, feat2_proj, kernel_size, stride_h, stride_w, nsample, distance,
            mlp, mlp2, is_training, scope, bn_decay=None, bn=True, pooling='max', radius=None, knn=True,
            random_hw=None, params=None, debug=None):
    """Set-upconv (utils/pointnet_util.py:254-316): features of the sparse grid (xyz2, feat2) are
    propagated to every pixel of the dense grid (xyz1), concatenated with feat1 and refined by mlp2.
    Returns (B, H*W, mlp2[-1])."""
    _check_training(is_training)
    store = current_store(params)
    B, H, W, _ = xyz1_proj.shape
    s1 = [scoped("%s/up_1_%d" % (scope, j)) for j in range(len(mlp))]
    s2 = [scoped("%s/up_2_%d" % (scope, j)) for j in range(len(mlp2))]
    if store.widths(s1) != list(mlp) or store.widths(s2) != list(mlp2):
        raise ValueError("%s: mlp/mlp2 do not match the stored weights" % scope)
    up = up_conv_group(xyz1_proj, xyz2_proj, [feat2_proj], kernel_size, stride_h, stride_w, nsample, distance,
                       [s1], store, [random_hw], debug=debug)[0]
    C1 = feat1_proj.shape[-1]
    outs, _ = row_mlp(B * H * W, [dict(sources=[[up.reshape(B * H * W, -1)], [feat1_proj.reshape(B * H * W, C1)]],
                                       channels=[up.shape[-1], C1], couts=list(mlp2))],
                      [store.stream(s2)], mlp2[-1], xyz1_proj.device)
    return outs[0].reshape(B, H * W, -1)


# ----------------------------------------------------------------------------------------------
def cost_volume(warped_xyz1_proj, xyz2_proj, points1_proj, points2_proj, kernel_size1, kernel_size2, nsample,
                nsample_q, distance, mlp1, mlp2, is_training, bn_decay, scope, bn=True, pooling='max', knn=True,
                corr_func='elementwise_product', random_hw_q=None, random_hw_p=None, params=None, debug=None,
                nbr_q=None, nbr_p=None, qrange1=None, qrange2=None):
    """Two-stage attentive cost volume (utils/pointnet_util.py:33-149).  Stage 1 correlates every
    (warped) frame-1 pixel with its nsample_q nearest frame-2 pixels inside kernel_size2 (select-K,
    distance fixed at 1000 as in :51); stage 2 aggregates the stage-1 embeddings of nsample frame-1
    neighbours inside kernel_size1 within `distance`.  Returns (B, H*W, mlp2[-1] = 64)."""
    _check_training(is_training)
    if list(mlp1) != [128, 64, 64] or list(mlp2) != [128, 64]:
        raise NotImplementedError("cost_volume: the fused kernels are built for mlp1=[128,64,64], mlp2=[128,64]")
    _lib.require_cuda("cost_volume", warped_xyz1_proj, xyz2_proj, points1_proj, points2_proj)
    store = current_store(params)
    B, H, W, _ = warped_xyz1_proj.shape
    dev = warped_xyz1_proj.device
    C = points1_proj.shape[-1]
    xyz1, xyz2 = _f32(warped_xyz1_proj), _f32(xyz2_proj)
    f1, f2 = _f32(points1_proj), _f32(points2_proj)
    sc = lambda n: scoped("%s/%s" % (scope, n))
    w1 = store.stream([sc("CV_0"), sc("CV_1"), sc("CV_2"), sc("CV_xyz"), sc("sum_CV_0"), sc("sum_CV_1")])
    w2 = store.stream([sc("sum_xyz_encoding"), sc("sum_cost_volume_0"), sc("sum_cost_volume_1")])
    if store.cin(sc("CV_0")) != 10 + 2 * C:
        raise ValueError("%s/CV_0 expects %d input channels, got 10 + 2*%d" % (scope, store.cin(sc("CV_0")), C))
    perm_q = _perm(random_hw_q, kernel_size2[0] * kernel_size2[1], dev)
    perm_p = _perm(random_hw_p, kernel_size1[0] * kernel_size1[1], dev)
    if nbr_q is None or nbr_p is None:
        # stand-alone call: both neighbour searches in one launch (pwclo_model batches them per level itself)
        allq = (H, W, 1, 1)
        nbr_q, nbr_p = multi_search([
            search_spec(True, xyz1, xyz2, allq, kernel_size2, nsample_q, 1000.0, 1, 1, perm_q),
            search_spec(False, xyz1, xyz1, allq, kernel_size1, nsample, distance, 1, 1, perm_p)])
    stage1 = torch.empty((B, H * W, 64), dtype=torch.float32, device=dev)
    out = torch.empty((B, H * W, 64), dtype=torch.float32, device=dev)
    dq = torch.full((B, H * W, nsample_q), -2, dtype=torch.int32, device=dev) if debug is not None else None
    dp = torch.full((B, H * W, nsample), -2, dtype=torch.int32, device=dev) if debug is not None else None
    d = _lib.CostVolumeDesc()
    d.batch_size, d.H, d.W, d.C = B, H, W, C
    d.window_q = _window(kernel_size2, nsample_q, 1000.0, 1, 1, H, W, perm_q)
    d.window_p = _window(kernel_size1, nsample, distance, 1, 1, H, W, perm_p)
    d.xyz1, d.xyz2, d.f1, d.f2 = xyz1.data_ptr(), xyz2.data_ptr(), f1.data_ptr(), f2.data_ptr()
    d.weights_1, d.weights_2 = w1.data_ptr(), w2.data_ptr()
    d.stage1_out, d.out = stage1.data_ptr(), out.data_ptr()
    d.dbg_nbr_q, d.dbg_nbr_p = _lib.ptr(dq), _lib.ptr(dp)
    d.nbr_q, d.nbr_p = _lib.ptr(nbr_q), _lib.ptr(nbr_p)
    # row bands: stage 1 on qrange1 (the band plus the rows stage 2's window reaches into), stage 2 on qrange2
    if qrange1 is not None:
        d.query_begin, d.query_end = int(qrange1[0]), int(qrange1[1])
    _lib.call("elo_cost_volume_1", d, dev)
    d.query_begin, d.query_end = (0, 0) if qrange2 is None else (int(qrange2[0]), int(qrange2[1]))
    _lib.call("elo_cost_volume_2", d, dev)
    if debug is not None:
        debug.update(nbr_q=dq, nbr_p=dp, stage1=stage1)
    return out


def flow_predictor(points_f1, upsampled_feat, cost_volume, mlp, is_training, bn_decay, scope, bn=True,
                   params=None):
    """Shared MLP on [points_f1, upsampled_feat, cost_volume] (utils/pointnet_util.py:153-175); either of
    the last two may be None.  Used for the embedding refinement and for the embedding-mask logits."""
    _check_training(is_training)
    store = current_store(params)
    parts = [t for t in (points_f1, upsampled_feat, cost_volume) if t is not None]
    _lib.require_cuda("flow_predictor", *parts)
    B, N = points_f1.shape[0], points_f1.shape[1]
    scopes = [scoped("%s/conv_predictor%d" % (scope, i)) for i in range(len(mlp))]
    if store.widths(scopes) != list(mlp):
        raise ValueError("%s: mlp does not match the stored weights" % scope)
    outs, _ = row_mlp(B * N, [dict(sources=[[t.reshape(B * N, t.shape[-1])] for t in parts],
                                   channels=[t.shape[-1] for t in parts], couts=list(mlp))],
                      [store.stream(scopes)], mlp[-1], points_f1.device)
    return outs[0].reshape(B, N, -1)


def warping_layers(xyz1, upsampled_flow):
    """utils/pointnet_util.py:18-20 (unused by the model; kept for API completeness)."""
    return xyz1 + upsampled_flow
