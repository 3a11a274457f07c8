"""Seeded input generators shared by the CPU and GPU parity tests of the two index ops."""
import math

import numpy as np

# (scope, op, H, W, npoints, kH, kW, K, distance, stride_h, stride_w) -- the 19 distinct call
# signatures of the model graph (pwclo_model.py:126-397).  Queries of the
# down_conv sites are the strided sub-grid (selected_idx), all others every cell.
MODEL_SITES = [
    ("sa1/layer0", "random", 64, 1800, 3600, 9, 15, 32, 0.5, 1, 1),
    ("sa1/layer1", "random", 16, 225, 904, 7, 11, 32, 3.0, 1, 1),
    ("sa1/layer2", "random", 8, 113, 228, 5, 9, 16, 6.0, 1, 1),
    ("sa1/layer3", "random", 4, 57, 116, 5, 9, 16, 12.0, 1, 1),
    ("flow_embedding_l2_origin/q", "select", 4, 57, 228, 5, 35, 32, 1000.0, 1, 1),
    ("flow_embedding_l2_origin/p", "random", 4, 57, 228, 3, 5, 4, 4.0, 1, 1),
    ("flow_embedding_l2/q", "select", 4, 57, 228, 5, 15, 6, 1000.0, 1, 1),
    ("up_sa_layer_layer_l2", "random", 4, 57, 228, 7, 15, 8, 9.0, 1, 2),
    ("flow_embedding_l1/q", "select", 8, 113, 904, 7, 25, 6, 1000.0, 1, 1),
    ("flow_embedding_l1/p", "random", 8, 113, 904, 3, 5, 4, 2.0, 1, 1),
    ("up_sa_layer_layer_l1", "random", 8, 113, 904, 7, 15, 8, 6.0, 2, 2),
    ("flow_embedding_l0/q", "select", 16, 225, 3600, 11, 41, 6, 1000.0, 1, 1),
    ("flow_embedding_l0/p", "random", 16, 225, 3600, 3, 5, 4, 1.0, 1, 1),
    ("up_sa_layer_layer_l0", "random", 16, 225, 3600, 7, 15, 8, 3.0, 2, 2),
]


def all_cells(B, H, W):
    hh, ww = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    q = np.stack([hh, ww], -1).reshape(1, H * W, 2).astype(np.int32)
    return np.ascontiguousarray(np.tile(q, (B, 1, 1)))


def strided_cells(B, H, W, n):
    """A strided sub-grid with exactly n queries (how down_conv samples its centres)."""
    for sh in (1, 2, 4, 8):
        for sw in (1, 2, 4, 8, 16):
            oh, ow = math.ceil(H / sh), math.ceil(W / sw)
            if oh * ow == n:
                hh, ww = np.meshgrid(np.arange(0, oh * sh, sh), np.arange(0, ow * sw, sw), indexing="ij")
                q = np.stack([hh, ww], -1).reshape(1, n, 2).astype(np.int32)
                return np.ascontiguousarray(np.tile(q, (B, 1, 1)))
    raise ValueError("no stride gives %d queries on %dx%d" % (n, H, W))


def range_image(rng, B, H, W, holes=0.2, integer=False, scale=10.0):
    """A smooth-ish random range image (B,H,W,3) with a fraction of empty (all-zero) pixels.
    integer=True makes small-integer coordinates: exact arithmetic and plenty of distance ties."""
    if integer:
        xyz = rng.integers(-3, 4, size=(B, H, W, 3)).astype(np.float32)
    else:
        az = np.linspace(-np.pi, np.pi, W, endpoint=False)[None, None, :]
        el = np.linspace(0.05, -0.4, H)[None, :, None]
        r = scale * (1.0 + 0.3 * rng.standard_normal((B, H, W)))
        r = np.abs(r) + 0.5
        xyz = np.stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az),
                        r * np.sin(el) * np.ones_like(az)], -1).astype(np.float32)
    keep = rng.random((B, H, W)) >= holes
    return np.ascontiguousarray(xyz * keep[..., None].astype(np.float32))


def random_case(rng, integer=None):
    """One randomised op invocation (kwargs of the 14-argument op + 'mode')."""
    B = int(rng.integers(1, 3))
    sh, sw = int(rng.choice([1, 1, 2])), int(rng.choice([1, 1, 2, 3]))
    H, W = int(rng.integers(1, 12)), int(rng.integers(4, 40))
    h2, w2 = math.ceil(H / sh), math.ceil(W / sw)
    kH = int(rng.choice([1, 3, 5, 7]))
    kW = int(rng.choice([1, 3, 5, 9, 15]))
    while kW // 2 > w2:          # only one wrap is defined behaviour
        kW -= 2
    kt = kH * kW
    K = int(rng.choice([1, 2, 4, 6, 8, 16, 32, 64]))
    if integer is None:
        integer = bool(rng.random() < 0.4)
    xyz1 = range_image(rng, B, H, W, holes=float(rng.choice([0.0, 0.2, 0.3, 0.9])), integer=integer)
    if rng.random() < 0.5 and sh == 1 and sw == 1:
        xyz2 = xyz1.copy()
    else:
        xyz2 = range_image(rng, B, h2, w2, holes=float(rng.choice([0.0, 0.2, 0.3, 1.0])), integer=integer)
    n = int(rng.integers(1, H * W + 1))
    idx = np.stack([rng.integers(0, H, size=(B, n)), rng.integers(0, W, size=(B, n))], -1).astype(np.int32)
    distance = float(rng.choice([0.5, 1.0, 2.0, 3.0, 12.0, 1000.0]))
    return dict(mode=str(rng.choice(["select", "random"])), xyz1=xyz1, xyz2=xyz2, idx_n2=idx,
                random_hw=rng.permutation(kt).astype(np.int32), H=H, W=W, npoints=n,
                kernel_size_H=kH, kernel_size_W=kW, K=K, flag_copy=int(rng.random() < 0.3),
                distance=distance, stride_h=sh, stride_w=sw)


def site_case(rng, site, B=1, holes=0.1):
    """Inputs shaped like one real call site of the model."""
    scope, op, H, W, n, kH, kW, K, dist, sh, sw = site
    h2, w2 = math.ceil(H / sh), math.ceil(W / sw)
    xyz1 = range_image(rng, B, H, W, holes=holes, scale=8.0)
    if sh == 1 and sw == 1 and op == "random":
        xyz2 = xyz1.copy()
    else:
        xyz2 = range_image(rng, B, h2, w2, holes=holes, scale=8.0)
    idx = all_cells(B, H, W) if n == H * W else strided_cells(B, H, W, n)
    return dict(mode=op, xyz1=xyz1, xyz2=xyz2, idx_n2=idx, random_hw=rng.permutation(kH * kW).astype(np.int32),
                H=H, W=W, npoints=n, kernel_size_H=kH, kernel_size_W=kW, K=K, flag_copy=0,
                distance=dist, stride_h=sh, stride_w=sw)


def call(fn, case, **extra):
    c = dict(case)
    mode = c.pop("mode")
    return fn(mode, c["xyz1"], c["xyz2"], c["idx_n2"], c["random_hw"], c["H"], c["W"], c["npoints"],
              c["kernel_size_H"], c["kernel_size_W"], c["K"], c["flag_copy"], c["distance"],
              c["stride_h"], c["stride_w"], **extra)


OUT_NAMES = ("selected_bhw_idx", "valid_idx", "valid_in_dis_idx", "selected_mask")


def assert_same(a, b, what=""):
    for name, x, y in zip(OUT_NAMES, a, b):
        x, y = np.asarray(x), np.asarray(y)
        assert x.shape == y.shape, "%s %s: shape %s vs %s" % (what, name, x.shape, y.shape)
        if not np.array_equal(x, y):
            bad = np.argwhere(x != y)
            raise AssertionError("%s %s: %d mismatches, first at %s: %s vs %s"
                                 % (what, name, len(bad), bad[0].tolist(), x[tuple(bad[0])], y[tuple(bad[0])]))
