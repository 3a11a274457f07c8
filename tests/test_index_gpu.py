"""GPU: the CUDA index ops, called through the C ABI, against the oracle -- bit-exact on all four
outputs (selected_bhw_idx, valid_idx, valid_in_dis_idx, selected_mask)."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import index_oracle as io

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "index_golden.npz")


@pytest.fixture(autouse=True, params=["tiled", "tiled+storewarp", "tiled+plainloads", "warp"])
def index_kernel(request, elo):
    """Every test of this module runs on all work decompositions of the index ops: the tile-staged
    thread-per-query kernel (fused_conv_tiled.cu) without and with its store warp (select-K: one more warp per CTA
    writes the count rows while the query warps walk), with its tile staged by bulk copies (default) or by plain loads,
    and one warp per query (fused_conv_index.cu)."""
    elo._lib.set_index_kernel(2 if request.param == "warp" else 1)
    elo._lib.set_store_warp_min_cells(0 if request.param == "tiled+storewarp" else 1 << 30)
    elo._lib.set_tile_staging(1 if request.param == "tiled+plainloads" else 0)
    yield request.param
    elo._lib.set_index_kernel(0)
    elo._lib.set_store_warp_min_cells(128)
    elo._lib.set_tile_staging(0)


def run_cuda(elo, cuda, mode, xyz1, xyz2, idx_n2, random_hw, H, W, npoints, kH, kW, K, flag_copy,
             distance, sh, sw):
    fn = elo.fused_conv_select_k if mode == "select" else elo.fused_conv_random_k
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(cuda)
    outs = fn(t(xyz1), t(xyz2), t(idx_n2), t(random_hw), H, W, npoints, kH, kW, K, flag_copy, distance, sh, sw)
    torch.cuda.synchronize()
    return tuple(o.cpu().numpy() for o in outs)


def cuda_impl(elo, cuda):
    return lambda mode, *a, **k: run_cuda(elo, cuda, mode, *a, **k)


def test_demo_known_answers(elo, cuda):
    H, W = 4, 7
    xyz = np.tile(np.arange(H * W, dtype=np.float32).reshape(1, H, W, 1), (1, 1, 1, 3))
    idx = np.array([[[0, 0], [0, 1]]], np.int32)
    for mode, rhw, cols in [("select", [0, 1, 2, 3, 4], [1, 2, 3, 6]), ("select", [3, 0, 4, 2, 1], [1, 2, 3, 6]),
                            ("random", [0, 1, 2, 3, 4], [6, 1, 2, 3]), ("random", [3, 0, 4, 2, 1], [2, 6, 3, 1])]:
        sel, valid, vdis, mask = run_cuda(elo, cuda, mode, xyz, xyz, idx, np.array(rhw, np.int32), H, W, 2, 1, 5, 8, 0, 200.0, 1, 1)
        assert not sel[0, 0].any() and not valid[0, 0].any() and not vdis[0, 0].any() and not mask[0, 0].any()
        assert sel[0, 1, :, 2].tolist() == cols + [0, 0, 0, 0]
        assert mask[0, 1, :, 0].tolist() == [1, 1, 1, 1, 0, 0, 0, 0]
        assert valid[0, 1, :, 0].tolist() == [1, 1, 1, 1, 0] and vdis[0, 1, :, 0].tolist() == [1, 1, 1, 1, 0]


def test_select_k_tie_order(elo, cuda):
    W = 8
    xyz = np.zeros((1, 1, W, 3), np.float32)
    xyz[0, 0, :, 0] = [11, 10, 9, 50, 60, 70, 80, 90]
    xyz[0, 0, :, 1] = 1
    idx = np.array([[[0, 1]]], np.int32)
    sel, _, _, mask = run_cuda(elo, cuda, "select", xyz, xyz, idx, np.array([0, 2, 1], np.int32), 1, W, 1, 1, 3, 3, 0, 1000.0, 1, 1)
    assert sel[0, 0, :, 2].tolist() == [1, 2, 0] and mask[0, 0, :, 0].tolist() == [1, 1, 1]


@pytest.mark.parametrize("chunk", range(8))
def test_random_cases_vs_oracle(elo, cuda, chunk):
    for seed in range(chunk * 40, chunk * 40 + 40):
        case = cases.random_case(np.random.default_rng(seed))
        cases.assert_same(cases.call(cuda_impl(elo, cuda), case), cases.call(io.port, case), "seed %d" % seed)


@pytest.mark.parametrize("chunk", range(4))
def test_tie_heavy_integer_cases_vs_oracle(elo, cuda, chunk):
    for seed in range(5000 + chunk * 40, 5000 + chunk * 40 + 40):
        case = cases.random_case(np.random.default_rng(seed), integer=True)
        cases.assert_same(cases.call(cuda_impl(elo, cuda), case), cases.call(io.port, case), "seed %d" % seed)


@pytest.mark.parametrize("site", cases.MODEL_SITES, ids=lambda s: s[0])
def test_model_sites_vs_oracle_and_reference_kernel(elo, cuda, site):
    rng = np.random.default_rng(abs(hash(site[0])) % (2 ** 31))
    case = cases.site_case(rng, site, B=2)
    got = cases.call(cuda_impl(elo, cuda), case)
    cases.assert_same(got, cases.call(io.port, case, nthreads=8), site[0] + " vs port")
    if io.have_ref_gpu():
        c = dict(case)
        for k in ("xyz1", "xyz2", "idx_n2", "random_hw"):
            c[k] = torch.as_tensor(c[k]).to(cuda)
        ref = tuple(o.cpu().numpy() for o in cases.call(io.ref_gpu, c))
        cases.assert_same(got, ref, site[0] + " vs reference .cu")


def raster_case(rng, mode, B, H, W, kH, kW, K, sh, sw, distance, flag_copy=0, holes=0.2, integer=False, n=None):
    """Queries in raster order (how the model and BASELINE.json configs[0] call the ops): the staged-tile path."""
    h2, w2 = -(-H // sh), -(-W // sw)
    xyz1 = cases.range_image(rng, B, H, W, holes=holes, integer=integer)
    xyz2 = xyz1.copy() if (sh == 1 and sw == 1 and rng.random() < 0.5) else cases.range_image(rng, B, h2, w2, holes=holes, integer=integer)
    idx = cases.all_cells(B, H, W)
    if n is not None:
        idx = np.ascontiguousarray(idx[:, :n])
    return dict(mode=mode, xyz1=xyz1, xyz2=xyz2, idx_n2=idx, random_hw=rng.permutation(kH * kW).astype(np.int32),
                H=H, W=W, npoints=idx.shape[1], kernel_size_H=kH, kernel_size_W=kW, K=K, flag_copy=flag_copy,
                distance=distance, stride_h=sh, stride_w=sw)


RASTER = [
    # B, H, W, kH, kW, K, sh, sw, distance, flag_copy, integer, n
    (1, 16, 225, 11, 41, 6, 1, 1, 1000.0, 0, False, None),
    (2, 8, 113, 7, 25, 6, 1, 1, 1000.0, 0, False, None),      # 904 queries: CTAs straddle the two samples
    (2, 8, 113, 7, 15, 8, 2, 2, 6.0, 0, False, None),         # strided window centres
    (1, 4, 57, 5, 35, 32, 1, 1, 1000.0, 1, False, None),
    (3, 5, 21, 3, 9, 16, 1, 1, 3.0, 1, False, 77),            # window wider than half the cylinder, ragged N
    (1, 9, 10, 5, 15, 4, 1, 1, 1000.0, 0, False, None),       # kW > W: every column seen more than once
    (2, 12, 64, 7, 9, 16, 1, 3, 2.0, 1, True, None),          # integer grid: ties everywhere -> exact replay
    (1, 32, 300, 9, 15, 32, 1, 1, 0.5, 0, False, None),
    (1, 6, 40, 1, 1, 1, 1, 1, 1000.0, 0, False, None),
    (1, 6, 40, 3, 3, 7, 1, 1, 1000.0, 1, True, 130),
]


@pytest.mark.parametrize("mode", ["select", "random"])
@pytest.mark.parametrize("spec", RASTER, ids=lambda s: "B%d_%dx%d_k%dx%d_K%d_s%d%d" % s[:8])
def test_raster_queries_vs_oracle(elo, cuda, mode, spec):
    B, H, W, kH, kW, K, sh, sw, distance, flag_copy, integer, n = spec
    rng = np.random.default_rng(abs(hash((mode,) + spec[:8])) % (2 ** 31))
    case = raster_case(rng, mode, B, H, W, kH, kW, K, sh, sw, distance, flag_copy, integer=integer, n=n)
    cases.assert_same(cases.call(cuda_impl(elo, cuda), case), cases.call(io.port, case, nthreads=8), "raster %s" % (spec,))


def test_select_k_near_ties_take_the_exact_path(elo, cuda):
    """Distances that differ only in the mantissa bits the packed keys give up for the walk position must
    still come out in the reference's order (the tiled kernel replays such queries exactly)."""
    H, W, kH, kW, K = 3, 96, 3, 31, 8
    rng = np.random.default_rng(11)
    xyz = np.zeros((1, H, W, 3), np.float32)
    base = 1.0 + rng.integers(0, 6, size=(H, W)).astype(np.float32)          # few distinct ranges
    xyz[0, :, :, 0] = base * (1.0 + rng.integers(0, 4, size=(H, W)).astype(np.float32) * 2.0 ** -21)
    xyz[0, :, :, 1] = 0.25
    idx = cases.all_cells(1, H, W)
    case = dict(mode="select", xyz1=xyz, xyz2=xyz, idx_n2=idx, random_hw=rng.permutation(kH * kW).astype(np.int32),
                H=H, W=W, npoints=H * W, kernel_size_H=kH, kernel_size_W=kW, K=K, flag_copy=0, distance=1000.0,
                stride_h=1, stride_w=1)
    cases.assert_same(cases.call(cuda_impl(elo, cuda), case), cases.call(io.port, case, nthreads=8), "near ties")


@pytest.mark.parametrize("mode", ["select", "random"])
def test_unaligned_output_buffers(elo, cuda, mode):
    """Output pointers that are only 4-byte aligned (views into larger buffers): the kernels fall back from
    16-byte to scalar stores and must write exactly the same rows, and nothing outside them."""
    rng = np.random.default_rng(21)
    case = raster_case(rng, mode, 2, 6, 50, 5, 9, 8, 1, 1, 3.0, flag_copy=1)
    want = cases.call(io.port, case, nthreads=8)
    B, N, K, kt = 2, case["npoints"], case["K"], 45
    t = {k: torch.as_tensor(case[k]).to(cuda) for k in ("xyz1", "xyz2", "idx_n2", "random_hw")}
    pad = 3                                           # elements in front: 12 bytes off a 16-byte boundary
    bufs = {"idx": torch.full((pad + B * N * K * 3 + 5,), -7, dtype=torch.int32, device=cuda),
            "valid": torch.full((pad + B * N * kt + 5,), -7.0, device=cuda),
            "vdis": torch.full((pad + B * N * kt + 5,), -7.0, device=cuda),
            "mask": torch.full((pad + B * N * K + 5,), -7.0, device=cuda)}
    lib = elo._lib.lib()
    fn = lib.elo_fused_conv_select_k if mode == "select" else lib.elo_fused_conv_random_k
    ptr = lambda name: bufs[name].data_ptr() + 4 * pad
    rc = fn(B, 6, 50, N, 5, 9, K, 1, 3.0, 1, 1, t["xyz1"].data_ptr(), t["xyz2"].data_ptr(), t["idx_n2"].data_ptr(),
            t["random_hw"].data_ptr(), ptr("idx"), ptr("valid"), ptr("vdis"), ptr("mask"), 6, 50,
            torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    got = (bufs["idx"][pad:-5].reshape(B, N, K, 3), bufs["valid"][pad:-5].reshape(B, N, kt, 1),
           bufs["vdis"][pad:-5].reshape(B, N, kt, 1), bufs["mask"][pad:-5].reshape(B, N, K, 1))
    cases.assert_same(tuple(g.cpu().numpy() for g in got), want, "unaligned " + mode)
    for b in bufs.values():                            # guard elements untouched
        assert (b[:pad] == -7).all() and (b[-5:] == -7).all()


@pytest.mark.parametrize("mode", ["select", "random"])
def test_unaligned_input_grid(elo, cuda, mode):
    """xyz2 that is only 4-byte aligned (a view one float into a buffer): bulk copies need 16-byte aligned sources, so the
    tiled kernel stages its tile with plain loads; same outputs."""
    rng = np.random.default_rng(22)
    case = raster_case(rng, mode, 2, 6, 50, 5, 9, 8, 1, 1, 3.0, flag_copy=1)
    want = cases.call(io.port, case, nthreads=8)
    t = {k: torch.as_tensor(case[k]).to(cuda) for k in ("xyz1", "idx_n2", "random_hw")}
    buf = torch.zeros(case["xyz2"].size + 4, device=cuda)
    buf[1:1 + case["xyz2"].size] = torch.as_tensor(case["xyz2"]).to(cuda).reshape(-1)
    xyz2 = buf[1:1 + case["xyz2"].size].reshape(case["xyz2"].shape)
    assert xyz2.data_ptr() % 16 == 4
    fn = elo.fused_conv_select_k if mode == "select" else elo.fused_conv_random_k
    outs = fn(t["xyz1"], xyz2, t["idx_n2"], t["random_hw"], 6, 50, case["npoints"], 5, 9, 8, 1, 3.0, 1, 1)
    torch.cuda.synchronize()
    cases.assert_same(tuple(o.cpu().numpy() for o in outs), want, "unaligned xyz2 " + mode)


@pytest.mark.parametrize("kH,kW", [(31, 33), (33, 33)])
def test_select_k_large_windows(elo, cuda, kH, kW):
    """Windows around the 1024-cell limit of the tiled kernel's parameter-resident walk table (1023 cells: tiled;
    1089: the warp-per-query kernel takes the call even when the tiled one is requested)."""
    rng = np.random.default_rng(31)
    case = raster_case(rng, "select", 1, 40, 48, kH, kW, 16, 1, 1, 1000.0, holes=0.3)
    cases.assert_same(cases.call(cuda_impl(elo, cuda), case), cases.call(io.port, case, nthreads=8), "window %dx%d" % (kH, kW))


def test_golden_vectors(elo, cuda):
    g = np.load(GOLDEN)
    for i in range(int(g["n_cases"])):
        p = "c%d_" % i
        case = dict(mode=str(g[p + "mode"]), xyz1=g[p + "xyz1"], xyz2=g[p + "xyz2"], idx_n2=g[p + "idx_n2"],
                    random_hw=g[p + "random_hw"], distance=float(g[p + "distance"]))
        for k, v in zip(("H", "W", "npoints", "kernel_size_H", "kernel_size_W", "K", "flag_copy",
                         "stride_h", "stride_w"), g[p + "ints"].tolist()):
            case[k] = int(v)
        cases.assert_same(cases.call(cuda_impl(elo, cuda), case), tuple(g[p + n] for n in cases.OUT_NAMES), "golden %d" % i)


def test_nonfinite_inputs_match_reference_kernel(elo, cuda):
    if not io.have_ref_gpu():
        pytest.skip("oracle/_ref/libref_gpu.so not built")
    rng = np.random.default_rng(3)
    case = cases.site_case(rng, cases.MODEL_SITES[6], B=1)
    for arr in (case["xyz1"], case["xyz2"]):
        flat = arr.reshape(-1)
        pos = rng.choice(flat.size, 40, replace=False)
        flat[pos[:20]] = np.nan
        flat[pos[20:30]] = np.inf
        flat[pos[30:]] = -np.inf
    got = cases.call(cuda_impl(elo, cuda), case)
    c = dict(case)
    for k in ("xyz1", "xyz2", "idx_n2", "random_hw"):
        c[k] = torch.as_tensor(c[k]).to(cuda)
    ref = tuple(o.cpu().numpy() for o in cases.call(io.ref_gpu, c))
    cases.assert_same(got, ref, "non-finite")


@pytest.mark.parametrize("kH,kW", [(7, 25), (11, 41)])
def test_config1_full_frame_select_k(elo, cuda, kH, kW):
    """BASELINE.json configs[0]: one 64x1800 synthetic frame, K=16, every cell a query."""
    H, W, K = 64, 1800, 16
    xyz = elo.synth.synth_scan(H, W, seed=0)[None]
    idx = elo.synth.hw_index(1, H, W)
    rhw = torch.randperm(kH * kW, generator=torch.Generator().manual_seed(0)).to(torch.int32)
    got = run_cuda(elo, cuda, "select", xyz.numpy(), xyz.numpy(), idx.numpy(), rhw.numpy(), H, W, H * W, kH, kW, K, 0, 1000.0, 1, 1)
    want = io.port("select", xyz.numpy(), xyz.numpy(), idx.numpy(), rhw.numpy(), H, W, H * W, kH, kW, K, 0, 1000.0, 1, 1,
                   nthreads=os.cpu_count() or 1)
    cases.assert_same(got, want, "config 1 %dx%d" % (kH, kW))
    # size-independent properties: masks are a prefix of ones, counts are run lengths with
    # nsel <= nvalid, every selected cell is a non-empty pixel, distances are non-decreasing,
    # and the nearest neighbour of a valid pixel searched in its own frame is itself.
    sel, valid, vdis, mask = got
    m = mask[0, :, :, 0]
    assert ((m[:, :-1] >= m[:, 1:]).all())
    v, d = valid[0, :, :, 0], vdis[0, :, :, 0]
    assert (v[:, :-1] >= v[:, 1:]).all() and (d[:, :-1] >= d[:, 1:]).all() and (d.sum(1) <= v.sum(1)).all()
    assert (np.minimum(d.sum(1), K) == m.sum(1)).all()
    g = xyz[0].numpy()
    pts = g[sel[0, :, :, 1], sel[0, :, :, 2]]
    centre = g.reshape(-1, 3)
    assert (np.abs(pts).sum(-1)[m > 0] > 0).all()
    dist = ((pts - centre[:, None, :]) ** 2).sum(-1)
    dist = np.where(m > 0, dist, np.inf)
    assert (dist[:, :-1] <= dist[:, 1:] * (1 + 1e-5) + 1e-9).all()
    ok = np.abs(centre).sum(-1) > 0
    own = np.stack([idx[0, :, 0].numpy(), idx[0, :, 1].numpy()], -1)
    assert (sel[0, ok, 0, 1:] == own[ok]).all()


def test_optional_count_outputs_and_graph_capture(elo, cuda):
    rng = np.random.default_rng(9)
    case = cases.site_case(rng, cases.MODEL_SITES[8], B=2)
    t = {k: torch.as_tensor(case[k]).to(cuda) for k in ("xyz1", "xyz2", "idx_n2", "random_hw")}
    full = elo.fused_conv_select_k(t["xyz1"], t["xyz2"], t["idx_n2"], t["random_hw"], case["H"], case["W"],
                                   case["npoints"], 7, 25, 6, 0, 1000.0, 1, 1)
    idx2, mask2 = elo.fused_conv_indices(True, t["xyz1"], t["xyz2"], t["idx_n2"], t["random_hw"], 7, 25, 6, 0, 1000.0, 1, 1)
    assert torch.equal(full[0], idx2) and torch.equal(full[3], mask2)
    # no allocation / sync / host read inside the C ABI call: it can be captured in a CUDA graph
    lib = elo._lib.lib()
    out_idx = torch.full_like(full[0], -7)
    out_mask = torch.full_like(full[3], -7)
    stream = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(graph, stream=stream):
        rc = lib.elo_fused_conv_select_k(2, case["H"], case["W"], case["npoints"], 7, 25, 6, 0, 1000.0, 1, 1,
                                         t["xyz1"].data_ptr(), t["xyz2"].data_ptr(), t["idx_n2"].data_ptr(),
                                         t["random_hw"].data_ptr(), out_idx.data_ptr(), None, None,
                                         out_mask.data_ptr(), case["H"], case["W"],
                                         torch.cuda.current_stream().cuda_stream)
        assert rc == 0
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out_idx, full[0]) and torch.equal(out_mask, full[3])


def test_argument_errors_follow_the_reference_op(elo, cuda):
    z = torch.zeros(1, 4, 8, 3, device=cuda)
    idx = torch.zeros(1, 2, 2, dtype=torch.int32, device=cuda)
    rhw = torch.arange(9, dtype=torch.int32, device=cuda)
    ok = dict(H=4, W=8, npoints=2, kernel_size_H=3, kernel_size_W=3, K=4, flag_copy=0, distance=1.0, stride_h=1, stride_w=1)
    elo.fused_conv_random_k(z, z, idx, rhw, **ok)
    for bad in (dict(K=0), dict(distance=0.0), dict(distance=-1.0), dict(flag_copy=-1), dict(stride_h=0),
                dict(kernel_size_H=0), dict(npoints=3), dict(kernel_size_W=5)):
        with pytest.raises(ValueError):
            elo.fused_conv_random_k(z, z, idx, rhw, **{**ok, **bad})
    with pytest.raises(ValueError):
        elo.fused_conv_select_k(z[..., :2], z, idx, rhw, **ok)
    with pytest.raises(ValueError):                      # xyz2 must have ceil(H/stride_h) rows
        elo.fused_conv_select_k(z, z, idx, rhw, **{**ok, "stride_h": 2})
