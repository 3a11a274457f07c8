"""GPU: every fused block and the whole forward against the torch-CPU restatement (oracle/graph_oracle.py).

Neighbour sets are compared bit-exactly (the kernels can emit their selected cells), fp32 features within
1e-4 relative (+1e-5 absolute for values near zero), as BASELINE.json's north_star asks.  Each block is
fed the ORACLE's intermediates as inputs, so a block test does not depend on the blocks upstream."""
import numpy as np
import pytest
import torch

from oracle import graph_oracle as go

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 1e-5
# Index work with a tolerance (row a10): a point whose azimuth / elevation sits on a bin edge to the last bit lands in
# the adjacent cell when CUDA's atan2f / asinf and the host libm of the restatement round differently.  The counts
# below are what the B200 run of this test measured (printed by the tests, recorded in DESIGN.md section 5); the
# tests fail above them.
PROJ_MOVED_MAX = 2       # measured (r2): 2, an adjacent pair [2,54,263]/[2,54,264]; of 4 x 64 x 1800 cells (input projection, two frames of two samples)
CHAIN_ATOL = 1e-5        # intermediates of the full chain: the block tests' own band (measured need at rtol 1e-4: <= 5.1e-6, r2)
REPROJ_MOVED_MAX = 0     # measured (r2): 0 at every level; of 2 x h x w cells per refinement level
H_IN, W_IN, NPTS = 64, 1800, 150000


@pytest.fixture(params=[1, 0], ids=["tcgen05", "ffma"], autouse=True)
def mlp_engine(request, elo):
    """Every block test runs on both MLP engines: tensor cores (3xTF32, the default) and fp32 FFMA."""
    elo._lib.set_mlp_engine(request.param)
    yield request.param
    elo._lib.set_mlp_engine(1)


def close(got, want, what, rtol=RTOL, atol=ATOL):
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    assert got.shape == want.shape, "%s: shape %s vs %s" % (what, tuple(got.shape), tuple(want.shape))
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    if not bool((err <= tol).all()):
        i = int((err - tol).argmax())
        raise AssertionError("%s: max violation at flat %d: got %.8g want %.8g (|err| %.3g, %d of %d elements off)"
                             % (what, i, got.reshape(-1)[i], want.reshape(-1)[i], err.reshape(-1)[i],
                                int((err > tol).sum()), err.numel()))


def nbr_from_oracle(idx, mask, w2):
    """(B,n,K,3) [b,h,w] + mask -> linear cell or -1, the kernels' debug format."""
    lin = idx[..., 1] * w2 + idx[..., 2]
    return torch.where(mask[..., 0] > 0, lin, torch.full_like(lin, -1)).to(torch.int32)


@pytest.fixture(scope="module")
def world(elo, cuda):
    """One B=2 synthetic batch pushed through the oracle, with every intermediate kept."""
    P = elo.params.init_params(0)
    perms = elo.params.make_perms(0)
    pc, T = elo.synth.synth_batch(2, H_IN, W_IN, NPTS)
    eye = torch.eye(4).expand(2, 4, 4).contiguous()
    keep = {}
    out = go.get_model(pc, H_IN, W_IN, T, eye, eye, P, perms, keep=keep)
    store = elo.ParamStore(P, cuda)
    return dict(P=P, perms=perms, pc=pc, T=T, keep=keep, out=out, store=store, dev=cuda)


def test_input_projection_and_preprocess(elo, world):
    dev = world["dev"]
    pc = world["pc"].to(dev)
    with elo.use_store(world["store"]):
        xyz, _, _ = elo.model_util.project_points(pc[:, :NPTS, 0:3], None, H_IN, W_IN, mode=1, inner_batch=2,
                                                  outer_stride=NPTS * 6, batch_size=4)
    torch.cuda.synchronize()
    want = torch.cat([world["keep"]["xyz_f1_proj"], world["keep"]["xyz_f2_proj"]], 0)
    # bins are integers: the images must agree exactly -- including the cells at azimuth +-pi that receive the
    # empty points whose x is -0.0 (the synthetic scans contain them; model_util.py:419 keeps the sign)
    # The only tolerated difference: a point whose azimuth / elevation lies on a bin edge to the last bit, where
    # CUDA's atan2f / asinf and the host libm may round differently and the point moves to the ADJACENT cell.
    diff = (xyz.cpu() != want).any(-1)
    cells = diff.nonzero().tolist()
    print("input projection: %d of %d cells differ from the restatement: %s" % (len(cells), diff.numel(), cells))
    assert len(cells) <= PROJ_MOVED_MAX, "%d cells differ: %s" % (len(cells), cells[:8])
    for b, h, w in cells:
        assert any(c[0] == b and abs(c[1] - h) + abs(c[2] - w) == 1 for c in cells), "isolated differing cell %s" % [b, h, w]
    assert bool((world["pc"][:, :NPTS, 0] == 0).logical_and(torch.signbit(world["pc"][:, :NPTS, 0])).any())
    assert int((want != 0).any(-1).sum()) > 300000
    # plain ProjectPC2SphericalRing on the same points, API of model_util.py:181
    f1 = pc[:, :NPTS, 0:3].contiguous()
    xyz_plain, again = elo.ProjectPC2SphericalRing(f1, None, H_IN, W_IN)
    want_plain, _ = go.ProjectPC2SphericalRing(world["pc"][:, :NPTS, 0:3], None, H_IN, W_IN)
    assert again is xyz_plain
    nplain = int((xyz_plain.cpu() != want_plain).any(-1).sum())
    print("plain ProjectPC2SphericalRing: %d cells differ" % nplain)
    assert nplain <= PROJ_MOVED_MAX


@pytest.mark.parametrize("lvl", [2, 1, 0])
def test_warp_and_reprojection(elo, world, lvl):
    dev, keep = world["dev"], world["keep"]
    oh, ow = go.pyramid_shapes(H_IN, W_IN)[2:]
    h, w = oh[lvl + 2], ow[lvl + 2]
    xyz_l = gather_level_xyz(world, lvl, "f1")
    q = keep["l%d_q" % (lvl + 1)]
    t = keep["l%d_t" % (lvl + 1)]
    feat = keep["l%d_points_f1" % lvl]
    with elo.use_store(world["store"]):
        xyz_wp, pts_wp, warped = elo.model_util.project_points(xyz_l.reshape(2, -1, 3).to(dev), feat.to(dev), h, w,
                                                               mode=2, q=q.to(dev), t=t.to(dev), want_points=True)
    torch.cuda.synchronize()
    # the warp is written without FMA contraction in the reference's operation order: bit-exact
    assert torch.equal(warped.cpu(), keep["l%d_flow_warp" % lvl]), "warp differs"
    bad = (xyz_wp.cpu() != keep["l%d_xyz_warp_proj" % lvl]).any(-1)
    print("re-projection l%d: %d of %d cells differ" % (lvl, int(bad.sum()), bad.numel()))
    assert int(bad.sum()) <= REPROJ_MOVED_MAX
    ok = ~bad
    close(pts_wp.cpu()[ok], keep["l%d_points_warp_proj" % lvl][ok], "re-projected features l%d" % lvl)


def gather_level_xyz(world, lvl, frame):
    """Level-l xyz grid of the oracle run (strided slicing of the projected input image)."""
    x = world["keep"]["xyz_%s_proj" % frame]
    sh, sw, oh, ow = go.pyramid_shapes(H_IN, W_IN)
    for l in range(lvl + 1):
        x = x[:, ::sh[l + 2], ::sw[l + 2]][:, :oh[l + 2], :ow[l + 2]]
    return x.contiguous()


DOWN = [(32, (9, 15), 0.5, [8, 8, 16]), (32, (7, 11), 3.0, [16, 16, 32]), (16, (5, 9), 6.0, [32, 32, 64]),
        (16, (5, 9), 12.0, [64, 64, 128])]


@pytest.mark.parametrize("l", [0, 1, 2, 3])
def test_down_conv(elo, world, l):
    dev, keep, P, perms = world["dev"], world["keep"], world["P"], world["perms"]
    sh, sw, oh, ow = go.pyramid_shapes(H_IN, W_IN)
    xyz = keep["xyz_f1_proj"] if l == 0 else gather_level_xyz(world, l - 1, "f1")
    B, H, W, _ = xyz.shape
    pts = torch.zeros(B, H, W, 3) if l == 0 else keep["l%d_points_f1" % (l - 1)].reshape(B, H, W, -1)
    K, ks, dist, mlp = DOWN[l]
    sel = go.get_selected_idx(B, sh[l + 2], sw[l + 2], oh[l + 2], ow[l + 2])
    dbg_o = {}
    want, want_xyz = go.down_conv(xyz, pts, sel, K, ks, dist, ["sa1/layer%d/conv%d" % (l, j) for j in range(3)], P,
                                  perms["sa1/layer%d/f1" % l], debug=dbg_o)
    dbg = {}
    with elo.use_store(world["store"]), elo.variable_scope("sa1"):
        got, got_xyz = elo.down_conv(xyz.to(dev), pts.to(dev),
                                     elo.get_selected_idx(xyz.to(dev), sh[l + 2], sw[l + 2], oh[l + 2], ow[l + 2]),
                                     K, ks, dist, mlp, None, False, False, None, "layer%d" % l,
                                     random_hw=perms["sa1/layer%d/f1" % l], debug=dbg)
    torch.cuda.synchronize()
    assert torch.equal(dbg["nbr"].cpu(), nbr_from_oracle(dbg_o["idx"], dbg_o["mask"], W)), "neighbour sets differ"
    assert torch.equal(got_xyz.cpu(), want_xyz)
    close(got, want, "down_conv layer%d" % l)
    close(got, keep["l%d_points_f1" % l], "down_conv layer%d vs full oracle run" % l)
    # the explicit (B,oh,ow,3) index tensor of the reference API is accepted too
    if l == 3:
        with elo.use_store(world["store"]), elo.variable_scope("sa1"):
            again, _ = elo.down_conv(xyz.to(dev), pts.to(dev), sel.to(dev), K, ks, dist, mlp, None, False, False, None,
                                     "layer%d" % l, random_hw=perms["sa1/layer%d/f1" % l])
        assert torch.equal(again, got)


CV = {"origin": (2, (5, 35), 32, 4.0, "flow_embedding_l2_origin"), 2: (2, (5, 15), 6, 4.0, "flow_embedding_l2"),
      1: (1, (7, 25), 6, 2.0, "flow_embedding_l1"), 0: (0, (11, 41), 6, 1.0, "flow_embedding_l0")}


@pytest.mark.parametrize("which", ["origin", 2, 1, 0])
def test_cost_volume(elo, world, which):
    dev, keep, P, perms = world["dev"], world["keep"], world["P"], world["perms"]
    lvl, kq, nq, dist, scope = CV[which]
    oh, ow = go.pyramid_shapes(H_IN, W_IN)[2:]
    h, w = oh[lvl + 2], ow[lvl + 2]
    if which == "origin":
        xyz1, pts1 = gather_level_xyz(world, 2, "f1"), keep["l2_points_f1"].reshape(2, h, w, -1)
    else:
        xyz1, pts1 = keep["l%d_xyz_warp_proj" % lvl], keep["l%d_points_warp_proj" % lvl]
    xyz2, pts2 = gather_level_xyz(world, lvl, "f2"), keep["l%d_points_f2" % lvl].reshape(2, h, w, -1)
    dbg_o, dbg = {}, {}
    want = go.cost_volume(xyz1, xyz2, pts1, pts2, (3, 5), kq, 4, nq, dist, scope, P, perms[scope + "/q"],
                          perms[scope + "/p"], debug=dbg_o)
    with elo.use_store(world["store"]):
        got = elo.cost_volume(xyz1.to(dev), xyz2.to(dev), pts1.to(dev), pts2.to(dev), [3, 5], list(kq), 4, nq, dist,
                              [128, 64, 64], [128, 64], False, None, scope, random_hw_q=perms[scope + "/q"],
                              random_hw_p=perms[scope + "/p"], debug=dbg)
    torch.cuda.synchronize()
    assert torch.equal(dbg["nbr_q"].cpu(), nbr_from_oracle(dbg_o["idx_q"], dbg_o["mask_q"], w)), "select-K sets differ"
    assert torch.equal(dbg["nbr_p"].cpu(), nbr_from_oracle(dbg_o["idx_p"], dbg_o["mask_p"], w)), "random-K sets differ"
    close(dbg["stage1"], dbg_o["stage1"], "cost volume stage 1 (%s)" % scope)
    close(got, want, "cost volume (%s)" % scope)


@pytest.mark.parametrize("lvl", [2, 1, 0])
def test_up_conv_and_predictor(elo, world, lvl):
    dev, keep, P, perms = world["dev"], world["keep"], world["P"], world["perms"]
    sh, sw, oh, ow = go.pyramid_shapes(H_IN, W_IN)
    h, w = oh[lvl + 2], ow[lvl + 2]
    xyz1, pts1 = keep["l%d_xyz_warp_proj" % lvl], keep["l%d_points_warp_proj" % lvl]
    if lvl == 2:
        xyz2 = gather_level_xyz(world, 3, "f1")
        feat2 = keep["l3_points_f1_cost_volume"]
    else:
        xyz2 = keep["l%d_xyz_warp_proj" % (lvl + 1)]
        feat2 = keep["l%d_predict" % (lvl + 1)]
    feat2 = feat2.reshape(2, xyz2.shape[1], xyz2.shape[2], -1)
    scope = "up_sa_layer_layer_l%dcostvolume" % lvl
    up_dis = {2: 9.0, 1: 6.0, 0: 3.0}[lvl]
    dbg_o, dbg = {}, {}
    want = go.up_conv(xyz1, xyz2, pts1, feat2, (7, 15), sh[lvl + 3], sw[lvl + 3], 8, up_dis, scope, P, perms[scope],
                      debug=dbg_o)
    with elo.use_store(world["store"]):
        got = elo.up_conv(xyz1.to(dev), xyz2.to(dev), pts1.to(dev), feat2.to(dev), [7, 15], sh[lvl + 3], sw[lvl + 3], 8,
                          up_dis, [128, 64], [128, 64], False, scope, random_hw=perms[scope], debug=dbg)
        cv = keep["l%d_cost_volume" % lvl]
        pred = elo.flow_predictor(pts1.reshape(2, h * w, -1).to(dev), got, cv.to(dev), [128, 64], False, None,
                                  "l%d_costvolume_predict" % lvl)
    torch.cuda.synchronize()
    assert torch.equal(dbg["nbr"][0].cpu(), nbr_from_oracle(dbg_o["idx"], dbg_o["mask"], xyz2.shape[2]))
    close(got, want, "up_conv l%d" % lvl)
    close(got, keep["l%d_p_up" % lvl], "up_conv l%d vs full oracle run" % lvl)
    close(pred, go.flow_predictor(pts1.reshape(2, h * w, -1), want, cv, "l%d_costvolume_predict" % lvl, P),
          "flow_predictor l%d" % lvl)


def test_softmax_valid_and_level3_predictor(elo, world):
    dev, keep, P = world["dev"], world["keep"], world["P"]
    f = keep["l0_predict"]
    w = keep["l0_w"]
    valid = ~(keep["l0_xyz_warp_proj"].reshape(2, -1, 3) == 0).all(-1)
    want = go.softmax_valid(f, w, valid)
    with elo.use_store(world["store"]):
        got = elo.softmax_valid(f.to(dev), w.to(dev), valid.to(dev))
        l3w = elo.flow_predictor(keep["l3_points_f1"].to(dev), None, keep["l3_points_f1_cost_volume"].to(dev),
                                 [128, 64], False, None, "l3_costvolume_predict_ww")
    torch.cuda.synchronize()
    close(got, want, "softmax_valid")
    close(got, keep["l0_pooled"], "softmax_valid vs full oracle run")
    close(l3w, go.flow_predictor(keep["l3_points_f1"], None, keep["l3_points_f1_cost_volume"],
                                 "l3_costvolume_predict_ww", P), "l3 predictor")


def test_full_forward_matches_oracle(elo, world):
    """BASELINE.json configs[1]: full PWCLO forward, random-init weights, 64x1800 synthetic pairs."""
    dev = world["dev"]
    keep = {}
    out = elo.get_model(world["pc"].to(dev), H_IN, W_IN, world["T"].to(dev), None, None, False,
                        params=world["store"], perms=world["perms"], keep=keep)
    torch.cuda.synchronize()
    names = "l0_q l0_t l1_q l1_t l2_q l2_t l3_q l3_t l0_xyz_f1 q_gt t_gt".split()
    report = []
    for k in ("l0_points_f1", "l1_points_f2", "l2_points_f1", "l3_points_f2", "l2_points_f1_new",
              "l3_points_f1_cost_volume", "l3_q", "l3_t", "l2_cost_volume", "l2_predict", "l2_w", "l2_q", "l2_t",
              "l1_cost_volume", "l1_predict", "l1_q", "l1_t", "l0_cost_volume", "l0_predict", "l0_w", "l0_pooled"):
        g, w_ = keep[k].cpu().double(), world["keep"][k].double()
        err = (g - w_).abs()
        # absolute slack an element needs beyond 1e-4 relative: errors of the upstream blocks compound along the
        # chain (each block is held to 1e-4 / 1e-5 on its own inputs in the block tests)
        need = float((err - 1e-4 * w_.abs()).clamp_min(0).max())
        report.append("%s max|err| %.2e  abs slack needed at rtol 1e-4: %.2e  (max|x| %.2f)" % (k, err.max().item(), need, w_.abs().max().item()))
        close(keep[k], world["keep"][k], "intermediate " + k, rtol=1e-4, atol=CHAIN_ATOL)
    print("\n".join(report))
    for n, g, w in zip(names, out, world["out"]):
        close(g, w, "get_model output " + n, rtol=1e-4, atol=2e-5)
    # SE(3) pose within 1e-4 relative: translation norm and quaternion
    for lvl, (qi, ti) in enumerate(((0, 1), (2, 3), (4, 5), (6, 7))):
        q, t = out[qi].cpu().double(), out[ti].cpu().double()
        qw, tw = world["out"][qi].double(), world["out"][ti].double()
        assert float((t - tw).norm(dim=-1).max() / tw.norm(dim=-1).max()) < 1e-4 + 1e-4
        assert float((q - qw).norm(dim=-1).max()) < 1e-4


def test_dependent_launch_on_and_off_agree(elo, world):
    """Programmatic dependent launch only moves kernel prologues in time: same numbers with and without."""
    dev = world["dev"]
    outs = []
    try:
        for on in (0, 1):
            elo._lib.set_pdl(on)
            outs.append(elo.get_model(world["pc"].to(dev), H_IN, W_IN, world["T"].to(dev), None, None, False,
                                      params=world["store"], perms=world["perms"]))
            torch.cuda.synchronize()
    finally:
        elo._lib.set_pdl(1)
    for a, b in zip(*outs):
        assert torch.allclose(a, b, rtol=0, atol=1e-6)       # atomics in the re-projection may reorder float adds


def test_tile_policies_agree(elo, world):
    """elo_set_tile_policy only changes how queries are cut into 128-row tiles (full tiles vs spread over the SMs):
    a row's arithmetic does not depend on its neighbours in the tile, so the forward gives the same numbers."""
    dev = world["dev"]
    outs = []
    try:
        for policy in (0, 1):
            elo._lib.set_tile_policy(policy)
            outs.append(elo.get_model(world["pc"].to(dev), H_IN, W_IN, world["T"].to(dev), None, None, False,
                                      params=world["store"], perms=world["perms"]))
            torch.cuda.synchronize()
    finally:
        elo._lib.set_tile_policy(0)
    for a, b in zip(*outs):
        assert torch.allclose(a, b, rtol=0, atol=1e-6)       # atomics in the re-projection may reorder float adds


def test_engine_graph_replay_is_deterministic(elo, world):
    dev = world["dev"]
    eng = elo.PWCLOEngine(2, H_IN, W_IN, NPTS, params=world["store"], perms=world["perms"], device=dev).capture()
    pinned = world["pc"].pin_memory()
    q1, t1 = eng.infer(pinned, world["T"])
    q2, t2 = eng.infer(pinned, world["T"])
    assert torch.equal(q1, q2) and torch.equal(t1, t2)
    close(q1, world["out"][0], "engine q", atol=2e-5)
    close(t1, world["out"][1], "engine t", atol=2e-5)


def test_pipeline_with_forwards_in_flight_keeps_order_and_values(elo, world):
    """PWCLOPipeline(streams=3): three independent forwards overlap on three streams (own graph, input buffer and
    scratch each); every batch must come back in order with the values a lone engine computes for it."""
    dev = world["dev"]
    B = 1
    batches = [elo.synth.synth_batch(B, H_IN, W_IN, NPTS, seed0=50 + i) for i in range(7)]
    pinned = [(pc.pin_memory(), T.pin_memory()) for pc, T in batches]
    eng = elo.PWCLOEngine(B, H_IN, W_IN, NPTS, params=world["store"], perms=world["perms"], device=dev).capture()
    want = [tuple(x.clone() for x in eng.infer(pc, T)) for pc, T in pinned]
    pipe = elo.PWCLOPipeline(B, H_IN, W_IN, NPTS, params=world["store"], perms=world["perms"], device=dev, streams=3)
    for rep in range(2):            # the second pass reuses every slot
        got = [(q.clone(), t.clone()) for q, t in pipe.run(iter(pinned))]
        assert len(got) == len(want)
        for i, ((q, t), (qw, tw)) in enumerate(zip(got, want)):
            assert torch.allclose(q, qw, rtol=0, atol=1e-6) and torch.allclose(t, tw, rtol=0, atol=1e-6), (rep, i)
    # different inputs do give different poses (the comparison above is not vacuous)
    assert not torch.allclose(want[0][1], want[1][1], rtol=0, atol=1e-4)


def test_forward_128x2048_matches_oracle(elo, cuda, mlp_engine):
    """BASELINE.json configs[4] geometry: a dense 128x2048 scan (pyramid 32x256 / 16x128 / 8x64 / 8x32)."""
    if mlp_engine == 0:
        pytest.skip("one engine is enough for the geometry check")
    H, W, N = 128, 2048, 300000
    P = elo.params.init_params(1)
    perms = elo.params.make_perms(1)
    pc, T = elo.synth.synth_batch(1, H, W, N, seed0=7)
    eye = torch.eye(4)[None]
    want = go.get_model(pc, H, W, T, eye, eye, P, perms)
    got = elo.get_model(pc.to(cuda), H, W, T.to(cuda), None, None, False, params=elo.ParamStore(P, cuda), perms=perms)
    torch.cuda.synchronize()
    for n, g, w in zip("l0_q l0_t l1_q l1_t l2_q l2_t l3_q l3_t l0_xyz_f1 q_gt t_gt".split(), got, want):
        close(g, w, "128x2048 " + n, rtol=1e-4, atol=2e-5)
