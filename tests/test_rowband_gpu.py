"""GPU: the banded blocks reproduce the single-GPU result bit for bit -- neighbour sets and features of every
owned row -- for 2, 4 and 8 bands.  The ranks are simulated one after the other on one device (the halo is cut
out of the full image with local_halo, which tests/test_rowband_cpu.py shows to be what exchange_halo delivers)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
H_IN, W_IN, NPTS = 64, 1800, 150000


@pytest.fixture(scope="module")
def scene(elo, cuda):
    P = elo.params.init_params(0)
    perms = elo.params.make_perms(0)
    store = elo.ParamStore(P, cuda)
    pc, T = elo.synth.synth_batch(1, H_IN, W_IN, NPTS)
    keep = {}
    elo.get_model(pc.to(cuda), H_IN, W_IN, T.to(cuda), None, None, False, params=store, perms=perms, keep=keep)
    torch.cuda.synchronize()
    return dict(store=store, perms=perms, keep=keep, dev=cuda)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_set_conv_layer0_by_row_bands(elo, scene, world):
    rb, pu = elo.rowband, elo.pointnet_util
    store, perms, dev = scene["store"], scene["perms"], scene["dev"]
    xyz = scene["keep"]["xyz_f1_proj"]                                   # (1, 64, 1800, 3)
    scopes = ["sa1/layer0/conv%d" % j for j in range(3)]
    sel = pu.SelectedIdx(1, 4, 8, 16, 225, dev)
    dbg = {}
    with elo.use_store(store):
        full = pu.set_conv(xyz, None, sel, 32, (9, 15), 0.5, scopes, store, [perms["sa1/layer0/f1"]], feat_channels=3,
                           debug=dbg).view(1, 16, 225, -1)
        full_nbr = dbg["nbr"].view(1, 16, 225, -1)
        for rank in range(world):
            r0, r1 = rb.band(H_IN, rank, world, align=4)
            sub, top = rb.local_halo(xyz, H_IN, 4, rank, world, align=4)
            out, nbr = rb.set_conv_band(sub, None, top, r1 - r0, 4, 8, 225, 32, (9, 15), 0.5, scopes, store,
                                        perms["sa1/layer0/f1"], feat_channels=3, want_nbr=True)
            want_nbr = full_nbr[:, r0 // 4:r1 // 4]
            want_nbr = torch.where(want_nbr >= 0, want_nbr - r0 * W_IN, want_nbr)
            assert torch.equal(nbr, want_nbr), "neighbour sets of band %d differ" % rank
            assert torch.equal(out, full[:, r0 // 4:r1 // 4]), "features of band %d differ" % rank


@pytest.mark.parametrize("world", [2, 4])
def test_cost_volume_level0_by_row_bands(elo, scene, world):
    rb, pu = elo.rowband, elo.pointnet_util
    store, perms, keep = scene["store"], scene["perms"], scene["keep"]
    h, w = 16, 225
    xyz1 = keep["l0_xyz_warp_proj"]
    f1 = keep["l0_points_warp_proj"]
    xyz2 = keep["xyz_f2_proj"][:, ::4, ::8][:, :h, :w].contiguous()
    f2 = keep["l0_points_f2"].view(1, h, w, -1)
    args = dict(kernel_size1=[3, 5], kernel_size2=[11, 41], nsample=4, nsample_q=6, distance=1.0,
                scope="flow_embedding_l0", random_hw_q=perms["flow_embedding_l0/q"],
                random_hw_p=perms["flow_embedding_l0/p"], store=store)
    full = rb.cost_volume_band(xyz1, xyz2, f1, f2, 0, h, **args)
    assert torch.allclose(full.reshape(1, h * w, -1), keep["l0_cost_volume"], rtol=0, atol=0)
    halo = 3 // 2 + 11 // 2
    for rank in range(world):
        r0, r1 = rb.band(h, rank, world)
        subs = [rb.local_halo(t, h, halo, rank, world) for t in (xyz1, xyz2, f1, f2)]
        top = subs[0][1]
        out = rb.cost_volume_band(subs[0][0], subs[1][0], subs[2][0], subs[3][0], top, r1 - r0, **args)
        assert torch.equal(out, full[:, r0:r1]), "cost volume of band %d differs" % rank


# ---- the whole forward by row bands (pwclo_model.RowBand): ranks emulated by threads on one device --------------------
class ThreadBand:
    """RowBand whose all-gather is played by threads of one process: every 'rank' is a thread running get_model on the
    same device with its own parameter store; gather_rows hands the owned rows to the other threads."""

    def __init__(self, elo, rank, world, shared, barrier, skip=()):
        self.inner = elo.RowBand(rank, world, skip=skip, min_points=0)       # band every level whose rows divide
        self.rank, self.world, self.shared, self.barrier = rank, world, shared, barrier
        self.exchanges = 0

    def rows(self, h, tag=None, w=None):
        return self.inner.rows(h, tag, w)

    def gather_rows(self, tensors, h, w):
        self.shared[self.rank] = tensors
        torch.cuda.synchronize()
        self.barrier.wait()
        per = h // self.world
        for r in range(self.world):
            if r == self.rank:
                continue
            for mine, theirs in zip(tensors, self.shared[r]):
                mine[:, r * per * w:(r + 1) * per * w].copy_(theirs[:, r * per * w:(r + 1) * per * w])
        torch.cuda.synchronize()
        self.barrier.wait()
        self.exchanges += 1


@pytest.mark.parametrize("world", [2, 4])
def test_forward_by_row_bands_equals_single_gpu(elo, cuda, world):
    """Every rank of a banded forward ends with the poses of the single-GPU forward (bands of layer 0 and of every
    level whose rows divide over the ranks; the other levels are computed by all)."""
    import threading
    P = elo.params.init_params(0)
    perms = elo.params.make_perms(0)
    pc, T = elo.synth.synth_batch(1, H_IN, W_IN, NPTS, seed0=3)
    pc, T = pc.to(cuda), T.to(cuda)
    want = elo.get_model(pc, H_IN, W_IN, T, None, None, False, params=elo.ParamStore(P, cuda), perms=perms)
    torch.cuda.synchronize()
    shared, barrier = [None] * world, threading.Barrier(world)
    results, errors = [None] * world, []

    def run(rank):
        try:
            torch.cuda.set_device(cuda)
            band = ThreadBand(elo, rank, world, shared, barrier)
            out = elo.get_model(pc, H_IN, W_IN, T, None, None, False, params=elo.ParamStore(P, cuda), perms=perms,
                                band=band)
            torch.cuda.synchronize()
            results[rank] = (out, band.exchanges)
        except Exception as e:          # a dead thread must not leave the others at the barrier forever
            errors.append(e)
            barrier.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    oh, _ = elo.pwclo_model.pyramid_shapes(H_IN, W_IN)
    banded_levels = sum(1 for lvl in (0, 1, 2) if oh[lvl + 2] % world == 0 and oh[lvl + 2] // world >= 2)
    for rank in range(world):
        out, exchanges = results[rank]
        assert exchanges == 1 + banded_levels          # layer 0 + one per banded level
        for name, g, w_ in zip("l0_q l0_t l1_q l1_t l2_q l2_t l3_q l3_t l0_xyz_f1 q_gt t_gt".split(), out, want):
            # same arithmetic per row; only the float atomics of the re-projection may reorder sums
            assert torch.allclose(g, w_, rtol=0, atol=1e-6), "rank %d: %s differs by %.3g" % (
                rank, name, float((g - w_).abs().max()))


@pytest.mark.parametrize("parts", [2, 3])
def test_query_ranges_write_the_same_rows(elo, scene, parts):
    """The C ABI's query_begin / query_end: a block run range by range writes, into its full-size outputs, exactly the
    rows the whole call writes (neighbour tables bit for bit, features bit for bit)."""
    pu = elo.pointnet_util
    store, perms, keep, dev = scene["store"], scene["perms"], scene["keep"], scene["dev"]
    h, w = 16, 225
    n = h * w
    cuts = [n * i // parts for i in range(parts + 1)]
    xyz1, f1 = keep["l0_xyz_warp_proj"], keep["l0_points_warp_proj"]
    xyz2 = keep["xyz_f2_proj"][:, ::4, ::8][:, :h, :w].contiguous()
    f2 = keep["l0_points_f2"].view(1, h, w, -1)

    def search(qr):
        return pu.multi_search([pu.search_spec(True, xyz1, xyz2, (h, w, 1, 1), (11, 41), 6, 1000.0, 1, 1,
                                               perms["flow_embedding_l0/q"], qrange=qr),
                                pu.search_spec(False, xyz1, xyz1, (h, w, 1, 1), (3, 5), 4, 1.0, 1, 1,
                                               perms["flow_embedding_l0/p"], qrange=qr)])

    def cv(nq, np_, q1, q2):
        with elo.use_store(store):
            return pu.cost_volume(xyz1, xyz2, f1, f2, [3, 5], [11, 41], 4, 6, 1.0, [128, 64, 64], [128, 64], False, None,
                                  "flow_embedding_l0", random_hw_q=perms["flow_embedding_l0/q"],
                                  random_hw_p=perms["flow_embedding_l0/p"], nbr_q=nq, nbr_p=np_, qrange1=q1, qrange2=q2)

    full_q, full_p = search(None)
    full_cv = cv(full_q, full_p, None, None)
    torch.cuda.synchronize()
    for a, b in zip(cuts[:-1], cuts[1:]):
        # stage 1 must cover the rows stage 2's 3x5 window reaches into: one image row either side
        a1, b1 = max(a - w, 0), min(b + w, n)
        nq, np_ = search((a1, b1))
        assert torch.equal(nq[:, a1:b1], full_q[:, a1:b1]) and torch.equal(np_[:, a1:b1], full_p[:, a1:b1])
        part = cv(nq, np_, (a1, b1), (a, b))
        torch.cuda.synchronize()
        assert torch.equal(part[:, a:b], full_cv[:, a:b]), "cost volume rows [%d, %d) differ" % (a, b)
