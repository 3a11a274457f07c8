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
