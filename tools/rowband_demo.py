"""Row-band sharding of one frame pair over N GPUs with an NCCL halo exchange per level (north_star / SURVEY 8(e)).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/rowband_demo.py [--hw 64x1800] [--iters 50]

Every rank owns a band of rows of (a) the input range image -> set-conv layer 0, (b) level 0 of the pyramid ->
attentive cost volume.  Per block: ONE halo exchange of the inputs (batched NCCL send/recv with the two
neighbours), then the unchanged fused kernels on the (halo + band + halo) sub-image.  Rank 0 checks the gathered
bands against its own single-GPU result (must be bit-identical) and prints one JSON line with the device-timed
step (max over ranks) next to the single-GPU time of the same two blocks."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import elo_b200 as elo  # noqa: E402


def timed(fn, iters, dev):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(dev)
    if dist.is_initialized():
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / iters, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hw", default="64x1800")
    ap.add_argument("--iters", type=int, default=50)
    a = ap.parse_args()
    H, W = (int(v) for v in a.hw.split("x"))
    npts = 150000 if H * W <= 150000 else 300000
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rb, pu = elo.rowband, elo.pointnet_util
    store = elo.ParamStore(elo.params.init_params(0), dev)
    perms = elo.params.make_perms(0)
    pc, T = elo.synth.synth_batch(1, H, W, npts)
    keep = {}
    elo.get_model(pc.to(dev), H, W, T.to(dev), None, None, False, params=store, perms=perms, keep=keep)
    oh, ow = elo.pwclo_model.pyramid_shapes(H, W)
    h0, w0 = oh[2], ow[2]
    xyz_in = keep["xyz_f1_proj"]
    xyz1, f1 = keep["l0_xyz_warp_proj"], keep["l0_points_warp_proj"]
    xyz2 = keep["xyz_f2_proj"][:, ::4, ::8][:, :h0, :w0].contiguous()
    f2 = keep["l0_points_f2"].view(1, h0, w0, -1)
    scopes = ["sa1/layer0/conv%d" % j for j in range(3)]
    kq = elo.pwclo_model.CV_KERNEL_Q[0]
    cv = dict(kernel_size1=[3, 5], kernel_size2=list(kq), nsample=4, nsample_q=6, distance=1.0, scope="flow_embedding_l0",
              random_hw_q=perms["flow_embedding_l0/q"], random_hw_p=perms["flow_embedding_l0/p"], store=store)
    halo_cv = 3 // 2 + kq[0] // 2

    def single():
        with elo.use_store(store):
            a_ = rb.set_conv_band(xyz_in, None, 0, H, 4, 8, w0, 32, (9, 15), 0.5, scopes, store, perms["sa1/layer0/f1"],
                                  feat_channels=3)
            b_ = rb.cost_volume_band(xyz1, xyz2, f1, f2, 0, h0, **cv)
        return a_, b_

    # this rank's bands (what it would hold if the stages before were banded too)
    r0, r1 = rb.band(H, rank, world, align=4)
    c0, c1 = rb.band(h0, rank, world)
    mine_in = xyz_in[:, r0:r1].contiguous()
    mine_cv = torch.cat([xyz1, xyz2, f1, f2], -1)[:, c0:c1].contiguous()       # one slab -> one exchange
    C = f1.shape[-1]

    def banded():
        with elo.use_store(store):
            sub, top = rb.exchange_halo(mine_in, H, 4, rank, world, align=4)
            a_ = rb.set_conv_band(sub, None, top, r1 - r0, 4, 8, w0, 32, (9, 15), 0.5, scopes, store,
                                  perms["sa1/layer0/f1"], feat_channels=3)
            slab, top2 = rb.exchange_halo(mine_cv, h0, halo_cv, rank, world)
            parts = [t.contiguous() for t in slab.split([3, 3, C, C], -1)]
            b_ = rb.cost_volume_band(parts[0], parts[1], parts[2], parts[3], top2, c1 - c0, **cv)
        return a_, b_

    t_band, (a_band, b_band) = timed(banded, a.iters, dev)
    if world > 1:
        t_band = elo.dist.max_over_ranks(t_band, device=dev)
    ok = None
    t_single = None
    if world > 1:
        ga = [torch.empty_like(a_band) for _ in range(world)] if all(
            rb.band(H, r, world, 4)[1] - rb.band(H, r, world, 4)[0] == r1 - r0 for r in range(world)) else None
        gb = [torch.empty_like(b_band) for _ in range(world)] if all(
            rb.band(h0, r, world)[1] - rb.band(h0, r, world)[0] == c1 - c0 for r in range(world)) else None
        if ga is not None and gb is not None:
            dist.all_gather(ga, a_band.contiguous())
            dist.all_gather(gb, b_band.contiguous())
    if rank == 0:
        t_single, (a_full, b_full) = timed(single, a.iters, dev) if world == 1 else (None, (None, None))
    if world > 1:
        # the single-GPU reference is timed with the other ranks idle at a barrier
        if rank == 0:
            for _ in range(3):
                single()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                a_full, b_full = single()
            e1.record()
            torch.cuda.synchronize(dev)
            t_single = e0.elapsed_time(e1) / a.iters
            if ga is not None and gb is not None:
                ok = bool(torch.equal(torch.cat(ga, 1), a_full)) and bool(torch.equal(torch.cat(gb, 1), b_full))
        dist.barrier()
    if rank == 0:
        print(json.dumps({"what": "row-band sharding of one frame pair: set-conv layer 0 + cost volume level 0",
                          "hw": a.hw, "n_gpus": world, "ms_banded_max_over_ranks": t_band, "ms_single_gpu": t_single,
                          "bit_identical_to_single_gpu": ok, "halo_rows": {"set_conv_l0": 4, "cost_volume_l0": halo_cv},
                          "exchanges_per_block": 1, "backend": "nccl" if world > 1 else None}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
