"""Row-band sharding of ONE frame pair across the GPUs of a box (BASELINE.json north_star).

A range image is cut into horizontal bands; rank r owns query rows [r0, r1) of a pyramid level.  A
projection-aware block only looks kH//2 rows above and below a query (the window never wraps vertically,
fused_conv_g.cu:96-99; it is full-width, so the horizontal handling stays local), hence one exchange of
`halo` rows with the two neighbouring ranks per level makes the band self-sufficient:

    set-conv layer l      halo = kH//2 rows of (xyz | feat) of the searched grid
    cost volume           halo = kH_p//2 + kH_q//2: stage 2 reads stage-1 embeddings of kH_p//2 neighbouring
                          rows, each of which looked kH_q//2 rows further -- those embeddings are recomputed
                          locally instead of being exchanged, so the level needs ONE exchange, of its inputs

The exchange is one batched NCCL send/recv pair per neighbour (torch.distributed P2P ops, NVLink); when the
halo is taller than a band (the deep, short levels) it degenerates to an all-gather of the level, as
SURVEY.md 8(e) prescribes.  The blocks then run unchanged on the (halo + band + halo) sub-image and the rows
outside the band are dropped: neighbour sets and features of the owned rows are bit-identical to the
single-GPU result (tests/test_rowband_*.py).

This module is the partitioning + exchange layer and the two banded blocks.  bench.py does NOT use it: at
64x1800 the levels have 16 / 8 / 4 / 4 rows, a whole forward is ~0.5 ms of dependent 10-40 us kernels, and
every exchange adds a collective's latency to that chain -- sharding frame pairs over the GPUs (dist.py)
is the faster way to use the box (DESIGN.md section 7 has the numbers).
"""
import torch
import torch.distributed as dist

from . import pointnet_util as pu


def band(rows, rank, world, align=1):
    """Rows [r0, r1) of a `rows`-high image owned by `rank`: contiguous, in units of `align` rows (the
    query stride of the level, so that strided centres stay on the grid), sizes differing by <= align."""
    if rows % align:
        raise ValueError("rows %d is not a multiple of align %d" % (rows, align))
    units = rows // align
    base, extra = divmod(units, world)
    u0 = rank * base + min(rank, extra)
    u1 = u0 + base + (1 if rank < extra else 0)
    return u0 * align, u1 * align


def halo_rows(r0, r1, rows, halo):
    """How many halo rows exist above / below the band inside the image."""
    return min(halo, r0), min(halo, rows - r1)


def exchange_halo(x_band, rows, halo, rank, world, group=None, align=1):
    """x_band: (B, r1 - r0, W, C) rows owned by this rank.  Returns ((B, top + (r1-r0) + bot, W, C), top)
    with the `halo` rows above and below fetched from the neighbouring ranks (fewer at the image border).
    Every rank must call it (collective)."""
    r0, r1 = band(rows, rank, world, align)
    assert x_band.shape[1] == r1 - r0, "band has %d rows, expected %d" % (x_band.shape[1], r1 - r0)
    top, bot = halo_rows(r0, r1, rows, halo)
    if world == 1 or halo == 0:
        return x_band, 0
    heights = [band(rows, r, world, align) for r in range(world)]
    if halo > min(b - a for a, b in heights):
        # a halo taller than a band would need several hops: gather the whole level instead
        sizes = [b - a for a, b in heights]
        hmax = max(sizes)
        pad = x_band if x_band.shape[1] == hmax else torch.cat(
            [x_band, x_band.new_zeros((x_band.shape[0], hmax - x_band.shape[1]) + tuple(x_band.shape[2:]))], 1)
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad.contiguous(), group=group)
        full = torch.cat([p[:, :s] for p, s in zip(parts, sizes)], 1)
        return full[:, r0 - top:r1 + bot].contiguous(), top
    x_band = x_band.contiguous()
    up = x_band.new_empty((x_band.shape[0], top) + tuple(x_band.shape[2:]))
    down = x_band.new_empty((x_band.shape[0], bot) + tuple(x_band.shape[2:]))
    ops, keep = [], []
    if rank > 0:
        # the rank above needs my first rows as its bottom halo; I need its last rows as my top halo
        need = halo_rows(*heights[rank - 1], rows, halo)[1]
        keep.append(x_band[:, :need].contiguous())
        ops.append(dist.P2POp(dist.isend, keep[-1], _peer(rank - 1, group), group))
        ops.append(dist.P2POp(dist.irecv, up, _peer(rank - 1, group), group))
    if rank < world - 1:
        need = halo_rows(*heights[rank + 1], rows, halo)[0]
        keep.append(x_band[:, x_band.shape[1] - need:].contiguous())
        ops.append(dist.P2POp(dist.isend, keep[-1], _peer(rank + 1, group), group))
        ops.append(dist.P2POp(dist.irecv, down, _peer(rank + 1, group), group))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    return torch.cat([up, x_band, down], 1), top


def _peer(rank_in_group, group):
    return rank_in_group if group is None else dist.get_global_rank(group, rank_in_group)


def local_halo(x_full, rows, halo, rank, world, align=1):
    """Single-process stand-in for exchange_halo (tests, and world == 1): cut band + halo out of the full image."""
    r0, r1 = band(rows, rank, world, align)
    top, bot = halo_rows(r0, r1, rows, halo)
    return x_full[:, r0 - top:r1 + bot].contiguous(), top


# ---------------------------------------------------------------------------------------------------------------
def set_conv_band(xyz_sub, feat_sub, top, own_rows, stride_h, stride_w, out_w, K_sample, kernel_size, distance,
                  layer_scopes, store, random_hw, feat_channels=None, want_nbr=False):
    """Set-conv (utils/pointnet_util.py:179-250) for the query rows of one band.

    xyz_sub / feat_sub: the (halo + band + halo) sub-image, `top` its halo rows above the band, `own_rows` the
    band height in pixels (a multiple of stride_h).  The sub-image's first row must sit on the centre grid
    (top % stride_h == 0; guaranteed when bands are aligned to the stride and halo is a multiple of it, else
    pass a taller halo).  Returns (own_rows/stride_h * out_w, C_out) features of the owned centres, plus their
    neighbour table in GLOBAL-row-independent form (row offsets relative to the band start) if want_nbr."""
    if top % stride_h or own_rows % stride_h:
        raise ValueError("band / halo not aligned to the centre stride")
    B, Hs = xyz_sub.shape[0], xyz_sub.shape[1]
    q_rows = (Hs + stride_h - 1) // stride_h
    sel = pu.SelectedIdx(B, stride_h, stride_w, q_rows, out_w, xyz_sub.device)
    dbg = {} if want_nbr else None
    out = pu.set_conv(xyz_sub, feat_sub, sel, K_sample, kernel_size, distance, layer_scopes, store, [random_hw],
                      feat_channels=feat_channels, debug=dbg)
    q0, q1 = top // stride_h, (top + own_rows) // stride_h
    out = out.view(B, q_rows, out_w, -1)[:, q0:q1]
    if not want_nbr:
        return out
    W = xyz_sub.shape[2]
    nbr = dbg["nbr"].view(B, q_rows, out_w, -1)[:, q0:q1]
    nbr = torch.where(nbr >= 0, nbr - top * W, nbr)            # cell index relative to the first owned row
    return out, nbr


def cost_volume_band(xyz1_sub, xyz2_sub, f1_sub, f2_sub, top, own_rows, kernel_size1, kernel_size2, nsample,
                     nsample_q, distance, scope, random_hw_q, random_hw_p, store):
    """Attentive cost volume (utils/pointnet_util.py:33-149) for the rows of one band; all four inputs are
    (halo + band + halo) sub-images with halo >= kernel_size1[0]//2 + kernel_size2[0]//2 (or the image border)."""
    B, Hs, W, _ = xyz1_sub.shape
    out = pu.cost_volume(xyz1_sub, xyz2_sub, f1_sub, f2_sub, kernel_size1, kernel_size2, nsample, nsample_q,
                         distance, [128, 64, 64], [128, 64], False, None, scope, random_hw_q=random_hw_q,
                         random_hw_p=random_hw_p, params=store)
    return out.view(B, Hs, W, -1)[:, top:top + own_rows]
