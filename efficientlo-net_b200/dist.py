"""Multi-GPU plumbing: one process per GPU, frame pairs sharded across ranks.

Frame pairs are independent at inference, so the path is data-parallel over pairs with NO collective on
the data path (batch-shard first).  torch.distributed (NCCL on GPUs, gloo in the
CPU tests) is used only around it: to agree on the shard boundaries, to reduce timings with MAX, and to
gather the regressed poses of all shards in order for evaluation.
Row-band sharding of a single pair (one halo exchange per pyramid level) lives in rowband.py; DESIGN.md
section 7 says where it pays (128x2048) and where it does not (64x1800).
"""
import torch
import torch.distributed as dist


def shard_range(num_items, rank, world_size):
    """Contiguous, balanced [begin, end) of `num_items` for `rank`: the first (num_items % world) ranks
    get one extra item."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world_size")
    base, extra = divmod(num_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def max_over_ranks(value, device="cpu"):
    """MAX-reduce a python float over all ranks (timings are reported as the slowest rank's)."""
    if not dist.is_available() or not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_poses(q, t, num_items):
    """All ranks' (q (n_r,4), t (n_r,3)) of their shard_range -> full (num_items,4), (num_items,3) in item
    order on every rank.  Shards may differ by one item; they are padded for the all_gather."""
    if not dist.is_available() or not dist.is_initialized():
        return q, t
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(num_items, r, world) for r in range(world)]
    cap = max(e - b for b, e in sizes)
    pad = torch.zeros(cap, 7, dtype=q.dtype, device=q.device)
    n = q.shape[0]
    pad[:n, :4], pad[:n, 4:] = q, t
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    full = torch.cat([p[: e - b] for p, (b, e) in zip(parts, sizes)], 0)
    return full[:, :4].contiguous(), full[:, 4:].contiguous()
