"""CPU: the multi-process plumbing (shard ranges, MAX-reduced timings, ordered pose gather) on the gloo
backend with world_size 2 and 3 -- the N > 1 path of bench.py / evaluation without GPUs."""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, num_items, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = importlib.import_module("efficientlo-net_b200.dist")
    b, e = d.shard_range(num_items, rank, world)
    # every item's "pose" encodes its global index, so ordering errors are visible
    idx = torch.arange(b, e, dtype=torch.float32)
    q = torch.stack([idx, idx + 0.25, idx + 0.5, idx + 0.75], 1)
    t = torch.stack([-idx, -idx - 1, -idx - 2], 1)
    fq, ft = d.gather_poses(q, t, num_items)
    slowest = d.max_over_ranks(10.0 + rank)
    results[rank] = (b, e, fq.clone(), ft.clone(), slowest)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,num_items", [(2, 8), (2, 7), (3, 10)])
def test_sharding_gather_and_max_reduce(world, num_items):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), num_items, results), nprocs=world, join=True)
    covered = []
    for r in range(world):
        b, e, fq, ft, slowest = results[r]
        covered += list(range(b, e))
        want = torch.arange(num_items, dtype=torch.float32)
        assert torch.equal(fq[:, 0], want) and torch.equal(fq[:, 3], want + 0.75)
        assert torch.equal(ft[:, 2], -want - 2)
        assert slowest == 10.0 + world - 1
    assert covered == list(range(num_items))


def test_shard_range_properties():
    d = importlib.import_module("efficientlo-net_b200.dist")
    for n in (0, 1, 5, 32, 33):
        for w in (1, 2, 4, 8):
            spans = [d.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        d.shard_range(4, 2, 2)
