"""CPU: row-band partitioning and the halo exchange on gloo (world size 2, 3 and 4, including the
all-gather fallback for halos taller than a band)."""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def rb():
    return importlib.import_module("efficientlo-net_b200.rowband")


def test_band_properties():
    m = rb()
    for rows, align in ((64, 4), (16, 1), (8, 1), (128, 4), (4, 1), (32, 2)):
        for world in (1, 2, 3, 4, 8):
            spans = [m.band(rows, r, world, align) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == rows
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(a % align == 0 and b % align == 0 for a, b in spans)
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= align
    with pytest.raises(ValueError):
        m.band(10, 0, 2, align=4)
    assert m.halo_rows(0, 8, 64, 4) == (0, 4)
    assert m.halo_rows(56, 64, 64, 4) == (4, 0)
    assert m.halo_rows(8, 16, 64, 4) == (4, 4)
    assert m.halo_rows(2, 4, 16, 6) == (2, 6)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _image(rows, W=5, C=3, B=2):
    return torch.arange(B * rows * W * C, dtype=torch.float32).view(B, rows, W, C)


def _worker(rank, world, port, rows, halo, align, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = rb()
    full = _image(rows)
    r0, r1 = m.band(rows, rank, world, align)
    got, top = m.exchange_halo(full[:, r0:r1].contiguous(), rows, halo, rank, world, align=align)
    want, top_w = m.local_halo(full, rows, halo, rank, world, align)
    results[rank] = (bool(torch.equal(got, want)), top, top_w, tuple(got.shape))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,rows,halo,align", [(2, 64, 4, 4), (3, 64, 4, 4), (2, 16, 6, 1), (4, 16, 6, 1), (4, 8, 3, 1)])
def test_halo_exchange_matches_a_cut_of_the_full_image(world, rows, halo, align):
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), rows, halo, align, results), nprocs=world, join=True)
    m = rb()
    for r in range(world):
        ok, top, top_w, shape = results[r]
        r0, r1 = m.band(rows, r, world, align)
        assert ok, "rank %d received the wrong halo" % r
        assert top == top_w == min(halo, r0)
        assert shape[1] == (r1 - r0) + min(halo, r0) + min(halo, rows - r1)
