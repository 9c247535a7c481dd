"""Host-side logic of the row-sharded tables on CPU: partition arithmetic and the all-to-all routing,
exercised with the gloo backend at world_size 2 (and 3)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xsdeepfwfm_deprecated_b200.sharded import exchange_rows, local_rows, owner_and_local, route_indices


def test_partition_arithmetic():
    for rows in (1, 2, 7, 200, 201, 1000003):
        for world in (1, 2, 3, 4, 8):
            assert sum(local_rows(rows, r, world) for r in range(world)) == rows
            idx = torch.arange(min(rows, 5000))
            owner, local = owner_and_local(idx, world)
            assert torch.equal(local * world + owner, idx)
            for r in range(world):
                sel = local[owner == r]
                assert sel.numel() == 0 or int(sel.max()) < local_rows(rows, r, world)


def test_route_indices_is_a_stable_owner_sort():
    g = torch.Generator().manual_seed(0)
    idx = torch.randint(0, 1000, (257,), generator=g)
    s, counts, inv = route_indices(idx, 4)
    assert torch.equal(s[inv], idx)
    assert int(counts.sum()) == idx.numel()
    assert torch.equal(torch.sort(s % 4).values, s % 4)           # grouped by owner
    assert torch.equal(torch.bincount(idx % 4, minlength=4), counts)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(123)
        rows, K = 1013, 10
        table = torch.randn(rows, K, generator=g)                 # every rank can rebuild the full table (the checker)
        shard = table[rank::world].contiguous()                   # what this rank really owns
        assert shard.shape[0] == local_rows(rows, rank, world)
        gi = torch.Generator().manual_seed(1000 + rank)
        for n in (0, 1, 33, 500):
            idx = torch.randint(0, rows, (n,), generator=gi)

            def gather(req):
                assert bool(((req % world) == rank).all())        # only rows this rank owns are ever requested
                return shard[torch.div(req, world, rounding_mode="floor")]

            got = exchange_rows(idx, gather, K)
            assert torch.equal(got, table[idx])                   # rows are copied, never summed: bit-exact
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_all_to_all_row_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret.get(r) == "ok" for r in range(world))
