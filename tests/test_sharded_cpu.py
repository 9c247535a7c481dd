"""Host-side logic of the row-sharded tables on CPU: partition arithmetic and the all-to-all routing,
exercised with the gloo backend at world_size 2 (and 3)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xsdeepfwfm_deprecated_b200.sharded import exchange_rows, local_rows, owner_and_local, pull_plan, route_indices


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


@pytest.mark.parametrize("world,c", [(2, 1), (4, 1), (8, 4), (3, 7)])
def test_pull_plan_stages_the_rows_the_unsharded_lookup_reads(world, c):
    """The p2p_pull exchange: pulling shard[owner][local] into staging row b and looking the rewritten index up in the
    staging table (QR: quotient row * remainder row) reads exactly what the unsharded table gives for the original id."""
    g = torch.Generator().manual_seed(world * 10 + c)
    n, K, B = 1003, 10, 257
    nq = -(-n // c)
    Wq = torch.randn(nq, K, generator=g)
    Wr = torch.randn(c, K, generator=g) if c > 1 else None
    shards = [Wq[r::world].contiguous() for r in range(world)]                 # row i on rank i mod P at local row i div P
    assert [s.shape[0] for s in shards] == [local_rows(nq, r, world) for r in range(world)]
    idx = torch.randint(0, n, (B,), generator=g)
    owner, local, rewritten = pull_plan(idx, c, world)
    staged = torch.stack([shards[int(o)][int(l)] for o, l in zip(owner, local)])      # what dfw_pull_rows copies
    assert torch.equal(staged, Wq[idx // c])
    # the fused kernel's view: a (B * c)-category table whose quotient rows are `staged`
    q2, r2 = rewritten // c, rewritten % c
    assert torch.equal(q2, torch.arange(B)) and torch.equal(r2, idx % c) and int(rewritten.max()) < B * c
    want = Wq[idx // c] * Wr[idx % c] if c > 1 else Wq[idx]
    got = staged[q2] * Wr[r2] if c > 1 else staged[q2]
    assert torch.equal(got, want)


def test_exchange_names():
    from xsdeepfwfm_deprecated_b200.sharded import ShardedDeepFMs
    with pytest.raises(ValueError):
        ShardedDeepFMs(39, [1] * 13 + [10] * 26, exchange="smoke_signals", use_cuda=False)
    for ex in ("p2p", "p2p_pull", "nccl"):
        assert ShardedDeepFMs(39, [1] * 13 + [10] * 26, exchange=ex, use_cuda=False, use_fwlw=True).exchange == ex


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


def test_index_rewriting_exchanges_refuse_first_order_tables():
    """'nccl' / 'p2p_pull' rewrite the sharded index columns, which the first-order tables (use_fwlw=0) are
    looked up with -- refused at construction; 'p2p' keeps the category ids and is allowed."""
    import pytest
    from xsdeepfwfm_deprecated_b200.sharded import ShardedDeepFMs
    sizes = [1] * 13 + [50] * 26
    kw = dict(use_fm=False, use_fwfm=True, use_deep=True, use_cuda=False, deep_nodes=16)
    for ex in ("nccl", "p2p_pull"):
        with pytest.raises(ValueError, match="use_fwlw"):
            ShardedDeepFMs(39, sizes, exchange=ex, use_fwlw=False, **kw)
        ShardedDeepFMs(39, sizes, exchange=ex, use_fwlw=True, **kw)
    ShardedDeepFMs(39, sizes, exchange="p2p", use_fwlw=False, **kw)
