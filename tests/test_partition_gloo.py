"""Multi-process (world_size 2 and 3, gloo, CPU) test of the column-block partitioning used for
N > 1 GPUs: the blocks tile the column axis exactly once, and the output gather restores column order.
The per-rank "kernel" here is a stand-in (a deterministic function of the column values): the data
path has no collective, so what needs testing across processes is only the sharding + gather."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xarray_parcel_b200 import partition


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, L, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        t = torch.rand((L, n), generator=g)
        p = torch.linspace(1000, 100, L)
        (p_loc, t_loc), (start, stop) = partition.shard_columns([p, t], rank, world)
        assert p_loc.shape == (L,) and t_loc.shape == (L, stop - start)
        assert t_loc.data_ptr() == t[:, start:].data_ptr()          # a view, no copy
        local = torch.stack([t_loc.sum(0), t_loc[0] * 2.0])          # "per-column outputs" [2, n_local]
        full = partition.gather_columns(local, n)
        expect = torch.stack([t.sum(0), t[0] * 2.0])
        ok = torch.equal(full, expect)
        counts = torch.zeros(n, dtype=torch.int64)
        counts[start:stop] += 1
        dist.all_reduce(counts)
        ok = ok and bool((counts == 1).all())
        if rank == 0:
            ret.put(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 1000), (2, 31), (3, 4097)])
def test_column_blocks_and_gather(world, n):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, 7, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=10) is True


def test_column_block_properties():
    for n in (0, 1, 31, 32, 33, 1000, 3_114_720):
        for world in (1, 2, 4, 8):
            blocks = [partition.column_block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
                assert a1 == b0 and a0 <= a1
            if world > 1:
                assert all(b[0] % 32 == 0 for b in blocks if b[0] < n)
