"""N>1 host logic on CPU: world_size-2 (and 3) gloo processes exercising the shard partition and the OFE all-gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from autotune.b200.epdf import gather_slices, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 100000, 100003):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, out_q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_total, rank, world)
        # stand-in for the per-rank kernel output: value = global repetition id (what keys the Philox streams)
        local = torch.arange(lo, hi, dtype=torch.float32)
        full = gather_slices(local, n_total)
        out_q.put((rank, full.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n_total', [(2, 10), (2, 11), (3, 10)])
def test_gather_slices_over_gloo(world, n_total):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, full in got:
        assert full == [float(i) for i in range(n_total)], (rank, full)


def test_gather_slices_single_process_checks_length():
    assert gather_slices(torch.arange(5.0), 5).tolist() == [0, 1, 2, 3, 4]
    with pytest.raises(ValueError):
        gather_slices(torch.arange(4.0), 5)
