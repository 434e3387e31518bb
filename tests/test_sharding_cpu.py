"""CPU, world_size 2 over gloo: the multi-rank plumbing of the sweep (index-range partition of
the scan points, per-rank slices of the wavelength grid, final gather of the result maps)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_partition_properties(fpa):
    sh = fpa.sharding
    for n, world in ((1_000_000, 8), (1000, 3), (7, 8), (0, 4), (10, 1)):
        parts = [sh.shard_range(n, world, r) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))          # contiguous, disjoint
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n
    with pytest.raises(ValueError):
        sh.shard_range(10, 2, 2)
    lam1 = np.linspace(1545e-9, 1555e-9, 10)
    got = np.concatenate([sh.shard_axis(lam1, 4, r) for r in range(4)])
    assert np.array_equal(got, lam1)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry
    fpa = entry.load_package()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n1, n3 = 5, 4
        lam1 = np.linspace(1545e-9, 1555e-9, n1)
        mine = fpa.sharding.shard_axis(lam1, world, rank)
        a, b = fpa.sharding.shard_range(n1, world, rank)
        # stand-in for the per-rank sweep result: a deterministic function of the rank's rows
        local = torch.from_numpy(np.outer(mine * 1e9, np.arange(1, n3 + 1)).copy())
        full = fpa.sharding.gather_rows(local, n1, dist, world, rank)
        expect = np.outer(lam1 * 1e9, np.arange(1, n3 + 1))
        ok = full is not None and np.array_equal(full.numpy(), expect) and (b - a) == mine.size
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_over_gloo(fpa):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == {0: True, 1: True}
