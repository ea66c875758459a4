"""Multi-rank host logic on CPU (gloo, world_size 2): sharding, counter-based draws, the gather buffer.
The solve itself needs a GPU; here a deterministic stand-in fills each rank's rows so the plumbing
(`dynode_b200.distributed`) is what is tested."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dynode_b200 import distributed as D


def test_shard_bounds_cover_the_batch_exactly():
    for B in (0, 1, 7, 8, 100_000, 100_003):
        for ws in (1, 2, 3, 4, 8):
            edges = [D.shard_bounds(B, ws, r) for r in range(ws)]
            assert edges[0][0] == 0 and edges[-1][1] == B
            assert all(edges[r][1] == edges[r + 1][0] for r in range(ws - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
            assert list(D.shard_counts(B, ws)) == sizes
    with pytest.raises(ValueError):
        D.shard_bounds(10, 2, 2)


def test_draws_do_not_depend_on_the_partition():
    full = D.ensemble_uniform(20260101, 0, 10_000, 5)
    assert full.shape == (10_000, 5) and 0.0 <= full.min() and full.max() < 1.0
    for ws in (2, 3, 8):
        parts = [D.ensemble_uniform(20260101, *D.shard_bounds(10_000, ws, r), 5) for r in range(ws)]
        assert np.array_equal(np.concatenate(parts), full)
    assert not np.array_equal(D.ensemble_uniform(1, 0, 64, 5), full[:64])
    assert D.ensemble_uniform(3, 5, 5, 2).shape == (0, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rows(lo, hi, T, ns):
    b = torch.arange(lo, hi, dtype=torch.float64)[:, None, None]
    t = torch.arange(T, dtype=torch.float64)[None, :, None]
    k = torch.arange(ns, dtype=torch.float64)[None, None, :]
    return b * 1000.0 + t + k / 16.0


def _worker(rank, ws, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        T, ns = 5, 3
        calls = []

        def local_solve(lo, hi, out):
            calls.append((lo, hi))
            out.copy_(_rows(lo, hi, T, ns))

        full = D.run_sharded(local_solve, B, (T, ns), gather=True)
        ok = full.shape == (B, T, ns) and torch.equal(full, _rows(0, B, T, ns))
        ok &= calls == [D.shard_bounds(B, ws, rank)]
        local = D.run_sharded(local_solve, B, (T, ns), gather=False)
        lo, hi = D.shard_bounds(B, ws, rank)
        ok &= torch.equal(local, _rows(lo, hi, T, ns))
        ok &= D.all_reduce_flags(rank + 1) == ws * (ws + 1) // 2
        # posterior draws / posterior-predictive rows, a different number per rank
        mine = {"r0": torch.full((3 + rank,), float(rank)), "inc": torch.full((3 + rank, 4, 2), float(rank))}
        got = D.gather_draws(mine)
        ok &= got["r0"].shape == (7,) and got["inc"].shape == (7, 4, 2)
        ok &= torch.equal(got["r0"], torch.tensor([0.0] * 3 + [1.0] * 4))
        ok &= bool((got["inc"][:3] == 0).all() and (got["inc"][3:] == 1).all())
        ok &= D.world() == (rank, ws)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 9])  # equal shards (in-place path) and ragged shards (padded path)
def test_gather_of_saved_trajectories_world_size_2(B):
    ws = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, B, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(ws))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def test_single_process_world_is_identity():
    assert D.world() == (0, 1)
    out = D.run_sharded(lambda lo, hi, o: o.copy_(_rows(lo, hi, 4, 2)), 6, (4, 2))
    assert torch.equal(out, _rows(0, 6, 4, 2))
    assert D.all_reduce_flags(3) == 3
