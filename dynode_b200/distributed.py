"""Multi-GPU layer: ensembles and chains shard over ranks with NO collective on the solve path.

One process per GPU (`torch.distributed`, backend nccl; gloo in the CPU tests).  Trajectories are
independent, so rank g integrates the contiguous block [lo_g, hi_g) of the batch axis.
Draws come from a counter-based generator keyed by the *global* draw index, so results do not depend
on the number of ranks.  The only collective is an optional all-gather of the saved trajectories /
posterior-predictive draws at the end: the kernel stores straight into the rank's slice of the gather
buffer and the all-gather runs in place over NVLink (nccl), so no staging copy precedes it.
"""

from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(B: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of rank `rank`: sizes differ by at most one, larger blocks first."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    q, r = divmod(int(B), world_size)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def shard_counts(B: int, world_size: int) -> Sequence[int]:
    return [shard_bounds(B, world_size, r)[1] - shard_bounds(B, world_size, r)[0] for r in range(world_size)]


_BLOCK = 4096


def ensemble_uniform(seed: int, lo: int, hi: int, n_cols: int) -> np.ndarray:
    """U(0,1) draws for global ensemble rows [lo, hi): row b is a function of (seed, b) only.

    Rows are generated in fixed global blocks of 4096 (Philox keyed by (seed, block)), then sliced, so
    any partition of the batch axis over ranks reproduces the same ensemble."""
    out = np.empty((max(hi - lo, 0), n_cols))
    b = lo
    while b < hi:
        blk = b // _BLOCK
        start = blk * _BLOCK
        gen = np.random.Generator(np.random.Philox(key=[int(seed) & (2**64 - 1), blk]))
        u = gen.random((_BLOCK, n_cols))
        take_hi = min(hi, start + _BLOCK)
        out[b - lo:take_hi - lo] = u[b - start:take_hi - start]
        b = take_hi
    return out


class GatherBuffer:
    """[B_total, *row] device buffer whose rows [lo, hi) belong to this rank.

    `local` is the view the kernel writes into; `all_gather()` fills the other ranks' rows in place.
    """

    def __init__(self, B_total: int, row_shape: Sequence[int], dtype=torch.float64, device=None, group=None):
        self.group = group
        self.rank, self.world = world()
        self.B = int(B_total)
        self.counts = list(shard_counts(self.B, self.world))
        self.lo, self.hi = shard_bounds(self.B, self.world, self.rank)
        # equal shards -> one in-place all_gather_into_tensor; ragged -> rows padded to the largest shard
        self.equal = len(set(self.counts)) == 1
        self.rows_padded = max(self.counts) * self.world
        self.full = torch.empty((self.B if self.equal else self.rows_padded, *row_shape), dtype=dtype,
                                device=device)
        if self.equal:
            self.local = self.full[self.lo:self.hi]
        else:
            m = max(self.counts)
            self.local = self.full[self.rank * m:self.rank * m + self.counts[self.rank]]

    def all_gather(self) -> torch.Tensor:
        """Returns the [B_total, *row] tensor with every rank's rows (collective; all ranks call it)."""
        if self.world == 1:
            return self.full[: self.B]
        m = max(self.counts)
        mine = self.full[self.rank * m:(self.rank + 1) * m]
        if self.full.is_cuda:
            dist.all_gather_into_tensor(self.full, mine, group=self.group)  # in place (NCCL semantics)
        else:
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(parts, mine.clone(), group=self.group)
            for r, p in enumerate(parts):
                self.full[r * m:(r + 1) * m].copy_(p)
        if self.equal:
            return self.full
        keep = torch.cat([torch.arange(r * m, r * m + c, device=self.full.device)
                          for r, c in enumerate(self.counts)])
        return self.full.index_select(0, keep)


def run_sharded(local_solve: Callable[[int, int, torch.Tensor], None], B_total: int, row_shape: Sequence[int],
                *, gather: bool = True, dtype=torch.float64, device=None, group=None) -> torch.Tensor:
    """Integrate this rank's block with `local_solve(lo, hi, out_rows)` and optionally all-gather.

    Returns the local rows (gather=False) or the whole [B_total, *row] ensemble on every rank."""
    buf = GatherBuffer(B_total, row_shape, dtype=dtype, device=device, group=group)
    if buf.hi > buf.lo:
        local_solve(buf.lo, buf.hi, buf.local)
    return buf.all_gather() if gather else buf.local


def gather_draws(local: dict, group=None) -> dict:
    """All-gather a dict of per-rank draws (posterior samples / posterior-predictive trajectories,
    {site: [n_local, ...]}) along the leading axis; every rank may hold a different number of rows.
    Uses the same padded in-place buffer as the trajectory gather (NCCL all_gather_into_tensor on CUDA)."""
    rank, ws = world()
    if ws == 1:
        return dict(local)
    out = {}
    for name in sorted(local):
        x = local[name].contiguous()
        n = torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device)
        counts = [torch.zeros_like(n) for _ in range(ws)]
        dist.all_gather(counts, n, group=group)
        counts = [int(c.item()) for c in counts]
        m = max(counts)
        full = torch.zeros((ws * m, *x.shape[1:]), dtype=x.dtype, device=x.device)
        full[rank * m:rank * m + x.shape[0]] = x
        mine = full[rank * m:(rank + 1) * m]
        if x.is_cuda:
            dist.all_gather_into_tensor(full, mine, group=group)
        else:
            parts = [torch.empty_like(mine) for _ in range(ws)]
            dist.all_gather(parts, mine.clone(), group=group)
            full = torch.cat(parts, 0)
        out[name] = torch.cat([full[r * m:r * m + c] for r, c in enumerate(counts)], 0)
    return out


def all_reduce_flags(n_failed: int, device=None, group=None) -> int:
    """Sum of per-rank failure counts (e.g. trajectories that hit max_steps)."""
    rank, ws = world()
    if ws == 1:
        return int(n_failed)
    t = torch.tensor([int(n_failed)], dtype=torch.int64, device=device)
    dist.all_reduce(t, group=group)
    return int(t.item())


def simulate_ensemble_sharded(ode, duration_days, initial_state, ode_parameters, solver_parameters,
                              sub_save_indices=None, save_step: int = 1, *, batch_size: int,
                              gather: bool = True, throw: bool = True):
    """`simulate_ensemble` over all ranks: every rank passes the SAME global inputs (device tensors with
    `batch_size` rows, or shared rows), solves its block in one launch and, when `gather`, receives the
    whole ensemble's saved trajectories [B, T, n_saved] through one NCCL all-gather.

    Returns (ys_flat, stats_local, (lo, hi))."""
    import dataclasses

    from .simulation import odes

    rank, ws = world()
    lo, hi = shard_bounds(batch_size, ws, rank)

    def cut(x):
        if isinstance(x, torch.Tensor) and x.ndim >= 1 and x.shape[0] == batch_size:
            return x[lo:hi]
        if dataclasses.is_dataclass(x) and not isinstance(x, type):
            return dataclasses.replace(x, **{f.name: cut(getattr(x, f.name)) for f in dataclasses.fields(x)})
        return x

    state_batched = all(c.ndim >= 2 and c.shape[0] == batch_size for c in initial_state)
    state_l = tuple(cut(c) if state_batched else c for c in initial_state)
    prm_l = cut(ode_parameters)
    dev = initial_state[0].device
    spec, model, _, _ = odes._resolve(ode, state_l, prm_l, hi - lo, state_batched)
    saveat = odes.build_saveat(0.0, duration_days, save_step, sub_save_indices)
    mask = odes._mask_from(saveat.indices, model.n_compartments)
    T, ns = len(saveat.times), model.saved_size(mask)
    stats_box = {}

    def local_solve(lo_, hi_, out_rows):
        sol = odes.simulate_ensemble(ode, duration_days, state_l, prm_l, solver_parameters, sub_save_indices,
                                     save_step, batch_size=hi_ - lo_, state_batched=state_batched, throw=False,
                                     out=out_rows)
        stats_box["result"] = sol.result

    ys = run_sharded(local_solve, batch_size, (T, ns), gather=gather, device=dev)
    failed = int((stats_box["result"] != 0).sum().item()) if "result" in stats_box else 0
    if throw and all_reduce_flags(failed, device=dev) > 0:
        raise RuntimeError(odes.MAX_STEPS_MESSAGE)
    return ys, stats_box.get("result"), (lo, hi)
