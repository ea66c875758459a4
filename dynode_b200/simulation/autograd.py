"""Differentiable, vmappable entry to the ensemble kernels.

In the reference, gradients reach NUTS by JAX reverse-mode AD through `diffeqsolve`
(RecursiveCheckpointAdjoint) and batching comes from `jax.vmap`.  Here the solve is a
`torch.autograd.Function` whose

  * forward makes ONE launch of the CUDA kernel for the whole batch, carrying forward sensitivities of the
    discrete scheme (`dynode_solve_sens_f64`) for exactly the inputs that require grad;
  * backward contracts the stored sensitivities with the incoming cotangent;
  * vmap rule folds the vmapped axis into the ensemble axis, so a per-draw model vmapped over 1024 chains is
    still one launch (the role of `vmap_method="broadcast_all"` in the XLA-FFI binding).

`PoissonLoglik` is the fused variant for the NUTS hot loop: solve + Poisson-incidence log-likelihood +
gradient in one launch, nothing but lp/grad written (`dynode_poisson_loglik_grad_f64`).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from .. import _lib, engine

_KINDS = ("beta", "gamma", "sigma", "omega", "season_amp", "season_phase")
_KIND_ID = {"beta": _lib.P_BETA, "gamma": _lib.P_GAMMA, "sigma": _lib.P_SIGMA, "omega": _lib.P_OMEGA,
            "season_amp": _lib.P_SEASON_AMP, "season_phase": _lib.P_SEASON_PHASE}


def unwrap(t):
    """Peel functorch wrappers (vmap / grad levels) off a tensor."""
    fc = torch._C._functorch
    while isinstance(t, torch.Tensor) and (fc.is_batchedtensor(t) or fc.is_gradtrackingtensor(t)):
        t = fc.get_unwrapped(t)
    return t


def is_transformed(t) -> bool:
    fc = torch._C._functorch
    return isinstance(t, torch.Tensor) and (fc.is_batchedtensor(t) or fc.is_gradtrackingtensor(t))


def needs_grad(t) -> bool:
    if not isinstance(t, torch.Tensor) or not torch.is_grad_enabled():
        return False
    if torch._C._functorch.is_gradtrackingtensor(t):
        return True
    return bool(unwrap(t).requires_grad)


@dataclass(frozen=True)
class SolveConfig:
    """Static (non-tensor) part of a solve: hashable, shared by forward / backward / the vmap rule."""

    model: engine.FlowModel
    opts_key: Tuple  # (t0, t1, rtol, atol, const_dt, max_steps, jump_ts)
    layout: Tuple[Tuple[str, int, int], ...]  # (kind, first column in theta, width)
    wrt_cols: Tuple[int, ...]  # theta columns that carry a tangent direction
    y0_grad: bool
    mask: int
    n_saved: int
    T: int
    # identity-compared payload (device tensors / arrays) lives outside the hash
    payload: "SolvePayload" = None

    def opts(self) -> engine.SolverOptions:
        t0, t1, rtol, atol, const_dt, max_steps, jumps = self.opts_key
        return engine.SolverOptions(t1=t1, t0=t0, rtol=rtol, atol=atol, const_dt=const_dt, max_steps=max_steps,
                                    jump_ts=jumps)

    def wrt_ids(self):
        ids = []
        for col in self.wrt_cols:
            for kind, first, width in self.layout:
                if first <= col < first + width:
                    strain = col - first if kind in ("beta", "gamma", "sigma", "omega") else 0
                    ids.append(_lib.wrt_id(_KIND_ID[kind], strain))
        return ids + [-1] * (self.model.state_size if self.y0_grad else 0)


class SolvePayload:
    """Shared device tensors of a solve (contact matrix, save grid, observations, period)."""

    def __init__(self, contact=None, save_ts=None, period=None, obs=None, obs_comp: int = -1,
                 lp_const: float = 0.0):
        self.contact, self.save_ts, self.period = contact, save_ts, period
        self.obs, self.obs_comp, self.lp_const = obs, obs_comp, lp_const

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other


def _kernel_params(cfg: SolveConfig, theta: torch.Tensor):
    p = {}
    for kind, first, width in cfg.layout:
        p[kind] = theta[:, first:first + width]  # a column block: passed as a strided DynodeArray, no copy
    if cfg.payload.period is not None:
        p["season_period"] = cfg.payload.period
    return p


def _seeds(cfg: SolveConfig, B: int, device):
    """dy0 [B][P][n]: zero for parameter directions, identity for the initial-state directions."""
    if not cfg.y0_grad:
        return None
    n, pw = cfg.model.state_size, len(cfg.wrt_cols)
    d = torch.zeros((B, pw + n, n), dtype=torch.float64, device=device)
    d[:, pw:, :] = torch.eye(n, dtype=torch.float64, device=device)
    return d


def _scatter_cols(g: torch.Tensor, cols: Tuple[int, ...], K: int) -> torch.Tensor:
    """[B, len(cols)] -> [B, K] with column j of g placed at cols[j] (no index tensors: graph-capturable)."""
    if cols == tuple(range(K)):
        return g[:, :K]
    out = g.new_zeros((g.shape[0], K))
    for j, c in enumerate(cols):
        out[:, c] = g[:, j]
    return out


def _fold(x: Optional[torch.Tensor], dim, B: int):
    """[.., vmapped axis at `dim`, ..] -> rows with the vmapped axis folded into the leading one."""
    if x is None:
        return None, 1
    if dim is None:
        return x, x.shape[0]
    x = x.movedim(dim, 0)
    inner = x.shape[1]
    return x.reshape(B * inner, *x.shape[2:]), inner


def _rows(y0: torch.Tensor, theta: torch.Tensor):
    """Broadcast the batch axes of y0 / theta against each other (a shared row stays shared)."""
    B = max(y0.shape[0], theta.shape[0])
    if theta.shape[0] != B:
        theta = theta.expand(B, theta.shape[1])
    return B, y0.contiguous(), theta.contiguous()


class EnsembleSolve(torch.autograd.Function):
    """(y0 [B|1, n], theta [B|1, K]) -> (ys [B, T, n_saved], stats [B, 4])."""

    @staticmethod
    def forward(y0, theta, cfg: SolveConfig):
        B, y0c, th = _rows(y0, theta)
        if cfg.y0_grad and y0c.shape[0] != B:
            y0c = y0c.expand(B, y0c.shape[1]).contiguous()
        wrt = cfg.wrt_ids()
        ys, dys, stats = engine.solve_ensemble(cfg.model, y0c, _kernel_params(cfg, th), cfg.payload.contact,
                                               cfg.opts(), cfg.payload.save_ts, cfg.mask, wrt=wrt,
                                               dy0=_seeds(cfg, B, th.device), B=B)
        if dys is None:
            dys = ys.new_empty((0,))
        return ys, stats, dys

    @staticmethod
    def setup_context(ctx, inputs, output):
        y0, theta, cfg = inputs
        ctx.cfg = cfg
        ctx.shapes = (y0.shape, theta.shape)
        ctx.save_for_backward(output[2])
        ctx.mark_non_differentiable(output[1], output[2])

    @staticmethod
    def backward(ctx, g_ys, _g_stats, _g_dys):
        cfg = ctx.cfg
        (dys,) = ctx.saved_tensors
        y0_shape, th_shape = ctx.shapes
        g_theta = g_y0 = None
        if dys.numel() > 0:
            # unreached slots hold +inf in ys and garbage in dys; a finite cotangent there is the caller's
            # business (NUTS sees a non-finite potential and rejects), so no masking here
            g = torch.einsum("btn,btnp->bp", g_ys, dys)
            pw = len(cfg.wrt_cols)
            if pw and ctx.needs_input_grad[1]:
                g_theta = _scatter_cols(g[:, :pw], cfg.wrt_cols, th_shape[1])
                if th_shape[0] == 1 and g.shape[0] > 1:
                    g_theta = g_theta.sum(0, keepdim=True)
            if cfg.y0_grad and ctx.needs_input_grad[0]:
                g_y0 = g[:, pw:]
                if y0_shape[0] == 1 and g.shape[0] > 1:
                    g_y0 = g_y0.sum(0, keepdim=True)
        return g_y0, g_theta, None

    @staticmethod
    def vmap(info, in_dims, y0, theta, cfg):
        B = info.batch_size
        y0f, i0 = _fold(y0, in_dims[0], B)
        thf, i1 = _fold(theta, in_dims[1], B)
        inner = max(i0, i1) if (in_dims[0] is not None or in_dims[1] is not None) else 1
        if in_dims[0] is not None and in_dims[1] is not None and i0 != i1:
            raise ValueError("vmapped y0 and theta must have the same inner batch size")
        if in_dims[0] is None and in_dims[1] is not None and y0f.shape[0] not in (1, B * inner):
            y0f = y0f.repeat(B, 1)
        if in_dims[1] is None and in_dims[0] is not None and thf.shape[0] not in (1, B * inner):
            thf = thf.repeat(B, 1)
        ys, stats, dys = EnsembleSolve.apply(y0f, thf, cfg)
        ys = ys.reshape(B, -1, *ys.shape[1:])
        stats = stats.reshape(B, -1, 4)
        dys = dys.reshape(B, -1, *dys.shape[1:]) if dys.numel() else dys.new_empty((B, 0))
        return (ys, stats, dys), (0, 0, 0)


def use_adjoint(model: engine.FlowModel, n_dir: int, opts: engine.SolverOptions, B: int = 1 << 30,
                n_rows: Optional[int] = None) -> bool:
    """Forward sensitivities or the discrete adjoint for the fused log-likelihood gradient?

    Measured on B200 (profiles/r1/adjoint_vs_forward.md), in units of one primal solve: the adjoint costs
    ~3.2 whatever the number of directions; forward mode ~1.45 per direction (4/5-element lanes, one direction
    per group) or ~1.9 per pair of directions (SIR lanes).  But the forward groups are independent work items of
    one launch, so while they do not fill the GPU (few NUTS chains) their latency is that of ONE group and they
    win up to ~8 groups.  DYNODE_B200_ADJOINT=0/1 forces a choice."""
    import os
    force = os.environ.get("DYNODE_B200_ADJOINT")
    if force is not None:
        return force == "1" and n_dir > 0
    if n_dir == 0:
        return False
    # the adjoint checkpoints every accepted step: [B][cap][n + 2] doubles of scratch
    if B * adjoint_capacity() * (model.state_size + 2) * 8 > ADJOINT_SCRATCH_LIMIT:
        return False
    chunk = 2 if model.flow == _lib.FLOW_SIR else 1
    groups = -(-n_dir // chunk)
    slots = max(1, 32 // (model.n_groups * model.n_strains))
    # rows a row mask leaves out retire at once: what counts is the rows that run (a NUTS run spends most of its
    # rounds on the few chains that are still building deep trees)
    rows = engine.rows_to_integrate(B) if n_rows is None else max(1, min(int(n_rows), B))
    warps = -(-rows // slots) * groups
    if warps <= RESIDENT_WARPS:  # latency regime
        return groups > 8
    return groups * (1.9 if chunk == 2 else 1.45) > ADJOINT_COST


ADJOINT_COST = 3.2      # primal-solve equivalents of one adjoint evaluation (forward + reverse sweep)
RESIDENT_WARPS = 1184   # 148 SMs x 8 warps of the 255-register tangent / adjoint kernels
ADJOINT_SCRATCH_LIMIT = 16 << 30  # bytes of checkpoint scratch beyond which forward mode is used instead


def adjoint_capacity() -> int:
    import os
    return int(os.environ.get("DYNODE_B200_ADJOINT_CAP", "512"))


_ADJ_COLS = {}


def _adjoint_columns(cfg: SolveConfig, device) -> torch.Tensor:
    """Columns of the adjoint's [4*S + 2] gradient row that correspond to cfg.wrt_cols (cached on device)."""
    key = (device.index, cfg.layout, cfg.wrt_cols, cfg.model.n_strains)
    t = _ADJ_COLS.get(key)
    if t is None:
        S = cfg.model.n_strains
        base = {"beta": 0, "gamma": S, "sigma": 2 * S, "omega": 3 * S, "season_amp": 4 * S, "season_phase": 4 * S + 1}
        cols = []
        for col in cfg.wrt_cols:
            for kind, first, width in cfg.layout:
                if first <= col < first + width:
                    cols.append(base[kind] + (col - first))
        t = torch.tensor(cols, dtype=torch.long, device=device)
        _ADJ_COLS[key] = t
    return t


class PoissonLoglik(torch.autograd.Function):
    """(y0 [B|1, n], theta [B|1, K]) -> (lp [B], stats [B, 4]); gradient from the same launch."""

    @staticmethod
    def forward(y0, theta, cfg: SolveConfig):
        B, y0c, th = _rows(y0, theta)
        if cfg.y0_grad and y0c.shape[0] != B:
            y0c = y0c.expand(B, y0c.shape[1]).contiguous()
        pl = cfg.payload
        n_dir = len(cfg.wrt_cols) + (cfg.model.state_size if cfg.y0_grad else 0)
        if use_adjoint(cfg.model, n_dir, cfg.opts(), B):
            # one reverse sweep gives d lp / d (every rate, y0); pick the columns that were asked for
            lp, g_all, g_y0, stats = engine.poisson_loglik_adjoint(
                cfg.model, y0c, _kernel_params(cfg, th), pl.contact, cfg.opts(), pl.save_ts, pl.obs_comp, pl.obs,
                pl.lp_const, with_y0_grad=cfg.y0_grad, B=B, cap=adjoint_capacity())
            grad = g_all.index_select(1, _adjoint_columns(cfg, th.device))
            if cfg.y0_grad:
                grad = torch.cat([grad, g_y0], dim=1)
            # Rows that accepted more steps than the checkpoint scratch holds come back NaN with result code
            # RESULT_ADJOINT_CAPACITY: they are re-evaluated right here by forward sensitivities, in a launch masked
            # to exactly those rows (no host sync: when nothing overflowed every warp scans its mask bytes and exits),
            # so neither NUTS nor SVI ever sees the capacity of a scratch buffer as a NaN log-density.
            over = stats[:, _lib.STAT_RESULT] == _lib.RESULT_ADJOINT_CAPACITY
            with engine.only_rows(over.view(torch.uint8)):
                lp_f, grad_f, stats_f = engine.poisson_loglik_grad(
                    cfg.model, y0c, _kernel_params(cfg, th), pl.contact, cfg.opts(), pl.save_ts, pl.obs_comp, pl.obs,
                    pl.lp_const, wrt=cfg.wrt_ids(), dy0=_seeds(cfg, B, th.device), B=B)
            lp = torch.where(over, lp_f, lp)
            grad = torch.where(over[:, None], grad_f, grad)
            stats = torch.where(over[:, None], stats_f, stats)
            return lp, stats, grad
        lp, grad, stats = engine.poisson_loglik_grad(cfg.model, y0c, _kernel_params(cfg, th), pl.contact,
                                                     cfg.opts(), pl.save_ts, pl.obs_comp, pl.obs, pl.lp_const,
                                                     wrt=cfg.wrt_ids(), dy0=_seeds(cfg, B, th.device), B=B)
        if grad is None:
            grad = lp.new_empty((0,))
        return lp, stats, grad

    @staticmethod
    def setup_context(ctx, inputs, output):
        y0, theta, cfg = inputs
        ctx.cfg = cfg
        ctx.shapes = (y0.shape, theta.shape)
        ctx.save_for_backward(output[2])
        ctx.mark_non_differentiable(output[1], output[2])
        ctx.set_materialize_grads(False)  # no zero-filled [B, 4] / [B, P] tensors for stats / grad every backward

    @staticmethod
    def backward(ctx, g_lp, _g_stats, _g_grad):
        cfg = ctx.cfg
        (grad,) = ctx.saved_tensors
        y0_shape, th_shape = ctx.shapes
        g_theta = g_y0 = None
        if g_lp is not None and grad.numel() > 0:
            g = g_lp[:, None] * grad
            pw = len(cfg.wrt_cols)
            if pw and ctx.needs_input_grad[1]:
                g_theta = _scatter_cols(g[:, :pw], cfg.wrt_cols, th_shape[1])
                if th_shape[0] == 1 and g.shape[0] > 1:
                    g_theta = g_theta.sum(0, keepdim=True)
            if cfg.y0_grad and ctx.needs_input_grad[0]:
                g_y0 = g[:, pw:]
                if y0_shape[0] == 1 and g.shape[0] > 1:
                    g_y0 = g_y0.sum(0, keepdim=True)
        return g_y0, g_theta, None

    @staticmethod
    def vmap(info, in_dims, y0, theta, cfg):
        B = info.batch_size
        y0f, i0 = _fold(y0, in_dims[0], B)
        thf, i1 = _fold(theta, in_dims[1], B)
        inner = max(i0, i1) if (in_dims[0] is not None or in_dims[1] is not None) else 1
        if in_dims[0] is None and in_dims[1] is not None and y0f.shape[0] not in (1, B * inner):
            y0f = y0f.repeat(B, 1)
        if in_dims[1] is None and in_dims[0] is not None and thf.shape[0] not in (1, B * inner):
            thf = thf.repeat(B, 1)
        lp, stats, grad = PoissonLoglik.apply(y0f, thf, cfg)
        lp = lp.reshape(B, -1)
        stats = stats.reshape(B, -1, 4)
        grad = grad.reshape(B, -1, grad.shape[-1]) if grad.numel() else grad.new_empty((B, 0))
        return (lp, stats, grad), (0, 0, 0)


_CONST_CACHE = {}


def const_to_device(t: torch.Tensor, device) -> torch.Tensor:
    """Device copy of a small constant host tensor, cached by content.

    Models rebuild their constants on every call (`config.initializer.get_initial_state()`, python-float
    rates, the contact matrix).  Caching the device copies keeps the per-call path free of host-to-device
    copies, which is what lets a whole model evaluation be captured in a CUDA graph."""
    if t.device == device:
        return t if t.dtype == torch.float64 else t.to(torch.float64)
    if is_transformed(t) or t.requires_grad or t.device.type != "cpu" or t.numel() > 65536:
        return t.to(device=device, dtype=torch.float64)
    tc = t.detach().to(torch.float64).contiguous()
    key = (device.index, tuple(tc.shape), tc.numpy().tobytes())
    hit = _CONST_CACHE.get(key)
    if hit is None:
        if len(_CONST_CACHE) > 4096:
            _CONST_CACHE.clear()
        hit = tc.to(device)
        _CONST_CACHE[key] = hit
    return hit


def pack_inputs(model: engine.FlowModel, params: dict, y0: torch.Tensor, batched: bool, device):
    """Kernel parameter dict -> (y0 [B|1, n], theta [B|1, K], layout, wrt_cols, y0_grad, period).

    `params[kind]` are per-draw tensors ([S] / scalar, or [B, S] when `batched`); functorch-wrapped and
    grad-requiring tensors pass through torch.cat, so vmap and autograd see ordinary differentiable ops."""
    S = model.n_strains
    cols, layout, wrt_cols, first = [], [], [], 0
    for kind in _KINDS:
        v = params.get(kind)
        if v is None:
            continue
        width = S if kind in ("beta", "gamma", "sigma", "omega") else 1
        v = const_to_device(v, device)
        v = v.reshape(-1, width) if batched else v.reshape(1, width) if v.numel() == width else v.reshape(-1, width)
        cols.append(v)
        layout.append((kind, first, width))
        if needs_grad(v):
            wrt_cols.extend(range(first, first + width))
        first += width
    B = max(c.shape[0] for c in cols)
    cols = [c if c.shape[0] == B else c.expand(B, c.shape[1]) for c in cols]
    theta = torch.cat(cols, dim=1)
    period = params.get("season_period")
    if period is not None:
        period = const_to_device(unwrap(period).detach(), device).reshape(-1, 1).contiguous()
    y0 = const_to_device(y0, device)
    y0 = y0.reshape(-1, model.state_size)
    return y0, theta, tuple(layout), tuple(wrt_cols), needs_grad(y0), period


def opts_key(opts: engine.SolverOptions) -> Tuple:
    return (float(opts.t0), float(opts.t1), float(opts.rtol), float(opts.atol), float(opts.const_dt),
            int(opts.max_steps), tuple(opts.jump_ts))


_TS_CACHE = {}


def device_grid(save_ts: np.ndarray, opts: engine.SolverOptions, device) -> torch.Tensor:
    """Device copy of a host save grid, tagged with the uniform-grid hint the kernel uses."""
    key = (device.index, len(save_ts), float(save_ts[0]), float(save_ts[-1]))
    t = _TS_CACHE.get(key)
    if t is None or not np.array_equal(t.host, save_ts):
        t = torch.as_tensor(save_ts, dtype=torch.float64, device=device)
        t.host = np.array(save_ts, copy=True)
        t.dynode_save_dt = engine.uniform_save_dt(t.host, float(opts.t0), float(opts.t1))
        _TS_CACHE[key] = t
    return t
