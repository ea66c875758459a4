"""Low-level ensemble engine: torch CUDA tensors in, one native launch, torch CUDA tensors out.

This is the device-resident layer under `dynode_b200.simulation.simulate`: it only validates
shapes, allocates outputs and forwards raw device pointers to the C ABI
(include/dynode_b200.h).  All arithmetic happens in the sm_100a kernels.
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

from . import _lib
from ._lib import DynodeError

_PARAM_FIELDS = ("beta", "gamma", "sigma", "omega", "season_amp", "season_phase", "season_period")


@dataclass(frozen=True)
class FlowModel:
    """A member of the compiled flow family (csrc/instances.def)."""

    flow: int
    flags: int
    n_groups: int
    n_strains: int

    def desc(self) -> _lib.ModelDesc:
        return _lib.ModelDesc(self.flow, self.flags, self.n_groups, self.n_strains)

    @property
    def n_compartments(self) -> int:
        return {0: 3, 1: 4, 2: 5}[self.flow]

    @property
    def state_size(self) -> int:
        return self.n_groups + (self.n_compartments - 1) * self.n_groups * self.n_strains

    def compartment_sizes(self) -> Tuple[int, ...]:
        gs = self.n_groups * self.n_strains
        return (self.n_groups,) + (gs,) * (self.n_compartments - 1)

    def saved_size(self, mask: int) -> int:
        return sum(sz for c, sz in enumerate(self.compartment_sizes()) if (mask >> c) & 1)

    def full_mask(self) -> int:
        return (1 << self.n_compartments) - 1

    def check_supported(self) -> None:
        d = self.desc()
        if not _lib.load().dynode_is_supported(ctypes.byref(d)):
            raise DynodeError("unsupported ODE: " + _lib.last_error())


@dataclass
class SolverOptions:
    """diffeqsolve arguments fixed by reference odes.py:107-144 / config/params.py:24-67."""

    t1: float
    t0: float = 0.0
    rtol: float = 1e-5
    atol: float = 1e-6
    const_dt: float = 0.0
    max_steps: int = 10**6
    jump_ts: Tuple[float, ...] = ()  # SolverParams.discontinuity_points

    def desc(self, save_dt: float = 0.0) -> _lib.SolverDesc:
        d = _lib.SolverDesc(float(self.t0), float(self.t1), float(self.rtol), float(self.atol),
                            float(self.const_dt), int(self.max_steps), float(save_dt), None, 0)
        if len(self.jump_ts) > 0:
            dev_t = _jump_tensor(tuple(sorted(float(x) for x in self.jump_ts)))
            d.jump_ts, d.n_jump = dev_t.data_ptr(), int(dev_t.numel())
            self._keep = dev_t  # keeps the device copy alive across the launch
        return d


_JUMP_CACHE: Dict[tuple, object] = {}

# ---- row mask (DynodeSolverDesc.only) -----------------------------------------------------------------
# per host thread: two samplers driven from two threads of one process must not see each other's masks
import threading


class _OnlyState(threading.local):
    mask = None
    n_rows = None  # how many rows of the mask are set, when the caller knows (only_rows(mask, n_rows=...))


_ONLY_TLS = _OnlyState()


class _OnlyProxy:
    """`_ONLY[0]` as the code below uses it, backed by thread-local storage."""

    def __getitem__(self, i):
        return _ONLY_TLS.mask

    def __setitem__(self, i, v):
        _ONLY_TLS.mask = v


_ONLY = _OnlyProxy()


class only_rows:
    """Context: ensemble launches made inside it integrate only the rows b with mask[b] != 0 (uint8/bool CUDA
    tensor of the ensemble size; launches of another size ignore it).  Rows left out cost nothing and come
    back as zeros.  `n_rows` (optional) is an upper bound of the number of rows set, see `rows_to_integrate`.  The many-chain NUTS wraps the model evaluation of a round in it with its "chain still
    running" flags; inside a CUDA-graph capture the mask's address is what gets recorded, so the flags may
    change between replays."""

    def __init__(self, mask, n_rows: Optional[int] = None):
        self.mask, self.n_rows, self.prev = mask, n_rows, None

    def __enter__(self):
        self.prev = (_ONLY[0], _ONLY_TLS.n_rows)
        _ONLY[0], _ONLY_TLS.n_rows = self.mask, self.n_rows
        return self

    def __exit__(self, *exc):
        _ONLY[0], _ONLY_TLS.n_rows = self.prev
        return False


def rows_to_integrate(B: int) -> int:
    """B, or the caller's count of the rows an active `only_rows` mask leaves in.  The mask itself lives on the device;
    this is host knowledge (the NUTS driver reads its active-chain count at its sync points), used where a launch
    is CHOSEN -- forward sensitivities against the adjoint depends on how many warps will really run."""
    n = _ONLY_TLS.n_rows
    m = _ONLY[0]
    if n is None or m is None or m.numel() != B:
        return B
    return max(1, min(int(n), B))


def current_row_mask(B: int):
    """The active `only_rows` mask when it fits an ensemble of B rows, else None."""
    m = _ONLY[0]
    if m is None or not m.is_cuda or m.numel() != B or m.element_size() != 1 or not m.is_contiguous():
        return None
    return m


def _row_mask(sd, B: int):
    """Point sd.only at the active context mask when it matches this launch; tells the caller to zero-fill."""
    m = _ONLY[0]
    if m is None or not m.is_cuda or m.numel() != B or m.element_size() != 1 or not m.is_contiguous():
        return False
    sd.only = m.data_ptr()
    return True



def _jump_tensor(ts: tuple):
    torch = _lib.require_cuda()
    key = (torch.cuda.current_device(), ts)
    if key not in _JUMP_CACHE:
        _JUMP_CACHE[key] = torch.tensor(ts, dtype=torch.float64, device="cuda")
    return _JUMP_CACHE[key]


def uniform_save_dt(save_ts, t0: float, t1: float) -> float:
    """save_dt hint for the C ABI: > 0 only if `save_ts` (host array) is bit-for-bit the grid
    t0 + k*dt (k < T-1), ts[T-1] = t1 that build_saveat's linspace produces; else 0."""
    import numpy as np

    if not isinstance(save_ts, np.ndarray) or save_ts.ndim != 1 or save_ts.size < 2:
        return 0.0
    ts = save_ts.astype(np.float64, copy=False)
    dt = float(ts[1] - ts[0])
    if not dt > 0.0 or ts[0] != t0 or ts[-1] != t1:
        return 0.0
    k = np.arange(ts.size - 1, dtype=np.float64)
    return dt if np.array_equal(ts[:-1], k * dt + t0) else 0.0


_GRID_CACHE: Dict[tuple, object] = {}


def _row_strided(t) -> bool:
    """[B, row] view whose rows are contiguous but lie batch_stride apart (a column block of a wider array):
    DynodeArray describes it as it is, no gather copy needed."""
    return t.dim() == 2 and t.shape[0] > 1 and t.stride(1) == 1 and t.stride(0) > t.shape[1]


def _dev_f64(torch, x, device):
    t = torch.as_tensor(x, dtype=torch.float64, device=device)
    return t if _row_strided(t) else t.contiguous()


def _as_array(t, row: int, B: int, name: str) -> _lib.Array:
    """(B,row) / (row,) / scalar tensor -> DynodeArray (shared rows get batch_stride 0)."""
    if t is None:
        return _lib.Array(None, 0)
    if t.numel() == row:
        return _lib.Array(t.data_ptr(), 0)
    if t.numel() == B * row:
        return _lib.Array(t.data_ptr(), t.stride(0) if _row_strided(t) and t.shape == (B, row) else row)
    raise ValueError(f"{name}: expected {row} or {B}x{row} values, got shape {tuple(t.shape)}")


class _Bound:
    """Validated device-side view of one ensemble call (keeps tensors alive across the launch)."""

    def __init__(self, model: FlowModel, y0, params: Dict[str, object], contact, save_ts, B=None,
                 opts: "SolverOptions" = None):
        torch = _lib.require_cuda()
        model.check_supported()
        dev = torch.device("cuda", torch.cuda.current_device())
        self.torch, self.dev, self.model = torch, dev, model
        n, S, G = model.state_size, model.n_strains, model.n_groups
        self.y0 = _dev_f64(torch, y0, dev)
        if self.y0.shape[-1] != n and self.y0.numel() % n != 0:
            raise ValueError(f"y0 must have {n} values per trajectory, got {tuple(self.y0.shape)}")
        self.p = {k: (_dev_f64(torch, params[k], dev) if params.get(k) is not None else None)
                  for k in _PARAM_FIELDS}
        # ensemble size = the largest leading dimension among batched inputs
        if B is None:
            B = max(1, self.y0.numel() // n)
            for k, t in self.p.items():
                if t is None:
                    continue
                row = S if k in ("beta", "gamma", "sigma", "omega") else 1
                B = max(B, t.numel() // row)
        self.B = int(B)
        self.contact = None
        if contact is not None:
            self.contact = _dev_f64(torch, contact, dev)
            if self.contact.numel() != G * G:
                raise ValueError(f"contact must be {G}x{G}, got {tuple(self.contact.shape)}")
        # host grids are checked for build_saveat's uniform pattern (lets the kernel skip the loads);
        # device-resident grids carry the hint from the caller (attribute `dynode_save_dt`)
        self.save_dt = getattr(save_ts, "dynode_save_dt", 0.0)
        on_device = isinstance(save_ts, torch.Tensor) and save_ts.is_cuda
        if opts is not None and not hasattr(save_ts, "dynode_save_dt") and not on_device:
            self.save_dt = uniform_save_dt(save_ts, float(opts.t0), float(opts.t1))
        key = None
        if self.save_dt > 0.0 and not on_device:  # uniform host grid: one device copy per (device, grid)
            key = (dev.index, len(save_ts), float(save_ts[0]), float(save_ts[-1]), self.save_dt)
        if key is not None and key in _GRID_CACHE:
            self.save_ts = _GRID_CACHE[key]
        else:
            self.save_ts = _dev_f64(torch, save_ts, dev)
            if key is not None:
                _GRID_CACHE[key] = self.save_ts
        self.T = int(self.save_ts.numel())
        self.c_params = _lib.Params()
        for k in _PARAM_FIELDS:
            row = S if k in ("beta", "gamma", "sigma", "omega") else 1
            setattr(self.c_params, k, _as_array(self.p[k], row, self.B, k))
        self.c_params.contact = self.contact.data_ptr() if self.contact is not None else None
        self.c_y0 = _as_array(self.y0, n, self.B, "y0")


def solve_ensemble(model: FlowModel, y0, params: Dict[str, object], contact, opts: SolverOptions,
                   save_ts, save_mask: Optional[int] = None, wrt: Sequence[int] = (), dy0=None,
                   out=None, stats_out=None, B: Optional[int] = None):
    """One launch for the whole ensemble.  Returns (ys[B,T,n_saved], dys or None, stats[B,4]).

    Everything stays on the current CUDA device and stream; nothing synchronises.
    """
    b = _Bound(model, y0, params, contact, save_ts, B, opts)
    torch = b.torch
    mask = model.full_mask() if save_mask is None else int(save_mask)
    ns = model.saved_size(mask)
    if out is not None:
        if (out.dtype != torch.float64 or not out.is_cuda or not out.is_contiguous()
                or out.numel() != b.B * b.T * ns):
            raise ValueError(f"out must be a contiguous float64 CUDA tensor of {b.B}x{b.T}x{ns} values")
    ys = out if out is not None else torch.empty((b.B, b.T, ns), dtype=torch.float64, device=b.dev)
    stats = stats_out if stats_out is not None else torch.empty((b.B, 4), dtype=torch.int32, device=b.dev)
    L = _lib.load()
    md, sd = model.desc(), opts.desc(b.save_dt)
    stream = ctypes.c_void_p(_lib.current_stream_ptr())
    P = len(wrt)
    masked = _row_mask(sd, b.B)
    if masked:
        ys.zero_(); stats.zero_()
    if P == 0:
        _lib.check(L.dynode_solve_f64(ctypes.byref(md), ctypes.byref(sd), b.B, b.c_y0,
                                      ctypes.byref(b.c_params), b.save_ts.data_ptr(), b.T, mask,
                                      ys.data_ptr(), stats.data_ptr(), stream))
        return ys, None, stats
    dys = (torch.zeros if masked else torch.empty)((b.B, b.T, ns, P), dtype=torch.float64, device=b.dev)
    d0 = None
    if dy0 is not None:
        d0 = _dev_f64(torch, dy0, b.dev)
        if d0.numel() != b.B * P * model.state_size:
            raise ValueError("dy0 must be [B][n_wrt][n]")
    _lib.check(L.dynode_solve_sens_f64(ctypes.byref(md), ctypes.byref(sd), b.B, b.c_y0,
                                       ctypes.byref(b.c_params), b.save_ts.data_ptr(), b.T, mask, P,
                                       _lib.i32_array(list(wrt)), d0.data_ptr() if d0 is not None else None,
                                       ys.data_ptr(), dys.data_ptr(), stats.data_ptr(), stream))
    return ys, dys, stats


def poisson_loglik_grad(model: FlowModel, y0, params: Dict[str, object], contact, opts: SolverOptions,
                        save_ts, obs_comp: int, obs, lp_const: float = 0.0, wrt: Sequence[int] = (),
                        dy0=None, B: Optional[int] = None, zero_masked: bool = True):
    """Fused solve + Poisson-incidence log-likelihood + gradient: returns (lp[B], grad[B,P], stats).
    `zero_masked=False`: rows left out by an `only_rows` mask are not zero-filled (the caller never reads them)."""
    b = _Bound(model, y0, params, contact, save_ts, B, opts)
    torch = b.torch
    m = model.compartment_sizes()[obs_comp]
    obs_t = _dev_f64(torch, obs, b.dev)
    if obs_t.numel() != (b.T - 1) * m:
        raise ValueError(f"obs must be [{b.T - 1}][{m}], got {tuple(obs_t.shape)}")
    P = len(wrt)
    md, sd = model.desc(), opts.desc(b.save_dt)
    # lp, grad and stats are views of ONE buffer: under a row mask (rows left out keep zeros) that is one fill
    # instead of three
    Pc = max(P, 1)
    masked = _row_mask(sd, b.B)
    buf = (torch.zeros if (masked and zero_masked) else torch.empty)((b.B * (Pc + 3),), dtype=torch.float64,
                                                                     device=b.dev)
    lp = buf[:b.B]
    grad = buf[b.B:b.B * (1 + Pc)].view(b.B, Pc)
    stats = buf[b.B * (1 + Pc):].view(torch.int32).view(b.B, 4)
    d0 = None
    if dy0 is not None:
        d0 = _dev_f64(torch, dy0, b.dev)
    _lib.check(_lib.load().dynode_poisson_loglik_grad_f64(
        ctypes.byref(md), ctypes.byref(sd), b.B, b.c_y0, ctypes.byref(b.c_params), b.save_ts.data_ptr(),
        b.T, int(obs_comp), obs_t.data_ptr(), float(lp_const), P, _lib.i32_array(list(wrt)),
        d0.data_ptr() if d0 is not None else None, lp.data_ptr(), grad.data_ptr(), stats.data_ptr(),
        ctypes.c_void_p(_lib.current_stream_ptr())))
    return lp, (grad[:, :P] if P else None), stats


_SCRATCH: Dict[tuple, object] = {}


def _scratch(torch, dev, name: str, numel: int):
    """Grow-only device scratch (checkpoints of the adjoint): reused across calls, so a NUTS run allocates
    it once and a CUDA-graph capture sees a stable address."""
    key = (dev.index, name)
    t = _SCRATCH.get(key)
    if t is None or t.numel() < numel:
        t = torch.empty(int(numel), dtype=torch.float64, device=dev)
        _SCRATCH[key] = t
    return t


def poisson_loglik_adjoint(model: FlowModel, y0, params: Dict[str, object], contact, opts: SolverOptions, save_ts,
                           obs_comp: int, obs, lp_const: float = 0.0, with_y0_grad: bool = False,
                           B: Optional[int] = None, cap: int = 512, zero_masked: bool = True):
    """Fused solve + Poisson-incidence log-likelihood + gradient w.r.t. ALL rates (and y0) by the discrete
    adjoint: returns (lp[B], grad[B, 4*S+2] ordered (beta_s, gamma_s, sigma_s, omega_s, amp, phase),
    grad_y0[B, n] or None, stats[B, 4]).  `cap` bounds the accepted steps per trajectory that fit the
    checkpoint scratch (stats result == 2 and NaN outputs beyond it)."""
    b = _Bound(model, y0, params, contact, save_ts, B, opts)
    torch = b.torch
    m = model.compartment_sizes()[obs_comp]
    obs_t = _dev_f64(torch, obs, b.dev)
    if obs_t.numel() != (b.T - 1) * m:
        raise ValueError(f"obs must be [{b.T - 1}][{m}], got {tuple(obs_t.shape)}")
    S, n = model.n_strains, model.state_size
    lp = torch.empty((b.B,), dtype=torch.float64, device=b.dev)
    grad = torch.empty((b.B, 4 * S + 2), dtype=torch.float64, device=b.dev)
    g0 = torch.empty((b.B, n), dtype=torch.float64, device=b.dev) if with_y0_grad else None
    stats = torch.empty((b.B, 4), dtype=torch.int32, device=b.dev)
    ckpt = _scratch(torch, b.dev, "ckpt", b.B * cap * (n + 2))
    vsave = _scratch(torch, b.dev, "vsave", b.B * b.T * m)
    md, sd = model.desc(), opts.desc(b.save_dt)
    if _row_mask(sd, b.B):
        stats.zero_()  # always: the overflow counter and the forward-mode fallback read the result code of every row
        if zero_masked:
            lp.zero_(); grad.zero_()
            if g0 is not None:
                g0.zero_()
    _lib.check(_lib.load().dynode_poisson_loglik_adjoint_f64(
        ctypes.byref(md), ctypes.byref(sd), b.B, b.c_y0, ctypes.byref(b.c_params), b.save_ts.data_ptr(), b.T,
        int(obs_comp), obs_t.data_ptr(), float(lp_const), lp.data_ptr(), grad.data_ptr(),
        g0.data_ptr() if g0 is not None else None, stats.data_ptr(), ckpt.data_ptr(), int(cap), vsave.data_ptr(),
        ctypes.c_void_p(_lib.current_stream_ptr())))
    _overflow_counter(torch, b.dev).add_((stats[:, _lib.STAT_RESULT] == _lib.RESULT_ADJOINT_CAPACITY).sum())
    return lp, grad, g0, stats


_OVERFLOW: Dict[int, object] = {}


def _overflow_counter(torch, dev):
    """Device counter of trajectories whose accepted steps did not fit the adjoint's checkpoint scratch since
    the last `adjoint_overflows(reset=True)`; updated with stream-ordered tensor ops (graph-capturable).  The
    differentiable path (simulation/autograd.py::PoissonLoglik) re-evaluates such rows by forward sensitivities in
    the same call, so the counter is a diagnostic (how often the fallback ran), not a failure count."""
    t = _OVERFLOW.get(dev.index)
    if t is None:
        t = torch.zeros((), dtype=torch.int64, device=dev)
        _OVERFLOW[dev.index] = t
    return t


def adjoint_overflows(reset: bool = False) -> int:
    torch = _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    t = _overflow_counter(torch, dev)
    n = int(t.item())
    if reset:
        t.zero_()
    return n
