"""`simulate`: the drop-in boundary (reference src/dynode/simulation/odes.py:35-198).

Same signature, validations and return conventions as the reference, but instead of tracing the
RHS through diffrax the call resolves the (registered) RHS to a compiled flow-family instance and
makes ONE launch of the sm_100a ensemble kernel.  `simulate_ensemble` is the batched entry the
reference would reach through jax.vmap(simulate) / numpyro Predictive (infer/inference.py:225-235).
"""

from __future__ import annotations

import dataclasses
from inspect import getfullargspec
from typing import Any, Dict, Optional, Sequence, Tuple, get_type_hints

import numpy as np
import torch

from .. import _lib, engine
from ..flows import FlowSpec, UnsupportedODEError, flow_spec_of, get_path
from ..typing import CompartmentState, ODE_Eqns

MAX_STEPS_MESSAGE = "The maximum number of solver steps was reached. Try increasing `max_steps`."


@dataclasses.dataclass
class AbstractODEParams:
    """Base class of the parameter containers passed to the ODEs (reference odes.py:25-32).

    Subclass it as a plain `@dataclasses.dataclass`; fields are floats or torch tensors.
    """


@dataclasses.dataclass
class SubSaveAt:
    ts: np.ndarray
    indices: Optional[Tuple[int, ...]] = None


@dataclasses.dataclass
class SaveAt:
    """Which times / compartments are saved (stands in for diffrax.SaveAt)."""

    ts: Optional[np.ndarray] = None
    subs: Optional[SubSaveAt] = None

    @property
    def times(self) -> np.ndarray:
        return self.subs.ts if self.subs is not None else self.ts

    @property
    def indices(self) -> Optional[Tuple[int, ...]]:
        return self.subs.indices if self.subs is not None else None


@dataclasses.dataclass
class Solution:
    """What `simulate` returns (the fields DynODE reads off diffrax.Solution).

    ts:    (T,) save times
    ys:    tuple, one tensor per compartment, leading time axis: (T, *shape); compartments left out
           by `sub_save_indices` come back with shape (T, 0) (reference odes.py:182-193).
           Ensemble solves carry a leading batch axis: (B, T, *shape).
    stats: num_steps, num_accepted_steps, num_rejected_steps, max_steps
    result: 0 = successful, 1 = max_steps reached (per trajectory for ensembles)
    """

    t0: float
    t1: float
    ts: torch.Tensor
    ys: Tuple[torch.Tensor, ...]
    stats: Dict[str, Any]
    result: Any


def build_saveat(start: float, stop, step=1, sub_save_indices: Optional[Tuple[int, ...]] = None) -> SaveAt:
    """The save grid of the reference (odes.py:148-198): linspace(start, stop, int(stop // step) + 1)."""
    if step <= 0:
        step = 1
    save_times = np.linspace(start, stop, int(stop // step) + 1)
    if sub_save_indices is not None:
        return SaveAt(subs=SubSaveAt(ts=save_times, indices=tuple(int(i) for i in sub_save_indices)))
    return SaveAt(ts=save_times)


def _solver_options(solver_parameters, duration_days) -> engine.SolverOptions:
    from ..config.params import Tsit5

    if not isinstance(solver_parameters.solver_method, Tsit5):
        raise UnsupportedODEError(
            f"solver {type(solver_parameters.solver_method).__name__} is not implemented on the device; "
            "only Tsit5 is (reference default, config/params.py:28-34)")
    const_dt = float(solver_parameters.constant_step_size)
    # the reference builds ConstantStepSize() with no ClipStepSizeController: discontinuity points are ignored in
    # constant-step mode (reference odes.py:113-131)
    jumps = () if const_dt > 0.0 else tuple(float(x) for x in solver_parameters.discontinuity_points)
    return engine.SolverOptions(
        jump_ts=jumps,
        t0=0.0, t1=float(duration_days), rtol=solver_parameters.ode_solver_rel_tolerance,
        atol=solver_parameters.ode_solver_abs_tolerance,
        const_dt=const_dt, max_steps=int(solver_parameters.max_steps))


def _numel_per_traj(t: torch.Tensor, batched: bool) -> int:
    shape = t.shape[1:] if batched else t.shape
    n = 1
    for d in shape:
        n *= int(d)
    return n


def _resolve(ode, initial_state, ode_parameters, batch_size: Optional[int], state_batched: bool):
    """(FlowSpec, FlowModel, kernel parameter dict, contact[target][source]) for this call."""
    spec: FlowSpec = flow_spec_of(ode)
    comps = spec.compartments
    if len(initial_state) != len(comps):
        raise UnsupportedODEError(
            f"flow '{spec.flow}' integrates compartments {comps}, got {len(initial_state)} compartments")
    G = _numel_per_traj(initial_state[0], state_batched)
    gs = _numel_per_traj(initial_state[1], state_batched)
    if G == 0 or gs % G != 0:
        raise UnsupportedODEError("compartment shapes are not (groups) / (groups, strains)")
    S = gs // G
    for c in initial_state[2:]:
        if _numel_per_traj(c, state_batched) != gs:
            raise UnsupportedODEError("all strain-stratified compartments must share one shape")
    model = engine.FlowModel(spec.flow_id, spec.flags, G, S)
    try:
        model.check_supported()
    except _lib.DynodeError as e:
        raise UnsupportedODEError(str(e)) from None
    from ..flows import verify_flow_body
    shapes = [tuple(c.shape[1:] if state_batched else c.shape) for c in initial_state]
    cshape = None
    if spec.contact is not None:
        cshape = tuple(torch.as_tensor(get_path(ode_parameters, spec.contact)).shape)
    verify_flow_body(ode, spec, shapes, ode_parameters, G, S, cshape)
    params = {}
    for kname, path in spec.fields.items():
        params[kname] = torch.as_tensor(get_path(ode_parameters, path), dtype=torch.float64)
    contact = None
    if spec.contact is not None:
        cm = torch.as_tensor(get_path(ode_parameters, spec.contact), dtype=torch.float64)
        if cm.numel() != G * G:
            raise UnsupportedODEError(f"contact matrix must have {G}x{G} entries, got {tuple(cm.shape)}")
        cm = cm.reshape(G, G)
        contact = cm.t().contiguous() if spec.contact_layout == "source_target" else cm
    return spec, model, params, contact


def _validate_call(ode, initial_state, ode_parameters, duration_days) -> None:
    # reference odes.py:93-112
    if any(not isinstance(compartment, torch.Tensor) for compartment in initial_state):
        raise TypeError("Please pass torch.Tensor compartments (the engine's array type) instead of np.array to ODEs")
    expected = get_type_hints(getattr(ode, "__wrapped__", ode))[getfullargspec(getattr(ode, "__wrapped__", ode)).args[2]]
    assert type(ode_parameters) is expected, (
        f"passed {type(ode_parameters)} ode parameters, but your ODE model expects {expected}")
    assert isinstance(duration_days, (int, float)), "tf must be of type int or float"


def _split_ys(ys_flat: torch.Tensor, model: engine.FlowModel, mask: int, shapes, lead: Tuple[int, ...]):
    """[..., T, n_saved] -> per-compartment views (..., T, *shape); unsaved -> (..., T, 0)."""
    out, off = [], 0
    for c, sz in enumerate(model.compartment_sizes()):
        if (mask >> c) & 1:
            out.append(ys_flat[..., off:off + sz].reshape(*lead, *shapes[c]))
            off += sz
        else:
            out.append(ys_flat.new_empty((*lead, 0)))
    return tuple(out)


def _mask_from(indices: Optional[Sequence[int]], ncomp: int) -> int:
    if indices is None:
        return (1 << ncomp) - 1
    mask = 0
    for i in indices:
        # reference odes.py:185-190: `y[i] if i in sub_save_indices` with i in range(len(y)) -- an index outside
        # [0, ncomp), negative ones included, never matches and saves nothing
        if 0 <= int(i) < ncomp:
            mask |= 1 << int(i)
    return mask


def simulate(
    ode: ODE_Eqns,
    duration_days,
    initial_state: CompartmentState,
    ode_parameters: AbstractODEParams,
    solver_parameters,
    sub_save_indices: Optional[Tuple[int, ...]] = None,
    save_step: int = 1,
) -> Solution:
    """Solve `ode` for `duration_days` days from `initial_state` (reference odes.py:35-145).

    Returns a `Solution` whose `ys` holds, per compartment, the state at every saved day including
    t=0 and t=duration_days.  Raises TypeError for non-tensor compartments, AssertionError when
    `ode_parameters` is not of the type the ODE annotates, UnsupportedODEError for an ODE outside the
    compiled flow family, RuntimeError when `max_steps` is exceeded (diffrax throw=True).
    """
    _validate_call(ode, initial_state, ode_parameters, duration_days)
    if _wants_autograd(ode, initial_state, ode_parameters):
        if flow_spec_of(ode).flow == "seip":
            raise UnsupportedODEError(
                "the immune-history ('seip') kernel integrates plain solves only: it carries no sensitivities, so "
                "`simulate` cannot be differentiated or vmapped for this flow (detach the inputs, or use "
                "`simulate_ensemble` for batches); there is no CPU fallback")
        return _run_differentiable(ode, duration_days, initial_state, ode_parameters, solver_parameters,
                                   sub_save_indices, save_step)
    sol = _run(ode, duration_days, initial_state, ode_parameters, solver_parameters, sub_save_indices,
               save_step, batch_size=None, state_batched=False, throw=True)
    return sol


def _wants_autograd(ode, initial_state, ode_parameters) -> bool:
    """True when the call sits inside torch.vmap / torch.func.grad or an input requires grad: the solve
    then goes through the differentiable, vmappable Function (simulation/autograd.py) -- the role JAX
    tracing plays for the reference (numpyro vmaps the model and differentiates through diffeqsolve)."""
    from . import autograd as ag

    spec = getattr(ode, "__dynode_flow__", None)
    if spec is None:
        return False
    tensors = list(initial_state)
    for path in list(spec.fields.values()):
        try:
            tensors.append(get_path(ode_parameters, path))
        except AttributeError:
            pass
    return any(ag.is_transformed(t) or ag.needs_grad(t) for t in tensors if isinstance(t, torch.Tensor))


def _differentiable_setup(ode, duration_days, initial_state, ode_parameters, solver_parameters,
                          sub_save_indices, save_step, **payload_kw):
    from . import autograd as ag

    spec, model, params, contact = _resolve(ode, initial_state, ode_parameters, None, False)
    opts = _solver_options(solver_parameters, duration_days)
    _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    saveat = build_saveat(opts.t0, duration_days, save_step, sub_save_indices)
    mask = _mask_from(saveat.indices, model.n_compartments)
    y0 = torch.cat([ag.const_to_device(c, dev).reshape(-1) for c in initial_state])
    y0r, theta, layout, wrt_cols, y0_grad, period = ag.pack_inputs(model, params, y0, False, dev)
    payload = ag.SolvePayload(contact=None if contact is None else ag.const_to_device(ag.unwrap(contact).detach(), dev),
                              save_ts=ag.device_grid(saveat.times, opts, dev), period=period, **payload_kw)
    cfg = ag.SolveConfig(model=model, opts_key=ag.opts_key(opts), layout=layout, wrt_cols=wrt_cols,
                         y0_grad=y0_grad, mask=mask, n_saved=model.saved_size(mask), T=len(saveat.times),
                         payload=payload)
    return cfg, y0r, theta, saveat, opts


def _run_differentiable(ode, duration_days, initial_state, ode_parameters, solver_parameters, sub_save_indices,
                        save_step) -> Solution:
    """Per-draw solve that torch.vmap batches into one launch and autograd differentiates (forward
    sensitivities of the frozen-step discrete scheme).  A trajectory that runs out of `max_steps` cannot
    raise here (the launch is asynchronous and possibly vmapped): its unreached save slots stay +inf, which
    makes any log-density built on them non-finite, and `Solution.result` carries the code."""
    from . import autograd as ag

    cfg, y0r, theta, saveat, opts = _differentiable_setup(ode, duration_days, initial_state, ode_parameters,
                                                          solver_parameters, sub_save_indices, save_step)
    ys, stats, _ = ag.EnsembleSolve.apply(y0r, theta, cfg)
    shapes = [tuple(c.shape) for c in initial_state]
    T = cfg.T
    st = stats[0]
    stats_d = {"num_steps": st[_lib.STAT_STEPS], "num_accepted_steps": st[_lib.STAT_ACCEPTED],
               "num_rejected_steps": st[_lib.STAT_REJECTED], "max_steps": int(solver_parameters.max_steps)}
    return Solution(t0=opts.t0, t1=opts.t1, ts=cfg.payload.save_ts,
                    ys=_split_ys(ys[0], cfg.model, cfg.mask, shapes, (T,)), stats=stats_d,
                    result=st[_lib.STAT_RESULT])


_OBS_CACHE: Dict[tuple, tuple] = {}


def _obs_key(obs, dev) -> tuple:
    """Identity AND content of the observations on one device: a tensor mutated in place bumps `_version`, an array's
    bytes are hashed (observation tables are a few hundred values), and every CUDA device keeps its own copy."""
    if isinstance(obs, torch.Tensor):
        return ("t", id(obs), obs._version, obs.data_ptr(), tuple(obs.shape), str(dev))
    a = np.ascontiguousarray(np.asarray(obs, dtype=np.float64))
    return ("a", a.shape, hash(a.tobytes()), str(dev))


def _observation_constants(obs, dev, rows: int):
    """(device observations [T-1][m], -sum lgamma(obs+1)) cached per observation content and device: the constant is
    reduced on the device once, so the hot loop never synchronises on it."""
    key = _obs_key(obs, dev)
    hit = _OBS_CACHE.get(key)
    if hit is not None and (not isinstance(obs, torch.Tensor) or hit[0] is obs):
        return hit[1], hit[2]
    obs_t = torch.as_tensor(obs, dtype=torch.float64).to(dev).reshape(rows, -1).contiguous()
    if obs_t.data_ptr() == (obs.data_ptr() if isinstance(obs, torch.Tensor) else 0):
        obs_t = obs_t.clone()  # never alias the caller's buffer: a later in-place edit must miss the cache, not change it
    lp_const = float(-torch.lgamma(obs_t + 1.0).sum())
    if len(_OBS_CACHE) > 64:
        _OBS_CACHE.clear()
    _OBS_CACHE[key] = (obs if isinstance(obs, torch.Tensor) else None, obs_t, lp_const)
    return obs_t, lp_const


def simulate_incidence_loglik(ode, duration_days, initial_state, ode_parameters, solver_parameters, *,
                              compartment: int, obs, save_step: int = 1):
    """Fused NUTS hot path: log-likelihood of `obs` under
    `Poisson(max(diff(ys[compartment], axis=0), 1e-6))` (the observation model of reference
    examples/sir_infer_parameters.py:30-38) WITHOUT materialising the trajectory: one launch solves, forms
    the increments, accumulates the log-probability and its gradient.  Differentiable w.r.t. the rates and
    the initial state, vmappable; returns a scalar tensor per draw."""
    from . import autograd as ag

    _validate_call(ode, initial_state, ode_parameters, duration_days)
    _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    ncomp = len(initial_state)
    comp = int(compartment) % ncomp
    T1 = len(build_saveat(0.0, duration_days, save_step).times) - 1
    obs_t, lp_const = _observation_constants(obs, dev, T1)
    cfg, y0r, theta, saveat, opts = _differentiable_setup(
        ode, duration_days, initial_state, ode_parameters, solver_parameters, None, save_step,
        obs=obs_t, obs_comp=comp, lp_const=lp_const)
    lp, stats, _ = ag.PoissonLoglik.apply(y0r, theta, cfg)
    out = lp[0]
    from ..infer import potential_plan  # a model being compiled remembers the solve behind its likelihood factor
    potential_plan.record_loglik_call(cfg, y0r, theta, out)
    return out


def simulate_ensemble(
    ode: ODE_Eqns,
    duration_days,
    initial_state: CompartmentState,
    ode_parameters: AbstractODEParams,
    solver_parameters,
    sub_save_indices: Optional[Tuple[int, ...]] = None,
    save_step: int = 1,
    *,
    batch_size: int,
    state_batched: Optional[bool] = None,
    throw: bool = True,
    out: Optional[torch.Tensor] = None,
    host_chunk: int = 2048,
) -> Solution:
    """`simulate` over an ensemble of `batch_size` parameter draws in one launch per device chunk.

    Parameter fields with `batch_size` rows are per-draw, others are shared.  `initial_state`
    compartments carry a leading batch axis when `state_batched` (default: inferred from the leading
    dimension).  Outputs carry a leading batch axis and live on the device of `initial_state[0]`;
    host-resident inputs are streamed (H2D / solve / D2H overlapped in `host_chunk`-sized pieces,
    `out` = optional pinned [B, T, n_saved] host buffer to reuse).
    """
    _validate_call(ode, initial_state, ode_parameters, duration_days)
    if state_batched is None:
        state_batched = all(c.ndim >= 1 and c.shape[0] == batch_size for c in initial_state) and (
            initial_state[0].ndim >= 2 or batch_size != initial_state[0].numel() or batch_size == 1)
    return _run(ode, duration_days, initial_state, ode_parameters, solver_parameters, sub_save_indices,
                save_step, batch_size=int(batch_size), state_batched=bool(state_batched), throw=throw,
                out=out, host_chunk=host_chunk)


def _run(ode, duration_days, initial_state, ode_parameters, solver_parameters, sub_save_indices, save_step,
         *, batch_size, state_batched, throw, out=None, host_chunk=2048) -> Solution:
    if flow_spec_of(ode).flow == "seip":
        return _run_seip(ode, duration_days, initial_state, ode_parameters, solver_parameters, sub_save_indices,
                         save_step, batch_size=batch_size, state_batched=state_batched, throw=throw)
    # unsupported ODEs / solver options fail loudly before anything touches the device
    spec, model, params, contact = _resolve(ode, initial_state, ode_parameters, batch_size, state_batched)
    opts = _solver_options(solver_parameters, duration_days)
    _lib.require_cuda()
    saveat = build_saveat(opts.t0, duration_days, save_step, sub_save_indices)
    mask = _mask_from(saveat.indices, model.n_compartments)
    B = 1 if batch_size is None else batch_size
    ensemble = batch_size is not None
    out_dev = initial_state[0].device
    shapes = [tuple(c.shape[1:] if state_batched else c.shape) for c in initial_state]
    n = model.state_size
    T = len(saveat.times)
    ns = model.saved_size(mask)
    cuda = torch.device("cuda", torch.cuda.current_device())
    streamed = out_dev.type != "cuda" and ensemble and B > host_chunk
    if streamed and state_batched:
        # host-resident ensemble: every compartment goes up on its own (asynchronously when it is page-locked) and
        # the state rows are assembled on the device -- no B x n concatenation on the host inside the call
        y0 = torch.cat([c.to(cuda, dtype=torch.float64, non_blocking=True).reshape(B, -1) for c in initial_state],
                       dim=1)
    elif state_batched:
        y0 = torch.cat([c.reshape(B, -1).to(torch.float64) for c in initial_state], dim=1)
    else:
        y0 = torch.cat([c.reshape(-1).to(torch.float64) for c in initial_state])
    assert y0.shape[-1] == n

    if not streamed:
        dev_out = out if (out is not None and out.is_cuda) else None  # e.g. a slice of a gather buffer
        ys, _, stats = engine.solve_ensemble(model, y0, params, contact, opts, saveat.times, mask, B=B,
                                             out=dev_out)
        if out_dev.type != "cuda":
            if out is not None:
                out.copy_(ys)
                ys = out
            else:
                ys = ys.cpu()
            stats_h = stats.cpu()
        else:
            stats_h = stats.cpu() if throw else None
    else:
        ys, stats_h = _host_pipeline(model, y0, params, contact, opts, saveat.times, mask, B, T, ns, out,
                                     host_chunk, cuda)
        stats = stats_h

    if throw and bool((stats_h[:, _lib.STAT_RESULT] != 0).any()):
        raise RuntimeError(MAX_STEPS_MESSAGE)
    st = stats if (out_dev.type == "cuda") else stats_h
    lead = (B, T) if ensemble else (T,)
    ys_view = ys if ensemble else ys[0]
    stats_d = {
        "num_steps": st[:, _lib.STAT_STEPS] if ensemble else st[0, _lib.STAT_STEPS],
        "num_accepted_steps": st[:, _lib.STAT_ACCEPTED] if ensemble else st[0, _lib.STAT_ACCEPTED],
        "num_rejected_steps": st[:, _lib.STAT_REJECTED] if ensemble else st[0, _lib.STAT_REJECTED],
        "max_steps": int(solver_parameters.max_steps),
    }
    result = st[:, _lib.STAT_RESULT] if ensemble else st[0, _lib.STAT_RESULT]
    ts = torch.as_tensor(saveat.times, dtype=torch.float64, device=out_dev)
    return Solution(t0=opts.t0, t1=opts.t1, ts=ts, ys=_split_ys(ys_view, model, mask, shapes, lead),
                    stats=stats_d, result=result)


def _run_seip(ode, duration_days, initial_state, ode_parameters, solver_parameters, sub_save_indices, save_step, *,
              batch_size, state_batched, throw) -> Solution:
    """`simulate` / `simulate_ensemble` for the immune-history / waning family: one thread block per trajectory
    (dynode_b200/seip.py).  Compartments (s, e, i, c) shaped (A, H, W), (A, H, K), (A, H, K), (A, H, K)."""
    from .. import seip

    spec = flow_spec_of(ode)
    if len(initial_state) != 4:
        raise UnsupportedODEError("flow 'seip' integrates compartments (s, e, i, c)")
    shapes = [tuple(c.shape[1:] if state_batched else c.shape) for c in initial_state]
    nd = len(shapes[0])
    if nd not in (3, 4) or len(shapes[1]) != nd or shapes[1] != shapes[2] or shapes[1] != shapes[3] \
            or shapes[0][:-1] != shapes[1][:-1] or shapes[0][1] != (1 << shapes[1][-1]):
        raise UnsupportedODEError("flow 'seip' needs s (ages, 2^strains, [vax,] wane) and e, i, c "
                                  "(ages, 2^strains, [vax,] strains)")
    A, H, W = shapes[0][0], shapes[0][1], shapes[0][-1]
    V = shapes[0][2] if nd == 4 else 1
    K = shapes[1][-1]
    vb = getattr(ode_parameters, "vax_base", None)
    NK = 0
    if vb is not None:
        kn = getattr(ode_parameters, "vax_knots", None)
        NK = 0 if kn is None else int(torch.as_tensor(kn).shape[-1])
    model = seip.SeipModel(A, K, W, V, NK)
    try:
        model.check_supported()
    except _lib.DynodeError as e:
        raise UnsupportedODEError(str(e)) from None
    opts = _solver_options(solver_parameters, duration_days)
    mask = _mask_from(sub_save_indices, 4)
    _lib.require_cuda()
    B = 1 if batch_size is None else batch_size
    ensemble = batch_size is not None
    if state_batched:
        y0 = torch.cat([c.reshape(B, -1).to(torch.float64) for c in initial_state], dim=1)
    else:
        y0 = torch.cat([c.reshape(-1).to(torch.float64) for c in initial_state])
    get = lambda k: torch.as_tensor(get_path(ode_parameters, spec.fields[k]), dtype=torch.float64)
    params = {k: get(k) for k in ("beta", "sigma", "gamma", "omega")}
    saveat = build_saveat(opts.t0, duration_days, save_step, None)
    opt = lambda name: getattr(ode_parameters, name, None)
    vaccination = None
    if vb is not None:
        z = torch.zeros((A, V, 0), dtype=torch.float64)
        vaccination = (vb, opt("vax_knots") if NK else z, opt("vax_coef") if NK else z)
    introductions = None
    if opt("intro_pct") is not None:
        introductions = dict(time=opt("intro_time"), scale=opt("intro_scale"), pct=opt("intro_pct"),
                             ages=opt("intro_ages"))
    ys, stats = seip.solve_ensemble(model, y0, params, get_path(ode_parameters, spec.contact), get("population"),
                                    get("immunity"), opts, saveat.times, B=B, vaccination=vaccination,
                                    introductions=introductions, season_tau=opt("season_tau"), save_mask=mask)
    out_dev = initial_state[0].device
    st = stats if out_dev.type == "cuda" else stats.cpu()
    if throw and bool((stats[:, _lib.STAT_RESULT] != 0).any()):
        raise RuntimeError(MAX_STEPS_MESSAGE)
    if out_dev.type != "cuda":
        ys = ys.cpu()
    T = len(saveat.times)
    lead = (B, T) if ensemble else (T,)
    flat = ys if ensemble else ys[0]
    outs, off = [], 0
    for c, shp in enumerate(shapes):
        if not (mask >> c) & 1:  # unsaved compartments come back as (T, 0), reference odes.py:182-193
            outs.append(flat.new_empty((*lead, 0)))
            continue
        sz = int(np.prod(shp))
        outs.append(flat[..., off:off + sz].reshape(*lead, *shp))
        off += sz
    pick = (lambda c: st[:, c]) if ensemble else (lambda c: st[0, c])
    stats_d = {"num_steps": pick(_lib.STAT_STEPS), "num_accepted_steps": pick(_lib.STAT_ACCEPTED),
               "num_rejected_steps": pick(_lib.STAT_REJECTED), "max_steps": int(solver_parameters.max_steps)}
    return Solution(t0=opts.t0, t1=opts.t1, ts=torch.as_tensor(saveat.times, dtype=torch.float64, device=out_dev),
                    ys=tuple(outs), stats=stats_d, result=pick(_lib.STAT_RESULT))


def _host_pipeline(model, y0, params, contact, opts, save_ts, mask, B, T, ns, out, chunk, cuda, depth: int = 3):
    """Host-resident ensemble.  The draws (B x (n + P) doubles, 0.4 % of the output) go up in one H2D; the kernel
    then solves `chunk` draws at a time into a ring of `depth` device buffers while the copy engine drains the ring
    into the page-locked output on a second stream, so the PCIe copy of chunk k overlaps the solve of chunk k+1 and
    the link never idles.  `stats` stay on the device until the end and cross once.

    The output buffer is what bounds the rate: it must be page-locked, and `hostmem.pinned_empty` backs it with
    2 MiB huge pages (pass one as `out=` to reuse it across calls -- pinning 7.6 GB takes longer than a step)."""
    from .. import hostmem

    S = model.n_strains
    n = model.state_size
    if out is None:
        out = hostmem.pinned_empty((B, T, ns))
    compute = torch.cuda.current_stream()
    copy = torch.cuda.Stream()
    depth = max(2, int(depth))
    bufs = [torch.empty((chunk, T, ns), dtype=torch.float64, device=cuda) for _ in range(depth)]
    stats_d = torch.empty((B, 4), dtype=torch.int32, device=cuda)
    free = [torch.cuda.Event() for _ in range(depth)]
    done = [torch.cuda.Event() for _ in range(depth)]
    contact_dev = None if contact is None else contact.to(cuda)

    def up(t, row):  # whole-ensemble H2D of one input (pinned inputs copy asynchronously)
        if t is None:
            return None, False
        d = t if t.is_cuda else t.to(cuda, non_blocking=True)
        return (d, False) if t.numel() == row else (d.reshape(B, row), True)

    p_dev = {name: up(t, S if name in ("beta", "gamma", "sigma", "omega") else 1) for name, t in params.items()}
    y_dev, y_batched = up(y0, n)

    k = 0
    for lo in range(0, B, chunk):
        hi = min(B, lo + chunk)
        j = k % depth
        if k >= depth:
            compute.wait_event(free[j])  # D2H of the chunk that used this buffer has finished
        p = {name: (None if d is None else (d[lo:hi] if batched else d)) for name, (d, batched) in p_dev.items()}
        y = y_dev[lo:hi] if y_batched else y_dev
        engine.solve_ensemble(model, y, p, contact_dev, opts, save_ts, mask, out=bufs[j][: hi - lo],
                              stats_out=stats_d[lo:hi], B=hi - lo)
        done[j].record(compute)
        copy.wait_event(done[j])
        with torch.cuda.stream(copy):
            out[lo:hi].copy_(bufs[j][: hi - lo], non_blocking=True)
            free[j].record(copy)
        k += 1
    stats_h = torch.empty((B, 4), dtype=torch.int32, pin_memory=True)
    stats_h.copy_(stats_d, non_blocking=True)  # one D2H, after the last kernel of the compute stream
    compute.synchronize()
    copy.synchronize()
    return out, stats_h
