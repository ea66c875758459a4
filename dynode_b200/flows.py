"""Flow-family registration: how a DynODE right-hand side reaches the CUDA kernels.

In the reference the RHS is an arbitrary Python callable traced by JAX
(src/dynode/typing/typing.py:18-21).  The B200 engine cannot execute Python on the device, so the
boundary adds ONE concept: a right-hand side is *registered* as a member of the compiled flow
family with `@flow_family(...)`, which records how the fields of its ODE-params dataclass map onto
the kernel's parameters.  An unregistered callable raises `UnsupportedODEError` -- there is no CPU
fallback (BASELINE.json north_star: "ODEs outside the supported flow family fail loudly").
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, Optional, Tuple

from . import _lib
from ._lib import DynodeError

# "seip": the immune-history / waning family (CTA-per-trajectory kernel, include/dynode_b200_seip.h)
_FLOWS = {"sir": _lib.FLOW_SIR, "seirs": _lib.FLOW_SEIRS, "seirs_c": _lib.FLOW_SEIRS_C, "seip": -1}
_COMPARTMENTS = {"sir": ("s", "i", "r"), "seirs": ("s", "e", "i", "r"), "seirs_c": ("s", "e", "i", "r", "c"),
                 "seip": ("s", "e", "i", "c")}


class UnsupportedODEError(DynodeError):
    """The ODE (or solver option) is outside what the compiled kernels implement."""


@dataclass(frozen=True)
class FlowSpec:
    flow: str
    # kernel parameter -> attribute path inside the ODE-params object (dots allowed)
    fields: Dict[str, str] = field(default_factory=dict)
    contact: Optional[str] = None
    # "target_source": contact[a, b] multiplies source b into target a  (sir_age_stratified.py:134-136,
    #                  seirs_multi_strain_age_stratified.py:229-231)
    # "source_target": einsum("ijkl,ij->kl") sums over the leading (source) index pair
    #                  (sir_age_risk_stratified.py:163-165)
    contact_layout: str = "target_source"
    seasonal: bool = False
    density_dependent: bool = False

    @property
    def flow_id(self) -> int:
        return _FLOWS[self.flow]

    @property
    def flags(self) -> int:
        return (_lib.FLAG_SEASONAL if self.seasonal else 0) | (_lib.FLAG_DENSITY_DEP if self.density_dependent else 0)

    @property
    def compartments(self) -> Tuple[str, ...]:
        return _COMPARTMENTS[self.flow]


def flow_family(flow: str, *, beta: str = "beta", gamma: str = "gamma", sigma: Optional[str] = None,
                omega: Optional[str] = None, contact: Optional[str] = None,
                contact_layout: str = "target_source", seasonal: Optional[Tuple[str, str, str]] = None,
                density_dependent: bool = False, population: Optional[str] = None,
                immunity: Optional[str] = None) -> Callable:
    """Register `ode(t, state, p)` as a member of the compiled flow family.

    flow: "sir" (s,i,r) | "seirs" (s,e,i,r) | "seirs_c" (s,e,i,r,c); the keyword arguments name the
    attributes of the ODE-params dataclass that hold each rate (`seasonal` = (amp, phase, period)).
    """
    if flow not in _FLOWS:
        raise ValueError(f"unknown flow {flow!r}; choose from {sorted(_FLOWS)}")
    if contact_layout not in ("target_source", "source_target"):
        raise ValueError("contact_layout must be 'target_source' or 'source_target'")
    fields = {"beta": beta, "gamma": gamma}
    if flow != "sir":
        if sigma is None:
            sigma = "sigma"
        if omega is None:
            omega = "omega"
    if sigma:
        fields["sigma"] = sigma
    if omega:
        fields["omega"] = omega
    if seasonal:
        fields["season_amp"], fields["season_phase"], fields["season_period"] = seasonal
    if flow == "seip":
        if not (contact and population and immunity):
            raise ValueError("flow 'seip' needs contact=, population= and immunity= attribute names")
        fields["population"], fields["immunity"] = population, immunity
    spec = FlowSpec(flow=flow, fields=fields, contact=contact, contact_layout=contact_layout,
                    seasonal=bool(seasonal), density_dependent=density_dependent)

    def deco(ode):
        target = getattr(ode, "__wrapped__", ode)
        try:
            target.__dynode_flow__ = spec
            ode.__dynode_flow__ = spec
        except AttributeError:
            pass
        return ode

    return deco


def flow_spec_of(ode) -> FlowSpec:
    spec = getattr(ode, "__dynode_flow__", None)
    if spec is None:
        name = getattr(ode, "__name__", repr(ode))
        raise UnsupportedODEError(
            f"ODE {name!r} is not registered with @flow_family: the B200 engine integrates only the "
            "compiled compartmental flow family (SIR / SEIRS / multi-strain SEIRS+C with contact "
            "matrix, waning, seasonal beta) and has no CPU fallback")
    return spec


def get_path(obj, path: str):
    for part in path.split("."):
        obj = getattr(obj, part)
    return obj


# ---------------------------------------------------------------------------------------------------------------
# The registration is a claim; the body is checked.  `simulate` never executes the Python right-hand side (the
# kernel integrates the compiled flow), so a user who edits the body of a registered function -- or registers a
# function under the wrong flow -- would silently get the compiled equations.  Before the first solve of every
# (function, shapes) pair the callable is therefore evaluated on a few random states and parameter values and
# compared with a host statement of what the kernel computes; a mismatch raises UnsupportedODEError.  This plays the
# role JAX tracing plays for the reference, where the callable itself is what runs (reference
# src/dynode/typing/typing.py:18-21).
_VERIFIED: Dict[tuple, bool] = {}


def compiled_flow_rhs(spec: FlowSpec, G: int, S: int, t: float, state, rates: Dict[str, "object"], contact):
    """What the kernels compute for one draw (csrc/lane_solver.cuh::rhs_sc), in torch on the host.
    state: s [G] and e/i/r/c [G, S]; rates: beta/gamma/sigma/omega [S] (+ season_amp/phase/period scalars);
    contact [target][source] or None (identity)."""
    import math

    import torch

    s = state[0].reshape(G)
    rest = [c.reshape(G, S) for c in state[1:]]
    has_e = spec.flow != "sir"
    e = rest[0] if has_e else None
    i = rest[1] if has_e else rest[0]
    r = rest[2] if has_e else rest[1]
    if spec.density_dependent:
        prop = i
    else:
        N = s + i.sum(1) + r.sum(1) + (e.sum(1) if has_e else 0.0)
        prop = i / N[:, None]
    acc = prop if contact is None else contact @ prop
    beta = rates["beta"].reshape(S)
    if spec.seasonal:
        beta = beta * (1.0 + rates["season_amp"] * torch.sin(2.0 * math.pi * t / rates["season_period"]
                                                             + rates["season_phase"]))
    newinf = beta * acc * s[:, None]
    rec = rates["gamma"].reshape(S) * i
    if spec.flow == "sir":
        return (-newinf.sum(1), newinf - rec, rec)
    sig, om = rates["sigma"].reshape(S) * e, rates["omega"].reshape(S) * r
    out = ((om - newinf).sum(1), newinf - sig, sig - rec, rec - om)
    return out + ((newinf,) if spec.flow == "seirs_c" else ())


def _set_path(obj, path: str, value):
    import copy

    parts = path.split(".")
    if len(parts) == 1:
        setattr(obj, parts[0], value)
        return
    child = copy.copy(getattr(obj, parts[0]))
    setattr(obj, parts[0], child)
    _set_path(child, ".".join(parts[1:]), value)


def verify_flow_body(ode, spec: FlowSpec, shapes, ode_parameters, G: int, S: int, contact_shape=None) -> None:
    """Raise UnsupportedODEError unless `ode` computes the flow it is registered as (checked once per function and
    shape, on random inputs; the caller's parameter values are not used)."""
    import copy

    import torch

    if spec.flow == "seip":
        return
    target = getattr(ode, "__wrapped__", ode)
    code = getattr(target, "__code__", None)
    key = (id(target), hash(code) if code is not None else 0, tuple(tuple(sh) for sh in shapes), id(spec), G, S,
           None if contact_shape is None else tuple(contact_shape))
    if _VERIFIED.get(key):
        return
    gen = torch.Generator().manual_seed(20260104)
    f64 = torch.float64
    name = getattr(ode, "__name__", repr(ode))
    for trial in range(3):
        state = tuple(torch.rand(tuple(sh), dtype=f64, generator=gen) * 50.0 + 1.0 for sh in shapes)
        rates = {k: torch.rand(S, dtype=f64, generator=gen) * 0.5 + 0.05 for k in ("beta", "gamma", "sigma", "omega")
                 if k in spec.fields}
        if spec.seasonal:
            rates["season_amp"] = torch.rand((), dtype=f64, generator=gen) * 0.4
            rates["season_phase"] = torch.rand((), dtype=f64, generator=gen) * 6.0
            rates["season_period"] = torch.tensor(365.0, dtype=f64)
        contact = None
        p = copy.copy(ode_parameters)
        for kname, path in spec.fields.items():
            v = rates[kname]
            _set_path(p, path, v.reshape(()) if (S == 1 and v.numel() == 1) else v)
        if spec.contact is not None:
            K = torch.rand((G, G), dtype=f64, generator=gen) + 0.1
            contact = K
            user_K = K.t().contiguous() if spec.contact_layout == "source_target" else K
            _set_path(p, spec.contact, user_K.reshape(tuple(contact_shape)) if contact_shape is not None else user_K)
        t = float(torch.rand((), generator=gen) * 300.0)
        want = compiled_flow_rhs(spec, G, S, t, state, rates, contact)
        try:
            got = ode(t, state, p)
        except Exception as exc:
            raise UnsupportedODEError(
                f"ODE {name!r} is registered as flow '{spec.flow}' but could not be evaluated on a probe state to "
                f"check its body ({type(exc).__name__}: {exc})") from exc
        ok = len(got) == len(want)
        if ok:
            for a, b in zip(got, want):
                a = torch.as_tensor(a, dtype=f64).reshape(-1)
                b = b.reshape(-1)
                ok = ok and a.shape == b.shape and bool(torch.allclose(a, b, rtol=1e-9, atol=1e-9 * float(b.abs().max() + 1)))
        if not ok:
            raise UnsupportedODEError(
                f"the body of ODE {name!r} does not compute the flow '{spec.flow}' it is registered as with "
                "@flow_family (compared on random states with the compiled kernel's equations): the B200 engine "
                "would integrate the compiled flow, not this function -- there is no CPU fallback that runs the "
                "Python body")
    _VERIFIED[key] = True
