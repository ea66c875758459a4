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
