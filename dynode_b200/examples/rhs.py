"""Right-hand sides and ODE-parameter dataclasses of the reference examples, registered with the
compiled flow family.  The function bodies are the model definitions in torch (usable for
inspection / host-side checks); `simulate` never calls them -- it runs the registered flow on the
device.
"""

from dataclasses import dataclass
from types import SimpleNamespace
from typing import Any

import math

import torch

from ..flows import flow_family
from ..simulation.odes import AbstractODEParams
from ..typing import CompartmentGradients, CompartmentState


# ------------------------------------------------------------------ examples/sir.py:70-84
@dataclass
class SIR_ODEParams(AbstractODEParams):
    beta: Any
    gamma: Any


@flow_family("sir")
def sir_ode(t: float, state: CompartmentState, p: SIR_ODEParams):
    s, i, r = state
    N = s + i + r
    ds = -p.beta * s * i / N
    di = p.beta * s * i / N - p.gamma * i
    dr = p.gamma * i
    return (ds, di, dr)


# ------------------------------------------------------------------ tests/test_simulation/test_odes.py:11-28
@dataclass
class DensitySIR_ODEParams(AbstractODEParams):
    beta: Any
    gamma: Any


@flow_family("sir", density_dependent=True)
def sir_density_ode(t: float, state: CompartmentState, p: DensitySIR_ODEParams) -> CompartmentGradients:
    s, i, _ = state
    s_to_i = p.beta * s * i
    i_to_r = i * p.gamma
    return (-s_to_i, s_to_i - i_to_r, i_to_r)


# ------------------------------------------------------------------ examples/seirs.py:80-95
@dataclass
class SEIRS_ODEParams(AbstractODEParams):
    beta: Any
    gamma: Any
    sigma: Any
    omega: Any


@flow_family("seirs")
def seirs_ode(t: float, state: CompartmentState, p: SEIRS_ODEParams):
    s, e, i, r = state
    N = s + e + i + r
    ds = -p.beta * s * i / N + p.omega * r
    de = p.beta * s * i / N - p.sigma * e
    di = p.sigma * e - p.gamma * i
    dr = p.gamma * i - p.omega * r
    return (ds, de, di, dr)


# ------------------------------------------------------------------ examples/seirs_seasonal_forcing.py:20-55
@dataclass
class SeasonalityParams:
    forcing_amp: Any
    forcing_phase: Any
    forcing_period: Any


@dataclass
class SeasonalSEIRS_ODEParams(AbstractODEParams):
    beta: Any
    gamma: Any
    sigma: Any
    omega: Any
    seasonality_params: SeasonalityParams


def seasonality(t, params: SeasonalityParams):
    return 1.0 + params.forcing_amp * torch.sin(
        torch.as_tensor(2 * math.pi * t / params.forcing_period + params.forcing_phase))


@flow_family("seirs", seasonal=("seasonality_params.forcing_amp", "seasonality_params.forcing_phase",
                                "seasonality_params.forcing_period"))
def seirs_ode_seasonal(t: float, state: CompartmentState, p: SeasonalSEIRS_ODEParams):
    s, e, i, r = state
    N = s + e + i + r
    beta_t = p.beta * seasonality(t, p.seasonality_params)
    ds = -beta_t * s * i / N + p.omega * r
    de = beta_t * s * i / N - p.sigma * e
    di = p.sigma * e - p.gamma * i
    dr = p.gamma * i - p.omega * r
    return (ds, de, di, dr)


# ------------------------------------------------------------------ examples/sir_age_stratified.py:103-142
@dataclass
class AgeSIR_ODEParams(AbstractODEParams):
    beta: Any  # r0 / infectious period
    gamma: Any  # 1 / infectious period
    contact_matrix: Any  # (age, age)


@flow_family("sir", contact="contact_matrix")
def sir_age_ode(t: float, state: CompartmentState, p: AgeSIR_ODEParams) -> CompartmentGradients:
    s, i, r = state
    pop_size = s + i + r
    force_of_infection = p.beta * torch.sum((p.contact_matrix * i) / pop_size, dim=1)
    s_to_i = s * force_of_infection
    i_to_r = i * p.gamma
    return (-s_to_i, s_to_i - i_to_r, i_to_r)


# ------------------------------------------------------------------ examples/sir_age_risk_stratified.py:134-173
@dataclass
class AgeRiskSIR_ODEParams(AbstractODEParams):
    beta: Any
    gamma: Any
    contact_matrix: Any  # (age, risk, age, risk): [i, j, k, l] couples source (i, j) to target (k, l)


@flow_family("sir", contact="contact_matrix", contact_layout="source_target")
def sir_age_risk_ode(t: float, state: CompartmentState, p: AgeRiskSIR_ODEParams) -> CompartmentGradients:
    s, i, r = state
    pop_size = s + i + r
    force_of_infection = p.beta * torch.einsum("ijkl,ij->kl", p.contact_matrix, i / pop_size)
    s_to_i = s * force_of_infection
    i_to_r = i * p.gamma
    return (-s_to_i, s_to_i - i_to_r, i_to_r)


# ------------------------------------------------------------------ examples/seirs_multi_strain_age_stratified.py:177-243
@dataclass
class SEIRS_MultiStrain_ODEParams(AbstractODEParams):
    beta: Any  # (num_strains,)
    gamma: Any  # (num_strains,)
    sigma: Any  # (num_strains,)
    omega: Any  # (num_strains,)
    contact_matrix: Any  # (age, age)
    idx: Any = None  # SimulationConfig.idx, kept for readability of user code


@flow_family("seirs_c", contact="contact_matrix")
def seirs_multi_strain_ode(t: float, state: CompartmentState, p: SEIRS_MultiStrain_ODEParams):
    s, e, i, r, c = state  # s: (age,)  e, i, r, c: (age, strain)
    N_age = s + e.sum(-1) + i.sum(-1) + r.sum(-1)
    fois = p.beta * (p.contact_matrix @ (i / N_age[:, None]))  # (age, strain)
    ds = -torch.sum(fois * s[:, None], dim=1) + torch.sum(p.omega * r, dim=1)
    de = fois * s[:, None] - p.sigma * e
    di = p.sigma * e - p.gamma * i
    dr = p.gamma * i - p.omega * r
    dc = fois * s[:, None]
    return (ds, de, di, dr, dc)


# ------------------------------------------------------------------ reference ode_model.md:15-53 (prose model)
@dataclass
class SEIP_ODEParams(AbstractODEParams):
    beta: Any  # (strains,)
    sigma: Any  # (strains,)
    gamma: Any  # (strains,)
    omega: Any  # (wane,) waning rates, the last stage absorbs
    contact_matrix: Any  # (age, age)
    population: Any  # (age,)
    immunity: Any  # (2^strains, wane, strains) protection in [0, 1]


@flow_family("seip", sigma="sigma", omega="omega", contact="contact_matrix", population="population",
             immunity="immunity")
def seip_ode(t: float, state: CompartmentState, p: SEIP_ODEParams):
    """S (age, hist, wane); E, I, C (age, hist, strain); hist = bit set of strains recovered from."""
    s, e, i, c = state
    K = e.shape[-1]
    foi = p.beta * (p.contact_matrix @ (i.sum(1) / p.population[:, None]))  # (age, strain)
    expo = foi[:, None, None, :] * (1.0 - p.immunity)[None] * s[..., None]  # (age, hist, wane, strain)
    ds = -expo.sum(-1)
    ds[..., 1:] += p.omega[:-1] * s[..., :-1]
    ds[..., :-1] -= p.omega[:-1] * s[..., :-1]
    for j in range(s.shape[1]):
        for k in range(K):
            if (j >> k) & 1:
                ds[:, j, 0] += p.gamma[k] * (i[:, j, k] + i[:, j ^ (1 << k), k])
    de = expo.sum(2) - p.sigma * e
    di = p.sigma * e - p.gamma * i
    return (ds, de, di, expo.sum(2))


__all__ = [n for n in dir() if not n.startswith("_") and n not in ("SimpleNamespace",)]
