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
    immunity: Any  # (2^strains, [vax,] wane, strains) protection in [0, 1]
    # vaccination (reference utils/splines.py:72-109): nu[age][vax](t) = cubic spline, proportion of the age group per day
    vax_base: Any = None  # (age, vax, 4)   a + b t + c t^2 + d t^3
    vax_knots: Any = None  # (age, vax, knots)
    vax_coef: Any = None  # (age, vax, knots)  coefficient of (t - knot)^3 for t > knot
    # external introductions (reference config/strains.py:59-109)
    intro_time: Any = None  # (strains,) day of the peak
    intro_scale: Any = None  # (strains,) standard deviation in days
    intro_pct: Any = None  # (strains,) external population relative to the tracked one (0 = none)
    intro_ages: Any = None  # (strains, age) age structure of the external population
    season_tau: Any = None  # seasonal reset of the top vaccination tier: phi(t) = sin(2 pi (t + tau) / 730)^1000


def evaluate_cubic_spline(t, knot_locations, base_equations, knot_coefficients):
    """reference utils/splines.py:72-109: base cubic + sum_i coef_i (t - knot_i)^3 [t > knot_i]."""
    t = torch.as_tensor(t, dtype=torch.float64)
    powers = torch.stack([torch.ones_like(t), t, t ** 2, t ** 3])
    base = (base_equations * powers).sum(-1)
    d = torch.clamp(t - knot_locations, min=0.0)
    return base + (d ** 3 * knot_coefficients).sum(-1)


@flow_family("seip", sigma="sigma", omega="omega", contact="contact_matrix", population="population",
             immunity="immunity")
def seip_ode(t: float, state: CompartmentState, p: SEIP_ODEParams):
    """S (age, hist, [vax,] wane); E, I, C (age, hist, [vax,] strain); hist = bit set of strains recovered from.
    The equations are this repository's reading of reference ode_model.md:15-53 (stated term by term in
    oracle/dynode_oracle.cpp FAM_SEIPV, of which this is the vectorised twin)."""
    s, e, i, c = state
    flat = s.dim() == 3  # no vaccination dimension
    if flat:
        s, e, i = s.unsqueeze(2), e.unsqueeze(2), i.unsqueeze(2)
    imm = p.immunity if p.immunity.dim() == 4 else p.immunity.unsqueeze(1)
    A, H, V, W = s.shape
    K = e.shape[-1]
    frac = i.sum((1, 2)) / p.population[:, None]  # (age, strain)
    if p.intro_pct is not None:
        z = (t - p.intro_time) / p.intro_scale
        pdf = torch.exp(-0.5 * z * z) / (p.intro_scale * math.sqrt(2.0 * math.pi))
        frac = frac + (pdf * p.intro_pct)[None, :] * p.intro_ages.T
    foi = p.beta * (p.contact_matrix @ frac)  # (age, strain)
    expo = foi[:, None, None, None, :] * (1.0 - imm)[None] * s[..., None]  # (age, hist, vax, wane, strain)
    ds = -expo.sum(-1)
    ds[..., 1:] += p.omega[:-1] * s[..., :-1]
    ds[..., :-1] -= p.omega[:-1] * s[..., :-1]
    for j in range(H):
        for k in range(K):
            if (j >> k) & 1:
                ds[:, j, :, 0] += p.gamma[k] * (i[:, j, :, k] + i[:, j ^ (1 << k), :, k])
    if p.vax_base is not None:
        nu = torch.clamp(evaluate_cubic_spline(t, p.vax_knots, p.vax_base, p.vax_coef), min=0.0)  # (age, vax)
        tot = s.sum((1, 3))
        ok = (nu > 0) & (tot > 0)
        rate = torch.where(ok, torch.clamp(nu * p.population[:, None] / torch.where(ok, tot, torch.ones_like(tot)),
                                           max=1.0), torch.zeros_like(tot))
        out = rate[:, None, :, None] * s
        out[:, :, V - 1, 0] = 0.0  # a dose in the top tier moves stages >= 1 back to stage 0; stage 0 stays
        ds = ds - out
        vin = torch.zeros((A, H, V), dtype=s.dtype)
        vin[:, :, 1:] += rate[:, None, :-1] * s[:, :, :-1, :].sum(-1)
        vin[:, :, V - 1] += rate[:, None, V - 1] * s[:, :, V - 1, 1:].sum(-1)
        ds[..., 0] += vin
    de = expo.sum(3) - p.sigma * e
    di = p.sigma * e - p.gamma * i
    if p.season_tau is not None and V >= 2:
        phi = torch.sin(torch.as_tensor(2.0 * math.pi * (t + p.season_tau) / 730.0, dtype=torch.float64)) ** 1000
        for d_, x in ((ds, s), (de, e), (di, i)):
            top = x[:, :, V - 1].clone()
            d_[:, :, V - 1] -= phi * top
            d_[:, :, V - 2] += phi * top
    dc = expo.sum(3)
    if flat:
        ds, de, di, dc = ds.squeeze(2), de.squeeze(2), di.squeeze(2), dc.squeeze(2)
    return (ds, de, di, dc)


__all__ = [n for n in dir() if not n.startswith("_") and n not in ("SimpleNamespace",)]
