"""BASELINE config 5: age(3) x risk(2) x strain(3) SEIRS + cumulative incidence, with NUTS inference of the
strains' r0 and infectious periods from daily incidence.

The reference has no such script (its multi-strain example, examples/seirs_multi_strain_age_stratified.py, has
2 ages x 3 strains and no inference); this composes that RHS with the age x risk contact structure of
examples/sir_age_risk_stratified.py:113-115 (kron of the matrices pinned by
tests/test_age_risk_groups/test_age_risk_groups.py:72-75).  State: s (6,), e/i/r/c (6, 3) -> n = 78.
"""

from datetime import date

import torch

from ..config import (
    Bin,
    Compartment,
    Dimension,
    Initializer,
    Params,
    SimulationConfig,
    SolverParams,
    Strain,
    TransmissionParams,
)
from ..infer import distributions as dist
from ..infer import ppl, sample_then_resolve
from ..simulation import simulate, simulate_incidence_loglik
from .rhs import SEIRS_MultiStrain_ODEParams, seirs_multi_strain_ode

AGE3 = torch.tensor([[0.8, 0.2, 0.0], [0.2, 0.8, 0.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
RISK2 = torch.tensor([[0.5, 0.5], [0.5, 0.5]], dtype=torch.float64)
TRUE_R0 = (2.2, 2.8, 1.9)
TRUE_INF = (6.0, 7.0, 8.0)


class GroupInitializer(Initializer):
    def __init__(self):
        super().__init__(description="age x risk SEIRS initializer", initialize_date=date(2022, 2, 11),
                         population_size=1000)

    def get_initial_state(self, **kwargs):
        demo = (torch.tensor([0.7, 0.2, 0.1], dtype=torch.float64)[:, None] * torch.full((3, 2), 0.5, dtype=torch.float64)).reshape(6)
        s0 = self.population_size * 0.99 * demo
        i0 = self.population_size * 0.01 * demo[:, None] * torch.tensor([[0.4, 0.35, 0.25]], dtype=torch.float64)
        z = torch.zeros(6, 3, dtype=torch.float64)
        return (s0, z, i0, z.clone(), z.clone())


def get_config(infer: bool = False) -> SimulationConfig:
    group = Dimension(name="group", bins=[Bin(name=f"a{a}_r{r}") for a in range(3) for r in range(2)])
    strain_dim = Dimension(name="strain", bins=[Bin(name=f"v{k}") for k in range(3)])
    comps = [Compartment(name="s", dimensions=[group])] + [
        Compartment(name=n, dimensions=[group, strain_dim]) for n in ("e", "i", "r", "c")]
    strains = []
    for k in range(3):
        if infer:
            strains.append(Strain(strain_name=f"v{k}", r0=dist.Uniform(1.2, 4.0),
                                  infectious_period=dist.TruncatedNormal(loc=7.0, scale=2.0, low=3.0, high=12.0)))
        else:
            strains.append(Strain(strain_name=f"v{k}", r0=TRUE_R0[k], infectious_period=TRUE_INF[k]))
    names = [s.strain_name for s in strains]
    tp = TransmissionParams(strains=strains, strain_interactions={a: {b: 1.0 for b in names} for a in names},
                            contact_matrix=torch.kron(AGE3, RISK2), latent_period=3.0, waning_period=60.0)
    return SimulationConfig(compartments=comps, initializer=GroupInitializer(),
                            parameters=Params(solver_params=SolverParams(), transmission_params=tp))


def get_odeparams(config: SimulationConfig) -> SEIRS_MultiStrain_ODEParams:
    tp = sample_then_resolve(config.parameters.transmission_params)
    r0 = torch.stack([torch.as_tensor(s.r0, dtype=torch.float64) for s in tp.strains])
    inf = torch.stack([torch.as_tensor(s.infectious_period, dtype=torch.float64) for s in tp.strains])
    dev = r0.device
    ones = torch.ones(3, dtype=torch.float64, device=dev)
    return SEIRS_MultiStrain_ODEParams(beta=r0 / inf, gamma=1.0 / inf, sigma=ones / tp.latent_period,
                                       omega=ones / tp.waning_period, contact_matrix=tp.contact_matrix)


def run_simulation(config: SimulationConfig, tf, sub_save_indices=None):
    return simulate(seirs_multi_strain_ode, tf, config.initializer.get_initial_state(), get_odeparams(config),
                    config.parameters.solver_params, sub_save_indices=sub_save_indices)


def model(config: SimulationConfig, tf, obs_data=None):
    """Poisson likelihood on the daily increments of the cumulative compartment (per group and strain)."""
    sol = run_simulation(config, tf, sub_save_indices=(int(config.idx.c),))
    incidence = torch.clamp(torch.diff(sol.ys[config.idx.c], dim=0), min=1e-6)
    ppl.sample("incidence", dist.Poisson(incidence), obs=obs_data)
    return sol


def model_fused(config: SimulationConfig, tf, obs_data):
    lp = simulate_incidence_loglik(seirs_multi_strain_ode, tf, config.initializer.get_initial_state(),
                                   get_odeparams(config), config.parameters.solver_params,
                                   compartment=int(config.idx.c), obs=obs_data)
    ppl.factor("incidence", lp)


def synthetic_incidence(tf=120):
    cfg = get_config(infer=False)
    sol = run_simulation(cfg, tf)
    return torch.diff(sol.ys[cfg.idx.c], dim=0)
