"""Inferring R0 and the infectious period of the age-stratified SIR model: the torch counterpart of
reference examples/sir_infer_parameters.py.

`model` is the reference's numpyro model line for line (numpyro.sample -> ppl.sample, jnp -> torch);
`model_fused` states the same posterior through the fused device path (solve + Poisson log-likelihood +
gradient in ONE launch, no trajectory written).
"""

import torch

from ..config import SimulationConfig, Strain
from ..infer import distributions as dist
from ..infer import ppl
from ..simulation import simulate_incidence_loglik
from .sir_age_stratified import get_config as get_static_config
from .sir_age_stratified import get_odeparams, run_simulation, sir_ode


def model(config: SimulationConfig, tf, obs_data=None):
    """Poisson likelihood on the daily increments of R (reference sir_infer_parameters.py:21-39)."""
    solution = run_simulation(config, tf)
    incidence = torch.diff(solution.ys[config.idx.r], dim=0)  # leading time axis
    incidence = torch.clamp(incidence, min=1e-6)
    ppl.sample("inf_incidence", dist.Poisson(incidence), obs=obs_data)
    return solution


def model_fused(config: SimulationConfig, tf, obs_data):
    """Same log-density as `model`, evaluated without materialising the trajectory."""
    lp = simulate_incidence_loglik(
        sir_ode, tf, config.initializer.get_initial_state(SIRConfig=config), get_odeparams(config),
        config.parameters.solver_params, compartment=int(config.idx.r), obs=obs_data)
    ppl.factor("inf_incidence", lp)


def get_config() -> SimulationConfig:
    """r0 = 1.5 + Beta(1/2, 1/2), infectious_period ~ TruncatedNormal(8, 2, 2, 15) (reference :42-60)."""
    cfg = get_static_config(r_0=2.0, infectious_period=7.0)
    cfg.parameters.transmission_params.strains = [
        Strain(strain_name="swo9",
               r0=dist.TransformedDistribution(dist.Beta(0.5, 0.5), dist.transforms.AffineTransform(1.5, 1)),
               infectious_period=dist.TruncatedNormal(loc=8, scale=2, low=2, high=15))
    ]
    return cfg


def synthetic_incidence(tf=100):
    """Un-noised incidence diff(R) from r0=2, infectious period 7 (reference :64-86)."""
    cfg = get_static_config()
    sol = run_simulation(cfg, tf=tf)
    return torch.diff(sol.ys[cfg.idx.r], dim=0)
