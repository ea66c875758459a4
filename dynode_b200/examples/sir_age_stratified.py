"""Age-stratified SIR model: the torch counterpart of reference examples/sir_age_stratified.py.

Same pieces as the reference script -- an Initializer, a SimulationConfig with an `age` dimension, the
translation of `TransmissionParams` into ODE parameters, `run_simulation` -- with the RHS registered in
the compiled flow family (`dynode_b200.examples.rhs.sir_age_ode`), so `simulate` runs on the device.
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
from ..infer import sample_then_resolve
from ..simulation import simulate
from ..typing import CompartmentState
from .rhs import AgeSIR_ODEParams, sir_age_ode

SIR_ODEParams = AgeSIR_ODEParams
sir_ode = sir_age_ode


class SIRInitializer(Initializer):
    """1000 people, 75 % young / 25 % old, 1 % infectious (reference sir_age_stratified.py:36-66)."""

    def __init__(self):
        super().__init__(description="An SIR initalizer", initialize_date=date(2022, 2, 11), population_size=1000)

    def get_initial_state(self, s0_prop=0.99, i0_prop=0.01, **kwargs) -> CompartmentState:
        assert s0_prop + i0_prop == 1.0, f"s0_prop and i0_prop must sum to 1.0, got {s0_prop} and {i0_prop}."
        demographics = torch.tensor([0.75, 0.25], dtype=torch.float64)
        s_0 = self.population_size * s0_prop * demographics
        i_0 = self.population_size * i0_prop * demographics
        return (s_0, i_0, torch.zeros(2, dtype=torch.float64))


def get_config(r_0=2.0, infectious_period=7.0) -> SimulationConfig:
    age = Dimension(name="age", bins=[Bin(name="young"), Bin(name="old")])
    compartments = [Compartment(name=n, dimensions=[age]) for n in ("s", "i", "r")]
    contact = torch.tensor([[0.7, 0.3], [0.3, 0.7]], dtype=torch.float64)
    contact = contact / torch.linalg.eigvals(contact).real.max()  # spectral-radius normalisation (:81-85)
    parameters = Params(
        solver_params=SolverParams(),
        transmission_params=TransmissionParams(
            strains=[Strain(strain_name="swo9", r0=r_0, infectious_period=infectious_period)],
            strain_interactions={"swo9": {"swo9": 1.0}},
            contact_matrix=contact,
        ),
    )
    return SimulationConfig(compartments=compartments, initializer=SIRInitializer(), parameters=parameters)


def get_odeparams(config: SimulationConfig) -> SIR_ODEParams:
    """beta = r0 / infectious_period, gamma = 1 / infectious_period (reference :112-124)."""
    tp = sample_then_resolve(config.parameters.transmission_params)
    strain = tp.strains[0]
    r0 = torch.as_tensor(strain.r0, dtype=torch.float64)
    inf = torch.as_tensor(strain.infectious_period, dtype=torch.float64)
    return SIR_ODEParams(beta=r0 / inf, gamma=1.0 / inf, contact_matrix=tp.contact_matrix)


def run_simulation(config: SimulationConfig, tf):
    return simulate(ode=sir_ode, duration_days=tf, initial_state=config.initializer.get_initial_state(SIRConfig=config),
                    ode_parameters=get_odeparams(config), solver_parameters=config.parameters.solver_params)
