"""dynode_b200 -- B200-native ensemble ODE engine for DynODE's simulate() hot path.

The package root re-exports the same flat names as the reference (`src/dynode/__init__.py:11-146`) for the
parts that sit on or next to the hot path: configuration classes, `simulate`, the inference processes and
the typing aliases, plus this engine's own entry points (`simulate_ensemble`, `flow_family`, ...).  The
reference's `utils` (plotting, logging, epi-week helpers) is out of scope and not provided.
"""

from . import config, infer, simulation, typing  # noqa: F401
from ._lib import DynodeError  # noqa: F401
from .config import (  # noqa: F401
    AgeBin,
    Bin,
    Compartment,
    DeterministicParameter,
    Dimension,
    DiscretizedPositiveIntBin,
    FullStratifiedImmuneHistoryDimension,
    Initializer,
    LastStrainImmuneHistoryDimension,
    Params,
    SimulationConfig,
    SolverParams,
    Strain,
    TransmissionParams,
    VaccinationDimension,
    WaneBin,
    WaneDimension,
    get_dynode_init_date_flag,
    set_dynode_init_date_flag,
    simulation_day,
)
from .engine import FlowModel, SolverOptions, poisson_loglik_adjoint, poisson_loglik_grad, solve_ensemble  # noqa: F401
from .flows import FlowSpec, UnsupportedODEError, flow_family  # noqa: F401
from .infer import (  # noqa: F401
    InferenceProcess,
    MCMCProcess,
    SVIProcess,
    checkpoint_compartment_sizes,
    resolve_deterministic,
    sample_distributions,
    sample_then_resolve,
)
from .simulation import AbstractODEParams, simulate, simulate_ensemble, simulate_incidence_loglik  # noqa: F401
from .typing import (  # noqa: F401
    CompartmentGradients,
    CompartmentState,
    CompartmentTimeseries,
    DynodeName,
    ObservedData,
    ODE_Eqns,
    UnitIntervalFloat,
)


def __getattr__(name):
    if name in ("PlaceholderSample", "SamplePlaceholderError"):
        return getattr(config, name)
    raise AttributeError(name)
