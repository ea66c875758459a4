"""DynODE configuration classes (host-side API kept from reference src/dynode/config/__init__.py:27-50)."""

from .bins import AgeBin, Bin, DiscretizedPositiveIntBin, WaneBin  # noqa: F401
from .deterministic_parameter import DeterministicParameter  # noqa: F401
from .dimension import (  # noqa: F401
    Dimension,
    FullStratifiedImmuneHistoryDimension,
    ImmuneHistoryDimension,
    LastStrainImmuneHistoryDimension,
    VaccinationDimension,
    WaneDimension,
)
from .initializer import Initializer  # noqa: F401
from .params import AbstractSolver, Params, SolverParams, TransmissionParams, Tsit5  # noqa: F401
from .simulation_config import Compartment, SimulationConfig  # noqa: F401
from .simulation_date import get_dynode_init_date_flag, set_dynode_init_date_flag, simulation_day  # noqa: F401
from .strains import Strain  # noqa: F401


def __getattr__(name):  # PlaceholderSample needs the distributions of dynode_b200.infer (import cycle)
    if name in ("PlaceholderSample", "SamplePlaceholderError"):
        from . import placeholder_sample
        return getattr(placeholder_sample, name)
    raise AttributeError(name)
