"""DynODE configuration classes (host-side API kept from reference src/dynode/config)."""

from .bins import AgeBin, Bin, DiscretizedPositiveIntBin, WaneBin  # noqa: F401
from .deterministic_parameter import DeterministicParameter  # noqa: F401
from .params import AbstractSolver, Params, SolverParams, TransmissionParams, Tsit5  # noqa: F401
from .strains import Strain  # noqa: F401
