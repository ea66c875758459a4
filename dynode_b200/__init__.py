"""dynode_b200 -- B200-native ensemble ODE engine for DynODE's simulate() hot path."""

from .engine import FlowModel, SolverOptions, poisson_loglik_grad, solve_ensemble  # noqa: F401
from ._lib import DynodeError  # noqa: F401
