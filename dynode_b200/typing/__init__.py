"""Type aliases mirroring reference src/dynode/typing/typing.py:11-39 (torch replaces jax.Array)."""

from .typing import (  # noqa: F401
    CompartmentGradients,
    CompartmentState,
    DynodeName,
    ODE_Eqns,
    ObservedData,
    UnitIntervalFloat,
)
