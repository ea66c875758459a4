"""Types of the DynODE-compatible host API (reference src/dynode/typing/__init__.py)."""

from .typing import (  # noqa: F401
    Array,
    CompartmentGradients,
    CompartmentState,
    CompartmentTimeseries,
    DynodeName,
    ObservedData,
    ODE_Eqns,
    UnitIntervalFloat,
)
