"""Types used across the DynODE-compatible host API.

Mirrors reference src/dynode/typing/typing.py:11-39; the array type is torch.Tensor (float64)
instead of jax.Array, since the engine's device memory is managed through torch.
"""

from typing import Annotated, Any, Callable, Tuple, Union

import torch
from annotated_types import Ge, Le
from pydantic import BeforeValidator

Array = torch.Tensor
CompartmentState = Tuple[torch.Tensor, ...]
CompartmentGradients = Tuple[torch.Tensor, ...]
CompartmentTimeseries = CompartmentState

UnitIntervalFloat = Annotated[float, Ge(0.0), Le(1.0)]

ODE_Eqns = Callable[[Any, CompartmentState, Any], CompartmentGradients]

ObservedData = Union[Tuple[torch.Tensor, ...], torch.Tensor]


def _verify_name(name: str) -> str:
    """Names are identifiers: no leading digit, no spaces, alphanumerics/underscores only."""
    if name[0].isnumeric():
        raise ValueError("Name can not start with a number.")
    if " " in name:
        raise ValueError("Name can not have spaces.")
    if not all(ch.isalnum() or ch == "_" for ch in name):
        raise ValueError("Name can only contain alphanumerics or underscores.")
    return name


DynodeName = Annotated[str, BeforeValidator(_verify_name)]
