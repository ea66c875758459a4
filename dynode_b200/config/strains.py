"""Strain definition (API of reference src/dynode/config/strains.py)."""

from datetime import date
from typing import Any, Dict, List, Optional

from pydantic import BaseModel, ConfigDict, NonNegativeFloat, PositiveFloat

from ..typing import DynodeName
from .bins import AgeBin


class Strain(BaseModel):
    """A pathogen strain.  `r0` / `infectious_period` (and the introduction fields) may be numbers,
    tensors, prior distributions or DeterministicParameters; they are resolved by
    `dynode_b200.infer.sample_then_resolve` before reaching the ODE parameters."""

    model_config = ConfigDict(arbitrary_types_allowed=True)

    strain_name: DynodeName
    r0: Any
    infectious_period: Any
    exposed_to_infectious: Optional[PositiveFloat] = None
    vaccine_efficacy: Optional[Dict[int, NonNegativeFloat]] = None
    is_introduced: bool = False
    introduction_time: Optional[Any] = None
    introduction_percentage: Optional[Any] = None
    introduction_scale: Optional[Any] = None
    introduction_ages: Optional[List[AgeBin]] = None
    introduction_ages_mask_vector: Optional[List[int]] = None


__all__ = ["Strain", "date"]
