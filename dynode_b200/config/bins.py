"""Bin types: the cells of a compartment dimension (API of reference src/dynode/config/bins.py)."""

from pydantic import BaseModel, NonNegativeFloat, NonNegativeInt, PositiveFloat, model_validator
from pydantic import Field

from ..typing import DynodeName


class Bin(BaseModel):
    """One named cell of a dimension."""

    name: DynodeName


class DiscretizedPositiveIntBin(Bin):
    """A bin covering the inclusive integer range [min_value, max_value]."""

    min_value: NonNegativeInt
    max_value: NonNegativeInt

    def __init__(self, min_value, max_value, name=None):
        super().__init__(name=name if name is not None else f"range_{min_value}_{max_value}",
                         min_value=min_value, max_value=max_value)

    @model_validator(mode="after")
    def _ordered(self):
        assert self.min_value <= self.max_value
        return self


class AgeBin(DiscretizedPositiveIntBin):
    """Inclusive age range; default name a{min}_{max}."""

    def __init__(self, min_value, max_value, name=None):
        super().__init__(min_value, max_value, name if name is not None else f"a{min_value}_{max_value}")


class WaneBin(Bin):
    """A waning stage: mean waiting time (days; inf = terminal) and retained protection in [0, 1]."""

    waiting_time: PositiveFloat
    base_protection: NonNegativeFloat = Field(le=1.0)
