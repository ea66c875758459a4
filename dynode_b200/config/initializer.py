"""Abstract initializer (API of reference src/dynode/config/initializer.py:12-47)."""

from datetime import date

from pydantic import BaseModel, PositiveInt

from ..typing import CompartmentState


class Initializer(BaseModel):
    """Produces the initial compartment state (a tuple of tensors shaped like the compartments)."""

    description: str
    initialize_date: date
    population_size: PositiveInt

    def get_initial_state(self, **kwargs) -> CompartmentState:
        raise NotImplementedError("implement functionality to get initial state")
