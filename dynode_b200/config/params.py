"""Parameter containers of a DynODE model (API of reference src/dynode/config/params.py).

`solver_method` is a tag object instead of a diffrax solver: the device kernels implement Tsit5
(the reference default, params.py:28-34); any other tag makes `simulate` fail loudly.
"""

from typing import Any, Dict, List, Union

from pydantic import (
    BaseModel,
    ConfigDict,
    Field,
    NonNegativeFloat,
    PositiveFloat,
    PositiveInt,
    field_validator,
    model_validator,
)

from .deterministic_parameter import DeterministicParameter
from .strains import Strain


class AbstractSolver:
    """Tag base class standing in for diffrax.AbstractSolver."""

    def __repr__(self) -> str:
        return f"{type(self).__name__}()"

    def __eq__(self, other) -> bool:
        return type(self) is type(other)

    def __hash__(self) -> int:
        return hash(type(self).__name__)


class Tsit5(AbstractSolver):
    """Tsitouras 5(4) explicit Runge-Kutta, FSAL, 4th-order dense output: the solver the kernels run."""


class SolverParams(BaseModel):
    """ODE solver settings; defaults equal the reference's (params.py:24-67)."""

    model_config = ConfigDict(arbitrary_types_allowed=True)
    solver_method: AbstractSolver = Field(default_factory=Tsit5)
    ode_solver_rel_tolerance: PositiveFloat = 1e-5
    ode_solver_abs_tolerance: PositiveFloat = 1e-6
    max_steps: PositiveInt = int(1e6)
    constant_step_size: NonNegativeFloat = 0
    discontinuity_points: List[float] = Field(default_factory=list)


class TransmissionParams(BaseModel):
    """Transmission parameters; free-form extras (contact_matrix, waning_period, ...) are allowed."""

    model_config = ConfigDict(arbitrary_types_allowed=True, extra="allow")
    strain_interactions: Dict[str, Dict[str, Any]]
    strains: List[Strain]

    @field_validator("strains", mode="before")
    @classmethod
    def _strains_not_empty(cls, strains):
        if not strains:
            raise ValueError("strains field must contain at least one Strain.")
        return strains

    @field_validator("strains", mode="after")
    @classmethod
    def _strains_consistent(cls, strains: List[Strain]) -> List[Strain]:
        intro_ages = [s.introduction_ages for s in strains if s.is_introduced]
        assert all(a == intro_ages[0] for a in intro_ages), (
            "currently DynODE requires all strains have matching introduction_ages.")
        for name in ("exposed_to_infectious", "vaccine_efficacy"):
            have = [getattr(s, name) is not None for s in strains]
            if any(have) and not all(have):
                raise AssertionError(f"if {name} is set within one strain it must be set in all of them.")
        return strains

    @model_validator(mode="after")
    def _interactions_cover_strains(self):
        names = {s.strain_name for s in self.strains}
        assert names == set(self.strain_interactions.keys()), (
            f"first dimension of strain_interactions must contain all strain names as keys. "
            f"Found {list(self.strain_interactions.keys())} but expected {sorted(names)}.")
        for strain_name, row in self.strain_interactions.items():
            assert names == set(row.keys()), (
                f"strain_interactions[{strain_name}] interactions must contain all strains as keys, "
                f"including itself, found {list(row.keys())}, expected {sorted(names)}.")
        return self


class Params(BaseModel):
    """Miscellaneous parameters of an ODE model."""

    model_config = ConfigDict(arbitrary_types_allowed=True)
    solver_params: SolverParams
    transmission_params: TransmissionParams


__all__ = ["AbstractSolver", "Tsit5", "SolverParams", "TransmissionParams", "Params",
           "DeterministicParameter", "Union"]
