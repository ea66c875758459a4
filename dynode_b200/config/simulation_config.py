"""Top-level model configuration (API of reference src/dynode/config/simulation_config.py:28-330).

Host-side only: it names the compartments and their axes (`Compartment.shape` is the state layout the
kernels flatten), carries the parameters, and offers the `config.idx.<compartment>.<dimension>.<bin>`
lookup that user code indexes `Solution.ys` with.
"""

from functools import cached_property
from types import SimpleNamespace
from typing import List

from pydantic import BaseModel, ConfigDict, model_validator

from ..typing import DynodeName
from .bins import AgeBin, Bin
from .dimension import (
    Dimension,
    FullStratifiedImmuneHistoryDimension,
    ImmuneHistoryDimension,
    LastStrainImmuneHistoryDimension,
)
from .initializer import Initializer
from .params import Params


class _Index(int):
    """An int (position of a compartment / dimension) that also carries the names nested below it."""

    def __new__(cls, value, **names):
        obj = super().__new__(cls, value)
        obj.__dict__.update(names)
        return obj

    def __str__(self):
        return str(self.__dict__)


class Compartment(BaseModel):
    model_config = ConfigDict(arbitrary_types_allowed=True)
    name: DynodeName
    dimensions: List[Dimension]

    @model_validator(mode="after")
    def _unique_dimension_names(self):
        names = [d.name for d in self.dimensions]
        assert len(set(names)) == len(names), "you can not have two identically named dimensions within a compartment"
        return self

    @property
    def shape(self):
        return tuple(len(d) for d in self.dimensions)

    @cached_property
    def idx(self):
        return SimpleNamespace(**{d.name: _Index(k, **vars(d.idx)) for k, d in enumerate(self.dimensions)})

    def __eq__(self, other) -> bool:
        return (isinstance(other, Compartment) and self.name == other.name
                and len(self.dimensions) == len(other.dimensions)
                and all(a == b for a, b in zip(self.dimensions, other.dimensions)))


class SimulationConfig(BaseModel):
    model_config = ConfigDict(arbitrary_types_allowed=True)
    initializer: Initializer
    compartments: List[Compartment]
    parameters: Params

    @cached_property
    def idx(self):
        return SimpleNamespace(**{c.name: _Index(k, **vars(c.idx)) for k, c in enumerate(self.compartments)})

    @model_validator(mode="after")
    def _validate(self):
        names = [c.name for c in self.compartments]
        dup = {n for n in names if names.count(n) > 1}
        assert not dup, f"you can not have two identically named compartments, found shared names: {dup}"
        seen = {}
        for d in self.flatten_dims():
            if d.name in seen:
                assert d == seen[d.name], (
                    f"dimension {d.name} has different definitions across different compartments, if this "
                    "intended, make the dimensions have different names")
            seen[d.name] = d
        strains = self.parameters.transmission_params.strains
        for d in self.flatten_dims():
            if isinstance(d, ImmuneHistoryDimension):
                assert isinstance(d, (FullStratifiedImmuneHistoryDimension, LastStrainImmuneHistoryDimension))
                assert type(d)(strains) == d, (
                    "Found immune states that dont correlate with strains from transmission_params")
        ages = [b for b in self.flatten_bins() if isinstance(b, AgeBin)]
        if any(s.introduction_ages is not None for s in strains):
            axis = next((d.bins for d in self.flatten_dims() if isinstance(d.bins[0], AgeBin)), [])
            assert len(axis) > 0, ("attempted to encode introduction_ages but could not find any age structure "
                                   "in the compartments")
            for s in strains:
                wanted = s.introduction_ages or []
                s.introduction_ages_mask_vector = [1 if b in wanted else 0 for b in axis]
        for s in strains:
            if s.is_introduced and s.introduction_ages is not None:
                assert all(a in ages for a in s.introduction_ages), (
                    f"{s.strain_name} attempts to introduce itself using {s.introduction_ages} age bins, but "
                    "those are not found within the age structure of the model.")
        return self

    def get_compartment(self, compartment_name: str) -> Compartment:
        for c in self.compartments:
            if c.name == compartment_name:
                return c
        raise AssertionError(f"Compartment with name {compartment_name} not found in model, found only these "
                             f"names: {[c.name for c in self.compartments]}")

    def flatten_bins(self) -> List[Bin]:
        return [b for c in self.compartments for d in c.dimensions for b in d.bins]

    def flatten_dims(self) -> List[Dimension]:
        return [d for c in self.compartments for d in c.dimensions]
