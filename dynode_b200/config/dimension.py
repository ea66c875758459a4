"""Compartment dimensions (API of reference src/dynode/config/dimension.py:21-244)."""

from itertools import combinations
from math import isinf
from types import SimpleNamespace
from typing import List

from pydantic import BaseModel, field_validator, model_validator

from ..typing import DynodeName
from .bins import Bin, DiscretizedPositiveIntBin, WaneBin
from .strains import Strain


class Dimension(BaseModel):
    """A named axis of a compartment: an ordered list of bins of one type with unique names."""

    name: DynodeName
    bins: List[Bin]

    def __len__(self):
        return len(self.bins)

    @property
    def idx(self):
        """bin name -> position along this axis."""
        return SimpleNamespace(**{b.name: k for k, b in enumerate(self.bins)})

    @field_validator("bins", mode="after")
    @classmethod
    def _check_bins(cls, bins):
        assert len(bins) > 0, "can not have dimension with no bins"
        kinds = [type(b) for b in bins]
        assert all(k is kinds[0] for k in kinds), (
            f"can not instantiate dimension with mixed type bins. Found list of types {kinds}")
        names = [b.name for b in bins]
        assert len(set(names)) == len(names), "Dimension of categorical bins must have unique bin names."
        if all(isinstance(b, DiscretizedPositiveIntBin) for b in bins):
            assert bins == sorted(bins, key=lambda b: b.min_value), (
                f"Any dimension made up of DiscretizedIntBins must be sorted, got {bins}")
            for lo, hi in zip(bins[:-1], bins[1:]):
                assert lo.max_value < hi.min_value, "DiscretizedPositiveIntBin within a dimension can not overlap."
                assert lo.max_value + 1 == hi.min_value, (
                    f"dimensions containing DiscretizedPositiveIntBin can not have gaps between them, "
                    f"found one between {lo} and {hi}")
        return bins


class VaccinationDimension(Dimension):
    """Ordinal dose counts v0..vN (+1 bin when a seasonal dose is tracked)."""

    seasonal_vaccination: bool = False

    def __init__(self, max_ordinal_vaccinations: int, seasonal_vaccination: bool = False, name: DynodeName = "vax"):
        n = max_ordinal_vaccinations + (1 if seasonal_vaccination else 0)
        super().__init__(name=name, bins=[DiscretizedPositiveIntBin(k, k, name=f"v{k}") for k in range(n + 1)])
        self.seasonal_vaccination = seasonal_vaccination

    @property
    def max_shots(self) -> int:
        return len(self.bins) - 1


class ImmuneHistoryDimension(Dimension):
    """Tracks which strains a population has recovered from."""


class FullStratifiedImmuneHistoryDimension(ImmuneHistoryDimension):
    """Every subset of the strains: none, x, y, x_y, ..."""

    def __init__(self, strains: List[Strain], name: DynodeName = "hist") -> None:
        assert len(strains) > 0, "Must pass at least one strain to immune history dimension."
        names = [s.strain_name for s in strains]
        bins = [Bin(name="none")]
        for k in range(1, len(names) + 1):
            bins += [Bin(name="_".join(c)) for c in combinations(names, k)]
        super().__init__(name=name, bins=bins)


class LastStrainImmuneHistoryDimension(ImmuneHistoryDimension):
    """Only the most recent infecting strain: none, x, y, ..."""

    def __init__(self, strains: List[Strain], name: DynodeName = "hist") -> None:
        assert len(strains) > 0, "Must pass at least one strain to immune history dimension."
        super().__init__(name=name, bins=[Bin(name="none")] + [Bin(name=s.strain_name) for s in strains])


class WaneDimension(Dimension):
    """Waning stages W0..Wk; the last stage never wanes (waiting_time = inf)."""

    def __init__(self, waiting_times, base_protections, name="wane"):
        assert len(waiting_times) > 0, "Wane dimension must have at least one bin."
        assert len(waiting_times) == len(base_protections), "must pass equal length wait times and base protections"
        super().__init__(name=name, bins=[WaneBin(name=f"W{k}", waiting_time=w, base_protection=p)
                                          for k, (w, p) in enumerate(zip(waiting_times, base_protections))])

    @model_validator(mode="after")
    def _terminal_stage(self):
        assert isinf(self.bins[-1].waiting_time), "last wane bin should have math.inf waiting time"
        return self
