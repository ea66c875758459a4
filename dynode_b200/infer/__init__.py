"""Inference layer (API of reference src/dynode/infer/__init__.py:3-19) on device-resident engines."""

from . import distributions  # noqa: F401
from .checkpointing import checkpoint_compartment_sizes  # noqa: F401
from .inference import (  # noqa: F401
    MCMC,
    NUTS,
    SVI,
    Adam,
    AutoMultivariateNormal,
    AutoNormal,
    InferenceProcess,
    MCMCProcess,
    SVIProcess,
    init_to_median,
    init_to_sample,
    log_likelihood,
)
from .model_density import ModelDensity, Predictive  # noqa: F401
from .nuts import BatchedNUTS, build_adaptation_schedule, effective_sample_size, split_rhat  # noqa: F401
from . import ppl  # noqa: F401  (`ppl.sample`, `ppl.deterministic`, `ppl.factor`, handlers: numpyro's role)
from .ppl import PRNGKey  # noqa: F401
from .sample import resolve_deterministic, sample_distributions, sample_then_resolve  # noqa: F401

__all__ = [
    "sample_then_resolve", "resolve_deterministic", "sample_distributions", "InferenceProcess", "MCMCProcess",
    "SVIProcess", "checkpoint_compartment_sizes", "MCMC", "NUTS", "SVI", "Predictive", "PRNGKey", "ppl",
    "distributions",
]
