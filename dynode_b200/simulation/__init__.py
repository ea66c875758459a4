"""Simulation entry points (API of reference src/dynode/simulation/__init__.py:3-8)."""

from .odes import (  # noqa: F401
    AbstractODEParams,
    SaveAt,
    Solution,
    build_saveat,
    simulate,
    simulate_ensemble,
    simulate_incidence_loglik,
)
