"""Sampling / resolving of parameter trees (API of reference src/dynode/infer/sample.py:18-197).

A DynODE parameter object (dict / pydantic model / nested lists) may hold prior distributions and
`DeterministicParameter` links.  `sample_then_resolve` returns a copy in which every distribution has
become a `sample` site and every link a `deterministic` site.  Site names follow the reference's
rules (pinned by its tests/test_infer/test_sample.py:49-151): keys are joined by "_", list positions
contribute their index, e.g. `strains_0_r0`.
"""

from __future__ import annotations

from copy import deepcopy
from typing import Any, Callable, Optional

import numpy as np
from pydantic import BaseModel

from ..config import DeterministicParameter
from . import ppl
from .distributions import Distribution


def _walk(node: Any, leaf: Callable[[Any, str], Any], path: str) -> Any:
    """Rebuild `node` with `leaf(value, site_name)` applied to every non-container value."""
    if isinstance(node, (BaseModel, dict)):
        items = {k: _walk(v, leaf, f"{path}{k}_") for k, v in dict(node).items()}
        return items if isinstance(node, dict) else type(node)(**items)
    if isinstance(node, (list, np.ndarray)):
        return [_walk(v, leaf, f"{path}{k}_") for k, v in enumerate(node)]
    return leaf(node, path[:-1] if path else path)


def sample_distributions(obj: Any, rng_key: Optional[ppl.PRNGKey] = None, _prefix: str = ""):
    """Replace every Distribution in `obj` by a draw from a sample site named after its position."""

    def leaf(value, name):
        if isinstance(value, Distribution):
            return ppl.sample(name, value, rng_key=rng_key)
        return value

    return _walk(obj, leaf, _prefix)


def resolve_deterministic(obj: Any, root_params, _prefix: str = ""):
    """Replace every DeterministicParameter in `obj` by the value it points to inside `root_params`
    (top-level keys only), recorded as a deterministic site."""
    scope = dict(root_params) if isinstance(root_params, BaseModel) else root_params

    def leaf(value, name):
        if isinstance(value, DeterministicParameter):
            return ppl.deterministic(name, value.resolve(scope))
        return value

    return _walk(obj, leaf, _prefix)


def sample_then_resolve(parameters: Any, rng_key: Optional[ppl.PRNGKey] = None, _prefix: str = ""):
    """Copy `parameters`, sample its distributions, then resolve its deterministic links."""
    sampled = sample_distributions(deepcopy(parameters), rng_key=rng_key, _prefix=_prefix)
    return resolve_deterministic(sampled, root_params=dict(sampled), _prefix=_prefix)
