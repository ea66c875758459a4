"""pytest plugin (test infrastructure): lets the REFERENCE's own CPU test files (read from /root/reference at
test time, never copied) import this repository's re-created API under the reference's module names.

    dynode, dynode.config, dynode.infer, dynode.simulation, dynode.typing  -> dynode_b200.*
    numpyro.distributions / numpyro.handlers / numpyro.sample|deterministic -> dynode_b200.infer (ppl, distributions)
    jax.random.PRNGKey -> dynode_b200.infer.PRNGKey;  jax.Array -> torch.Tensor;  jax.numpy -> a thin torch facade
"""
import sys
import types

import numpy as np
import torch

import dynode_b200
from dynode_b200 import config, infer, simulation, typing as dtyping
from dynode_b200.infer import distributions, ppl


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    sys.modules["dynode"] = dynode_b200
    for sub, mod in (("config", config), ("infer", infer), ("simulation", simulation), ("typing", dtyping)):
        sys.modules[f"dynode.{sub}"] = mod
    for leaf in ("bins", "dimension", "params", "strains", "simulation_config", "deterministic_parameter",
                 "placeholder_sample", "simulation_date", "initializer"):
        sys.modules[f"dynode.config.{leaf}"] = __import__(f"dynode_b200.config.{leaf}", fromlist=["x"])
    handlers = _module("numpyro.handlers", trace=ppl.trace, substitute=ppl.substitute, seed=ppl.seed,
                       condition=ppl.condition)
    numpyro = _module("numpyro", distributions=distributions, handlers=handlers, sample=ppl.sample,
                      deterministic=ppl.deterministic, factor=ppl.factor)
    sys.modules["numpyro.distributions"] = distributions

    def asarray(x, dtype=None):
        return torch.as_tensor(np.asarray(x, dtype=np.float64))

    jnp = _module("jax.numpy", array=asarray, asarray=asarray, ndarray=torch.Tensor, float64=torch.float64,
                  allclose=lambda a, b, **kw: bool(np.allclose(np.asarray(a), np.asarray(b), **kw)),
                  sum=lambda x, axis=None: torch.as_tensor(x).sum() if axis is None else torch.as_tensor(x).sum(axis),
                  einsum=lambda spec, *ops: torch.einsum(spec, *[torch.as_tensor(o) for o in ops]))
    random = _module("jax.random", PRNGKey=infer.PRNGKey)
    _module("jax", numpy=jnp, random=random, Array=torch.Tensor, jit=lambda f: f)


install()
