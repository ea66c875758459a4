"""Probabilistic-programming primitives for DynODE models, in torch.

The reference writes its models against numpyro (`numpyro.sample`, `numpyro.deterministic`, effect
handlers; reference src/dynode/infer/sample.py:72-79,158, examples/sir_infer_parameters.py:21-39), which is
absent from this image.  This module provides the same primitives with the same semantics so that a
DynODE model keeps its shape: it is written for ONE parameter draw; the inference engines evaluate it
for many chains / posterior draws at once with `torch.vmap`, and the ODE solve inside it becomes one
ensemble launch of the CUDA kernel (the vmap rule of `dynode_b200.simulation.autograd`).
"""

from __future__ import annotations

from collections import OrderedDict
from typing import Any, Callable, Dict, List, Optional

import torch

from . import distributions as dist

_STACK: List["Messenger"] = []


class PRNGKey:
    """Seed holder standing in for jax.random.PRNGKey: hands out torch generators per device.

    Unlike a JAX key it is stateful: every generator request advances a counter, so successive uses of
    one key give independent streams while a fresh `PRNGKey(seed)` reproduces the sequence."""

    def __init__(self, seed: int):
        self.seed = int(seed)
        self._count = 0

    def generator(self, device) -> torch.Generator:
        g = torch.Generator(device=device)
        g.manual_seed((self.seed * 1_000_003 + self._count) % (2**63 - 1))
        self._count += 1
        return g

    def fold_in(self, data: int) -> "PRNGKey":
        return PRNGKey((self.seed * 7_919 + int(data) + 1) % (2**63 - 1))

    def __repr__(self):
        return f"PRNGKey({self.seed})"


class Messenger:
    """Effect handler: a context manager that sees every primitive call made inside it."""

    def __init__(self, fn: Optional[Callable] = None):
        self.fn = fn

    def __enter__(self):
        _STACK.append(self)
        return self

    def __exit__(self, *exc):
        assert _STACK.pop() is self
        return False

    def process_message(self, msg: Dict[str, Any]) -> None:
        pass

    def postprocess_message(self, msg: Dict[str, Any]) -> None:
        pass

    def __call__(self, *args, **kwargs):
        with self:
            return self.fn(*args, **kwargs)


class trace(Messenger):
    """Record every site: `trace(fn).get_trace(*args)` -> OrderedDict name -> message."""

    def __enter__(self):
        self.trace: "OrderedDict[str, Dict[str, Any]]" = OrderedDict()
        return super().__enter__()

    def postprocess_message(self, msg):
        if msg["name"] in self.trace and msg["type"] in ("sample", "deterministic"):
            raise ValueError(f"all sites must have unique names but got `{msg['name']}` duplicated")
        self.trace[msg["name"]] = msg.copy()

    def get_trace(self, *args, **kwargs):
        self(*args, **kwargs)
        return self.trace


class seed(Messenger):
    """Provide randomness to sample sites that have no value yet."""

    def __init__(self, fn=None, rng_seed=None):
        super().__init__(fn)
        self.key = rng_seed if isinstance(rng_seed, PRNGKey) else PRNGKey(0 if rng_seed is None else rng_seed)

    def process_message(self, msg):
        if msg["type"] == "sample" and msg["value"] is None and msg["rng_key"] is None:
            msg["rng_key"] = self.key


class substitute(Messenger):
    """Fix the value of latent sample sites (constrained space) from `data`."""

    def __init__(self, fn=None, data: Optional[Dict[str, Any]] = None):
        super().__init__(fn)
        self.data = data or {}

    def process_message(self, msg):
        if msg["type"] in ("sample", "deterministic") and not msg["is_observed"] and msg["name"] in self.data:
            msg["value"] = self.data[msg["name"]]


class condition(Messenger):
    """Turn sample sites into observed sites with the given values."""

    def __init__(self, fn=None, data: Optional[Dict[str, Any]] = None):
        super().__init__(fn)
        self.data = data or {}

    def process_message(self, msg):
        if msg["type"] == "sample" and msg["name"] in self.data:
            msg["value"] = self.data[msg["name"]]
            msg["is_observed"] = True


def _total(x):
    """Sum over a site's own axes; a scalar site (0-dim, also under vmap) needs no reduction launch."""
    return x if isinstance(x, torch.Tensor) and x.dim() == 0 else x.sum()


class log_density_handler(Messenger):
    """Evaluate the joint log-density at unconstrained latent values.

    For every latent site: x = biject_to(support)(z), log p(x) + log|dx/dz|; observed sites and
    factors add their log-probability (numpyro.infer.util.log_density / potential_energy)."""

    def __init__(self, fn=None, unconstrained: Optional[Dict[str, torch.Tensor]] = None):
        super().__init__(fn)
        self.z = unconstrained or {}
        self.logp = 0.0
        self.constrained: Dict[str, torch.Tensor] = {}

    def process_message(self, msg):
        if msg["type"] == "sample" and not msg["is_observed"] and msg["name"] in self.z:
            z = self.z[msg["name"]]
            whole = dist.fused_site(msg["fn"], z)
            if whole is not None:  # bijector + log-Jacobian + prior log-density from one kernel
                x, lp = whole
                msg["prior_done"] = True
            else:
                x, lp = dist.constrain_with_ladj(msg["fn"].support, z)
            msg["value"] = x
            self.logp = self.logp + _total(lp)
            self.constrained[msg["name"]] = x

    def postprocess_message(self, msg):
        if msg["type"] == "sample":
            if not msg.get("prior_done", False):
                self.logp = self.logp + _total(msg["fn"].log_prob(msg["value"]))
        elif msg["type"] == "factor":
            self.logp = self.logp + _total(msg["value"])


def _apply_stack(msg: Dict[str, Any]) -> Dict[str, Any]:
    for h in reversed(_STACK):
        h.process_message(msg)
    if msg["type"] == "sample" and msg["value"] is None:
        key = msg["rng_key"]
        fn = msg["fn"]
        if getattr(fn, "raises_on_sample", False):
            fn.sample(None, msg["sample_shape"])  # PlaceholderSample: its own error, whatever the context
        if key is None and not _STACK:
            raise ValueError(
                f"site `{msg['name']}`: sampling outside an inference context needs rng_key=PRNGKey(...)")
        gen = None
        if key is not None and not torch._C._functorch.is_batchedtensor(fn._params()[0]):
            gen = key.generator(fn._device())
        msg["value"] = fn.sample(gen, msg["sample_shape"])
    for h in _STACK:
        h.postprocess_message(msg)
    return msg


def sample(name: str, fn: dist.Distribution, obs=None, rng_key: Optional[PRNGKey] = None, sample_shape=()):
    """numpyro.sample: draw (or observe) the site `name` ~ `fn`."""
    if not isinstance(fn, dist.Distribution):
        raise TypeError(f"site `{name}`: expected a Distribution, got {type(fn).__name__}")
    if obs is not None and not isinstance(obs, torch.Tensor):
        obs = torch.as_tensor(obs, dtype=torch.float64, device=fn._device())
    msg = {"type": "sample", "name": name, "fn": fn, "value": obs, "is_observed": obs is not None,
           "rng_key": rng_key, "sample_shape": tuple(sample_shape)}
    return _apply_stack(msg)["value"]


def deterministic(name: str, value):
    """numpyro.deterministic: record a derived quantity."""
    msg = {"type": "deterministic", "name": name, "value": value, "is_observed": False, "fn": None,
           "rng_key": None, "sample_shape": ()}
    return _apply_stack(msg)["value"]


def factor(name: str, log_factor):
    """numpyro.factor: add an arbitrary term to the log-density (used by the fused ODE likelihood)."""
    msg = {"type": "factor", "name": name, "value": log_factor, "is_observed": True, "fn": None,
           "rng_key": None, "sample_shape": ()}
    _apply_stack(msg)


def in_inference_context() -> bool:
    return bool(_STACK)


__all__ = ["PRNGKey", "Messenger", "trace", "seed", "substitute", "condition", "log_density_handler",
           "sample", "deterministic", "factor", "in_inference_context"]
