"""Turn a per-draw DynODE model into batched device functions: potential energy + gradient over chains,
initial values, constrained samples, predictive draws.

This is the glue numpyro provides through `initialize_model` / `potential_energy` / `Predictive`
(used by the reference at src/dynode/infer/inference.py:149-163, 225-235).  The model is plain Python
written for ONE draw; `torch.vmap` evaluates it for all chains at once, so the ODE solve inside it is a
single ensemble launch, and reverse-mode autograd through the vmapped model yields the gradient NUTS asks
for (the ODE part of that gradient comes out of the CUDA kernel's forward sensitivities).
"""

from __future__ import annotations

from collections import OrderedDict
from typing import Any, Callable, Dict, Optional, Tuple

import torch

from . import distributions as dist
from . import ppl


class _Discover(ppl.Messenger):
    """Run the model once: latent sites get the median of `n_med` prior draws (init_to_median); the
    distribution objects are kept so that more draws can be taken without re-running the model."""

    def __init__(self, key: ppl.PRNGKey, n_med: int = 15):
        super().__init__(None)
        self.key, self.n_med = key, n_med
        self.latent: "OrderedDict[str, Dict[str, Any]]" = OrderedDict()

    def process_message(self, msg):
        if msg["type"] == "sample" and not msg["is_observed"] and msg["value"] is None:
            fn = msg["fn"]
            draws = fn.sample(self.key.generator(fn._device()), (self.n_med,))
            msg["value"] = draws.median(dim=0).values
            self.latent[msg["name"]] = {"fn": fn, "shape": tuple(msg["value"].shape)}


class ModelDensity:
    def __init__(self, model: Callable, model_args: Tuple = (), model_kwargs: Optional[Dict] = None,
                 rng_key: Optional[ppl.PRNGKey] = None, device=None, allow_discrete: bool = False):
        self.model, self.args, self.kwargs = model, tuple(model_args), dict(model_kwargs or {})
        self.key = rng_key or ppl.PRNGKey(0)
        disc = _Discover(self.key.fold_in(1))
        with disc, ppl.trace() as tr:
            model(*self.args, **self.kwargs)
        self.prototype = tr.trace
        self.sites = disc.latent
        if not self.sites:
            raise ValueError("the model has no latent sample sites: nothing to infer")
        self.discrete = [n for n, info in self.sites.items() if info["fn"].is_discrete]
        if self.discrete and not allow_discrete:
            raise NotImplementedError(f"latent sites {self.discrete} are discrete; NUTS/SVI need continuous sites")
        for n in self.discrete:  # predictive mode: sampled inside the model run, never part of z
            del self.sites[n]
        self.device = device
        if self.device is None:
            self.device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() \
                else torch.device("cpu")
        off = 0
        for info in self.sites.values():
            n = 1
            for d in info["shape"]:
                n *= d
            info["slice"] = (off, off + n)
            off += n
        self.dim = off
        self._plan, self._plan_tried, self.plan_reason = None, False, "not compiled yet"

    # ------------------------------------------------------------------ packing
    def unpack(self, z: torch.Tensor) -> Dict[str, torch.Tensor]:
        """z [..., D] -> {site: [..., *shape]} (unconstrained)."""
        out = {}
        for name, info in self.sites.items():
            a, b = info["slice"]
            out[name] = z[..., a:b].reshape(tuple(z.shape[:-1]) + tuple(info["shape"]))
        return out

    def init_to_median(self, num_chains: int, num_draws: int = 15) -> torch.Tensor:
        """Per chain, the median of `num_draws` prior draws of every site, in unconstrained space
        (numpyro init_to_median(num_samples=15), the reference's nuts_init_strategy)."""
        cols = []
        for name, info in self.sites.items():
            fn = info["fn"].to(self.device)
            draws = fn.sample(self.key.generator(self.device), (num_chains, num_draws))
            med = draws.median(dim=1).values
            t = dist.biject_to(fn.support)
            cols.append(t.inv(med).reshape(num_chains, -1))
        return torch.cat(cols, dim=1).to(torch.float64)

    def init_to_sample(self, num_chains: int) -> torch.Tensor:
        cols = []
        for name, info in self.sites.items():
            fn = info["fn"].to(self.device)
            x = fn.sample(self.key.generator(self.device), (num_chains,))
            cols.append(dist.biject_to(fn.support).inv(x).reshape(num_chains, -1))
        return torch.cat(cols, dim=1).to(torch.float64)

    def prior_draws(self, num_samples: int) -> Dict[str, torch.Tensor]:
        """Constrained draws of every latent site from its prior (sites are taken as independent given
        the prototype run, which is how DynODE configs declare priors)."""
        return {name: info["fn"].to(self.device).sample(self.key.generator(self.device), (num_samples,))
                for name, info in self.sites.items()}

    # ------------------------------------------------------------------ densities
    def _potential_one(self, z: torch.Tensor) -> torch.Tensor:
        h = ppl.log_density_handler(unconstrained=self.unpack(z))
        with h:
            self.model(*self.args, **self.kwargs)
        lp = h.logp
        return -lp if isinstance(lp, torch.Tensor) else -torch.as_tensor(lp, dtype=z.dtype, device=z.device)

    def potential(self, Z: torch.Tensor) -> torch.Tensor:
        return torch.vmap(self._potential_one)(Z)

    def potential_and_grad(self, Z: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """U [C], dU/dz [C, D] for unconstrained positions Z [C, D].  Models of the compiled form (constant-parameter
        priors on scalar sites, monomial rates, one fused ODE likelihood) take three launches
        (infer/potential_plan.py); everything else one vmapped evaluation of the Python model."""
        if self._plan is None and not self._plan_tried and Z.is_cuda:
            self._plan_tried = True
            from .potential_plan import compile_plan
            try:
                self._plan, self.plan_reason = compile_plan(self)
            except Exception as exc:  # discovery runs user code: whatever it raises, the composed path still stands
                import warnings
                self._plan, self.plan_reason = None, f"compiling the potential failed: {type(exc).__name__}: {exc}"
                warnings.warn(f"dynode_b200: {self.plan_reason}; evaluating the model through torch.vmap instead")
        if self._plan is not None and Z.is_cuda:
            return self._plan.potential_and_grad(Z.detach())
        return self.potential_and_grad_composed(Z)

    def launch_key(self, C: int, n_rows: int):
        """Hashable description of the launches `potential_and_grad` makes for C rows of which only n_rows run under
        an `engine.only_rows` mask (None: no choice depends on it)."""
        return None if self._plan is None else self._plan.launch_key(C, n_rows)

    def potential_and_grad_composed(self, Z: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """The same through the vmapped Python model and autograd (~35 launches for a DynODE model)."""
        Zr = Z.detach().requires_grad_(True)
        with torch.enable_grad():
            U = self.potential(Zr)
            (g,) = torch.autograd.grad(U.sum(), Zr)
        return U.detach(), g

    # ------------------------------------------------------------------ constrained values
    def _constrain_one(self, z: torch.Tensor, with_deterministic: bool):
        h = ppl.log_density_handler(unconstrained=self.unpack(z))
        with h, ppl.trace() as tr:
            self.model(*self.args, **self.kwargs)
        out = dict(h.constrained)
        if with_deterministic:
            for name, msg in tr.trace.items():
                if msg["type"] == "deterministic" and isinstance(msg["value"], torch.Tensor):
                    out[name] = msg["value"]
        return out

    def constrain(self, Z: torch.Tensor, with_deterministic: bool = False, chunk: int = 65536):
        """Unconstrained Z [N, D] -> {site: [N, *shape]} in the model's (constrained) space."""
        parts = []
        with torch.no_grad():
            for lo in range(0, Z.shape[0], chunk):
                parts.append(torch.vmap(lambda z: self._constrain_one(z, with_deterministic))(Z[lo:lo + chunk]))
        return {k: torch.cat([p[k] for p in parts], 0) for k in parts[0]}

    def unconstrain(self, samples: Dict[str, torch.Tensor]) -> torch.Tensor:
        cols = []
        for name, info in self.sites.items():
            x = samples[name].to(self.device)
            N = x.shape[0]
            cols.append(dist.biject_to(info["fn"].to(self.device).support).inv(x).reshape(N, -1))
        return torch.cat(cols, 1)


class Predictive:
    """numpyro.infer.Predictive for DynODE models: run the model for every posterior (or prior) draw in
    ONE vmapped pass and return the sites that were not given (observation sites are sampled)."""

    def __init__(self, model: Callable, posterior_samples: Optional[Dict[str, torch.Tensor]] = None,
                 num_samples: Optional[int] = None, exclude_deterministic: bool = True,
                 return_sites: Optional[list] = None, chunk: int = 16384):
        if posterior_samples is None and num_samples is None:
            raise ValueError("either posterior_samples or num_samples must be given")
        self.model, self.posterior_samples, self.num_samples = model, posterior_samples, num_samples
        self.exclude_deterministic, self.return_sites, self.chunk = exclude_deterministic, return_sites, chunk

    def __call__(self, rng_key: Optional[ppl.PRNGKey] = None, *args, **kwargs) -> Dict[str, torch.Tensor]:
        key = rng_key or ppl.PRNGKey(0)
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() \
            else torch.device("cpu")
        given = {k: torch.as_tensor(v, dtype=torch.float64).to(device)
                 for k, v in (self.posterior_samples or {}).items()}
        N = next(iter(given.values())).shape[0] if given else int(self.num_samples)
        torch.manual_seed(key.fold_in(3).seed % (2**31))

        def one(_, values):
            # sites without a given value are drawn inside the run (vmap randomness="different"), so priors
            # that depend on other sites and observation noise are sampled from the right conditional
            with ppl.substitute(data=values), ppl.trace() as tr:
                self.model(*args, **kwargs)
            out = {}
            for name, msg in tr.trace.items():
                if not isinstance(msg["value"], torch.Tensor):
                    continue
                if self.return_sites is not None:
                    if name in self.return_sites:
                        out[name] = msg["value"]
                elif msg["type"] == "sample" and name not in given:  # observed ones too, like numpyro
                    out[name] = msg["value"]
                elif msg["type"] == "deterministic" and not self.exclude_deterministic:
                    out[name] = msg["value"]
            return out

        parts = []
        with torch.no_grad():
            for lo in range(0, N, self.chunk):
                hi = min(N, lo + self.chunk)
                sl = {k: v[lo:hi] for k, v in given.items()}
                parts.append(torch.vmap(one, randomness="different")(torch.arange(lo, hi, device=device), sl))
        return {k: torch.cat([p[k] for p in parts], 0) for k in parts[0]}
