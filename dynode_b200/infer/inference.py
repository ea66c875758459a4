"""Inference processes (API of reference src/dynode/infer/inference.py:27-405).

`MCMCProcess` / `SVIProcess` keep the reference's fields and methods; underneath, the numpyro objects are
replaced by device-resident engines: `MCMC` drives `BatchedNUTS` (all chains in lock-step, one ODE-kernel
launch per leapfrog round) and `SVI` fits a full-rank normal guide with Adam on the single-particle ELBO.
"""

from __future__ import annotations

from typing import Any, Callable, Dict, Optional

import os
import torch
from pydantic import BaseModel, ConfigDict, Field, PositiveInt, PrivateAttr

from . import ppl
from .model_density import ModelDensity, Predictive
from .nuts import BatchedNUTS, effective_sample_size, split_rhat


def init_to_median(md: ModelDensity, num_chains: int) -> torch.Tensor:
    return md.init_to_median(num_chains)


def init_to_sample(md: ModelDensity, num_chains: int) -> torch.Tensor:
    return md.init_to_sample(num_chains)


class NUTS:
    """Kernel description (numpyro.infer.NUTS signature subset)."""

    def __init__(self, model: Callable, dense_mass: bool = True, max_tree_depth: int = 10,
                 init_strategy: Callable = init_to_median, target_accept_prob: float = 0.8,
                 step_size: float = 1.0, adapt_step_size: bool = True, adapt_mass_matrix: bool = True):
        self.model = model
        self.dense_mass, self.max_tree_depth, self.init_strategy = dense_mass, max_tree_depth, init_strategy
        self.target_accept_prob, self.step_size = target_accept_prob, step_size
        self.adapt_step_size, self.adapt_mass_matrix = adapt_step_size, adapt_mass_matrix


class MCMC:
    """numpyro.infer.MCMC stand-in: `run`, `get_samples`, `get_extra_fields`, `last_state`, `print_summary`."""

    def __init__(self, sampler: NUTS, num_warmup: int, num_samples: int, num_chains: int = 1,
                 progress_bar: bool = True, chain_method: str = "vectorized", cuda_graph: Optional[bool] = None,
                 sync_every: int = 4):
        self.sampler, self.num_warmup, self.num_samples, self.num_chains = sampler, num_warmup, num_samples, num_chains
        self.progress_bar = progress_bar
        self.chain_method = chain_method
        self.cuda_graph, self.sync_every = cuda_graph, sync_every
        self._samples_z: Optional[torch.Tensor] = None
        self._states: Dict[str, Dict[str, torch.Tensor]] = {}
        self._states_flat: Dict[str, Dict[str, torch.Tensor]] = {}
        self._sample_field = "z"
        self._extra: Dict[str, torch.Tensor] = {}
        self.last_state = None
        self.density: Optional[ModelDensity] = None
        self.engine: Optional[BatchedNUTS] = None

    def run(self, rng_key: Optional[ppl.PRNGKey] = None, *args, init_z: Optional[torch.Tensor] = None, **kwargs):
        import time
        key = rng_key or ppl.PRNGKey(0)
        t_start = time.perf_counter()
        md = ModelDensity(self.sampler.model, args, kwargs, rng_key=key.fold_in(11))
        self.density = md
        z0 = init_z if init_z is not None else self.sampler.init_strategy(md, self.num_chains)
        z0 = z0.to(md.device)
        t_setup = time.perf_counter()
        gen = key.fold_in(12).generator(md.device)
        if md.device.type == "cuda":
            torch.cuda.manual_seed(key.fold_in(13).seed % (2**63 - 1))  # graph-replayed rounds draw from it
        s = self.sampler
        eng = BatchedNUTS(md.potential_and_grad, max_tree_depth=s.max_tree_depth,
                          target_accept_prob=s.target_accept_prob, dense_mass=s.dense_mass,
                          step_size=s.step_size, adapt_step_size=s.adapt_step_size,
                          adapt_mass_matrix=s.adapt_mass_matrix, generator=gen, cuda_graph=self.cuda_graph,
                          sync_every=self.sync_every,
                          # DYNODE_B200_FIXED_LAUNCH=1: keep the launches chosen for all chains running (A/B knob)
                          launch_key=None if os.environ.get("DYNODE_B200_FIXED_LAUNCH") else md.launch_key)
        self.engine = eng
        progress = None
        if self.progress_bar:
            total = self.num_warmup + self.num_samples

            def progress(t, e):
                phase = "warmup" if t < self.num_warmup else "sample"
                print(f"[dynode_b200.infer] {phase} {t + 1}/{total}  chains={self.num_chains}  "
                      f"rounds={e.rounds}  mean step_size={float(e.step_size.mean()):.3g}  "
                      f"mean accept={float(e.stats['accept_prob'].mean()):.2f}  "
                      f"mean leapfrogs={float(e.stats['num_steps'].mean()):.1f}  graph={e.graph_used}")

        if md.device.type == "cuda":
            from .. import engine as _engine
            _engine.adjoint_overflows(reset=True)
        z, extra, last = eng.run(z0, self.num_warmup, self.num_samples, progress)
        if md.device.type == "cuda":
            # rows beyond the adjoint's checkpoint capacity were re-evaluated by forward sensitivities where they
            # occurred (simulation/autograd.py); the count is kept as a diagnostic
            self.adjoint_fallbacks = _engine.adjoint_overflows(reset=True)
        self._samples_z, self._extra, self.last_state = z, extra, last
        t_sampled = time.perf_counter()  # eng.run ends on host reads of device counters: the device is idle here
        C, N, D = z.shape
        flat = md.constrain(z.reshape(C * N, D), with_deterministic=True)
        latent = set(md.sites)
        self._states_flat = {"z": flat}
        self._states = {"z": {k: v.reshape(C, N, *v.shape[1:]) for k, v in flat.items()}}
        self._latent = latent
        if md.device.type == "cuda":
            torch.cuda.synchronize()
        # where a run's wall time went (seconds): model trace + initial positions | first evaluation (compiles the
        # potential) | graph capture | the rounds | constraining the draws
        self.timing = {"setup_s": t_setup - t_start, **eng.timing, "constrain_s": time.perf_counter() - t_sampled}
        return self

    def get_samples(self, group_by_chain: bool = False) -> Dict[str, torch.Tensor]:
        src = self._states["z"] if group_by_chain else self._states_flat["z"]
        return {k: v for k, v in src.items() if k in self._latent}

    def get_extra_fields(self, group_by_chain: bool = False) -> Dict[str, torch.Tensor]:
        return {k: (v if group_by_chain else v.reshape(-1)) for k, v in self._extra.items()}

    def summary(self) -> Dict[str, Dict[str, float]]:
        out = {}
        for name, v in self._states["z"].items():
            if name not in self._latent:
                continue
            x = v.reshape(v.shape[0], v.shape[1], -1)
            for j in range(x.shape[2]):
                col = x[:, :, j]
                key = name if x.shape[2] == 1 else f"{name}[{j}]"
                q = torch.quantile(col.reshape(-1), torch.tensor([0.05, 0.5, 0.95], dtype=col.dtype, device=col.device))
                out[key] = {"mean": float(col.mean()), "std": float(col.std()), "median": float(q[1]),
                            "5.0%": float(q[0]), "95.0%": float(q[2]),
                            "n_eff": float(effective_sample_size(col)) if col.shape[1] >= 4 else float("nan"),
                            "r_hat": float(split_rhat(col)) if col.shape[1] >= 4 else float("nan")}
        return out

    def print_summary(self) -> None:
        rows = self.summary()
        print(f"{'':>32s} {'mean':>10s} {'std':>10s} {'median':>10s} {'5.0%':>10s} {'95.0%':>10s} {'n_eff':>10s} {'r_hat':>7s}")
        for k, r in rows.items():
            print(f"{k:>32s} {r['mean']:10.4f} {r['std']:10.4f} {r['median']:10.4f} {r['5.0%']:10.4f} "
                  f"{r['95.0%']:10.4f} {r['n_eff']:10.1f} {r['r_hat']:7.3f}")
        div = int(self._extra["diverging"].sum()) if "diverging" in self._extra else 0
        print(f"Number of divergences: {div}")


class Adam:
    """numpyro.optim.Adam stand-in (step_size = learning rate)."""

    def __init__(self, step_size: float = 0.1, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
        self.step_size, self.b1, self.b2, self.eps = step_size, b1, b2, eps

    def build(self, params):
        return torch.optim.Adam(params, lr=self.step_size, betas=(self.b1, self.b2), eps=self.eps)


class AutoContinuous:
    """Base of the automatic guides over the unconstrained latent vector."""

    def __init__(self, model: Callable, init_loc_fn: Callable = init_to_median, **kwargs):
        self.model, self.init_loc_fn, self.kwargs = model, init_loc_fn, kwargs
        self.loc: Optional[torch.Tensor] = None

    def setup(self, md: ModelDensity):
        raise NotImplementedError

    def parameters(self):
        raise NotImplementedError

    def rsample(self, n: int, generator) -> torch.Tensor:
        raise NotImplementedError

    def entropy(self) -> torch.Tensor:
        raise NotImplementedError


class AutoMultivariateNormal(AutoContinuous):
    """q(z) = N(loc, L L^T), L lower triangular with positive diagonal (init_scale = 0.1)."""

    def setup(self, md: ModelDensity):
        D = md.dim
        init_scale = float(self.kwargs.get("init_scale", 0.1))
        self.loc = self.init_loc_fn(md, 1)[0].clone().requires_grad_(True)
        self.log_diag = torch.full((D,), float(torch.log(torch.tensor(init_scale))), dtype=torch.float64,
                                   device=md.device, requires_grad=True)
        self.off = torch.zeros((D, D), dtype=torch.float64, device=md.device, requires_grad=True)

    def scale_tril(self) -> torch.Tensor:
        return torch.tril(self.off, -1) + torch.diag(torch.exp(self.log_diag))

    def parameters(self):
        return [self.loc, self.log_diag, self.off]

    def rsample(self, n: int, generator) -> torch.Tensor:
        eps = torch.randn((n, self.loc.shape[0]), dtype=torch.float64, device=self.loc.device, generator=generator)
        return self.loc + eps @ self.scale_tril().T

    def entropy(self) -> torch.Tensor:
        D = self.loc.shape[0]
        return self.log_diag.sum() + 0.5 * D * (1.0 + torch.log(torch.tensor(2.0 * torch.pi, dtype=torch.float64)))


class AutoNormal(AutoMultivariateNormal):
    """Mean-field normal guide."""

    def scale_tril(self) -> torch.Tensor:
        return torch.diag(torch.exp(self.log_diag))

    def parameters(self):
        return [self.loc, self.log_diag]


class SVIRunResult:
    def __init__(self, params: Dict[str, torch.Tensor], losses: torch.Tensor):
        self.params, self.losses = params, losses


class SVI:
    """Stochastic variational inference with the single-particle Trace_ELBO."""

    def __init__(self, model: Callable, guide: AutoContinuous, optim: Adam, num_particles: int = 1):
        self.model, self.guide, self.optim, self.num_particles = model, guide, optim, num_particles
        self.density: Optional[ModelDensity] = None

    def run(self, rng_key: Optional[ppl.PRNGKey], num_steps: int, *args, progress_bar: bool = True, **kwargs):
        key = rng_key or ppl.PRNGKey(0)
        md = ModelDensity(self.model, args, kwargs, rng_key=key.fold_in(21))
        self.density = md
        self.guide.setup(md)
        opt = self.optim.build(self.guide.parameters())
        gen = key.fold_in(22).generator(md.device)
        losses = torch.empty(num_steps, dtype=torch.float64, device=md.device)
        every = max(1, num_steps // 10)
        for it in range(num_steps):
            opt.zero_grad(set_to_none=True)
            z = self.guide.rsample(self.num_particles, gen)
            # -ELBO = E_q[U(z)] - H[q] with z = loc + L eps.  U and dU/dz come from the model's own evaluation path
            # (three launches for a model of the compiled form, infer/potential_plan.py -- not ~35 forward and as
            # many backward through autograd); the guide's parameters get their gradient through the surrogate
            # sum(z * dU/dz) / n, whose derivative is E_q[dU/dz . dz/dparams].
            U, dU = md.potential_and_grad(z.detach())
            entropy = self.guide.entropy()
            loss = U.mean() - entropy.detach()
            ((z * dU).sum() / z.shape[0] - entropy).backward()
            # A non-finite loss (a solve that ran out of max_steps reports NaN) must not reach Adam: one NaN
            # gradient poisons loc / scale for good.  Its gradients are zeroed on the device (no sync), the step is
            # recorded, and the run stops at the next check below.
            ok = torch.isfinite(loss) & torch.isfinite(dU).all()
            loss = torch.where(ok, loss, torch.full_like(loss, float("nan")))
            for p in self.guide.parameters():
                if p.grad is not None:
                    p.grad = torch.where(ok, p.grad, torch.zeros_like(p.grad))
            opt.step()
            losses[it] = loss.detach()
            if (it + 1) % every == 0 or it + 1 == num_steps:
                bad = int((~torch.isfinite(losses[:it + 1])).sum())
                if bad:
                    raise RuntimeError(
                        f"SVI: {bad} of the first {it + 1} ELBO estimates were not finite (a model evaluation "
                        "returned NaN/inf: e.g. an ODE solve that reached `max_steps`); the optimiser state was "
                        "protected, the run is stopped")
                if progress_bar:
                    print(f"[dynode_b200.infer] svi {it + 1}/{num_steps}  loss={float(loss):.4f}")
        params = {"auto_loc": self.guide.loc.detach().clone(), "auto_scale_tril": self.guide.scale_tril().detach().clone()}
        return SVIRunResult(params, losses)

    def sample_posterior(self, rng_key: ppl.PRNGKey, num_samples: int, with_deterministic: bool = False):
        gen = rng_key.fold_in(23).generator(self.density.device)
        with torch.no_grad():
            z = self.guide.rsample(num_samples, gen)
        return self.density.constrain(z, with_deterministic=with_deterministic)


def log_likelihood(model: Callable, posterior_samples: Dict[str, torch.Tensor], *args, **kwargs):
    """Per-draw log-probability of every observed site (numpyro.infer.util.log_likelihood)."""
    md = ModelDensity(model, args, kwargs)
    given = {k: torch.as_tensor(v, dtype=torch.float64).to(md.device) for k, v in posterior_samples.items()
             if k in md.sites}

    def one(values):
        with ppl.substitute(data=values), ppl.trace() as tr:
            model(*args, **kwargs)
        return {n: m["fn"].log_prob(m["value"]) for n, m in tr.trace.items()
                if m["type"] == "sample" and m["is_observed"]}

    with torch.no_grad():
        return torch.vmap(one)(given)


def _to_inference_data(posterior, prior, posterior_predictive, sample_stats=None, log_lik=None):
    groups = {"posterior": posterior, "prior": prior, "posterior_predictive": posterior_predictive}
    if sample_stats is not None:
        groups["sample_stats"] = sample_stats
    if log_lik is not None:
        groups["log_likelihood"] = log_lik
    try:  # arviz is not in this image; where it is, hand back a real InferenceData
        import arviz as az  # type: ignore

        def np_(d, chains):
            return {k: v.detach().cpu().numpy().reshape(chains, -1, *v.shape[1:]) if chains else
                    v.detach().cpu().numpy()[None] for k, v in d.items()}

        return az.from_dict(**{g: np_(d, 0) for g, d in groups.items()})
    except ImportError:
        return groups


class InferenceProcess(BaseModel):
    """Abstract inference process fitting a DynODE model to data (reference inference.py:27-116)."""

    model_config = ConfigDict(arbitrary_types_allowed=True)
    numpyro_model: Callable = Field(description="Model that samples and resolves parameters, solves the "
                                                "ODEs and optionally compares to observed data.")
    inference_prngkey: Any = Field(default_factory=lambda: ppl.PRNGKey(8675314))
    _inference_complete: bool = PrivateAttr(default=False)
    _inferer: Optional[Any] = PrivateAttr(default=None)
    _inference_state: Optional[Any] = PrivateAttr(default=None)
    _inferer_kwargs: Optional[dict] = PrivateAttr(default_factory=dict)

    def infer(self, **kwargs):
        raise NotImplementedError("Inference process not implemented, please use a subclass.")

    def get_samples(self, group_by_chain=False, exclude_deterministic=True) -> Dict[str, torch.Tensor]:
        raise NotImplementedError("get_samples() process not implemented, please use a subclass.")

    def to_arviz(self):
        raise NotImplementedError("to_arviz not implemented for abstract InferenceProcess, use subclass")

    def _require_fit(self):
        if not self._inference_complete:
            raise AssertionError("Inference process not completed, please call infer() first.")


class MCMCProcess(InferenceProcess):
    """Fit with NUTS (reference inference.py:119-241): dense mass matrix, init_to_median."""

    num_samples: PositiveInt
    num_warmup: PositiveInt
    num_chains: PositiveInt
    nuts_max_tree_depth: PositiveInt
    nuts_init_strategy: Callable = init_to_median
    mcmc_kwargs: dict = Field(default_factory=dict)
    nuts_kwargs: dict = Field(default_factory=dict)
    progress_bar: bool = True

    def infer(self, **kwargs) -> MCMC:
        inferer = MCMC(
            NUTS(self.numpyro_model, dense_mass=True, max_tree_depth=self.nuts_max_tree_depth,
                 init_strategy=self.nuts_init_strategy, **self.nuts_kwargs),
            num_warmup=self.num_warmup, num_samples=self.num_samples, num_chains=self.num_chains,
            progress_bar=self.progress_bar, **self.mcmc_kwargs)
        inferer.run(self.inference_prngkey, **kwargs)
        self._inference_complete = True
        self._inferer = inferer
        self._inference_state = inferer.last_state
        self._inferer_kwargs = kwargs
        return inferer

    def get_samples(self, group_by_chain=False, exclude_deterministic=True) -> Dict[str, torch.Tensor]:
        self._require_fit()
        if exclude_deterministic:
            return self._inferer.get_samples(group_by_chain=group_by_chain)
        if group_by_chain:
            return self._inferer._states[self._inferer._sample_field]
        return self._inferer._states_flat[self._inferer._sample_field]

    def to_arviz(self):
        self._require_fit()
        post = self.get_samples()
        pp = Predictive(self.numpyro_model, posterior_samples=post)(self.inference_prngkey, **self._inferer_kwargs)
        prior = Predictive(self.numpyro_model, num_samples=self.num_samples)(self.inference_prngkey,
                                                                              **self._inferer_kwargs)
        return _to_inference_data(post, prior, pp, sample_stats=self._inferer.get_extra_fields())


class SVIProcess(InferenceProcess):
    """Fit with SVI (reference inference.py:244-405): AutoMultivariateNormal guide, Adam(0.1), Trace_ELBO."""

    num_iterations: PositiveInt
    num_samples: PositiveInt
    guide_class: Any = AutoMultivariateNormal
    guide_init_strategy: Callable = init_to_median
    optimizer: Any = Field(default_factory=lambda: Adam(step_size=0.1))
    progress_bar: bool = True
    guide_kwargs: dict = Field(default_factory=dict)

    def infer(self, **kwargs) -> SVI:
        guide = self.guide_class(self.numpyro_model, init_loc_fn=self.guide_init_strategy, **self.guide_kwargs)
        inferer = SVI(model=self.numpyro_model, guide=guide, optim=self.optimizer)
        self._inference_state = inferer.run(self.inference_prngkey, self.num_iterations,
                                            progress_bar=self.progress_bar, **kwargs)
        self._inference_complete = True
        self._inferer = inferer
        self._inferer_kwargs = kwargs
        return inferer

    def get_samples(self, _: bool = False, exclude_deterministic: bool = True) -> Dict[str, torch.Tensor]:
        self._require_fit()
        samples = self._inferer.sample_posterior(self.inference_prngkey, self.num_samples,
                                                 with_deterministic=not exclude_deterministic)
        return {k: v for k, v in samples.items() if not k.startswith("_auto_")}

    def to_arviz(self):
        self._require_fit()
        post = self.get_samples()
        pp = Predictive(self.numpyro_model, posterior_samples=post)(self.inference_prngkey, **self._inferer_kwargs)
        prior = Predictive(self.numpyro_model, num_samples=self.num_iterations)(self.inference_prngkey,
                                                                                 **self._inferer_kwargs)
        ll = log_likelihood(self.numpyro_model, post, **self._inferer_kwargs)
        return _to_inference_data(post, prior, pp, log_lik=ll)
