"""Inference layer on CPU: distributions/bijectors against scipy, site naming rules pinned by the reference's
tests/test_infer/test_sample.py, the process wrappers on a toy model (as the reference's
tests/test_infer/test_inference_processes.py does), NUTS against analytic posteriors."""
import math

import numpy as np
import pytest
import scipy.stats as st
import torch

from dynode_b200.config import DeterministicParameter
from dynode_b200.infer import (MCMCProcess, ModelDensity, Predictive, PRNGKey, SVIProcess, build_adaptation_schedule,
                               effective_sample_size, ppl, resolve_deterministic, sample_distributions,
                               sample_then_resolve, split_rhat)
from dynode_b200.infer import distributions as dist
from dynode_b200.infer.nuts import BatchedNUTS

X = torch.tensor([0.3, 1.7, 4.2], dtype=torch.float64)


@pytest.mark.parametrize("d,ref,x", [
    (dist.Normal(1.0, 2.0), st.norm(1.0, 2.0), X),
    (dist.LogNormal(0.5, 0.7), st.lognorm(0.7, scale=math.exp(0.5)), X),
    (dist.HalfNormal(1.5), st.halfnorm(scale=1.5), X),
    (dist.Exponential(0.8), st.expon(scale=1 / 0.8), X),
    (dist.Uniform(0.0, 5.0), st.uniform(0.0, 5.0), X),
    (dist.Gamma(2.5, 1.5), st.gamma(2.5, scale=1 / 1.5), X),
    (dist.Beta(0.5, 0.5), st.beta(0.5, 0.5), torch.tensor([0.1, 0.5, 0.93], dtype=torch.float64)),
    (dist.TruncatedNormal(8.0, 2.0, low=2.0, high=15.0), st.truncnorm(-3.0, 3.5, loc=8.0, scale=2.0),
     torch.tensor([2.5, 8.0, 14.0], dtype=torch.float64)),
])
def test_log_prob_matches_scipy(d, ref, x):
    assert np.allclose(d.log_prob(x).numpy(), ref.logpdf(x.numpy()), rtol=1e-12, atol=1e-12)
    g = torch.Generator().manual_seed(1)
    s = d.sample(g, (20000,))
    assert abs(float(s.mean()) - ref.mean()) < 5 * ref.std() / math.sqrt(20000) + 1e-3


def test_poisson_log_prob_accepts_non_integer_observations():
    rate = torch.tensor([0.5, 3.0, 10.0], dtype=torch.float64)
    k = torch.tensor([0.0, 2.0, 12.0], dtype=torch.float64)
    assert np.allclose(dist.Poisson(rate).log_prob(k).numpy(), st.poisson(rate.numpy()).logpmf(k.numpy()))
    v = torch.tensor([0.25, 2.5, 11.1], dtype=torch.float64)
    ref = v * torch.log(rate) - rate - torch.lgamma(v + 1)
    assert torch.allclose(dist.Poisson(rate).log_prob(v), ref)


def test_transformed_beta_prior_of_the_reference_example():
    # r0 = 1.5 + 1 * Beta(1/2, 1/2)  (reference examples/sir_infer_parameters.py:50-53)
    d = dist.TransformedDistribution(dist.Beta(0.5, 0.5), dist.transforms.AffineTransform(1.5, 1))
    x = torch.tensor([1.6, 2.0, 2.45], dtype=torch.float64)
    assert np.allclose(d.log_prob(x).numpy(), st.beta(0.5, 0.5, loc=1.5, scale=1.0).logpdf(x.numpy()))
    assert (d.support.lower_bound, d.support.upper_bound) == (1.5, 2.5)


@pytest.mark.parametrize("support", [dist.constraints.real, dist.constraints.positive, dist.constraints.unit_interval,
                                     dist.constraints.interval(2.0, 15.0), dist.constraints.greater_than(3.0),
                                     dist.constraints.less_than(-1.0)])
def test_bijectors_round_trip_and_jacobian(support):
    t = dist.biject_to(support)
    z = torch.tensor([-2.0, 0.1, 3.0], dtype=torch.float64, requires_grad=True)
    x = t(z)
    assert torch.allclose(t.inv(x), z, atol=1e-10)
    (jac,) = torch.autograd.grad(x.sum(), z)
    assert torch.allclose(t.log_abs_det_jacobian(z, x), torch.log(jac.abs()), atol=1e-10)


def _names(fn):
    return list(ppl.trace(fn).get_trace().keys())


def test_site_naming_rules():
    # reference tests/test_infer/test_sample.py:49-117
    tree = {"a": dist.Normal(), "b": [1, dist.Normal()], "c": np.array([dist.Normal(), 1], dtype=object),
            "d": {"nested_dict": dist.Normal()}, "e": DeterministicParameter("a")}
    names = _names(lambda: sample_then_resolve(tree, rng_key=PRNGKey(0)))
    assert names == ["a", "b_1", "c_0", "d_nested_dict", "e"]
    names = _names(lambda: sample_then_resolve(tree, rng_key=PRNGKey(0), _prefix="test_"))
    assert names == ["test_a", "test_b_1", "test_c_0", "test_d_nested_dict", "test_e"]
    out = sample_then_resolve(tree, rng_key=PRNGKey(0))
    assert isinstance(out["a"], torch.Tensor) and out["a"] == out["e"]
    assert isinstance(out["b"][1], torch.Tensor) and out["b"][0] == 1


def test_resolve_deterministic_lists_and_slices():
    t = {"a": [0, 1, 3, 4], "b": DeterministicParameter("a", index=1), "c": DeterministicParameter("a", index=slice(0, 2))}
    r = resolve_deterministic(t, root_params=t)
    assert r["b"] == 1 and r["c"] == [0, 1]
    with pytest.raises(Exception, match="unable to find"):
        resolve_deterministic({"x": DeterministicParameter("nope")}, root_params={})


def test_sampling_outside_context_needs_a_key():
    with pytest.raises(ValueError, match="rng_key"):
        sample_distributions({"a": dist.Normal()})
    with ppl.seed(rng_seed=3):
        v = sample_distributions({"a": dist.Normal()})["a"]
    with ppl.seed(rng_seed=3):
        assert sample_distributions({"a": dist.Normal()})["a"] == v
    with pytest.raises(ValueError, match="unique names"):
        ppl.trace(ppl.seed(lambda: (ppl.sample("x", dist.Normal()), ppl.sample("x", dist.Normal())), 0)).get_trace()


def test_adaptation_schedule_is_stans():
    assert build_adaptation_schedule(10) == [(0, 9)]
    assert build_adaptation_schedule(500) == [(0, 74), (75, 99), (100, 149), (150, 249), (250, 449), (450, 499)]
    s = build_adaptation_schedule(100)
    assert s[0] == (0, 14) and s[-1] == (90, 99) and s[1][0] == 15 and s[-2][1] == 89


def test_nuts_recovers_a_correlated_gaussian():
    cov = torch.tensor([[1.0, 0.8], [0.8, 1.5]], dtype=torch.float64)
    mu = torch.tensor([1.0, -2.0], dtype=torch.float64)
    prec = torch.linalg.inv(cov)

    def pg(z):
        d = z - mu
        g = d @ prec
        return 0.5 * (d * g).sum(1), g

    eng = BatchedNUTS(pg, max_tree_depth=6, generator=torch.Generator().manual_seed(5))
    z, extra, last = eng.run(torch.zeros(48, 2, dtype=torch.float64), 120, 80)
    x = z.reshape(-1, 2)
    assert torch.allclose(x.mean(0), mu, atol=0.08)
    assert torch.allclose(torch.cov(x.T), cov, atol=0.15)
    assert float(extra["diverging"].sum()) == 0
    assert 0.7 < float(extra["accept_prob"].mean()) < 0.98
    assert float(split_rhat(z[:, :, 0])) < 1.05 and float(effective_sample_size(z[:, :, 0])) > 500
    assert 0 < eng.grad_evals <= eng.launched_evals  # tree leapfrogs vs rounds x chains


def _toy_model(obs=None):
    mu = ppl.sample("mu", dist.Normal(0.0, 10.0))
    sigma = ppl.sample("sigma", dist.HalfNormal(5.0))
    ppl.deterministic("mu_plus_one", mu + 1.0)
    ppl.sample("y", dist.Normal(mu, sigma), obs=obs)


def test_mcmc_process_on_a_toy_model():
    # reference tests/test_infer/test_inference_processes.py:15-40: runs and returns num_samples draws
    obs = torch.tensor(np.random.default_rng(0).normal(3.0, 2.0, 60))
    proc = MCMCProcess(numpyro_model=_toy_model, num_samples=60, num_warmup=100, num_chains=8,
                       nuts_max_tree_depth=6, progress_bar=False)
    with pytest.raises(AssertionError, match="call infer"):
        proc.get_samples()
    proc.infer(obs=obs)
    s = proc.get_samples()
    assert set(s) == {"mu", "sigma"} and s["mu"].shape == (8 * 60,)
    assert proc.get_samples(group_by_chain=True)["mu"].shape == (8, 60)
    assert "mu_plus_one" in proc.get_samples(exclude_deterministic=False)
    assert abs(float(s["mu"].mean()) - float(obs.mean())) < 0.3
    assert abs(float(s["sigma"].mean()) - float(obs.std())) < 0.4
    az = proc.to_arviz()
    assert az["posterior_predictive"]["y"].shape == (480, 60) and az["prior"]["mu"].shape == (60,)
    summ = proc._inferer.summary()
    assert summ["mu"]["r_hat"] < 1.1


def test_svi_process_on_a_toy_model():
    obs = torch.tensor(np.random.default_rng(1).normal(-1.0, 0.5, 80))
    proc = SVIProcess(numpyro_model=_toy_model, num_iterations=300, num_samples=200, progress_bar=False)
    proc.infer(obs=obs)
    s = proc.get_samples()
    assert s["mu"].shape == (200,) and abs(float(s["mu"].mean()) + 1.0) < 0.2
    assert abs(float(s["sigma"].mean()) - 0.5) < 0.15
    assert "mu_plus_one" in proc.get_samples(exclude_deterministic=False)
    losses = proc._inference_state.losses
    assert float(losses[-20:].mean()) < float(losses[:20].mean())
    az = proc.to_arviz()
    assert az["log_likelihood"]["y"].shape == (200, 80)


def test_predictive_samples_observation_sites():
    post = {"mu": torch.full((50,), 2.0, dtype=torch.float64), "sigma": torch.full((50,), 0.1, dtype=torch.float64)}
    out = Predictive(_toy_model, posterior_samples=post)(PRNGKey(0), obs=None)
    assert out["y"].shape == (50,) and abs(float(out["y"].mean()) - 2.0) < 0.1
    prior = Predictive(_toy_model, num_samples=30, exclude_deterministic=False)(PRNGKey(0), obs=None)
    assert prior["mu"].shape == (30,) and torch.allclose(prior["mu_plus_one"], prior["mu"] + 1.0)


def test_potential_matches_a_hand_written_density():
    obs = torch.tensor([0.5, 1.5], dtype=torch.float64)
    md = ModelDensity(_toy_model, (), {"obs": obs}, device=torch.device("cpu"))
    assert md.dim == 2 and list(md.sites) == ["mu", "sigma"]
    z = torch.tensor([[0.3, -0.2], [1.0, 0.5]], dtype=torch.float64)
    U, g = md.potential_and_grad(z)
    mu, sig = z[:, 0], torch.exp(z[:, 1])
    lp = (st.norm(0, 10).logpdf(mu.numpy()) + st.halfnorm(scale=5).logpdf(sig.numpy()) + z[:, 1].numpy()
          + st.norm(mu.numpy()[:, None], sig.numpy()[:, None]).logpdf(obs.numpy()[None, :]).sum(1))
    assert np.allclose(U.numpy(), -lp, rtol=1e-12)
    eps = 1e-6
    for j in range(2):
        dz = torch.zeros_like(z)
        dz[:, j] = eps
        num = (md.potential(z + dz) - md.potential(z - dz)) / (2 * eps)
        assert torch.allclose(g[:, j], num, rtol=1e-6, atol=1e-8)


def test_transition_schedule_flags_follow_the_adaptation_windows():
    """build_transition_schedule: what each chain does after its t-th transition (numpyro warmup_adapter: dual
    averaging throughout warm-up, Welford in the slow windows, mass-matrix update + step-size search at their
    ends, averaged step size after the last warm-up transition, draws stored afterwards)."""
    from dynode_b200.infer import nuts as N
    for nw, ns in ((150, 50), (100, 10), (10, 5), (0, 7), (1000, 3)):
        fl, wl = N.build_transition_schedule(nw, ns, True, True)
        assert len(fl) == len(wl) == nw + ns
        wins = N.build_adaptation_schedule(nw) if nw > 0 else []
        assert all(f & N.ADAPT for f in fl[:nw]) and not any(f & N.ADAPT for f in fl[nw:])
        assert all(f == N.SAMPLING for f in fl[nw:])
        for w, (a, e) in enumerate(wins):
            middle = 0 < w < len(wins) - 1
            assert all(bool(f & N.WELFORD) == middle for f in fl[a:e + 1])
            assert bool(fl[e] & N.END_SLOW) == middle
            assert not any(f & N.END_SLOW for f in fl[a:e])
            assert all(x == float(e - a + 1) for x in wl[a:e + 1])
        if nw > 0:
            assert fl[nw - 1] & N.END_WARMUP and sum(1 for f in fl if f & N.END_WARMUP) == 1
        fl2, _ = N.build_transition_schedule(nw, ns, False, False)
        assert not any(f & (N.ADAPT | N.WELFORD) for f in fl2)


def test_fused_site_specs_reproduce_log_prob():
    """The (family, p0, p1, c, affine) description handed to dynode_site_logdensity_f64 restates each prior's
    log_prob: evaluated here with numpy from the spec alone (the kernel's formulas, include/dynode_b200_ppl.h)."""
    import numpy as np
    from dynode_b200.infer import distributions as D

    def from_spec(spec, x):
        fam, p0, p1, c, loc, sc = spec
        u = (x - loc) / sc
        xlogy = lambda a, y: np.where(a == 0.0, 0.0, a * np.log(y))
        f = {D.FAM_NORMAL: lambda: -0.5 * ((u - p0) / p1) ** 2, D.FAM_UNIFORM: lambda: np.zeros_like(u),
             D.FAM_BETA: lambda: xlogy(p0 - 1.0, u) + xlogy(p1 - 1.0, 1.0 - u),
             D.FAM_GAMMA: lambda: (p0 - 1.0) * np.log(u) - p1 * u,
             D.FAM_LOGNORMAL: lambda: -0.5 * ((np.log(u) - p0) / p1) ** 2 - np.log(u),
             D.FAM_HALFNORMAL: lambda: -0.5 * (u / p0) ** 2, D.FAM_EXPONENTIAL: lambda: -p0 * u}[fam]()
        return f + c

    rng = np.random.default_rng(1)
    cases = [(D.Normal(0.3, 1.7), rng.normal(size=50)), (D.TruncatedNormal(loc=8, scale=2, low=2, high=15), rng.uniform(2, 15, 50)),
             (D.TruncatedNormal(1.0, 0.5, low=0.0), rng.uniform(0, 4, 50)), (D.Uniform(-1.0, 3.0), rng.uniform(-1, 3, 50)),
             (D.Beta(0.5, 0.5), rng.uniform(0.01, 0.99, 50)), (D.Beta(2.0, 5.0), rng.uniform(0.01, 0.99, 50)),
             (D.Gamma(3.0, 0.5), rng.uniform(0.1, 20, 50)), (D.LogNormal(0.2, 0.8), rng.uniform(0.1, 9, 50)),
             (D.HalfNormal(2.5), rng.uniform(0, 9, 50)), (D.Exponential(0.7), rng.uniform(0, 9, 50)),
             (D.TransformedDistribution(D.Beta(0.5, 0.5), D.transforms.AffineTransform(1.5, 1)), rng.uniform(1.51, 2.49, 50)),
             (D.TransformedDistribution(D.Gamma(2.0, 3.0), D.transforms.AffineTransform(0.25, 2.0)), rng.uniform(0.3, 6, 50))]
    for fn, x in cases:
        spec = D._family_spec(fn)
        assert spec is not None, type(fn).__name__
        ref = fn.log_prob(torch.as_tensor(x, dtype=torch.float64)).numpy()
        assert np.allclose(from_spec(spec, x), ref, rtol=1e-12, atol=1e-12), type(fn).__name__
        assert D._bijector_spec(fn.support) is not None
    # parameters that are another site's value (a tensor with more than one element) cannot be folded into a constant
    assert D._family_spec(D.Normal(torch.zeros(3, dtype=torch.float64), 1.0)) is None


def test_svi_stops_on_a_non_finite_loss_without_poisoning_the_guide():
    """A model evaluation that returns NaN (e.g. an ODE solve that reached max_steps) must not flow into Adam: the
    gradients of that step are zeroed, and the run raises at the next check instead of returning NaN parameters."""
    calls = {"n": 0}

    def model():
        x = ppl.sample("x", dist.Normal(0.0, 1.0))
        calls["n"] += 1
        bad = float("nan") if calls["n"] == 8 else 0.0
        ppl.factor("lik", -0.5 * (x - 1.0) ** 2 + bad)

    from dynode_b200.infer.inference import SVI, Adam, AutoMultivariateNormal, init_to_median
    guide = AutoMultivariateNormal(model, init_loc_fn=init_to_median)
    svi = SVI(model=model, guide=guide, optim=Adam(step_size=0.1))
    with pytest.raises(RuntimeError, match="not finite"):
        svi.run(PRNGKey(3), 50, progress_bar=False)
    assert calls["n"] < 20  # stopped at the first check after the bad step, not after all 50
    assert all(bool(torch.isfinite(p).all()) for p in guide.parameters())


def test_nuts_kernel_limits_match_the_header():
    import os
    import re

    from dynode_b200 import _lib
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include",
                            "dynode_b200_nuts.h")).read()
    assert int(re.search(r"#define DYNODE_NUTS_MAX_DIM (\d+)", hdr).group(1)) == _lib.NUTS_MAX_DIM
    assert int(re.search(r"#define DYNODE_NUTS_MAX_DEPTH (\d+)", hdr).group(1)) == _lib.NUTS_MAX_DEPTH


def test_row_count_hint_reaches_the_forward_versus_adjoint_choice():
    """engine.only_rows(mask, n_rows=) is host knowledge about a device mask; simulation.autograd.use_adjoint sizes its
    choice by the rows that run, and ModelDensity.launch_key lets a sampler ask whether a captured round is stale."""
    import torch

    from dynode_b200 import _lib, engine
    from dynode_b200.simulation import autograd as ag
    model = engine.FlowModel(_lib.FLOW_SEIRS_C, 0, 6, 3)
    opts = engine.SolverOptions(t1=120.0)
    assert ag.use_adjoint(model, 6, opts, 1024) is True          # 6144 warps: the adjoint's ~3.2 solves win
    assert ag.use_adjoint(model, 6, opts, 128) is False          # 768 warps: one direction per warp, latency regime
    assert ag.use_adjoint(model, 6, opts, 1024, n_rows=40) is False
    mask = torch.ones(1024, dtype=torch.uint8)
    with engine.only_rows(mask, n_rows=40):
        assert engine.rows_to_integrate(1024) == 40 and engine.rows_to_integrate(512) == 512
        assert ag.use_adjoint(model, 6, opts, 1024) is False
        with engine.only_rows(mask):                             # a nested mask without a count: all rows again
            assert engine.rows_to_integrate(1024) == 1024
        assert engine.rows_to_integrate(1024) == 40
    assert engine.rows_to_integrate(1024) == 1024
    assert ag.use_adjoint(model, 6, engine.SolverOptions(t1=120.0, jump_ts=(30.0,)), 1024) is True  # jumps: no longer forward-only
