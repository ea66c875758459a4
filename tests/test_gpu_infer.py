"""GPU: the differentiable / vmappable solve, the fused NUTS log-density and the inference processes on
the reference's SIR inference example (examples/sir_infer_parameters.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda", 0)


def test_simulate_is_differentiable_and_matches_oracle_tangents():
    from dynode_b200.config import SolverParams
    from dynode_b200.examples import rhs as ex
    from dynode_b200.simulation import simulate
    from oracle import oracle as orc
    from tests.cases import make_case
    case = make_case("sir_age2", 1)
    beta = torch.tensor(case["params"]["beta"][0], dtype=torch.float64, device=_dev(), requires_grad=True)
    gamma = torch.tensor(case["params"]["gamma"][0], dtype=torch.float64, device=_dev(), requires_grad=True)
    y0 = torch.tensor(case["y0"], dtype=torch.float64, device=_dev())
    state = (y0[0:2], y0[2:4], y0[4:6])
    p = ex.AgeSIR_ODEParams(beta=beta, gamma=gamma, contact_matrix=torch.tensor(case["contact"], device=_dev()))
    sol = simulate(ex.sir_age_ode, 100, state, p, SolverParams())
    assert sol.ys[2].shape == (101, 2) and sol.ys[2].requires_grad
    w = torch.linspace(0.5, 1.5, 101 * 2, dtype=torch.float64, device=_dev()).reshape(101, 2)
    loss = (w * sol.ys[2]).sum() + sol.ys[1][50, 0] * 3.0
    gb, gg = torch.autograd.grad(loss, (beta, gamma))
    fam, dims, theta, shared = case["oracle"]
    ys, dys, _ = orc.solve(fam, dims, case["y0"], theta, shared, t1=100, wrt=[0, 1])
    ref = (w.cpu().numpy()[:, :, None] * dys[0][:, 4:6, :]).sum((0, 1)) + 3.0 * dys[0][50, 2, :]
    assert np.allclose([float(gb), float(gg)], ref, rtol=1e-8)
    assert np.allclose(sol.ys[2].detach().cpu().numpy(), ys[0][:, 4:6], rtol=1e-9, atol=1e-9)


def test_simulate_with_discontinuity_points_is_differentiable():
    """SolverParams.discontinuity_points (reference odes.py:120-131) under autograd: the gradient through
    `simulate` equals the oracle's tangents of the same clipped step sequence."""
    from dynode_b200.config import SolverParams
    from dynode_b200.examples import rhs as ex
    from dynode_b200.simulation import simulate
    from oracle import oracle as orc
    from tests.cases import make_case
    case = make_case("sir_age2", 1)
    beta = torch.tensor(case["params"]["beta"][0], dtype=torch.float64, device=_dev(), requires_grad=True)
    gamma = torch.tensor(case["params"]["gamma"][0], dtype=torch.float64, device=_dev(), requires_grad=True)
    y0 = torch.tensor(case["y0"], dtype=torch.float64, device=_dev())
    p = ex.AgeSIR_ODEParams(beta=beta, gamma=gamma, contact_matrix=torch.tensor(case["contact"], device=_dev()))
    jumps = [30.0, 61.5]
    sol = simulate(ex.sir_age_ode, 100, (y0[0:2], y0[2:4], y0[4:6]), p, SolverParams(discontinuity_points=jumps))
    w = torch.linspace(0.5, 1.5, 101 * 2, dtype=torch.float64, device=_dev()).reshape(101, 2)
    gb, gg = torch.autograd.grad((w * sol.ys[2]).sum(), (beta, gamma))
    fam, dims, theta, shared = case["oracle"]
    ys, dys, st = orc.solve(fam, dims, case["y0"], theta, shared, t1=100, wrt=[0, 1], jump_ts=jumps)
    ref = (w.cpu().numpy()[:, :, None] * dys[0][:, 4:6, :]).sum((0, 1))
    assert np.allclose([float(gb), float(gg)], ref, rtol=1e-8)
    assert int(sol.stats["num_accepted_steps"]) == int(st[0, 1])


def test_gradient_with_respect_to_initial_state():
    from dynode_b200.config import SolverParams
    from dynode_b200.examples import rhs as ex
    from dynode_b200.simulation import simulate
    y0 = torch.tensor([0.9, 0.1, 0.0], dtype=torch.float64, device=_dev(), requires_grad=True)
    p = ex.SIR_ODEParams(beta=torch.tensor(0.3, dtype=torch.float64, device=_dev()), gamma=torch.tensor(0.1, device=_dev(), dtype=torch.float64))

    def final_r(y):
        return simulate(ex.sir_ode, 60, (y[0:1], y[1:2], y[2:3]), p, SolverParams()).ys[2][-1, 0]

    (g,) = torch.autograd.grad(final_r(y0), y0)
    eps = 1e-6
    for j in range(3):
        d = torch.zeros(3, dtype=torch.float64, device=_dev())
        d[j] = eps
        with torch.no_grad():
            num = (final_r((y0 + d).detach().requires_grad_(True)) - final_r((y0 - d).detach().requires_grad_(True))) / (2 * eps)
        assert abs(float(g[j]) - float(num)) < 2e-4 * max(1.0, abs(float(num)))  # adaptive steps move under FD


def _example():
    from dynode_b200.examples import sir_infer_parameters as m
    cfg = m.get_config()
    obs = m.synthetic_incidence(100).to(_dev())
    return m, cfg, obs


def test_vmapped_potential_equals_per_draw_and_fused_equals_general():
    from dynode_b200.infer import ModelDensity
    m, cfg, obs = _example()
    md = ModelDensity(m.model, (), dict(config=cfg, tf=100, obs_data=obs))
    mdf = ModelDensity(m.model_fused, (), dict(config=cfg, tf=100, obs_data=obs))
    assert list(md.sites) == ["strains_0_r0", "strains_0_infectious_period"] and md.dim == 2
    g = torch.Generator(device=_dev()).manual_seed(0)
    Z = torch.randn(33, 2, dtype=torch.float64, device=_dev(), generator=g)
    U, G = md.potential_and_grad(Z)
    Uf, Gf = mdf.potential_and_grad(Z)
    assert torch.allclose(U, Uf, rtol=1e-10) and torch.allclose(G, Gf, rtol=1e-7, atol=1e-7)
    for k in (0, 7, 32):
        u1, g1 = md.potential_and_grad(Z[k:k + 1])
        assert torch.allclose(u1[0], U[k], rtol=1e-12) and torch.allclose(g1[0], G[k], rtol=1e-9, atol=1e-9)
    # the posterior mode sits near r0 = 2, infectious period = 7
    truth = mdf.unconstrain({"strains_0_r0": torch.tensor([2.0]), "strains_0_infectious_period": torch.tensor([7.0])})
    Ut, Gt = mdf.potential_and_grad(truth.to(torch.float64))
    assert float(Ut) < float(U.min()) and float(Gt.abs().max()) < 50.0


def test_potential_gradient_against_complex_free_finite_differences():
    from dynode_b200.infer import ModelDensity
    m, cfg, obs = _example()
    md = ModelDensity(m.model_fused, (), dict(config=cfg, tf=100, obs_data=obs))
    Z = torch.tensor([[0.1, -0.3], [0.8, 0.4]], dtype=torch.float64, device=_dev())
    U, G = md.potential_and_grad(Z)
    eps = 1e-5
    for j in range(2):
        d = torch.zeros_like(Z)
        d[:, j] = eps
        num = (md.potential(Z + d) - md.potential(Z - d)) / (2 * eps)
        # finite differences see the adaptive step sequence move; the kernel's gradient is the frozen-step one
        assert torch.allclose(G[:, j], num, rtol=5e-3, atol=1e-3)


def test_mcmc_process_recovers_the_generating_parameters():
    from dynode_b200.infer import MCMCProcess
    m, cfg, obs = _example()
    proc = MCMCProcess(numpyro_model=m.model_fused, num_warmup=150, num_samples=100, num_chains=64,
                       nuts_max_tree_depth=6, progress_bar=False)
    proc.infer(config=cfg, tf=100, obs_data=obs)
    s = proc.get_samples()
    assert s["strains_0_r0"].shape == (6400,)
    # the posterior means of this model are 2.0454 and 7.197 (262144 chains, profiles/r1/nuts_async.md): the
    # generating values 2 and 7 shifted by the priors; 64 chains x 100 draws scatter around them by ~0.01 / 0.05
    assert abs(float(s["strains_0_r0"].mean()) - 2.045) < 0.03
    assert abs(float(s["strains_0_infectious_period"].mean()) - 7.2) < 0.25
    summ = proc._inferer.summary()
    # 64 chains x 100 draws after 150 warm-up transitions: split r_hat scatters over 1.03..1.06 from seed to seed,
    # the same with the compiled and the composed model evaluation (scripts/plan_compile_time.py on B200)
    assert summ["strains_0_r0"]["r_hat"] < 1.1
    # general (trajectory-materialising) model through the same process, fewer chains
    proc2 = MCMCProcess(numpyro_model=m.model, num_warmup=100, num_samples=50, num_chains=16,
                        nuts_max_tree_depth=6, progress_bar=False)
    proc2.infer(config=cfg, tf=100, obs_data=obs)
    s2 = proc2.get_samples()
    assert abs(float(s2["strains_0_r0"].mean()) - float(s["strains_0_r0"].mean())) < 0.05
    pp = proc2.to_arviz()
    assert pp["posterior_predictive"]["inf_incidence"].shape == (16 * 50, 100, 2)


def test_svi_process_and_predictive_projection():
    from dynode_b200.infer import Predictive, PRNGKey, SVIProcess
    m, cfg, obs = _example()
    proc = SVIProcess(numpyro_model=m.model_fused, num_iterations=300, num_samples=200, progress_bar=False)
    proc.infer(config=cfg, tf=100, obs_data=obs)
    s = proc.get_samples()
    assert abs(float(s["strains_0_r0"].mean()) - 2.0) < 0.1
    # project forward 200 days without data (reference sir_infer_parameters.py:156-169)
    out = Predictive(m.model, posterior_samples=s)(PRNGKey(1), config=cfg, tf=200, obs_data=None)
    assert out["inf_incidence"].shape == (200, 200, 2) and bool((out["inf_incidence"] >= 0).all())


def test_sharded_ensemble_single_rank():
    from dynode_b200.config import SolverParams
    from dynode_b200.distributed import simulate_ensemble_sharded
    from dynode_b200.examples import rhs as ex
    from dynode_b200.simulation import simulate_ensemble
    from tests.cases import make_case
    B = 300
    case = make_case("seirs_multi_a2s3", B)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64, device=_dev())
    prm = case["params"]
    p = ex.SEIRS_MultiStrain_ODEParams(beta=t(prm["beta"]), gamma=t(prm["gamma"]), sigma=t(prm["sigma"]),
                                       omega=t(prm["omega"]), contact_matrix=t(case["contact"]))
    y0 = t(case["y0"])
    state = (y0[:, :2], y0[:, 2:8].reshape(B, 2, 3), y0[:, 8:14].reshape(B, 2, 3),
             y0[:, 14:20].reshape(B, 2, 3), y0[:, 20:26].reshape(B, 2, 3))
    ys, res, (lo, hi) = simulate_ensemble_sharded(ex.seirs_multi_strain_ode, 120, state, p, SolverParams(), batch_size=B)
    ref = simulate_ensemble(ex.seirs_multi_strain_ode, 120, state, p, SolverParams(), batch_size=B, state_batched=True)
    flat = torch.cat([c.reshape(B, 121, -1) for c in ref.ys], dim=2)
    assert (lo, hi) == (0, B) and torch.equal(ys, flat) and int((res != 0).sum()) == 0


def test_config5_age_risk_strain_nuts():
    """BASELINE config 5 (n = 78, six inferred parameters): fused and trajectory-materialising log-densities
    agree, and a short many-chain NUTS run concentrates on the generating values."""
    from dynode_b200.examples import seirs_age_risk_strain as m5
    from dynode_b200.infer import MCMC, NUTS, ModelDensity, PRNGKey
    tf = 120
    obs = m5.synthetic_incidence(tf).to(_dev())
    assert obs.shape == (tf, 6, 3)
    cfg = m5.get_config(infer=True)
    md = ModelDensity(m5.model, (), dict(config=cfg, tf=tf, obs_data=obs))
    mdf = ModelDensity(m5.model_fused, (), dict(config=cfg, tf=tf, obs_data=obs))
    assert md.dim == 6
    Z = torch.randn(9, 6, dtype=torch.float64, device=_dev(), generator=torch.Generator(device=_dev()).manual_seed(1)) * 0.5
    U, G = md.potential_and_grad(Z)
    Uf, Gf = mdf.potential_and_grad(Z)
    assert torch.allclose(U, Uf, rtol=1e-9) and torch.allclose(G, Gf, rtol=1e-6, atol=1e-6 * float(G.abs().max()))
    mc = MCMC(NUTS(m5.model_fused, max_tree_depth=6), num_warmup=120, num_samples=60, num_chains=32, progress_bar=False)
    mc.run(PRNGKey(3), config=cfg, tf=tf, obs_data=obs)
    s = mc.get_samples()
    for k in range(3):
        assert abs(float(s[f"strains_{k}_r0"].mean()) - m5.TRUE_R0[k]) < 0.25
        assert abs(float(s[f"strains_{k}_infectious_period"].mean()) - m5.TRUE_INF[k]) < 1.0


def test_cuda_nuts_round_equals_the_torch_round():
    """dynode_nuts_round_pre/post (one thread per chain) against the masked-tensor round of infer/nuts.py, in
    lock-step from the same state with the same random numbers: every state field agrees round by round
    (compared over 120 rounds -- beyond a few hundred, rounding differences between cuBLAS bmm and the
    per-thread dot products are amplified by warm-up trajectories that run near the stability limit)."""
    from dynode_b200.infer import nuts as N
    from dynode_b200.infer.nuts import BatchedNUTS
    dev = _dev()
    cov = torch.tensor([[1.0, 0.6, 0.0], [0.6, 2.0, -0.4], [0.0, -0.4, 0.5]], dtype=torch.float64, device=dev)
    mu = torch.tensor([0.5, -1.0, 2.0], dtype=torch.float64, device=dev)
    prec = torch.linalg.inv(cov)

    def pg(z):
        d = z - mu
        g = d @ prec
        return 0.5 * (d * g).sum(1), g

    C = 96
    engs = []
    for kernels in (False, True):
        e = BatchedNUTS(pg, max_tree_depth=5, generator=torch.Generator(device=dev).manual_seed(11),
                        cuda_graph=False, cuda_kernels=kernels)
        e._allocate(torch.zeros(C, 3, dtype=torch.float64, device=dev), 30)
        e._g = e.gen
        U, g = e._eval(e.b.z)
        e.b.U.copy_(U)
        e.b.g.copy_(g)
        e.b.need_tree.fill_(True)
        # 12 transitions per chain, every kind of bookkeeping on each of them, a slow-window end (mass-matrix
        # update by the chain's own thread + dual-averaging restart) after the 6th and the warm-up end after the 12th
        fl = [N.ADAPT | N.WELFORD | N.SAMPLING] * 12
        fl[5] |= N.END_SLOW
        fl[11] |= N.END_WARMUP
        e.set_schedule(fl, [6.0] * 12, 0)
        e.b.searching.fill_(True)  # step-size search before the first transition (and again after the 6th)
        e._prepare_round_fn()
        engs.append(e)
    fields = ["z", "U", "g", "eps", "k", "active", "need_tree", "searching", "fr_dir", "fr_last", "energy0", "zL", "rL", "gL", "zR", "rR", "gR", "zP",
              "gP", "r_sum", "UP", "weight", "sum_acc", "depth", "nprop", "turning", "diverging", "s_n", "s_right",
              "s_turn", "s_div", "s_z", "s_r", "s_g", "s_zP", "s_gP", "s_rsum", "s_UP", "s_w", "s_acc", "da_x",
              "da_xavg", "da_gavg", "da_t", "da_prox", "wf_n", "wf_mean", "wf_m2", "imm", "msqrt", "out_z"]
    for rnd in range(120):
        for e in engs:
            e._round_fn()
        a, b = engs[0].b, engs[1].b
        for f in fields:
            x, y = getattr(a, f).double(), getattr(b, f).double()
            assert torch.allclose(x, y, rtol=1e-8, atol=1e-9, equal_nan=True), (rnd, f)
    assert int(engs[0].b.k.min()) >= 3 and not bool(engs[0].b.active.all())  # trees were built, chains finished
    assert int(engs[0].b.n_useful) == int(engs[1].b.n_leap.sum())
    for name in ("accept_prob", "num_steps", "diverging", "potential_energy", "tree_depth"):
        assert torch.allclose(engs[0].b.out_stats[name], engs[1].b.out_stats[name], rtol=1e-8, atol=1e-9)
    # and with graph replay the sampler recovers the target
    eng = BatchedNUTS(pg, max_tree_depth=7, cuda_graph=True, cuda_kernels=True)
    z, extra, st = eng.run(torch.zeros(512, 3, dtype=torch.float64, device=dev), 150, 100)
    assert eng.graph_used and eng.kernels_used
    x = z.reshape(-1, 3)
    assert torch.allclose(x.mean(0), mu, atol=0.05) and torch.allclose(torch.cov(x.T), cov, atol=0.08)


def test_fused_bijector_equals_the_composed_transforms():
    """dynode_bijector_f64 / _vjp_f64 (one launch each way) against biject_to(support) written with tensor
    operations: value, log-Jacobian and the gradient of an arbitrary function of both, plain and under vmap."""
    from dynode_b200.infer import distributions as D
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(3)
    z0 = torch.randn(4097, dtype=torch.float64, device=dev, generator=g) * 6.0
    z0[:4] = torch.tensor([-745.0, -40.0, 40.0, 700.0], dtype=torch.float64, device=dev)  # tails stay finite
    w1, w2 = (torch.randn(4097, dtype=torch.float64, device=dev, generator=g) for _ in range(2))
    for sup in (D.constraints.unit_interval, D.constraints.interval(2.0, 15.0), D.constraints.interval(1.5, 2.5),
                D.constraints.positive, D.constraints.greater_than(-3.0), D.constraints.less_than(4.0)):
        outs = []
        for fused in (True, False):
            z = z0.clone().requires_grad_(True)
            if fused:
                x, l = D.constrain_with_ladj(sup, z)
            else:
                t = D.biject_to(sup)
                x = t(z)
                l = t.log_abs_det_jacobian(z, x)
            (gz,) = torch.autograd.grad((w1 * x).sum() + (w2 * l).sum(), z)
            outs.append((x.detach(), l.detach(), gz))
        for a, b in zip(*outs):
            fin = torch.isfinite(b)
            assert torch.equal(torch.isfinite(a), fin)
            assert torch.allclose(a[fin], b[fin], rtol=1e-13, atol=1e-13), sup
        # under vmap (how ModelDensity evaluates the model for all chains at once)
        f = lambda zz: D.constrain_with_ladj(sup, zz)
        xv, lv = torch.vmap(f)(z0)
        assert torch.equal(xv, outs[0][0]) and torch.equal(lv, outs[0][1])


def test_fused_site_kernel_equals_bijector_plus_log_prob():
    """dynode_site_logdensity_f64 / _vjp_f64 against biject_to(support) + log|J| + fn.log_prob written with tensor
    operations, for every prior family the kernel states (incl. an affine-transformed and a truncated one)."""
    from dynode_b200.infer import distributions as D
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(4)
    z0 = torch.randn(2049, dtype=torch.float64, device=dev, generator=g) * 3.0
    w1, w2 = (torch.randn(2049, dtype=torch.float64, device=dev, generator=g) for _ in range(2))
    priors = [D.Normal(0.3, 1.7), D.TruncatedNormal(loc=8, scale=2, low=2, high=15), D.TruncatedNormal(1.0, 0.5, low=0.0),
              D.Uniform(-1.0, 3.0), D.Beta(0.5, 0.5), D.Beta(2.0, 5.0), D.Gamma(3.0, 0.5), D.LogNormal(0.2, 0.8),
              D.HalfNormal(2.5), D.Exponential(0.7),
              D.TransformedDistribution(D.Beta(0.5, 0.5), D.transforms.AffineTransform(1.5, 1)),
              D.TransformedDistribution(D.Gamma(2.0, 3.0), D.transforms.AffineTransform(0.25, 2.0))]
    for fn in priors:
        z = z0.clone().requires_grad_(True)
        got = D.fused_site(fn, z)
        assert got is not None, type(fn).__name__
        x, lp = got
        (gz,) = torch.autograd.grad((w1 * x).sum() + (w2 * lp).sum(), z)
        z2 = z0.clone().requires_grad_(True)
        t = D.biject_to(fn.support)
        x2 = t(z2)
        lp2 = t.log_abs_det_jacobian(z2, x2) + fn.log_prob(x2)
        (gz2,) = torch.autograd.grad((w1 * x2).sum() + (w2 * lp2).sum(), z2)
        for a, b in ((x, x2), (lp, lp2), (gz, gz2)):
            fin = torch.isfinite(b) & torch.isfinite(a)
            assert fin.double().mean() > 0.99
            assert torch.allclose(a[fin], b[fin].detach(), rtol=1e-11, atol=1e-11), type(fn).__name__
    # a prior whose parameter is another site's value keeps the composed path
    assert D.fused_site(D.Normal(z0[:3], 1.0), z0[:3]) is None


@pytest.mark.parametrize("which", ["c2", "c5"])
def test_compiled_potential_equals_the_composed_model_evaluation(which, monkeypatch):
    """dynode_potential_pre_f64 -> fused log-likelihood -> dynode_potential_post_f64 (three launches) against the
    vmapped Python model + autograd (~35 launches): same potential, same gradient, with and without the sampler's row
    mask, in forward mode and through the adjoint (incl. rows that overflow its checkpoint capacity)."""
    from dynode_b200 import engine
    from dynode_b200.infer import ModelDensity
    dev = torch.device("cuda", 0)
    if which == "c2":
        from dynode_b200.examples import sir_infer_parameters as m
        kw = dict(config=m.get_config(), tf=100, obs_data=m.synthetic_incidence(100).to(dev))
    else:
        from dynode_b200.examples import seirs_age_risk_strain as m
        kw = dict(config=m.get_config(infer=True), tf=120, obs_data=m.synthetic_incidence(120).to(dev))
    md = ModelDensity(m.model_fused, (), kw, device=dev)
    g = torch.Generator(device=dev).manual_seed(11)
    Z = md.init_to_median(1) + 0.6 * torch.randn(301, md.dim, dtype=torch.float64, device=dev, generator=g)
    U, dU = md.potential_and_grad(Z)
    assert md._plan is not None, md.plan_reason
    for force in ("0", "1"):
        monkeypatch.setenv("DYNODE_B200_ADJOINT", force)
        if force == "1":
            monkeypatch.setenv("DYNODE_B200_ADJOINT_CAP", "24")  # some rows overflow and take the forward fallback
        U_c, dU_c = md.potential_and_grad_composed(Z)
        U_p, dU_p = md._plan.potential_and_grad(Z)
        assert torch.allclose(U_p, U_c, rtol=1e-10), (force, float((U_p - U_c).abs().max()))
        assert torch.allclose(dU_p, dU_c, rtol=1e-7, atol=1e-8 * float(dU_c.abs().max()))
        mask = (torch.arange(301, device=dev) % 3 != 0)
        with engine.only_rows(mask.view(torch.uint8)):
            U_m, dU_m = md._plan.potential_and_grad(Z)
        assert torch.equal(U_m[mask], U_p[mask]) and torch.equal(dU_m[mask], dU_p[mask])
        assert bool((U_m[~mask] == 0).all()) and bool((dU_m[~mask] == 0).all())


def test_models_outside_the_compiled_form_keep_the_composed_path():
    """The trajectory-materialising model (an observed `sample` site on diff(R), not the fused factor) is not of the
    compiled form: the plan says why and the vmapped evaluation runs."""
    from dynode_b200.examples import sir_infer_parameters as m
    from dynode_b200.infer import ModelDensity
    dev = torch.device("cuda", 0)
    md = ModelDensity(m.model, (), dict(config=m.get_config(), tf=100, obs_data=m.synthetic_incidence(100).to(dev)),
                      device=dev)
    Z = md.init_to_median(5)
    U, dU = md.potential_and_grad(Z)
    assert md._plan is None and "fused log-likelihood" in md.plan_reason
    assert bool(torch.isfinite(U).all()) and bool(torch.isfinite(dU).all())


def test_cuda_nuts_rounds_reproduce_the_numpy_oracle_on_a_gaussian():
    """dynode_nuts_round_pre / _post (one thread per chain, state machines) against oracle/nuts_np.py (numpyro's
    sequential tree building restated in numpy) on a shared random tape: same tree depth, leapfrog count and
    divergence flag in every transition, same acceptance statistic, step size and draws."""
    from tests.nuts_tape import compare_with_oracle
    from tests.test_nuts_oracle import Z0, pg_batched, pg_single
    dev = "cuda"
    compare_with_oracle(pg_batched, pg_single, Z0, 30, 20, seed=77, max_tree_depth=6, device=dev, cuda_kernels=True,
                        atol=1e-8)
    compare_with_oracle(pg_batched, pg_single, Z0, 120, 40, seed=5, max_tree_depth=6, device=dev, cuda_kernels=True,
                        atol=1e-11, adapt_step_size=False, step_size=0.5)
    compare_with_oracle(pg_batched, pg_single, Z0[:2], 150, 30, seed=77, max_tree_depth=6, device=dev,
                        cuda_kernels=True, atol=5e-2)


def test_cuda_nuts_rounds_reproduce_the_numpy_oracle_on_config_2():
    """The same on the posterior of reference examples/sir_infer_parameters.py (fused ODE log-likelihood): the CUDA
    sampler evaluates all chains in one compiled-potential launch per round, the oracle one chain and one leapfrog at a
    time through the same potential."""
    from dynode_b200.examples import sir_infer_parameters as m
    from dynode_b200.infer import ModelDensity
    from tests.nuts_tape import compare_with_oracle
    dev = torch.device("cuda", 0)
    md = ModelDensity(m.model_fused, (), dict(config=m.get_config(), tf=100, obs_data=m.synthetic_incidence(100).to(dev)),
                      device=dev)
    z0 = md.init_to_median(3).cpu().numpy() + np.array([[0.0, 0.0], [0.3, -0.2], [-0.4, 0.1]])

    def pg_single(z):
        U, g = md.potential_and_grad(torch.as_tensor(z[None, :], device=dev))
        return float(U[0]), g[0].cpu().numpy()

    # the potential is an adaptive ODE solve: a 1e-16 change of z can move a solver step and U by 1e-9, which the
    # step-size feedback amplifies -- so the draws are compared to 1e-3 (the discrete decisions: exactly)
    compare_with_oracle(md.potential_and_grad, pg_single, z0, 30, 15, seed=11, max_tree_depth=6, device="cuda",
                        cuda_kernels=True, atol=1e-3)


def test_posterior_of_config_2_matches_an_independent_sampler_on_an_independent_potential():
    """BASELINE: "posterior summaries statistically indistinguishable".  Left side: the product -- MCMCProcess on the
    fused CUDA log-likelihood, 256 chains.  Right side: nothing of the product's numerics -- the numpy NUTS oracle
    (oracle/nuts_np.py) on the host statement of numpyro's config-2 potential whose ODE part is the C++ oracle's
    frozen-step tangents (tests/test_oracle.py::config2_potential_on_host).  Means agree within Monte-Carlo error."""
    from dynode_b200.examples import sir_infer_parameters as m
    from dynode_b200.infer import MCMC, NUTS, PRNGKey
    from oracle.nuts_np import NutsChain, Tape
    from tests.test_oracle import config2_potential_on_host
    dev = torch.device("cuda", 0)
    obs = m.synthetic_incidence(100).to(dev)
    mc = MCMC(NUTS(m.model_fused, max_tree_depth=8), num_warmup=200, num_samples=100, num_chains=256, progress_bar=False)
    mc.run(PRNGKey(3), config=m.get_config(), tf=100, obs_data=obs)
    s = mc.get_samples()
    names = list(mc.density.sites)  # column order of z
    obs_h = obs.cpu().numpy()

    def pg(z):
        U, dU, *_ = config2_potential_on_host(z[None, :], obs_h, names)
        return float(U[0]), dU[0]

    draws = []
    for c in range(3):
        chain = NutsChain(pg, 2, Tape(41, c, 3, 2), max_tree_depth=8)
        rec = chain.run(np.zeros(2), 150, 250)[150:]
        draws.append(np.array([r["z"] for r in rec]))
    z = torch.as_tensor(np.concatenate(draws), device=dev)
    ref = mc.density.constrain(z)
    for name, tol in (("strains_0_r0", 0.02), ("strains_0_infectious_period", 0.12)):
        a, b = s[name].cpu().numpy(), ref[name].cpu().numpy()
        # 750 correlated oracle draws: allow 4 standard errors at an effective sample size of ~150
        se = b.std() / np.sqrt(150.0)
        assert abs(a.mean() - b.mean()) < max(4 * se, tol), (name, a.mean(), b.mean(), se)
        assert 0.6 < a.std() / b.std() < 1.6, (name, a.std(), b.std())


def test_compiled_potential_over_prior_families():
    """Every pair of prior families on (r0, infectious period) of the config-2 model -- bounded, half-bounded, affine
    transformed, truncated on one side or both -- compiles into the three-launch evaluation and equals the composed
    one (bijector + log|J| + log_prob + get_odeparams + fused likelihood + autograd)."""
    from dynode_b200.config import Strain
    from dynode_b200.examples import sir_infer_parameters as m
    from dynode_b200.infer import ModelDensity
    from dynode_b200.infer import distributions as D
    dev = torch.device("cuda", 0)
    aff = lambda base, loc, sc: D.TransformedDistribution(base, D.transforms.AffineTransform(loc, sc))
    r0_priors = [aff(D.Beta(0.5, 0.5), 1.5, 1), D.Uniform(1.2, 3.0), D.LogNormal(0.7, 0.2), D.Gamma(20.0, 10.0),
                 D.TruncatedNormal(2.0, 0.5, low=1.1, high=4.0), aff(D.HalfNormal(0.8), 1.1, 1.0),
                 aff(D.Exponential(1.5), 1.2, 1.0), D.TruncatedNormal(2.0, 0.4, low=1.05)]
    ip_priors = [D.TruncatedNormal(loc=8, scale=2, low=2, high=15), D.Uniform(3.0, 12.0), D.LogNormal(2.0, 0.2),
                 D.Gamma(16.0, 2.0), aff(D.Exponential(0.25), 2.5, 1.0), aff(D.Beta(2.0, 3.0), 3.0, 9.0),
                 D.Normal(7.5, 0.4)]
    obs = m.synthetic_incidence(100).to(dev)
    g = torch.Generator(device=dev).manual_seed(23)
    n = 0
    for i, pr in enumerate(r0_priors):
        for j, pi in enumerate(ip_priors):
            if (i + j) % 2 and i and j:  # every family meets every other side's first entry, half of the rest
                continue
            cfg = m.get_config()
            cfg.parameters.transmission_params.strains = [Strain(strain_name="swo9", r0=pr, infectious_period=pi)]
            md = ModelDensity(m.model_fused, (), dict(config=cfg, tf=100, obs_data=obs), device=dev)
            Z = md.init_to_median(1) + 0.3 * torch.randn(65, md.dim, dtype=torch.float64, device=dev, generator=g)
            U, dU = md.potential_and_grad(Z)
            assert md._plan is not None, (type(pr).__name__, type(pi).__name__, md.plan_reason)
            U_c, dU_c = md.potential_and_grad_composed(Z)
            ok = torch.isfinite(U_c)
            assert ok.double().mean() > 0.9
            assert torch.equal(torch.isfinite(U), ok)
            assert torch.allclose(U[ok], U_c[ok], rtol=1e-10), (i, j, float((U[ok] - U_c[ok]).abs().max()))
            assert torch.allclose(dU[ok], dU_c[ok], rtol=1e-7, atol=1e-8 * float(dU_c[ok].abs().max())), (i, j)
            n += 1
    assert n >= 30


def test_sampler_recaptures_its_round_when_the_models_launch_choice_changes(monkeypatch):
    """Most rounds of a run belong to a few straggler chains.  The sampler tells the model's launches how many chains
    still run (engine.only_rows(n_rows=...)); when that moves the forward-vs-adjoint choice (simulation.autograd.
    use_adjoint), the captured round is captured again.  Forced here with a tiny 'resident warps' figure on config 5;
    the posterior must come out the same as with the launches fixed."""
    from dynode_b200 import engine
    from dynode_b200.examples import seirs_age_risk_strain as m5
    from dynode_b200.infer import MCMC, NUTS, PRNGKey
    from dynode_b200.simulation import autograd as ag
    dev = torch.device("cuda", 0)
    assert engine.rows_to_integrate(64) == 64
    mask = torch.ones(64, dtype=torch.uint8, device=dev)
    with engine.only_rows(mask, n_rows=5):
        assert engine.rows_to_integrate(64) == 5 and engine.rows_to_integrate(63) == 63
    obs = m5.synthetic_incidence(60).to(dev)
    cfg = m5.get_config(infer=True)
    monkeypatch.setattr(ag, "RESIDENT_WARPS", 6 * 20)  # forward mode only once fewer than ~20 of the 96 chains run

    def run():
        mc = MCMC(NUTS(m5.model_fused, max_tree_depth=5), num_warmup=40, num_samples=30, num_chains=96,
                  progress_bar=False)
        mc.run(PRNGKey(3), config=cfg, tf=60, obs_data=obs)
        return mc

    mc = run()
    assert mc.engine.graph_used and mc.engine.recaptures == 1, mc.engine.recaptures
    assert set(mc.timing) >= {"setup_s", "first_eval_s", "capture_s", "rounds_s", "constrain_s"}
    monkeypatch.setenv("DYNODE_B200_FIXED_LAUNCH", "1")
    mc_fixed = run()
    assert mc_fixed.engine.recaptures == 0
    a, b = mc.get_samples(), mc_fixed.get_samples()
    for k in a:  # same chains, same random numbers; the two gradient modes agree to ~1e-9, the draws stay close
        sd = float(b[k].std())
        assert abs(float(a[k].mean()) - float(b[k].mean())) < 0.25 * sd, k


def test_models_beyond_the_round_kernels_limits_run_the_tensor_round_on_the_device():
    """More than 16 latent dimensions (or more than 12 doublings) do not fit the one-thread-per-chain kernels: the
    sampler then runs the same round as masked tensor operations on the GPU and says so; forcing the kernels fails
    loudly."""
    from dynode_b200._lib import DynodeError
    from dynode_b200.infer.nuts import BatchedNUTS
    dev = torch.device("cuda", 0)
    D = 20
    mu = torch.linspace(-2.0, 2.0, D, dtype=torch.float64, device=dev)
    sd = torch.linspace(0.5, 2.0, D, dtype=torch.float64, device=dev)

    def pg(Z):
        d = (Z - mu) / sd
        return 0.5 * (d * d).sum(1), d / sd

    eng = BatchedNUTS(pg, max_tree_depth=6, generator=torch.Generator(device=dev).manual_seed(3))
    zs, stats, _ = eng.run(torch.zeros(64, D, dtype=torch.float64, device=dev), 150, 100)
    assert not eng.kernels_used
    flat = zs.reshape(-1, D)
    assert float(((flat.mean(0) - mu).abs() / sd).max()) < 0.15
    assert float((flat.std(0) / sd - 1.0).abs().max()) < 0.15
    forced = BatchedNUTS(pg, max_tree_depth=6, cuda_kernels=True, cuda_graph=False)
    with pytest.raises(DynodeError, match="dimension"):
        forced.run(torch.zeros(4, D, dtype=torch.float64, device=dev), 5, 5)
