"""GPU parity: the sm_100a kernels (through the C ABI) against the CPU oracle on the same seeded
inputs.  Bar (BASELINE.json north_star): 1e-6 relative in FP64; we hold the kernels to 1e-9 with
identical accepted/rejected step counts (the canary for step-sequence parity)."""
import numpy as np
import pytest

from tests.cases import ALL_CASES, EXTRA_CASES, make_case

pytestmark = pytest.mark.gpu

RTOL = 1e-9  # relative, on top of ATOL_SCALE * max|y| absolute (values decaying through ~0)
ATOL_SCALE = 1e-12


def _run_engine(case, t1, **kw):
    import torch
    from dynode_b200.engine import SolverOptions, solve_ensemble
    opts = SolverOptions(t1=t1, **kw.pop("opts", {}))
    save_ts = kw.pop("save_ts", np.linspace(0.0, t1, int(t1 // 1) + 1))
    ys, dys, stats = solve_ensemble(case["model"], case["y0"], case["params"], case["contact"], opts,
                                    save_ts, **kw)
    torch.cuda.synchronize()
    return ys.cpu().numpy(), (None if dys is None else dys.cpu().numpy()), stats.cpu().numpy()


def _run_oracle(case, t1, **kw):
    from oracle import oracle as orc
    fam, dims, theta, shared = case["oracle"]
    return orc.solve(fam, dims, case["y0"], theta, shared, t1=t1, **kw)


def _assert_close(got, ref, rtol=RTOL, atol_scale=ATOL_SCALE):
    scale = np.max(np.abs(ref[np.isfinite(ref)])) if np.isfinite(ref).any() else 1.0
    bad = ~(np.abs(got - ref) <= atol_scale * scale + rtol * np.abs(ref))
    bad &= ~((got == ref))  # inf == inf
    assert not bad.any(), f"max rel err {np.nanmax(np.abs(got - ref) / (np.abs(ref) + atol_scale * scale)):.3e}"


@pytest.mark.parametrize("name", ALL_CASES + EXTRA_CASES)
def test_saved_trajectories_match_oracle(name):
    B = 257  # ragged: not a multiple of trajectories-per-warp/CTA
    case = make_case(name, B)
    ys, _, st = _run_engine(case, case["t1"])
    ref, _, rst = _run_oracle(case, case["t1"])
    assert ys.shape == ref.shape
    assert np.array_equal(st, rst), "accepted/rejected step counts differ from the oracle"
    assert np.all(st[:, 0] == 0)
    _assert_close(ys, ref)
    # ys[0] == y0 exactly (reference tests/test_simulation/test_odes.py:63-74)
    y0 = np.broadcast_to(case["y0"], (B, ys.shape[2]))
    assert np.array_equal(ys[:, 0, :], y0)


def test_partial_save_mask_and_nonuniform_grid():
    """sub_save_indices / save_step paths (reference odes.py:177-193): generic store path + loaded grid."""
    case = make_case("seirs_multi_a2s3", 101)
    t1 = 100
    ts = np.linspace(0.0, t1, int(t1 // 3) + 1)  # save_step=3 -> 34 points spaced 100/33
    from oracle import oracle as orc
    fam, dims, theta, shared = case["oracle"]
    for mask, idx in ((0b10001, list(range(0, 2)) + list(range(20, 26))), (0b00100, list(range(8, 14)))):
        ys, _, st = _run_engine(case, t1, save_ts=ts, save_mask=mask)
        ref, _, rst = orc.solve(fam, dims, case["y0"], theta, shared, t1=t1, save_ts=ts, save_idx=idx)
        assert ys.shape == ref.shape and np.array_equal(st, rst)
        _assert_close(ys, ref)


def test_constant_step_and_max_steps():
    case = make_case("sir_age2", 40)
    ys, _, st = _run_engine(case, 60, opts=dict(const_dt=0.5))
    ref, _, rst = _run_oracle(case, 60, const_dt=0.5)
    assert np.array_equal(st, rst) and np.all(st[:, 1] == 120)
    _assert_close(ys, ref)
    ys, _, st = _run_engine(case, 100, opts=dict(max_steps=7))
    ref, _, rst = _run_oracle(case, 100, max_steps=7)
    assert np.array_equal(st, rst) and np.all(st[:, 0] == 1)
    assert np.array_equal(np.isinf(ys), np.isinf(ref))  # unreached slots keep +inf
    fin = np.isfinite(ref)
    _assert_close(ys[fin], ref[fin])


def test_zero_infection_edge_and_tight_tolerance():
    """f == 0 initial-step edge (reference tests/test_sir_dynamics/test_sir.py:75) and rtol=1e-10."""
    import dynode_b200._lib as L
    from dynode_b200.engine import FlowModel
    model = FlowModel(L.FLOW_SIR, 0, 1, 1)
    case = dict(model=model, params=dict(beta=np.array([[2 / 7]]), gamma=np.array([[1 / 7]])), contact=None,
                y0=np.array([0.8, 0.0, 0.2]), oracle=(0, (1, 1, 1), np.array([[2 / 7, 1 / 7]]), None))
    ys, _, st = _run_engine(case, 120)
    ref, _, rst = _run_oracle(case, 120)
    assert np.array_equal(st, rst) and np.all(np.isfinite(ys))
    _assert_close(ys, ref)
    case2 = make_case("seirs_seasonal", 64)
    ys, _, st = _run_engine(case2, 200, opts=dict(rtol=1e-10, atol=1e-12))
    ref, _, rst = _run_oracle(case2, 200, rtol=1e-10, atol=1e-12)
    assert np.array_equal(st, rst)
    _assert_close(ys, ref, rtol=1e-8)


WRT = {
    # case -> (engine wrt ids, oracle theta indices)
    "sir_age2": ([0 * 16 + 0, 1 * 16 + 0], [0, 1]),
    "sir_1bin": ([0 * 16 + 0, 1 * 16 + 0], [0, 1]),
    "seirs_seasonal": ([0, 16, 32, 48, 64, 80], [0, 1, 2, 3, 4, 5]),
    "seirs_multi_a2s3": ([0, 2, 16 + 1, 32 + 2, 48 + 0], [0, 2, 4, 8, 9]),
    "sir_age_risk32": ([0, 16], [0, 1]),
}


@pytest.mark.parametrize("name", sorted(WRT))
def test_forward_sensitivities_match_oracle(name):
    B = 37
    case = make_case(name, B)
    t1 = min(case["t1"], 120)
    wrt_e, wrt_o = WRT[name]
    ys, dys, st = _run_engine(case, t1, wrt=wrt_e)
    ref, dref, rst = _run_oracle(case, t1, wrt=wrt_o)
    assert np.array_equal(st, rst)
    _assert_close(ys, ref)
    assert dys.shape == dref.shape
    for p in range(len(wrt_e)):
        _assert_close(dys[..., p], dref[..., p], rtol=1e-8, atol_scale=1e-11)


def test_sensitivities_with_initial_state_tangents():
    B = 19
    case = make_case("seirs_multi_a2s3", B)
    rng = np.random.default_rng(0)
    dy0 = rng.normal(size=(B, 2, 26))
    ys, dys, st = _run_engine(case, 90, wrt=[-1, 0], dy0=dy0)
    ref, dref, rst = _run_oracle(case, 90, wrt=[-1, 0], dy0=dy0)
    assert np.array_equal(st, rst)
    _assert_close(dys, dref, rtol=1e-8, atol_scale=1e-11)


@pytest.mark.parametrize("name,obs_comp,wrt_e,wrt_o", [
    ("sir_age2", 2, [0, 16], [0, 1]),            # NUTS config: Poisson on diff(R), d/d(beta, gamma)
    ("seirs_multi_a2s3", 4, [0, 1, 2, 16, 17, 18], [0, 1, 2, 3, 4, 5]),  # Poisson on diff(C)
    ("sir_age2", 0, [], []),
    ("seirs_seasonal", 3, [], []),   # P == 0 on a shared-memory-offload instance (coefficients stay in registers)
    ("seirs_1bin", 2, [], []),
    ("seirs_seasonal", 1, [0, 16, 64], [0, 1, 4]),
])
def test_fused_poisson_loglik_and_gradient(name, obs_comp, wrt_e, wrt_o):
    import torch
    from scipy.special import gammaln
    from dynode_b200.engine import SolverOptions, poisson_loglik_grad
    from oracle import oracle as orc
    B = 53
    case = make_case(name, B)
    t1 = 100
    fam, dims, theta, shared = case["oracle"]
    sizes = case["model"].compartment_sizes()
    lo = sum(sizes[:obs_comp])
    idx = list(range(lo, lo + sizes[obs_comp]))
    # synthetic observations: increments of trajectory 0 (non-integer, as in the reference example)
    truth, _, _ = orc.solve(fam, dims, case["y0"][:1] if np.ndim(case["y0"]) == 2 else case["y0"], theta[:1],
                            shared, t1=t1, save_idx=idx)
    obs = np.abs(np.diff(truth[0], axis=0)) + 0.05
    lp_const = float(-gammaln(obs + 1).sum())
    ts = np.linspace(0.0, t1, t1 + 1)
    lp, grad, st = poisson_loglik_grad(case["model"], case["y0"], case["params"], case["contact"],
                                       SolverOptions(t1=t1), ts, obs_comp, obs, lp_const, wrt=wrt_e)
    torch.cuda.synchronize()
    ys, dys, rst = orc.solve(fam, dims, case["y0"], theta, shared, t1=t1, save_idx=idx, wrt=wrt_o)
    lp_ref, g_ref = orc.poisson_incidence(ys, dys, obs)
    assert np.array_equal(st.cpu().numpy(), rst)
    assert np.allclose(lp.cpu().numpy(), lp_ref, rtol=1e-10, atol=0)
    if wrt_e:
        g = grad.cpu().numpy()
        assert np.all(np.abs(g - g_ref) <= 1e-8 * np.abs(g_ref) + 1e-9 * np.abs(g_ref).max())


@pytest.mark.parametrize("name", ["seirs_seasonal", "seirs_multi_a2s3", "sir_age2"])
def test_discontinuity_points_match_oracle(name):
    """SolverParams.discontinuity_points -> ClipStepSizeController(jump_ts) (reference odes.py:120-131)."""
    B = 131
    case = make_case(name, B)
    t1 = min(case["t1"], 200)
    jumps = (30.0, 100.5, 150.0)
    ys, _, st = _run_engine(case, t1, opts=dict(jump_ts=jumps))
    ref, _, rst = _run_oracle(case, t1, jump_ts=jumps)
    assert np.array_equal(st, rst) and np.all(st[:, 0] == 0)
    _assert_close(ys, ref)
    ys0, _, st0 = _run_engine(case, t1)
    assert not np.array_equal(st, st0)


@pytest.mark.parametrize("name,obs_comp", [("sir_age2", 2), ("seirs_multi_a2s3", 4), ("seirs_seasonal", 3),
                                           ("seirs_multi_g6s3", 4), ("sir_density", 0), ("seirs_multi_g3s2", 2)])
def test_discrete_adjoint_equals_forward_sensitivities(name, obs_comp):
    """dynode_poisson_loglik_adjoint_f64 (one reverse sweep, all parameters + y0) against
    dynode_poisson_loglik_grad_f64 (forward tangents, one direction at a time): same lp, same gradient."""
    import torch
    from dynode_b200 import _lib
    from dynode_b200.engine import SolverOptions, poisson_loglik_adjoint, poisson_loglik_grad
    from oracle import oracle as orc
    B = 41
    case = make_case(name, B)
    model = case["model"]
    S, n = model.n_strains, model.state_size
    t1 = 90
    fam, dims, theta, shared = case["oracle"]
    sizes = model.compartment_sizes()
    lo = sum(sizes[:obs_comp])
    idx = list(range(lo, lo + sizes[obs_comp]))
    truth, _, _ = orc.solve(fam, dims, case["y0"][:1] if np.ndim(case["y0"]) == 2 else case["y0"], theta[:1], shared,
                            t1=t1, save_idx=idx)
    obs = np.abs(np.diff(truth[0], axis=0)) + 0.05
    ts = np.linspace(0.0, t1, t1 + 1)
    kinds = [_lib.P_BETA, _lib.P_GAMMA] + ([_lib.P_SIGMA, _lib.P_OMEGA] if model.flow != _lib.FLOW_SIR else [])
    wrt = [_lib.wrt_id(k, s) for k in kinds for s in range(S)]
    cols = [k * S + s for k in kinds for s in range(S)]
    if model.flags & _lib.FLAG_SEASONAL:
        wrt += [_lib.wrt_id(_lib.P_SEASON_AMP, 0), _lib.wrt_id(_lib.P_SEASON_PHASE, 0)]
        cols += [4 * S, 4 * S + 1]
    y0 = np.broadcast_to(case["y0"], (B, n)).copy()
    # forward mode, including initial-state directions (a subset when the state is large: 64 directions max)
    y0_dirs = list(range(n)) if n <= 40 else [0, 5, 6, 17, 24, 33, 42, 60, 77]
    dy0 = np.zeros((B, len(wrt) + len(y0_dirs), n))
    for j, e in enumerate(y0_dirs):
        dy0[:, len(wrt) + j, e] = 1.0
    lp_f, g_f, st_f = poisson_loglik_grad(model, y0, case["params"], case["contact"], SolverOptions(t1=t1), ts,
                                          obs_comp, obs, 1.5, wrt=wrt + [-1] * len(y0_dirs), dy0=dy0)
    lp_a, g_a, g0_a, st_a = poisson_loglik_adjoint(model, y0, case["params"], case["contact"], SolverOptions(t1=t1),
                                                   ts, obs_comp, obs, 1.5, with_y0_grad=True)
    torch.cuda.synchronize()
    assert torch.equal(st_f, st_a) and int((st_a[:, 0] != 0).sum()) == 0
    assert torch.allclose(lp_a, lp_f, rtol=1e-12)
    gf = g_f.cpu().numpy()
    ga = g_a.cpu().numpy()[:, cols]
    scale = np.abs(gf[:, :len(wrt)]).max(axis=0, keepdims=True) + 1e-300
    assert np.all(np.abs(ga - gf[:, :len(wrt)]) <= 1e-8 * np.abs(gf[:, :len(wrt)]) + 1e-9 * scale)
    g0f = gf[:, len(wrt):]
    g0a = g0_a.cpu().numpy()[:, y0_dirs]
    assert np.all(np.abs(g0a - g0f) <= 1e-8 * np.abs(g0f) + 1e-9 * np.abs(g0f).max())
    # parameters the flow does not have get a zero gradient
    others = [c for c in range(4 * S + 2) if c not in cols]
    assert np.all(g_a.cpu().numpy()[:, others] == 0.0)
    # checkpoint capacity exceeded -> flagged, NaN, no crash
    lp_c, g_c, _, st_c = poisson_loglik_adjoint(model, y0, case["params"], case["contact"], SolverOptions(t1=t1),
                                                ts, obs_comp, obs, 0.0, cap=3)
    assert bool((st_c[:, 0] == 2).all()) and bool(torch.isnan(lp_c).all()) and bool(torch.isnan(g_c).all())


@pytest.mark.parametrize("A,K,W", [(3, 2, 3), (4, 3, 4), (1, 1, 2), (5, 2, 1), (2, 4, 3)])
def test_seip_cta_per_trajectory_kernel_matches_oracle(A, K, W):
    """CTA-per-trajectory kernel of the immune-history / waning family (include/dynode_b200_seip.h)."""
    import torch
    from dynode_b200 import seip
    from dynode_b200.engine import SolverOptions
    from tests.cases import make_seip_case
    B = 37
    case = make_seip_case(B, A=A, K=K, W=W, t1=150)
    fam, dims, theta, shared = case["oracle"]
    ts = np.linspace(0.0, 150.0, 151)
    ys, st = seip.solve_ensemble(case["model"], case["y0"], case["params"], case["contact"], case["pop"],
                                 case["immunity"], SolverOptions(t1=150.0), ts)
    torch.cuda.synchronize()
    ref, _, rst = _run_oracle(case, 150)
    assert np.array_equal(st.cpu().numpy(), rst)
    _assert_close(ys.cpu().numpy(), ref)
    assert np.array_equal(ys[:, 0, :].cpu().numpy(), np.broadcast_to(case["y0"], (B, ref.shape[2])))
    # max_steps and a non-uniform grid
    ts2 = np.array([0.0, 0.5, 7.25, 100.0, 150.0])
    ys2, st2 = seip.solve_ensemble(case["model"], case["y0"], case["params"], case["contact"], case["pop"],
                                   case["immunity"], SolverOptions(t1=150.0, max_steps=9), ts2)
    ref2, _, rst2 = _run_oracle(case, 150, save_ts=ts2, max_steps=9)
    assert np.array_equal(st2.cpu().numpy(), rst2) and np.all(rst2[:, 0] == 1)
    got2 = ys2.cpu().numpy()
    assert np.array_equal(np.isinf(got2), np.isinf(ref2))
    fin = np.isfinite(ref2)
    _assert_close(got2[fin], ref2[fin])


@pytest.mark.gpu
def test_row_mask_leaves_masked_trajectories_out():
    """DynodeSolverDesc.only: masked rows are neither computed nor written, the others are bit-identical to
    the unmasked launch -- through the plain solve (one generation per warp), the persistent-slot fused
    log-likelihood with tangents (candidates scanned by ballot) and the adjoint."""
    import torch
    from dynode_b200 import _lib, engine
    from tests.cases import make_case

    dev = torch.device("cuda", 0)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    g = torch.Generator(device=dev).manual_seed(5)
    for name, B in (("seirs_multi_a2s3", 333), ("sir_age2", 1000), ("seirs_seasonal", 777)):
        case = make_case(name, B)
        model, t1 = case["model"], float(min(case["t1"], 120))
        prm = {k: t(v) for k, v in case["params"].items()}
        y0 = t(case["y0"])
        contact = None if case["contact"] is None else t(case["contact"])
        ts = np.linspace(0.0, t1, int(t1) + 1)
        opts = engine.SolverOptions(t1=t1)
        for frac in (0.5, 0.03, 0.0):
            mask = (torch.rand(B, device=dev, generator=g) < frac)
            ys0, _, st0 = engine.solve_ensemble(model, y0, prm, contact, opts, ts, B=B)
            with engine.only_rows(mask.view(torch.uint8)):
                ys1, _, st1 = engine.solve_ensemble(model, y0, prm, contact, opts, ts, B=B)
            assert torch.equal(ys1[mask], ys0[mask]) and torch.equal(st1[mask], st0[mask])
            assert not ys1[~mask].any() and not st1[~mask].any()
            obs_comp = model.n_compartments - 1
            m = model.compartment_sizes()[obs_comp]
            obs = torch.rand(len(ts) - 1, m, dtype=torch.float64, device=dev, generator=g) + 0.1
            wrt = [_lib.wrt_id(_lib.P_BETA, 0), _lib.wrt_id(_lib.P_GAMMA, 0)]
            lp0, g0, s0 = engine.poisson_loglik_grad(model, y0, prm, contact, opts, ts, obs_comp, obs, wrt=wrt, B=B)
            with engine.only_rows(mask.view(torch.uint8)):
                lp1, g1, s1 = engine.poisson_loglik_grad(model, y0, prm, contact, opts, ts, obs_comp, obs, wrt=wrt, B=B)
            assert torch.equal(lp1[mask], lp0[mask]) and torch.equal(g1[mask], g0[mask]) and torch.equal(s1[mask], s0[mask])
            assert not lp1[~mask].any() and not g1[~mask].any()
            la0, ga0, _, sa0 = engine.poisson_loglik_adjoint(model, y0, prm, contact, opts, ts, obs_comp, obs, B=B)
            with engine.only_rows(mask.view(torch.uint8)):
                la1, ga1, _, sa1 = engine.poisson_loglik_adjoint(model, y0, prm, contact, opts, ts, obs_comp, obs, B=B)
            assert torch.equal(la1[mask], la0[mask]) and torch.equal(ga1[mask], ga0[mask])
            assert not la1[~mask].any() and not ga1[~mask].any()


@pytest.mark.parametrize("name", ["seirs_multi_a2s3", "seirs_seasonal", "sir_age2"])
def test_tiny_ensembles_and_degenerate_grids(name):
    """Ensemble sizes around the slot / warp granularity (1, 2, one short of and one past a warp's slots), a
    save grid that is t0 alone, a two-point grid, two save times 1e-9 apart, and an empty horizon (t1 == t0:
    nothing to integrate, zero steps) -- all against the oracle."""
    for B in (1, 2, 4, 6, 31, 33):
        case = make_case(name, B)
        ys, _, st = _run_engine(case, 30.0)
        ref, _, rst = _run_oracle(case, 30.0)
        assert np.array_equal(st, rst)
        _assert_close(ys, ref)
    case = make_case(name, 37)
    for grid in (np.array([0.0]), np.array([0.0, 12.5]), np.array([0.0, 3.0, 3.0 + 1e-9, 12.5])):
        ys, _, st = _run_engine(case, 12.5, save_ts=grid)
        ref, _, rst = _run_oracle(case, 12.5, save_ts=grid)
        assert ys.shape == ref.shape == (37, len(grid), ref.shape[2])
        assert np.array_equal(st, rst)
        _assert_close(ys, ref)
    # empty horizon: the loop body never runs (tprev < t1 is false from the start), so no save is reached and the
    # output keeps diffrax's +inf fill -- whatever the restated loop does, kernel and oracle must agree
    ys, _, st = _run_engine(case, 0.0, save_ts=np.array([0.0]))
    ref, _, rst = _run_oracle(case, 0.0, save_ts=np.array([0.0]))
    assert np.array_equal(st, rst) and np.all(st[:, 3] == 0)
    assert np.array_equal(ys, ref)


@pytest.mark.parametrize("name", ["sir_age2", "seirs_seasonal", "seirs_multi_a2s3"])
def test_discontinuity_points_with_sensitivities_and_loglik(name):
    """jump_ts together with forward tangents (dynode_solve_sens_f64) and with the fused Poisson log-likelihood +
    gradient: the tangents ride the clipped step sequence like any other (the clip times are constants), so both
    must agree with the oracle run with the same jump_ts -- identical step counts included."""
    import torch
    from scipy.special import gammaln
    from dynode_b200.engine import SolverOptions, poisson_loglik_grad
    from oracle import oracle as orc
    B = 41
    case = make_case(name, B)
    t1 = 100
    jumps = (20.0, 55.5, 80.0)
    wrt_e, wrt_o = WRT[name]
    ys, dys, st = _run_engine(case, t1, wrt=wrt_e, opts=dict(jump_ts=jumps))
    ref, dref, rst = _run_oracle(case, t1, wrt=wrt_o, jump_ts=jumps)
    assert np.array_equal(st, rst) and np.all(st[:, 0] == 0)
    _assert_close(ys, ref)
    for p in range(len(wrt_e)):
        _assert_close(dys[..., p], dref[..., p], rtol=1e-8, atol_scale=1e-11)
    _, _, st_plain = _run_engine(case, t1, wrt=wrt_e)
    assert not np.array_equal(st, st_plain)  # the jumps did clip steps
    # fused log-likelihood on the last compartment
    model = case["model"]
    obs_comp = model.n_compartments - 1
    sizes = model.compartment_sizes()
    lo = sum(sizes[:obs_comp])
    idx = list(range(lo, lo + sizes[obs_comp]))
    fam, dims, theta, shared = case["oracle"]
    ys_o, dys_o, rst2 = orc.solve(fam, dims, case["y0"], theta, shared, t1=t1, save_idx=idx, wrt=wrt_o, jump_ts=jumps)
    obs = np.abs(np.diff(ys_o[0], axis=0)) + 0.05
    lp_const = float(-gammaln(obs + 1).sum())
    ts = np.linspace(0.0, t1, t1 + 1)
    for w_e in (wrt_e, []):
        lp, grad, st2 = poisson_loglik_grad(model, case["y0"], case["params"], case["contact"],
                                            SolverOptions(t1=t1, jump_ts=jumps), ts, obs_comp, obs, lp_const, wrt=w_e)
        torch.cuda.synchronize()
        lp_ref, g_ref = orc.poisson_incidence(ys_o, dys_o, obs)
        assert np.array_equal(st2.cpu().numpy(), rst2)
        assert np.allclose(lp.cpu().numpy(), lp_ref, rtol=1e-10, atol=0)
        if w_e:
            g = grad.cpu().numpy()
            assert np.all(np.abs(g - g_ref) <= 1e-8 * np.abs(g_ref) + 1e-9 * np.abs(g_ref).max())


def test_fused_loglik_on_a_nonuniform_grid():
    """The fused log-likelihood with observation times that are not build_saveat's uniform grid (the kernel then
    loads the save times instead of generating them)."""
    import torch
    from scipy.special import gammaln
    from dynode_b200.engine import SolverOptions, poisson_loglik_grad
    from oracle import oracle as orc
    B = 29
    for name, obs_comp, wrt_e, wrt_o in (("seirs_multi_a2s3", 4, [0, 17], [0, 4]), ("seirs_seasonal", 3, [], [])):
        case = make_case(name, B)
        t1 = 90.0
        ts = np.array([0.0, 0.5, 1.75, 4.0, 4.0 + 1e-7, 11.0, 30.0, 30.5, 62.0, 89.0, 90.0])
        model = case["model"]
        sizes = model.compartment_sizes()
        lo = sum(sizes[:obs_comp])
        idx = list(range(lo, lo + sizes[obs_comp]))
        fam, dims, theta, shared = case["oracle"]
        ys_o, dys_o, rst = orc.solve(fam, dims, case["y0"], theta, shared, t1=t1, save_ts=ts, save_idx=idx, wrt=wrt_o)
        obs = np.abs(np.diff(ys_o[0], axis=0)) + 0.05
        lp_const = float(-gammaln(obs + 1).sum())
        lp, grad, st = poisson_loglik_grad(model, case["y0"], case["params"], case["contact"], SolverOptions(t1=t1),
                                           ts, obs_comp, obs, lp_const, wrt=wrt_e)
        torch.cuda.synchronize()
        lp_ref, g_ref = orc.poisson_incidence(ys_o, dys_o, obs)
        assert np.array_equal(st.cpu().numpy(), rst)
        assert np.allclose(lp.cpu().numpy(), lp_ref, rtol=1e-9, atol=0)
        if wrt_e:
            g = grad.cpu().numpy()
            assert np.all(np.abs(g - g_ref) <= 1e-8 * np.abs(g_ref) + 1e-9 * np.abs(g_ref).max())
