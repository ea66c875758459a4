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
    # the discrete adjoint over the same clipped step sequence (its forward sweep clips, its reverse sweep rebuilds
    # every step from its checkpoint): same counts, same log-likelihood, same gradient, initial-state gradient finite
    from dynode_b200 import _lib
    from dynode_b200.engine import poisson_loglik_adjoint
    lp_a, g_a, g0_a, st_a = poisson_loglik_adjoint(model, case["y0"], case["params"], case["contact"],
                                                   SolverOptions(t1=t1, jump_ts=jumps), ts, obs_comp, obs, lp_const,
                                                   with_y0_grad=True)
    torch.cuda.synchronize()
    assert np.array_equal(st_a.cpu().numpy(), rst2)
    assert np.allclose(lp_a.cpu().numpy(), lp_ref, rtol=1e-10, atol=0)
    S = model.n_strains
    cols = [(w >> 4) * S + (w & 15) if (w >> 4) < 4 else 4 * S + ((w >> 4) - 4) for w in wrt_e]
    ga = g_a.cpu().numpy()[:, cols]
    assert np.all(np.abs(ga - g_ref) <= 1e-6 * np.abs(g_ref) + 1e-8 * np.abs(g_ref).max())
    assert bool(torch.isfinite(g0_a).all())


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


def _public_api_solver(name, B, jump_ts=(), const_dt=0.0, save_step=1, sub_save=None, rtol=1e-5, atol=1e-6):
    """tests/golden_check.py solver: the CUDA path through the PUBLIC API -- `simulate_ensemble` with the registered
    example RHS, its ODEParams dataclass and a `SolverParams`, i.e. the call the golden file's producer made on the
    reference (`dynode.simulation.simulate` under `jax.vmap`)."""
    import torch
    from dynode_b200.config import SolverParams
    from dynode_b200.examples import rhs as ex
    from dynode_b200.simulation import simulate_ensemble
    case = make_case(name, B)
    dev = torch.device("cuda", 0)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    p = {k: t(v) for k, v in case["params"].items()}
    m = case["model"]
    G, S = m.n_groups, m.n_strains
    if name == "sir_1bin":
        ode, prm, shapes = ex.sir_ode, ex.SIR_ODEParams(beta=p["beta"], gamma=p["gamma"]), [(1,)] * 3
    elif name == "sir_density":
        ode, prm, shapes = ex.sir_density_ode, ex.DensitySIR_ODEParams(beta=p["beta"], gamma=p["gamma"]), [(1,)] * 3
    elif name == "seirs_1bin":
        ode, prm, shapes = ex.seirs_ode, ex.SEIRS_ODEParams(**p), [(1,)] * 4
    elif name == "seirs_seasonal":
        sp = ex.SeasonalityParams(forcing_amp=p["season_amp"], forcing_phase=p["season_phase"],
                                  forcing_period=p["season_period"])
        ode, shapes = ex.seirs_ode_seasonal, [(1,)] * 4
        prm = ex.SeasonalSEIRS_ODEParams(beta=p["beta"], gamma=p["gamma"], sigma=p["sigma"], omega=p["omega"],
                                         seasonality_params=sp)
    elif name.startswith("sir_age_risk"):
        ode, shapes = ex.sir_age_risk_ode, [(3, 2)] * 3
        prm = ex.AgeRiskSIR_ODEParams(beta=p["beta"], gamma=p["gamma"], contact_matrix=t(case["oracle"][3]))
    elif name.startswith("sir_age"):
        ode, shapes = ex.sir_age_ode, [(G,)] * 3
        prm = ex.AgeSIR_ODEParams(beta=p["beta"], gamma=p["gamma"], contact_matrix=t(case["contact"]))
    else:
        ode, shapes = ex.seirs_multi_strain_ode, [(G,)] + [(G, S)] * 4
        prm = ex.SEIRS_MultiStrain_ODEParams(beta=p["beta"], gamma=p["gamma"], sigma=p["sigma"], omega=p["omega"],
                                             contact_matrix=t(case["contact"]))
    sizes = [int(np.prod(s)) for s in shapes]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    y0 = t(np.broadcast_to(case["y0"], (B, int(offs[-1]))))
    state = tuple(y0[:, offs[i]:offs[i + 1]].reshape(B, *shapes[i]) for i in range(len(shapes)))
    sp = SolverParams(discontinuity_points=list(jump_ts), constant_step_size=const_dt,
                      ode_solver_rel_tolerance=rtol, ode_solver_abs_tolerance=atol)
    sub = (0, len(shapes) - 1) if sub_save == "first_last" else sub_save  # or a tuple of compartment indices
    sol = simulate_ensemble(ode, case["t1"], state, prm, sp, sub_save_indices=sub, save_step=save_step,
                            batch_size=B, state_batched=True)
    torch.cuda.synchronize()
    T = sol.ts.shape[0]
    if sub is not None:
        for c, y in enumerate(sol.ys):
            assert (y.shape == (B, T, 0)) == (c not in sub)  # unsaved compartments: (T, 0), reference odes.py:182-193
    ys = torch.cat([y.reshape(B, T, -1) for y in sol.ys], dim=2).cpu().numpy()
    st = np.stack([np.zeros(B, np.int64)] + [sol.stats[k].cpu().numpy() for k in
                                             ("num_accepted_steps", "num_rejected_steps", "num_steps")], axis=1)
    st[:, 0] = sol.result.cpu().numpy()
    return ys, st


@pytest.mark.parametrize("kind", ["diffrax", "standin"])
def test_cuda_path_matches_golden(kind):
    """GPU twin of tests/test_oracle.py::test_oracle_matches_golden: every key of the golden file against the CUDA
    kernels reached through `simulate_ensemble` (all cases; discontinuity points, constant step, save_step 2/3/7,
    sub-save, tight tolerances; accepted / rejected / total step counts)."""
    from tests import golden_check as gc
    gold = gc.load(kind)
    if gold is None:
        assert kind == "diffrax", "tests/golden/standin_golden.npz must be committed"
        pytest.skip("PARITY UNPINNED: tests/golden/diffrax_golden.npz absent (see baseline/dump_diffrax_golden.py)")
    assert gc.check(gold, _public_api_solver) == 9 + 3 * 7
    if "c2/potential" in gold:
        import torch
        from dynode_b200.examples import sir_infer_parameters as c2
        from dynode_b200.infer.model_density import ModelDensity
        dev = torch.device("cuda", 0)
        obs = torch.as_tensor(gold["c2/obs"], dtype=torch.float64, device=dev)
        for model in (c2.model_fused, c2.model):
            md = ModelDensity(model, model_kwargs=dict(config=c2.get_config(), tf=100, obs_data=obs), device=dev)
            cols = [list(md.sites).index(str(nm)) for nm in gold["c2/site_names"]]
            assert all(md.sites[str(nm)]["slice"] == (c, c + 1) for nm, c in zip(gold["c2/site_names"], cols))
            Z = torch.zeros((gold["c2/z"].shape[0], md.dim), dtype=torch.float64, device=dev)
            Z[:, cols] = torch.as_tensor(gold["c2/z"], dtype=torch.float64, device=dev)
            U, dU = md.potential_and_grad(Z)
            assert np.allclose(U.cpu().numpy(), gold["c2/potential"], rtol=1e-6)
            g = dU[:, cols].cpu().numpy()
            assert np.allclose(g, gold["c2/grad"], rtol=1e-6, atol=1e-6 * np.abs(gold["c2/grad"]).max())


def test_forward_loglik_reports_nan_when_max_steps_is_reached():
    """A trajectory that runs out of `max_steps` (the reference raises: diffrax throw=True) must not return a
    plausible truncated log-likelihood: lp and its gradient are NaN, stats carry the result code -- forward-mode
    kernel and adjoint kernel alike."""
    import torch
    from dynode_b200 import _lib
    from dynode_b200.engine import SolverOptions, poisson_loglik_adjoint, poisson_loglik_grad
    B = 37
    case = make_case("sir_age2", B)
    t1 = 100
    obs = np.full((t1, 2), 3.0)
    ts = np.linspace(0.0, t1, t1 + 1)
    wrt = [_lib.wrt_id(_lib.P_BETA, 0), _lib.wrt_id(_lib.P_GAMMA, 0)]
    ok = poisson_loglik_grad(case["model"], case["y0"], case["params"], case["contact"], SolverOptions(t1=t1), ts, 2,
                             obs, 0.0, wrt=wrt)
    few = int(ok[2][:, _lib.STAT_STEPS].min().item())  # every trajectory needs more attempts than this - 1
    lp, grad, st = poisson_loglik_grad(case["model"], case["y0"], case["params"], case["contact"],
                                       SolverOptions(t1=t1, max_steps=few - 1), ts, 2, obs, 0.0, wrt=wrt)
    lp_a, g_a, _, st_a = poisson_loglik_adjoint(case["model"], case["y0"], case["params"], case["contact"],
                                                SolverOptions(t1=t1, max_steps=few - 1), ts, 2, obs, 0.0)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(ok[0]).all()) and bool(torch.isfinite(ok[1]).all())
    assert bool((st[:, _lib.STAT_RESULT] == _lib.RESULT_MAX_STEPS).all())
    assert bool(torch.isnan(lp).all()) and bool(torch.isnan(grad).all())
    assert bool((st_a[:, _lib.STAT_RESULT] == _lib.RESULT_MAX_STEPS).all()) and bool(torch.isnan(lp_a).all())
    # P == 0 instance too
    lp0, _, st0 = poisson_loglik_grad(case["model"], case["y0"], case["params"], case["contact"],
                                      SolverOptions(t1=t1, max_steps=few - 1), ts, 2, obs, 0.0)
    assert bool(torch.isnan(lp0).all())


def test_adjoint_capacity_overflow_falls_back_to_forward_sensitivities(monkeypatch):
    """More accepted steps than the adjoint's checkpoint scratch holds: the differentiable path re-evaluates those rows
    by forward sensitivities in the same call -- the caller sees the same finite lp and gradient as forward mode."""
    import torch
    from dynode_b200 import engine
    from dynode_b200.examples import sir_infer_parameters as c2
    from dynode_b200.infer.model_density import ModelDensity
    dev = torch.device("cuda", 0)
    obs = c2.synthetic_incidence(100).to(dev)
    Z = torch.randn((64, 2), dtype=torch.float64, device=dev, generator=torch.Generator(dev).manual_seed(5)) * 0.7
    md = ModelDensity(c2.model_fused, model_kwargs=dict(config=c2.get_config(), tf=100, obs_data=obs), device=dev)
    monkeypatch.setenv("DYNODE_B200_ADJOINT", "0")
    U_f, g_f = md.potential_and_grad(Z)
    monkeypatch.setenv("DYNODE_B200_ADJOINT", "1")
    monkeypatch.setenv("DYNODE_B200_ADJOINT_CAP", "20")  # config 2 accepts 15..30 steps: some rows fit, some do not
    engine.adjoint_overflows(reset=True)
    U_a, g_a = md.potential_and_grad(Z)
    n_over = engine.adjoint_overflows(reset=True)
    assert 0 < n_over < 64, n_over
    assert bool(torch.isfinite(U_a).all()) and bool(torch.isfinite(g_a).all())
    assert torch.allclose(U_a, U_f, rtol=1e-10) and torch.allclose(g_a, g_f, rtol=1e-7, atol=1e-7 * float(g_f.abs().max()))


@pytest.mark.parametrize("A,K,W,V,NK,kw", [
    (3, 2, 3, 3, 2, {}),                       # everything on
    (2, 2, 4, 2, 0, dict(season=False)),       # base cubic only, no reset
    (4, 1, 2, 4, 3, dict(intro=False)),        # one strain, four tiers
    (2, 3, 3, 2, 1, dict(vaccinate=False)),    # tiers present but nobody is vaccinated; reset + introductions only
    (3, 2, 3, 1, 0, dict(season=False)),       # V = 1: boosters within the single tier
    (4, 3, 4, 3, 2, {}),                       # n = 1248 (age 4 x hist 8 x vax 3 x wane 4): 10 elements per thread
])
def test_seip_vaccination_kernel_matches_oracle(A, K, W, V, NK, kw):
    """The vaccination extension of the CTA-per-trajectory kernel (tiers, spline rates with the min(.,1) cap, the
    seasonal reset, external introductions; include/dynode_b200_seip.h [V]) against oracle FAM_SEIPV: 1e-9 and the
    same accepted / rejected step counts, directly and through `simulate_ensemble`."""
    import torch
    from dynode_b200 import seip
    from dynode_b200.config import SolverParams
    from dynode_b200.engine import SolverOptions
    from dynode_b200.examples import rhs as ex
    from dynode_b200.simulation import simulate_ensemble
    from tests.cases import make_seipv_case
    B, t1 = (29, 180) if A * (1 << K) * V * (W + 3 * K) < 1000 else (7, 120)
    case = make_seipv_case(B, A=A, K=K, W=W, V=V, NK=NK, t1=t1, **kw)
    ts = np.linspace(0.0, t1, t1 + 1)
    ys, st = seip.solve_ensemble(case["model"], case["y0"], case["params"], case["contact"], case["pop"],
                                 case["immunity"], SolverOptions(t1=float(t1)), ts, vaccination=case["vaccination"],
                                 introductions=case["introductions"], season_tau=case["season_tau"])
    torch.cuda.synchronize()
    ref, _, rst = _run_oracle(case, t1)
    assert np.array_equal(st.cpu().numpy(), rst) and np.all(rst[:, 0] == 0)
    _assert_close(ys.cpu().numpy(), ref)
    # the public API: compartments shaped (A, H, [V,] W) / (A, H, [V,] K), optional fields of SEIP_ODEParams
    H = 1 << K
    dev = torch.device("cuda", 0)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    nS, nX = A * H * V * W, A * H * V * K
    y0 = case["y0"]
    shp_s, shp_x = ((A, H, V, W), (A, H, V, K)) if V > 1 else ((A, H, W), (A, H, K))
    state = (t(y0[:nS]).reshape(shp_s), t(y0[nS:nS + nX]).reshape(shp_x), t(y0[nS + nX:nS + 2 * nX]).reshape(shp_x),
             t(y0[nS + 2 * nX:]).reshape(shp_x))
    prm = case["params"]
    intro, vac = case["introductions"], case["vaccination"]
    imm = case["immunity"] if V > 1 else case["immunity"].reshape(H, W, K)
    p = ex.SEIP_ODEParams(
        beta=t(prm["beta"]), sigma=t(prm["sigma"]), gamma=t(prm["gamma"]), omega=t(prm["omega"]),
        contact_matrix=t(case["contact"]), population=t(case["pop"]), immunity=t(imm),
        vax_base=None if vac is None else t(vac[0]), vax_knots=None if vac is None else t(vac[1]),
        vax_coef=None if vac is None else t(vac[2]),
        intro_time=None if intro is None else t(intro["time"]), intro_scale=None if intro is None else t(intro["scale"]),
        intro_pct=None if intro is None else t(intro["pct"]), intro_ages=None if intro is None else t(intro["ages"]),
        season_tau=case["season_tau"])
    sol = simulate_ensemble(ex.seip_ode, t1, state, p, SolverParams(), batch_size=B, state_batched=False)
    got = torch.cat([c.reshape(B, t1 + 1, -1) for c in sol.ys], dim=2).cpu().numpy()
    assert np.array_equal(got, ys.cpu().numpy())
    assert sol.ys[0].shape == (B, t1 + 1) + shp_s


def test_seip_discontinuity_points_and_sub_save():
    """SolverParams.discontinuity_points and sub_save_indices on the CTA-per-trajectory kernel: steps end at
    prevbefore(jump) and restart at the jump with a fresh f0 (same accepted / rejected counts as the oracle, which
    differ from the unclipped solve); unsaved compartments are neither computed at the save times nor written, and come
    back as (T, 0) through `simulate_ensemble`."""
    import torch
    from dynode_b200 import seip
    from dynode_b200.config import SolverParams
    from dynode_b200.engine import SolverOptions
    from dynode_b200.examples import rhs as ex
    from dynode_b200.simulation import simulate_ensemble
    from tests.cases import make_seipv_case
    B, t1 = 23, 150
    A, K, W, V, NK = 3, 2, 3, 2, 1
    H = 1 << K
    case = make_seipv_case(B, A=A, K=K, W=W, V=V, NK=NK, t1=t1)
    ts = np.linspace(0.0, t1, t1 + 1)
    jumps = (30.0, 75.5, 120.0)
    kw = dict(vaccination=case["vaccination"], introductions=case["introductions"], season_tau=case["season_tau"])
    ys, st = seip.solve_ensemble(case["model"], case["y0"], case["params"], case["contact"], case["pop"],
                                 case["immunity"], SolverOptions(t1=float(t1), jump_ts=jumps), ts, **kw)
    ys0, st0 = seip.solve_ensemble(case["model"], case["y0"], case["params"], case["contact"], case["pop"],
                                   case["immunity"], SolverOptions(t1=float(t1)), ts, **kw)
    torch.cuda.synchronize()
    ref, _, rst = _run_oracle(case, t1, jump_ts=jumps)
    assert np.array_equal(st.cpu().numpy(), rst) and not np.array_equal(st.cpu().numpy(), st0.cpu().numpy())
    _assert_close(ys.cpu().numpy(), ref)
    # constant-step mode ignores the list (reference odes.py:113-131)
    yc, stc = seip.solve_ensemble(case["model"], case["y0"], case["params"], case["contact"], case["pop"],
                                  case["immunity"], SolverOptions(t1=float(t1), const_dt=0.5, jump_ts=jumps), ts, **kw)
    refc, _, rstc = _run_oracle(case, t1, const_dt=0.5)
    assert np.array_equal(stc.cpu().numpy(), rstc)
    _assert_close(yc.cpu().numpy(), refc)
    # sub-save: S and C only (mask 0b1001), on a coarse grid
    nS, nX = A * H * V * W, A * H * V * K
    ts7 = np.linspace(0.0, t1, int(t1 // 7) + 1)
    ym, stm = seip.solve_ensemble(case["model"], case["y0"], case["params"], case["contact"], case["pop"],
                                  case["immunity"], SolverOptions(t1=float(t1)), ts7, save_mask=0b1001, **kw)
    idx = list(range(nS)) + list(range(nS + 2 * nX, nS + 3 * nX))
    refm, _, rstm = _run_oracle(case, t1, save_ts=ts7, save_idx=idx)
    assert ym.shape == (B, len(ts7), nS + nX) and np.array_equal(stm.cpu().numpy(), rstm)
    _assert_close(ym.cpu().numpy(), refm)
    # ... and through the public API
    dev = torch.device("cuda", 0)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    y0 = case["y0"]
    state = (t(y0[:nS]).reshape(A, H, V, W), t(y0[nS:nS + nX]).reshape(A, H, V, K),
             t(y0[nS + nX:nS + 2 * nX]).reshape(A, H, V, K), t(y0[nS + 2 * nX:]).reshape(A, H, V, K))
    prm, intro, vac = case["params"], case["introductions"], case["vaccination"]
    p = ex.SEIP_ODEParams(beta=t(prm["beta"]), sigma=t(prm["sigma"]), gamma=t(prm["gamma"]), omega=t(prm["omega"]),
                          contact_matrix=t(case["contact"]), population=t(case["pop"]), immunity=t(case["immunity"]),
                          vax_base=t(vac[0]), vax_knots=t(vac[1]), vax_coef=t(vac[2]), intro_time=t(intro["time"]),
                          intro_scale=t(intro["scale"]), intro_pct=t(intro["pct"]), intro_ages=t(intro["ages"]),
                          season_tau=case["season_tau"])
    sol = simulate_ensemble(ex.seip_ode, t1, state, p, SolverParams(discontinuity_points=list(jumps)),
                            sub_save_indices=(0, 3), save_step=7, batch_size=B, state_batched=False)
    assert sol.ys[1].shape == (B, len(ts7), 0) and sol.ys[2].shape == (B, len(ts7), 0)
    refj, _, rstj = _run_oracle(case, t1, save_ts=ts7, save_idx=idx, jump_ts=jumps)
    got = torch.cat([sol.ys[0].reshape(B, len(ts7), -1), sol.ys[3].reshape(B, len(ts7), -1)], dim=2).cpu().numpy()
    _assert_close(got, refj)
    assert np.array_equal(sol.stats["num_accepted_steps"].cpu().numpy(), rstj[:, 1])


def test_single_direction_work_items_equal_the_two_direction_kernel():
    """Few chains take the fused log-likelihood with ONE tangent per work item (shorter instruction stream per warp:
    lower latency), many chains the production kernel with two per item: same lp, same gradient, same step counts."""
    import torch
    from dynode_b200 import _lib
    from dynode_b200.engine import SolverOptions, poisson_loglik_grad
    big = make_case("sir_age2", 20000)
    t1 = 100
    obs = np.full((t1, 2), 3.0) + np.linspace(0, 2, t1)[:, None]
    ts = np.linspace(0.0, t1, t1 + 1)
    wrt = [_lib.wrt_id(_lib.P_BETA, 0), _lib.wrt_id(_lib.P_GAMMA, 0)]
    run = lambda prm: poisson_loglik_grad(big["model"], big["y0"], prm, big["contact"], SolverOptions(t1=t1), ts, 2,
                                          obs, 1.25, wrt=wrt)
    lp_b, g_b, st_b = run(big["params"])                                   # 20000 x 2 work items: two per item
    small = {k: v[:61] for k, v in big["params"].items()}
    lp_s, g_s, st_s = run(small)                                           # 61 chains: one direction per item
    torch.cuda.synchronize()
    assert torch.equal(st_s, st_b[:61])
    assert torch.allclose(lp_s, lp_b[:61], rtol=1e-13, atol=0)
    assert torch.allclose(g_s, g_b[:61], rtol=1e-12, atol=1e-12 * float(g_b.abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["seirs_1bin", "seirs_seasonal", "seirs_multi_a2s3"])
def test_output_buffer_that_is_only_8_byte_aligned(name):
    """The C ABI takes any double* for ys.  The one-lane-per-row instances store 16-byte pairs when they can; a
    buffer offset by one double must give the same bits through 8-byte stores, and nothing outside it is written."""
    import torch
    from dynode_b200.engine import SolverOptions, solve_ensemble
    B, t1 = 515, 90
    case = make_case(name, B)
    model = case["model"]
    ts = np.linspace(0.0, t1, t1 + 1)
    numel = B * len(ts) * model.state_size
    opts = SolverOptions(t1=t1)
    ys0, _, st0 = solve_ensemble(model, case["y0"], case["params"], case["contact"], opts, ts, B=B)
    buf = torch.full((numel + 4,), -7.25, dtype=torch.float64, device="cuda")
    assert buf.data_ptr() % 16 == 0
    out = buf[1:1 + numel]
    ys1, _, st1 = solve_ensemble(model, case["y0"], case["params"], case["contact"], opts, ts, out=out, B=B)
    torch.cuda.synchronize()
    assert ys1.data_ptr() % 16 == 8
    assert torch.equal(ys1.reshape(-1), ys0.reshape(-1)) and torch.equal(st1, st0)
    assert float(buf[0]) == -7.25 and bool((buf[1 + numel:] == -7.25).all())
