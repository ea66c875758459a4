"""GPU: the reference's own behavioural and physics tests, run through the drop-in `simulate`
(reference tests/test_simulation/test_odes.py, tests/test_sir_dynamics/test_sir.py,
tests/test_seirs_dynamics/test_seirs.py, tests/test_seirs_seasonality_dynamics/), plus the single-draw
result against the oracle."""
import math

import numpy as np
import pytest
import torch
from scipy.optimize import root_scalar

pytestmark = pytest.mark.gpu

from dynode_b200.config import SolverParams  # noqa: E402
from dynode_b200.examples import rhs as ex  # noqa: E402
from dynode_b200.simulation import simulate, simulate_ensemble  # noqa: E402

DEV = "cuda:0"


def t(x):
    return torch.as_tensor(x, dtype=torch.float64, device=DEV)


def _density_case():
    # reference tests/test_simulation/test_odes.py:31-42
    return (t([99.0]), t([1.0]), t([0.0])), ex.DensitySIR_ODEParams(beta=t(0.003), gamma=t(0.1))


@pytest.mark.parametrize("duration_days", [50, 100, 200, 300.0])
def test_simulate_shapes(duration_days):
    state, p = _density_case()
    sol = simulate(ex.sir_density_ode, duration_days, state, p, SolverParams())
    assert len(sol.ys) == len(state)
    for c in sol.ys:
        assert c.shape == (int(duration_days) + 1, 1)
    assert sol.ts.shape == (int(duration_days) + 1,)


def test_simulate_initial_state_preserved():
    state, p = _density_case()
    sol = simulate(ex.sir_density_ode, 100, state, p, SolverParams())
    for k, comp in enumerate(sol.ys):
        assert torch.equal(comp[0], state[k])  # exactly, as the reference asserts with allclose


@pytest.mark.parametrize("save_step", [1, 2, 3, 7])
def test_simulate_save_step(save_step):
    state, p = _density_case()
    sol = simulate(ex.sir_density_ode, 100, state, p, SolverParams(), save_step=save_step)
    for c in sol.ys:
        assert c.shape == (int(100 / save_step) + 1, 1)
    full = simulate(ex.sir_density_ode, 100, state, p, SolverParams())
    if 100 % save_step == 0:  # grid points coincide with daily ones: same dense-output values
        assert torch.allclose(sol.ys[1][:, 0], full.ys[1][::save_step, 0], rtol=1e-12, atol=1e-12)


def test_simulate_sub_save_indices():
    state, p = _density_case()
    for idx in ((0,), (0, 2), (1,)):
        sol = simulate(ex.sir_density_ode, 100, state, p, SolverParams(), sub_save_indices=idx)
        for k, c in enumerate(sol.ys):
            assert c.shape == ((101, 1) if k in idx else (101, 0))
    full = simulate(ex.sir_density_ode, 100, state, p, SolverParams())
    part = simulate(ex.sir_density_ode, 100, state, p, SolverParams(), sub_save_indices=(2,))
    assert torch.equal(part.ys[2], full.ys[2])


def test_max_steps_raises_like_diffrax_throw():
    state, p = _density_case()
    with pytest.raises(RuntimeError, match="maximum number of solver steps"):
        simulate(ex.sir_density_ode, 100, state, p, SolverParams(max_steps=5))
    sol = simulate_ensemble(ex.sir_density_ode, 100, state, p, SolverParams(max_steps=5), batch_size=3, throw=False)
    assert bool((sol.result == 1).all()) and bool(torch.isinf(sol.ys[0][:, -1]).all())


def test_stats_and_single_draw_against_oracle():
    from oracle import oracle as orc
    state, p = _density_case()
    sol = simulate(ex.sir_density_ode, 100, state, p, SolverParams())
    ref, _, st = orc.solve(orc.SIR_DENSITY, (1, 1, 1), np.array([99.0, 1.0, 0.0]), np.array([0.003, 0.1]), t1=100)
    got = torch.cat(sol.ys, dim=1).cpu().numpy()
    assert np.allclose(got, ref[0], rtol=1e-9, atol=1e-9)
    assert int(sol.stats["num_accepted_steps"]) == st[0, 1] and int(sol.stats["num_rejected_steps"]) == st[0, 2]
    assert int(sol.stats["num_steps"]) == st[0, 3] and int(sol.result) == 0


@pytest.mark.parametrize("s0,i0,r0", [(0.99, 0.01, 0.0), (0.95, 0.05, 0.0), (0.90, 0.10, 0.0), (0.80, 0.20, 0.0)])
def test_final_epidemic_size_matches_theory(s0, i0, r0):
    # reference tests/test_sir_dynamics/test_sir.py:9-65 (r0 = 2, infectious period 7: examples/sir.py:43)
    R0 = 2.0
    p = ex.SIR_ODEParams(beta=t(R0 / 7.0), gamma=t(1 / 7.0))
    sol = simulate(ex.sir_ode, 300, (t([s0]), t([i0]), t([r0])), p, SolverParams())
    root = root_scalar(lambda x: x - s0 * math.exp(-R0 * (1 - x)), bracket=[0.0, s0], method="bisect", xtol=1e-8).root
    assert float(sol.ys[2][-1, 0]) == pytest.approx(1 - root, abs=2e-2)


@pytest.mark.parametrize("s0,i0,r0", [(0.99, 0.01, 0.0), (0.95, 0.05, 0.0), (0.90, 0.10, 0.0), (0.80, 0.20, 0.0),
                                      (0.8, 0.0, 0.2), (0.75, 0.1, 0.15)])
def test_sir_mass_conservation(s0, i0, r0):
    # reference test_sir.py:68-100, including the i0 = 0 (f == 0) initial-step edge
    p = ex.SIR_ODEParams(beta=t(2 / 7.0), gamma=t(1 / 7.0))
    sol = simulate(ex.sir_ode, 120, (t([s0]), t([i0]), t([r0])), p, SolverParams())
    total = sum(c.squeeze() for c in sol.ys)
    assert torch.isfinite(total).all() and torch.allclose(total, total[0], atol=1e-6)


@pytest.mark.parametrize("r_0,inf,lat,wan", [(2.0, 7.0, 3.0, 60.0), (3.0, 5.0, 2.0, 100.0)])
def test_seirs_endemic_equilibrium(r_0, inf, lat, wan):
    # reference tests/test_seirs_dynamics/test_seirs.py:21-65
    beta, gamma, sigma, omega = r_0 / inf, 1 / inf, 1 / lat, 1 / wan
    p = ex.SEIRS_ODEParams(beta=t(beta), gamma=t(gamma), sigma=t(sigma), omega=t(omega))
    sol = simulate(ex.seirs_ode, 1000, (t([0.99]), t([0.0]), t([0.01]), t([0.0])), p, SolverParams())
    s, e, i, r = [c.squeeze() for c in sol.ys]
    for c in (s, e, i, r):
        assert float(c[-100:].std()) < 1e-4
    S = gamma / beta
    I = (1 - S) / (1 + gamma / sigma + gamma / omega)
    assert float(s[-1]) == pytest.approx(S, rel=1e-2) and float(i[-1]) == pytest.approx(I, rel=1e-2)
    assert float(e[-1]) == pytest.approx(gamma / sigma * I, rel=1e-2)
    assert float(r[-1]) == pytest.approx(gamma / omega * I, rel=1e-2)


def test_seasonal_seirs_keeps_oscillating():
    # reference tests/test_seirs_seasonality_dynamics: the last 100 days still move (std > 1e-4)
    p = ex.SeasonalSEIRS_ODEParams(beta=t(2 / 7.0), gamma=t(1 / 7.0), sigma=t(1 / 3.0), omega=t(1 / 60.0),
                                   seasonality_params=ex.SeasonalityParams(forcing_amp=t(0.2), forcing_phase=t(0.0),
                                                                           forcing_period=t(365.0)))
    sol = simulate(ex.seirs_ode_seasonal, 1500, (t([0.99]), t([0.0]), t([0.01]), t([0.0])), p, SolverParams())
    assert float(sol.ys[2][-100:, 0].std()) > 1e-4 and bool(torch.isfinite(sol.ys[2]).all())


def test_age_risk_contact_tensor_layout():
    # reference examples/sir_age_risk_stratified.py: CM[i,j,k,l] couples source (i,j) into target (k,l)
    from oracle import oracle as orc
    from tests.cases import make_case
    case = make_case("sir_age_risk32", 1)
    CM = torch.as_tensor(case["oracle"][3], dtype=torch.float64, device=DEV)
    y0 = t(case["y0"])
    state = (y0[0:6].reshape(3, 2), y0[6:12].reshape(3, 2), y0[12:18].reshape(3, 2))
    p = ex.AgeRiskSIR_ODEParams(beta=t(case["params"]["beta"][0, 0]), gamma=t(case["params"]["gamma"][0, 0]),
                                contact_matrix=CM)
    sol = simulate(ex.sir_age_risk_ode, 150, state, p, SolverParams())
    assert sol.ys[1].shape == (151, 3, 2)
    fam, dims, theta, shared = case["oracle"]
    ref, _, _ = orc.solve(fam, dims, case["y0"], theta, shared, t1=150)
    got = torch.cat([c.reshape(151, -1) for c in sol.ys], dim=1).cpu().numpy()
    assert np.allclose(got, ref[0], rtol=1e-9, atol=1e-9 * np.abs(ref).max())


def test_ensemble_matches_single_draws_and_host_pipeline():
    from tests.cases import make_case
    B = 70
    case = make_case("seirs_seasonal", B)
    prm = {k: t(v) for k, v in case["params"].items()}
    mk = lambda sl: ex.SeasonalSEIRS_ODEParams(
        beta=prm["beta"][sl], gamma=prm["gamma"][sl], sigma=prm["sigma"][sl], omega=prm["omega"][sl],
        seasonality_params=ex.SeasonalityParams(forcing_amp=prm["season_amp"][sl], forcing_phase=prm["season_phase"][sl],
                                                forcing_period=prm["season_period"][sl]))
    y0 = tuple(t([v]) for v in case["y0"])
    ens = simulate_ensemble(ex.seirs_ode_seasonal, 365, y0, mk(slice(None)), SolverParams(), batch_size=B)
    assert ens.ys[0].shape == (B, 366, 1) and ens.stats["num_accepted_steps"].shape == (B,)
    for b in (0, 17, 69):
        one = simulate(ex.seirs_ode_seasonal, 365, y0, mk(b), SolverParams())
        for k in range(4):
            assert torch.equal(one.ys[k], ens.ys[k][b])
    # host-resident inputs stream through the chunked H2D / solve / D2H pipeline and give the same bits
    cpu = lambda x: x.cpu()
    p_h = ex.SeasonalSEIRS_ODEParams(
        beta=cpu(prm["beta"]), gamma=cpu(prm["gamma"]), sigma=cpu(prm["sigma"]), omega=cpu(prm["omega"]),
        seasonality_params=ex.SeasonalityParams(forcing_amp=cpu(prm["season_amp"]), forcing_phase=cpu(prm["season_phase"]),
                                                forcing_period=cpu(prm["season_period"])))
    y0_h = tuple(c.cpu() for c in y0)
    host = simulate_ensemble(ex.seirs_ode_seasonal, 365, y0_h, p_h, SolverParams(), batch_size=B, host_chunk=16)
    assert host.ys[0].device.type == "cpu"
    for k in range(4):
        assert torch.equal(host.ys[k], ens.ys[k].cpu())


def test_discontinuity_points_through_simulate():
    from oracle import oracle as orc
    state, p = _density_case()
    sol = simulate(ex.sir_density_ode, 100, state, p, SolverParams(discontinuity_points=[25.0, 60.0]))
    ref, _, st = orc.solve(orc.SIR_DENSITY, (1, 1, 1), np.array([99.0, 1.0, 0.0]), np.array([0.003, 0.1]), t1=100,
                           jump_ts=[25.0, 60.0])
    assert np.allclose(torch.cat(sol.ys, dim=1).cpu().numpy(), ref[0], rtol=1e-9, atol=1e-9)
    assert int(sol.stats["num_accepted_steps"]) == st[0, 1]


def test_seip_family_through_simulate():
    """The immune-history / waning family behind the same drop-in `simulate` / `simulate_ensemble`."""
    from oracle import oracle as orc
    from tests.cases import make_seip_case
    B = 5
    case = make_seip_case(B, A=3, K=2, W=3)
    A, W, K = 3, 3, 2
    H = 4
    nS, nX = A * H * W, A * H * K
    y0 = t(case["y0"])
    state = (y0[:nS].reshape(A, H, W), y0[nS:nS + nX].reshape(A, H, K), y0[nS + nX:nS + 2 * nX].reshape(A, H, K),
             y0[nS + 2 * nX:].reshape(A, H, K))
    prm = case["params"]
    mk = lambda sl: ex.SEIP_ODEParams(beta=t(prm["beta"][sl]), sigma=t(prm["sigma"][sl]), gamma=t(prm["gamma"][sl]),
                                      omega=t(prm["omega"][sl]), contact_matrix=t(case["contact"]),
                                      population=t(case["pop"]), immunity=t(case["immunity"]))
    sol = simulate(ex.seip_ode, 120, state, mk(0), SolverParams())
    assert sol.ys[0].shape == (121, A, H, W) and sol.ys[3].shape == (121, A, H, K)
    fam, dims, theta, shared = case["oracle"]
    ref, _, st = orc.solve(fam, dims, case["y0"], theta, shared, t1=120)
    got = torch.cat([c.reshape(121, -1) for c in sol.ys], dim=1).cpu().numpy()
    assert np.allclose(got, ref[0], rtol=1e-9, atol=1e-9 * np.abs(ref).max())
    assert int(sol.stats["num_accepted_steps"]) == st[0, 1]
    ens = simulate_ensemble(ex.seip_ode, 120, state, mk(slice(None)), SolverParams(), batch_size=B)
    assert ens.ys[1].shape == (B, 121, A, H, K)
    flat = torch.cat([c.reshape(B, 121, -1) for c in ens.ys], dim=2).cpu().numpy()
    assert np.allclose(flat, ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())
