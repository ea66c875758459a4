"""Host-side contract of `simulate` (reference src/dynode/simulation/odes.py:35-198) that needs no GPU:
validation order and messages, the save grid, flow registration, and that the product path refuses to
run without the CUDA engine (no CPU fallback)."""
from dataclasses import dataclass
from typing import Any

import numpy as np
import pytest
import torch

from dynode_b200 import DynodeError
from dynode_b200.config import SolverParams
from dynode_b200.examples import rhs as ex
from dynode_b200.flows import UnsupportedODEError, flow_family, flow_spec_of
from dynode_b200.simulation import AbstractODEParams, build_saveat, simulate, simulate_ensemble

Y0 = (torch.tensor([99.0]), torch.tensor([1.0]), torch.tensor([0.0]))
P = ex.DensitySIR_ODEParams(beta=0.003, gamma=0.1)


@pytest.mark.parametrize("stop,step,count", [(100, 1, 101), (100, 2, 51), (100, 3, 34), (100, 7, 15),
                                             (300.0, 1, 301), (150, 0, 151), (150, -4, 151)])
def test_build_saveat_is_the_reference_linspace(stop, step, count):
    # reference tests/test_simulation/test_odes.py:77-92 pins the count; odes.py:177-179 the rule
    sa = build_saveat(0.0, stop, step)
    assert len(sa.times) == count and sa.times[0] == 0.0 and sa.times[-1] == stop
    assert np.array_equal(sa.times, np.linspace(0.0, stop, count))
    assert sa.indices is None
    sub = build_saveat(0.0, stop, step, (0, 2))
    assert sub.indices == (0, 2) and np.array_equal(sub.times, sa.times)


def test_numpy_compartments_raise_type_error_first():
    # odes.py:93-98
    with pytest.raises(TypeError):
        simulate(ex.sir_density_ode, 100, (np.array([99.0]), np.array([1.0]), np.array([0.0])), P, SolverParams())


def test_wrong_parameter_type_raises_assertion():
    # odes.py:100-106
    with pytest.raises(AssertionError):
        simulate(ex.sir_density_ode, 100, Y0, ex.SIR_ODEParams(beta=0.3, gamma=0.1), SolverParams())
    with pytest.raises(AssertionError):
        simulate(ex.sir_density_ode, "100", Y0, P, SolverParams())


def test_unregistered_ode_fails_loudly():
    @dataclass
    class MyParams(AbstractODEParams):
        k: Any

    def my_ode(t: float, state, p: MyParams):
        return tuple(-p.k * c for c in state)

    with pytest.raises(UnsupportedODEError, match="no CPU fallback"):
        simulate(my_ode, 10, Y0, MyParams(k=1.0), SolverParams())
    with pytest.raises(UnsupportedODEError):
        flow_spec_of(my_ode)


def test_model_outside_compiled_instances_fails_loudly():
    # 5 age groups x 7 strains is not in csrc/instances.def
    p = ex.SEIRS_MultiStrain_ODEParams(beta=torch.ones(7), gamma=torch.ones(7), sigma=torch.ones(7),
                                       omega=torch.ones(7), contact_matrix=torch.eye(5))
    state = (torch.ones(5),) + tuple(torch.zeros(5, 7) for _ in range(4))
    with pytest.raises(UnsupportedODEError, match="not compiled in"):
        simulate(ex.seirs_multi_strain_ode, 10, state, p, SolverParams())
    with pytest.raises(UnsupportedODEError):  # wrong number of compartments for the flow
        simulate(ex.seirs_multi_strain_ode, 10, state[:3], p, SolverParams())


def test_unsupported_solver_options_fail_loudly():
    from dynode_b200.config.params import AbstractSolver

    class Dopri5(AbstractSolver):
        pass

    with pytest.raises(UnsupportedODEError, match="Dopri5"):
        simulate(ex.sir_density_ode, 10, Y0, P, SolverParams(solver_method=Dopri5()))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour of a box without a GPU")
def test_no_cpu_fallback_without_cuda():
    with pytest.raises(DynodeError, match="no CPU fallback"):
        simulate(ex.sir_density_ode, 10, Y0, P, SolverParams())
    with pytest.raises(DynodeError, match="no CPU fallback"):
        simulate_ensemble(ex.sir_density_ode, 10, Y0, P, SolverParams(), batch_size=4)


def test_flow_registration_records_the_parameter_map():
    spec = flow_spec_of(ex.seirs_ode_seasonal)
    assert spec.flow == "seirs" and spec.seasonal and not spec.density_dependent
    assert spec.fields["season_amp"] == "seasonality_params.forcing_amp"
    assert spec.compartments == ("s", "e", "i", "r")
    assert flow_spec_of(ex.sir_age_risk_ode).contact_layout == "source_target"
    assert flow_spec_of(ex.seirs_multi_strain_ode).compartments == ("s", "e", "i", "r", "c")
    with pytest.raises(ValueError):
        flow_family("seir")
    with pytest.raises(ValueError):
        flow_family("sir", contact_layout="diag")


def test_product_package_never_imports_the_oracle():
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dynode_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "libdynode_oracle" not in txt, f


def test_constant_step_mode_ignores_discontinuity_points():
    # reference odes.py:113-131: ConstantStepSize() is built WITHOUT a ClipStepSizeController, so the list is unused
    from dynode_b200.simulation.odes import _solver_options
    o = _solver_options(SolverParams(constant_step_size=0.5, discontinuity_points=[10.0, 20.0]), 100)
    assert o.jump_ts == () and o.const_dt == 0.5
    o = _solver_options(SolverParams(discontinuity_points=[10.0, 20.0]), 100)
    assert o.jump_ts == (10.0, 20.0) and o.const_dt == 0.0


def test_sub_save_indices_outside_the_compartments_save_nothing():
    # reference odes.py:185-190: `y[i] if i in sub_save_indices ... for i in range(len(y))` -- negative or too-large
    # indices never match (they do NOT wrap around)
    from dynode_b200.simulation.odes import _mask_from
    assert _mask_from(None, 3) == 0b111
    assert _mask_from((0, 2), 3) == 0b101
    assert _mask_from((-1,), 3) == 0
    assert _mask_from((1, 7, -3), 3) == 0b010


def test_observation_cache_sees_in_place_edits_and_devices():
    from dynode_b200.simulation.odes import _obs_key
    obs = torch.arange(6.0, dtype=torch.float64)
    k0 = _obs_key(obs, "cuda:0")
    assert _obs_key(obs, "cuda:0") == k0 and _obs_key(obs, "cuda:1") != k0
    obs[2] = 9.0  # in-place edit bumps the version counter
    assert _obs_key(obs, "cuda:0") != k0
    a = np.arange(6.0)
    ka = _obs_key(a, "cuda:0")
    a[1] = 5.0
    assert _obs_key(a, "cuda:0") != ka


def test_registered_bodies_are_checked_against_the_compiled_flow():
    """`simulate` integrates the compiled flow, never the Python body: a registered function whose body computes
    something else (an edited copy of the SIR example, a function registered under the wrong flow) must fail loudly
    instead of silently getting the compiled equations.  The shipped examples all pass the same check."""
    from dynode_b200.flows import verify_flow_body

    @dataclass
    class MyParams(AbstractODEParams):
        beta: Any
        gamma: Any

    @flow_family("sir")
    def edited_sir(t: float, state, p: MyParams):
        s, i, r = state
        N = s + i + r
        new = p.beta * s * i / N * 0.5  # a user "improves" the force of infection
        return (-new, new - p.gamma * i, p.gamma * i)

    y0 = (torch.tensor([0.9]), torch.tensor([0.1]), torch.tensor([0.0]))
    with pytest.raises(UnsupportedODEError, match="does not compute the flow 'sir'"):
        simulate(edited_sir, 10, y0, MyParams(beta=0.3, gamma=0.1), SolverParams())

    @flow_family("sir")  # frequency-dependent registration, density-dependent body
    def wrong_flag(t: float, state, p: MyParams):
        s, i, r = state
        return (-p.beta * s * i, p.beta * s * i - p.gamma * i, p.gamma * i)

    with pytest.raises(UnsupportedODEError, match="does not compute"):
        simulate(wrong_flag, 10, y0, MyParams(beta=0.3, gamma=0.1), SolverParams())

    @flow_family("sir")
    def crashes(t: float, state, p: MyParams):
        raise ZeroDivisionError("boom")

    with pytest.raises(UnsupportedODEError, match="could not be evaluated"):
        simulate(crashes, 10, y0, MyParams(beta=0.3, gamma=0.1), SolverParams())

    # every shipped example body equals the flow it is registered as
    G, S = 2, 3
    one = lambda *sh: torch.ones(sh, dtype=torch.float64)
    cases = [
        (ex.sir_ode, ex.SIR_ODEParams(beta=0.3, gamma=0.1), [(1,)] * 3, 1, 1, None),
        (ex.sir_density_ode, ex.DensitySIR_ODEParams(beta=0.3, gamma=0.1), [(1,)] * 3, 1, 1, None),
        (ex.seirs_ode, ex.SEIRS_ODEParams(beta=0.3, gamma=0.1, sigma=0.2, omega=0.01), [(1,)] * 4, 1, 1, None),
        (ex.seirs_ode_seasonal,
         ex.SeasonalSEIRS_ODEParams(beta=0.3, gamma=0.1, sigma=0.2, omega=0.01,
                                    seasonality_params=ex.SeasonalityParams(forcing_amp=0.1, forcing_phase=0.0,
                                                                            forcing_period=365.0)),
         [(1,)] * 4, 1, 1, None),
        (ex.sir_age_ode, ex.AgeSIR_ODEParams(beta=0.3, gamma=0.1, contact_matrix=one(4, 4)), [(4,)] * 3, 4, 1, (4, 4)),
        (ex.sir_age_risk_ode, ex.AgeRiskSIR_ODEParams(beta=0.3, gamma=0.1, contact_matrix=one(3, 2, 3, 2)),
         [(3, 2)] * 3, 6, 1, (3, 2, 3, 2)),
        (ex.seirs_multi_strain_ode,
         ex.SEIRS_MultiStrain_ODEParams(beta=one(S), gamma=one(S), sigma=one(S), omega=one(S), contact_matrix=one(G, G)),
         [(G,)] + [(G, S)] * 4, G, S, (G, G)),
    ]
    for ode, prm, shapes, g, s_, cshape in cases:
        verify_flow_body(ode, flow_spec_of(ode), shapes, prm, g, s_, cshape)


def test_differentiating_the_immune_history_flow_fails_with_a_clear_message():
    A, K, W, H = 2, 2, 2, 4
    one = lambda *sh: torch.ones(sh, dtype=torch.float64)
    beta = torch.tensor([0.3, 0.4], dtype=torch.float64, requires_grad=True)
    p = ex.SEIP_ODEParams(beta=beta, sigma=one(K), gamma=one(K), omega=one(W), contact_matrix=torch.eye(A, dtype=torch.float64),
                          population=8 * one(A), immunity=torch.zeros(H, W, K, dtype=torch.float64))
    e = torch.zeros(A, H, K, dtype=torch.float64)
    with pytest.raises(UnsupportedODEError, match="no sensitivities"):
        simulate(ex.seip_ode, 10, (one(A, H, W), e, e.clone(), e.clone()), p, SolverParams())
