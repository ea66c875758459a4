#!/usr/bin/env python
"""Write the golden file that pins this repository's solver to the REFERENCE's own diffrax path.

It drives the reference itself -- `dynode.simulation.simulate` (src/dynode/simulation/odes.py:35-145), the RHS
callables and ODEParams dataclasses of `examples/*.py` and `tests/test_simulation/test_odes.py`, the pydantic
`SolverParams` (src/dynode/config/params.py:24-67), and the numpyro model of `examples/sir_infer_parameters.py:21-59`
-- under `jax.vmap` with `jax_enable_x64`, on the seeded inputs of `tests/cases.py`.  Nothing of the solve is restated
here: what `diffrax.diffeqsolve` returns through the reference's call is what gets written.

    # anywhere DynODE's stack is installed (diffrax 0.7.*, jax >=0.6.1,<0.7, numpyro 0.15.*), with this repo checked out:
    python baseline/dump_diffrax_golden.py --reference /path/to/DynODE
        -> tests/golden/diffrax_golden.npz      (commit it: tests/test_oracle.py and tests/test_gpu_parity.py pick it up
                                                 and "parity unpinned" becomes pinned with no code change)

    # in this image (no jax / diffrax / numpyro, no network):
    python baseline/dump_diffrax_golden.py --standin --draws 8
        -> tests/golden/standin_golden.npz      (same keys; `diffeqsolve` is oracle/oracle_np.py -- see
                                                 baseline/standin_stack.py for what that does and does not pin)

Keys (per case `c` of tests.cases.ALL_CASES, B = draws):
  c/ys_full [Bf,T,n]  c/ys_rows [B,R,n] + c/rows [R]   saved states (all days for the first Bf draws, every 13th day
                                                       and the last one for all draws), compartments concatenated
  c/accepted c/rejected c/num_steps [B]                diffrax `stats`
  c/ts [T]                                             `Solution.ts`
for c in OPTION_CASES additionally, `o` in jump | const | step2 | step3 | step7 | sub | tight:
  c/o/ys [Bo,To,n_o]  c/o/ts  c/o/accepted  c/o/rejected  c/o/num_steps   (+ c/sub/empty_shapes for `(T,0)` outputs)
config 2 (sir_infer_parameters.py model):
  c2/obs [100,2]  c2/z [K,2]  c2/potential [K]  c2/grad [K,2]     numpyro potential energy in unconstrained space
  c2/site_names                                                   latent site order of the z columns
  c2/loglik [K] c2/loglik_grad [K,2]                              log p(obs | r0, infectious_period) and d/d(r0, T_inf)
  c2/r0 c2/infectious_period [K]                                  constrained values of the z rows
meta/backend, meta/diffrax, meta/jax, meta/numpyro, meta/draws
"""
import argparse
import os
import sys
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

OPTION_CASES = ("sir_age2", "seirs_seasonal", "seirs_multi_a2s3")
OPTIONS = {  # name -> (SolverParams kwargs, simulate kwargs)
    "jump": (dict(discontinuity_points=[30.0, 61.5]), {}),
    "const": (dict(constant_step_size=0.25), {}),
    "step2": ({}, dict(save_step=2)),
    "step3": ({}, dict(save_step=3)),
    "step7": ({}, dict(save_step=7)),
    "sub": ({}, dict(sub_save_indices="first_last")),
    "tight": (dict(ode_solver_rel_tolerance=1e-8, ode_solver_abs_tolerance=1e-10), {}),
}
ROW_STRIDE = 13
N_FULL = 2
N_OPTION_DRAWS = 4
N_Z = 16


def reference_bridge(jnp):
    """{case name: (ode, make_params(dict of per-draw arrays, shared) -> ODEParams, state shapes)} built from the
    reference's own modules (imported, not restated)."""
    import seirs
    import sir
    import sir_age_risk_stratified as risk
    import sir_age_stratified as age
    import seirs_multi_strain_age_stratified as multi
    from examples import seirs_seasonal_forcing as seas  # uses a relative import of .seirs

    sys.path.insert(0, os.path.join(REF, "tests", "test_simulation"))
    import test_odes as tst

    idx_multi = SimpleNamespace(e=SimpleNamespace(strain=1), i=SimpleNamespace(strain=1), r=SimpleNamespace(strain=1))

    def scalar(p, k):
        return p[k][0]

    def sir_p(p, sh):
        return sir.SIR_ODEParams(beta=scalar(p, "beta"), gamma=scalar(p, "gamma"))

    def dens_p(p, sh):
        return tst.TestingODEParams(beta=scalar(p, "beta"), gamma=scalar(p, "gamma"))

    def seirs_p(p, sh):
        return seirs.SEIRS_ODEParams(beta=scalar(p, "beta"), gamma=scalar(p, "gamma"), sigma=scalar(p, "sigma"),
                                     omega=scalar(p, "omega"))

    def seas_p(p, sh):
        sp = seas.SeasonalityParams(forcing_amp=scalar(p, "season_amp"), forcing_phase=scalar(p, "season_phase"),
                                    forcing_period=scalar(p, "season_period"))
        return seas.SEIRS_ODEParams(beta=scalar(p, "beta"), gamma=scalar(p, "gamma"), sigma=scalar(p, "sigma"),
                                    omega=scalar(p, "omega"), seasonality_params=sp)

    def age_p(p, sh):
        return age.SIR_ODEParams(beta=scalar(p, "beta"), gamma=scalar(p, "gamma"), contact_matrix=jnp.asarray(sh))

    def risk_p(p, sh):
        return risk.SIR_ODEParams(beta=scalar(p, "beta"), gamma=scalar(p, "gamma"), contact_matrix=jnp.asarray(sh))

    def multi_p(p, sh):
        return multi.SEIRS_MultiStrain_ODEParams(beta=p["beta"], gamma=p["gamma"], sigma=p["sigma"], omega=p["omega"],
                                                 contact_matrix=jnp.asarray(sh), idx=idx_multi)

    def multi_shapes(G, S):
        return [(G,)] + [(G, S)] * 4

    return {
        "sir_1bin": (sir.sir_ode, sir_p, [(1,)] * 3),
        "sir_density": (tst.sir_ode, dens_p, [(1,)] * 3),
        "seirs_1bin": (seirs.seirs_ode, seirs_p, [(1,)] * 4),
        "seirs_seasonal": (seas.seirs_ode_seasonal, seas_p, [(1,)] * 4),
        "sir_age2": (age.sir_ode, age_p, [(2,)] * 3),
        "sir_age4": (age.sir_ode, age_p, [(4,)] * 3),
        "sir_age_risk32": (risk.sir_ode, risk_p, [(3, 2)] * 3),
        "seirs_multi_a2s3": (multi.seirs_multi_strain_ode, multi_p, multi_shapes(2, 3)),
        "seirs_multi_g6s3": (multi.seirs_multi_strain_ode, multi_p, multi_shapes(6, 3)),
    }


def flat(ys_tuple):
    """(B, T, *shape) per compartment -> (B, T, n) in `jnp.concatenate([c.ravel() ...])` order; (T,0) outputs vanish."""
    cols = [np.asarray(c).reshape(np.asarray(c).shape[0], np.asarray(c).shape[1], -1) for c in ys_tuple]
    return np.concatenate(cols, axis=2)


def main():
    global REF
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--draws", type=int, default=32)
    ap.add_argument("--reference", default=os.environ.get("DYNODE_REFERENCE", "/root/reference"),
                    help="checkout of CDCgov/DynODE (its examples/ and tests/ are needed even when dynode is pip-installed)")
    ap.add_argument("--standin", action="store_true",
                    help="no jax/diffrax here: drive the reference over baseline/standin_stack.py (plumbing pin only)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    REF = args.reference

    if args.standin:
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        import standin_stack
        standin_stack.install(REF)
        sys.path.insert(0, REF)
    else:
        for sub in ("", "examples", "src"):
            sys.path.insert(0, os.path.join(REF, sub))
    import jax
    jax.config.update("jax_enable_x64", True)
    import jax.numpy as jnp
    import diffrax

    from dynode.config import SolverParams  # the reference's
    from dynode.simulation import simulate  # the reference's

    from tests.cases import ALL_CASES, make_case

    bridge = reference_bridge(jnp)
    out = {}

    def run(name, B, solver_kw, sim_kw):
        """The reference's simulate() vmapped over B draws of tests.cases.make_case(name, B)."""
        case = make_case(name, B)
        ode, make_params, shapes = bridge[name]
        sizes = [int(np.prod(s)) for s in shapes]
        offs = np.concatenate([[0], np.cumsum(sizes)])
        shared = None
        if case["contact"] is not None:
            shared = case["oracle"][3]  # (A,A) contact[target][source]; the age x risk case keeps CM[i,j,k,l]
        sim_kw = dict(sim_kw)
        if sim_kw.get("sub_save_indices") == "first_last":
            sim_kw["sub_save_indices"] = (0, len(shapes) - 1)
        y0 = np.broadcast_to(case["y0"], (B, int(offs[-1])))
        prm = {k: jnp.asarray(v) for k, v in case["params"].items()}
        sp = SolverParams(**solver_kw)

        def solve_one(y0_b, p_b):
            state = tuple(y0_b[offs[i]:offs[i + 1]].reshape(shapes[i]) for i in range(len(shapes)))
            sol = simulate(ode, case["t1"], state, make_params(p_b, shared), sp, **sim_kw)
            st = sol.stats
            return (sol.ys, sol.ts, jnp.asarray(st["num_accepted_steps"]), jnp.asarray(st["num_rejected_steps"]),
                    jnp.asarray(st["num_steps"]))

        ys, ts, acc, rej, steps = jax.jit(jax.vmap(solve_one))(jnp.asarray(y0), prm)
        as_int = lambda a: np.asarray(a).astype(np.int64)
        return ys, np.asarray(ts)[0], as_int(acc), as_int(rej), as_int(steps)

    for name in ALL_CASES:
        ys, ts, acc, rej, steps = run(name, args.draws, {}, {})
        full = flat(ys)
        rows = np.unique(np.concatenate([np.arange(0, full.shape[1], ROW_STRIDE), [full.shape[1] - 1]]))
        out[f"{name}/ys_full"] = full[:N_FULL]
        out[f"{name}/ys_rows"] = full[:, rows]
        out[f"{name}/rows"] = rows
        out[f"{name}/ts"] = ts
        out[f"{name}/accepted"], out[f"{name}/rejected"], out[f"{name}/num_steps"] = acc, rej, steps
        print(f"{name}: ys {full.shape}, accepted {acc.min()}..{acc.max()}, rejected {rej.min()}..{rej.max()}")
        if name in OPTION_CASES:
            for opt, (solver_kw, sim_kw) in OPTIONS.items():
                ys, ts, acc, rej, steps = run(name, N_OPTION_DRAWS, solver_kw, sim_kw)
                k = f"{name}/{opt}"
                out[f"{k}/ys"], out[f"{k}/ts"] = flat(ys), ts
                out[f"{k}/accepted"], out[f"{k}/rejected"], out[f"{k}/num_steps"] = acc, rej, steps
                if opt == "sub":
                    out[f"{k}/empty_shapes"] = np.array([np.asarray(c).shape[1:] for c in ys
                                                         if np.asarray(c).size == 0])
                print(f"  {opt}: ys {out[f'{k}/ys'].shape}, accepted {acc.min()}..{acc.max()}")

    numpyro_version = "absent"
    if not args.standin:
        numpyro_version = dump_config2(out, jax, jnp)
    out["meta/backend"] = np.array("standin:oracle_np" if args.standin else "diffrax")
    out["meta/diffrax"] = np.array(diffrax.__version__)
    out["meta/jax"] = np.array(jax.__version__)
    out["meta/numpyro"] = np.array(numpyro_version)
    out["meta/draws"] = np.array(args.draws)
    path = args.out or os.path.join(ROOT, "tests", "golden",
                                    "standin_golden.npz" if args.standin else "diffrax_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def dump_config2(out, jax, jnp):
    """numpyro's potential energy (priors + bijector Jacobians + Poisson likelihood) of the reference's config-2 model
    and its gradient, at N_Z points of unconstrained space: what NUTS evaluates per leapfrog
    (reference src/dynode/infer/inference.py:149-163 -> examples/sir_infer_parameters.py:21-59)."""
    import numpyro
    from numpyro.infer.util import constrain_fn, potential_energy

    import sir_infer_parameters as c2  # the reference's example module
    from sir_age_stratified import get_config as get_static_config
    from sir_age_stratified import run_simulation

    sol = run_simulation(get_static_config(), tf=100)
    idx = get_static_config().idx
    obs = jnp.diff(sol.ys[idx.r], axis=0)
    cfg = c2.get_config()
    margs, mkw = (), dict(config=cfg, tf=100, obs_data=obs)
    names = ["strains_0_r0", "strains_0_infectious_period"]  # tests/test_infer/test_sample.py:49-71 naming
    rng = np.random.Generator(np.random.PCG64(20260102))
    z = rng.normal(0.0, 1.0, size=(N_Z, 2))

    def pot(zrow):
        return potential_energy(c2.model, margs, mkw, {names[0]: zrow[0], names[1]: zrow[1]})

    val, grad = jax.jit(jax.vmap(jax.value_and_grad(pot)))(jnp.asarray(z))
    cons = jax.vmap(lambda zr: constrain_fn(c2.model, margs, mkw, {names[0]: zr[0], names[1]: zr[1]}))(jnp.asarray(z))

    def loglik(theta):  # log p(obs | r0, T_inf) through the reference's own simulate + RHS
        from dynode.simulation import simulate
        from sir_age_stratified import SIR_ODEParams, sir_ode
        tp = get_static_config().parameters.transmission_params
        p = SIR_ODEParams(beta=theta[0] / theta[1], gamma=1.0 / theta[1], contact_matrix=tp.contact_matrix)
        y0 = get_static_config().initializer.get_initial_state()
        s = simulate(sir_ode, 100, y0, p, get_static_config().parameters.solver_params)
        inc = jnp.maximum(jnp.diff(s.ys[idx.r], axis=0), 1e-6)
        return jnp.sum(numpyro.distributions.Poisson(inc).log_prob(obs))

    theta = jnp.stack([cons[names[0]], cons[names[1]]], axis=1)
    ll, llg = jax.jit(jax.vmap(jax.value_and_grad(loglik)))(theta)
    out["c2/obs"], out["c2/z"] = np.asarray(obs), z
    out["c2/potential"], out["c2/grad"] = np.asarray(val), np.asarray(grad)
    out["c2/site_names"] = np.array(names)
    out["c2/r0"], out["c2/infectious_period"] = np.asarray(cons[names[0]]), np.asarray(cons[names[1]])
    out["c2/loglik"], out["c2/loglik_grad"] = np.asarray(ll), np.asarray(llg)
    print(f"config 2: potential {np.asarray(val).min():.3f}..{np.asarray(val).max():.3f}")
    return numpyro.__version__


if __name__ == "__main__":
    main()
