#!/usr/bin/env python
"""Write tests/golden/diffrax_golden.npz from the REAL reference stack (diffrax 0.7 + jax x64).

This image has neither jax nor diffrax (no network), so the solver arithmetic of the oracle is a restatement
("parity unpinned").  Run this script anywhere DynODE's own stack is installed
(`pip install diffrax==0.7.* "jax>=0.6.1,<0.7"`), commit the resulting .npz, and tests/test_oracle.py
::test_oracle_matches_diffrax_golden turns the restatement into a pinned oracle: saved states, accepted /
rejected step counts and (for the NUTS configuration) the log-density gradient are compared to what
`diffrax.diffeqsolve` returns for the exact call DynODE makes (reference src/dynode/simulation/odes.py:107-144).

    python baseline/dump_diffrax_golden.py [--draws 8]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--draws", type=int, default=8)
    args = ap.parse_args()

    import jax
    jax.config.update("jax_enable_x64", True)
    import jax.numpy as jnp
    from diffrax import ClipStepSizeController, ODETerm, PIDController, SaveAt, Tsit5, diffeqsolve

    from tests.cases import ALL_CASES, make_case

    def rhs_for(name, case):
        """The reference right-hand sides (examples/*.py) on a flat state vector."""
        m = case["model"]
        G, S = m.n_groups, m.n_strains
        K = None if case["contact"] is None else jnp.asarray(case["contact"])  # contact[target][source]

        def f(t, y, p):
            if name in ("sir_1bin", "sir_density", "sir_age2", "sir_age4", "sir_age_risk32"):
                s, i, r = y[:G], y[G:2 * G], y[2 * G:3 * G]
                beta, gamma = p["beta"][0], p["gamma"][0]
                if name == "sir_density":
                    new = beta * s * i
                elif K is None:
                    new = beta * s * i / (s + i + r)
                else:
                    new = s * beta * (K @ (i / (s + i + r)))
                return jnp.concatenate([-new, new - gamma * i, gamma * i])
            if name in ("seirs_1bin", "seirs_seasonal"):
                s, e, i, r = y
                beta = p["beta"][0]
                if name == "seirs_seasonal":
                    beta = beta * (1 + p["season_amp"][0] * jnp.sin(2 * jnp.pi * t / p["season_period"][0]
                                                                    + p["season_phase"][0]))
                N = s + e + i + r
                return jnp.stack([-beta * s * i / N + p["omega"][0] * r, beta * s * i / N - p["sigma"][0] * e,
                                  p["sigma"][0] * e - p["gamma"][0] * i, p["gamma"][0] * i - p["omega"][0] * r])
            s = y[:G]
            e, i, r, c = (y[G + k * G * S:G + (k + 1) * G * S].reshape(G, S) for k in range(4))
            N = s + e.sum(1) + i.sum(1) + r.sum(1)
            foi = p["beta"] * (K @ (i / N[:, None]))
            new = foi * s[:, None]
            ds = -new.sum(1) + (p["omega"] * r).sum(1)
            return jnp.concatenate([ds, (new - p["sigma"] * e).ravel(), (p["sigma"] * e - p["gamma"] * i).ravel(),
                                    (p["gamma"] * i - p["omega"] * r).ravel(), new.ravel()])
        return f

    out = {}
    for name in ALL_CASES:
        case = make_case(name, args.draws)
        f = rhs_for(name, case)
        t1 = float(case["t1"])
        ts = jnp.linspace(0.0, t1, int(t1 // 1) + 1)
        y0 = np.broadcast_to(case["y0"], (args.draws, case["model"].state_size))

        def solve_one(y0_b, p_b):
            sol = diffeqsolve(ODETerm(f), Tsit5(), 0.0, t1, None, y0_b, args=p_b,
                              stepsize_controller=ClipStepSizeController(PIDController(rtol=1e-5, atol=1e-6), jump_ts=None),
                              saveat=SaveAt(ts=ts), max_steps=int(1e6))
            return sol.ys, sol.stats["num_accepted_steps"], sol.stats["num_rejected_steps"]

        prm = {k: jnp.asarray(v) for k, v in case["params"].items()}
        ys, acc, rej = jax.jit(jax.vmap(solve_one))(jnp.asarray(y0), prm)
        out[f"{name}/ys"] = np.asarray(ys)
        out[f"{name}/accepted"] = np.asarray(acc)
        out[f"{name}/rejected"] = np.asarray(rej)
        if name == "sir_age2":  # the NUTS configuration: d/d(beta, gamma) of sum_t w_t . R(t)
            w = jnp.linspace(0.5, 1.5, ys.shape[1] * 2).reshape(ys.shape[1], 2)

            def loss(p_b, y0_b):
                return jnp.sum(w * solve_one(y0_b, p_b)[0][:, 4:6])

            g = jax.jit(jax.vmap(jax.grad(loss), in_axes=(0, 0)))(prm, jnp.asarray(y0))
            out[f"{name}/grad_beta"] = np.asarray(g["beta"])
            out[f"{name}/grad_gamma"] = np.asarray(g["gamma"])
            out[f"{name}/grad_weights"] = np.asarray(w)
        print(f"{name}: ys {ys.shape}, accepted {np.asarray(acc).min()}..{np.asarray(acc).max()}")
    import diffrax
    out["meta/diffrax"] = np.array(diffrax.__version__)
    out["meta/jax"] = np.array(jax.__version__)
    out["meta/draws"] = np.array(args.draws)
    path = os.path.join(ROOT, "tests", "golden", "diffrax_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
