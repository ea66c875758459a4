"""Generate tests/golden/rhs_golden.npz by executing the REFERENCE's own right-hand-side functions.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_rhs_golden.py

The reference package itself cannot be imported here (jax, diffrax, chex, numpyro are absent), but
its RHS callables are pure array functions.  This script loads the reference's example modules
*from where they lie* under /root/reference with thin stand-ins for the missing third-party
modules (jax.numpy -> numpy, chex.dataclass -> dataclasses.dataclass, jax.jit -> identity), then
evaluates each RHS on seeded random states/parameters and, using oracle/oracle_np.py (a numpy
restatement of diffrax's Tsit5 loop - PARITY UNPINNED for that part), integrates the reference's
RHS to produce golden trajectories.  No reference source is copied into this repo.

Golden content per case `<name>`:
  <name>/t, y, theta, shared, dy      RHS evaluations  (pinned to the reference's own code)
  <name>/traj_*                       trajectories: reference RHS + numpy-restated solver
"""

from __future__ import annotations

import dataclasses
import importlib.util
import os
import sys
import types
from types import SimpleNamespace

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


# --------------------------------------------------------------------------- stand-in modules
class _AtIndexer:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Setter:
            def set(self, v):
                out = np.array(arr, copy=True).view(_JArray)
                out[idx] = v
                return out

        return _Setter()


class _JArray(np.ndarray):
    """numpy array with jax's functional `.at[idx].set(v)`."""

    @property
    def at(self):
        return _AtIndexer(self)


def _install_shims():
    jnp = types.ModuleType("jax.numpy")
    for name in dir(np):
        if not name.startswith("_"):
            setattr(jnp, name, getattr(np, name))
    jnp.array = lambda x, dtype=None: np.array(x, dtype=dtype or np.float64).view(_JArray)
    jnp.zeros_like = lambda x: np.zeros_like(np.asarray(x)).view(_JArray)
    jnp.zeros = lambda s, dtype=None: np.zeros(s).view(_JArray)
    jax = types.ModuleType("jax")
    jax.numpy = jnp
    jax.jit = lambda f=None, **kw: f if f is not None else (lambda g: g)
    jax.Array = np.ndarray
    chex = types.ModuleType("chex")

    def _dc(cls=None, **kw):
        def wrap(c):
            return dataclasses.dataclass(c)
        return wrap if cls is None else wrap(cls)

    chex.dataclass = _dc
    chex.ArrayDevice = np.ndarray

    def _dummy_module(name, attrs):
        m = types.ModuleType(name)
        for a in attrs:
            setattr(m, a, type(a, (), {"__init__": lambda self, *a, **k: None}))
        return m

    cfg_names = ["Bin", "AgeBin", "Compartment", "Dimension", "Initializer", "Params",
                 "SimulationConfig", "SolverParams", "Strain", "TransmissionParams"]
    dynode = _dummy_module("dynode", cfg_names + ["MCMCProcess", "SVIProcess"])
    dynode.config = _dummy_module("dynode.config", cfg_names)
    dynode.config.__path__ = []
    bins = _dummy_module("dynode.config.bins", ["AgeBin", "Bin"])
    dynode.config.bins = bins
    dynode.simulation = _dummy_module("dynode.simulation", ["AbstractODEParams"])
    dynode.simulation.simulate = lambda *a, **k: None
    dynode.typing = _dummy_module("dynode.typing", [])
    dynode.typing.CompartmentState = tuple
    dynode.typing.CompartmentGradients = tuple
    dynode.utils = _dummy_module("dynode.utils", [])
    dynode.utils.vectorize_objects = lambda objs, target, filter=None: [getattr(o, target) for o in objs]
    dynode.infer = _dummy_module("dynode.infer", [])
    dynode.infer.sample_then_resolve = lambda x, **k: x
    diffrax = _dummy_module("diffrax", ["Solution"])
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    mods = {"jax": jax, "jax.numpy": jnp, "chex": chex, "dynode": dynode,
            "dynode.config": dynode.config, "dynode.config.bins": bins, "dynode.simulation": dynode.simulation,
            "dynode.typing": dynode.typing, "dynode.utils": dynode.utils,
            "dynode.infer": dynode.infer, "diffrax": diffrax, "matplotlib": mpl,
            "matplotlib.pyplot": plt}
    sys.modules.update(mods)


def _load(relpath, modname):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, relpath))
    m = importlib.util.module_from_spec(spec)
    sys.modules[modname] = m
    spec.loader.exec_module(m)
    return m


def load_reference_rhs():
    """Returns {family_name: (ode_callable, params_class)} from the reference sources."""
    _install_shims()
    pkg = types.ModuleType("examples")
    pkg.__path__ = [os.path.join(REF, "examples")]
    sys.modules["examples"] = pkg
    sir = _load("examples/sir.py", "examples.sir")
    seirs = _load("examples/seirs.py", "examples.seirs")
    seas = _load("examples/seirs_seasonal_forcing.py", "examples.seirs_seasonal_forcing")
    age = _load("examples/sir_age_stratified.py", "examples.sir_age_stratified")
    risk = _load("examples/sir_age_risk_stratified.py", "examples.sir_age_risk_stratified")
    multi = _load("examples/seirs_multi_strain_age_stratified.py", "examples.seirs_multi_strain")
    tst = _load("tests/test_simulation/test_odes.py", "ref_test_odes")
    return dict(sir=sir, seirs=seirs, seas=seas, age=age, risk=risk, multi=multi, tst=tst)


def J(x):
    return np.asarray(x, dtype=np.float64).view(_JArray)


# --------------------------------------------------------------------------- case builders
# Each returns (ode, args, unpack(y)->tuple, theta_flat, shared_flat, dims(A,R,S), family id)
def make_cases(m):
    from oracle import oracle as orc

    idx_multi = SimpleNamespace(e=SimpleNamespace(strain=1), i=SimpleNamespace(strain=1),
                                r=SimpleNamespace(strain=1))

    def sir_args(th, sh):
        return m["sir"].SIR_ODEParams(beta=th[0], gamma=th[1])

    def dens_args(th, sh):
        return m["tst"].TestingODEParams(beta=th[0], gamma=th[1])

    def seirs_args(th, sh):
        return m["seirs"].SEIRS_ODEParams(beta=th[0], gamma=th[1], sigma=th[2], omega=th[3])

    def seas_args(th, sh):
        sp = m["seas"].SeasonalityParams(forcing_amp=th[4], forcing_phase=th[5], forcing_period=th[6])
        return m["seas"].SEIRS_ODEParams(beta=th[0], gamma=th[1], sigma=th[2], omega=th[3],
                                         seasonality_params=sp)

    def age_args(A):
        def f(th, sh):
            return m["age"].SIR_ODEParams(beta=J(th[0]), gamma=J(th[1]), contact_matrix=J(sh).reshape(A, A))
        return f

    def risk_args(A, R):
        def f(th, sh):
            return m["risk"].SIR_ODEParams(beta=J(th[0]), gamma=J(th[1]),
                                           contact_matrix=J(sh).reshape(A, R, A, R))
        return f

    def multi_args(A, S):
        def f(th, sh):
            return m["multi"].SEIRS_MultiStrain_ODEParams(
                beta=J(th[0:S]), gamma=J(th[S:2 * S]), sigma=J(th[2 * S:3 * S]),
                omega=J(th[3 * S:4 * S]), contact_matrix=J(sh).reshape(A, A), idx=idx_multi)
        return f

    def shapes_for(fam, A, R, S):
        if fam in (orc.SIR_1BIN, orc.SIR_DENSITY):
            return [(1,)] * 3
        if fam in (orc.SEIRS_1BIN, orc.SEIRS_SEASONAL):
            return [(1,)] * 4
        if fam == orc.SIR_AGE:
            return [(A,)] * 3
        if fam == orc.SIR_AGE_RISK:
            return [(A, R)] * 3
        return [(A,)] + [(A, S)] * 4

    cases = {
        "sir_1bin": (m["sir"].sir_ode, sir_args, orc.SIR_1BIN, (1, 1, 1)),
        "sir_density": (m["tst"].sir_ode, dens_args, orc.SIR_DENSITY, (1, 1, 1)),
        "seirs_1bin": (m["seirs"].seirs_ode, seirs_args, orc.SEIRS_1BIN, (1, 1, 1)),
        "seirs_seasonal": (m["seas"].seirs_ode_seasonal, seas_args, orc.SEIRS_SEASONAL, (1, 1, 1)),
        "sir_age2": (m["age"].sir_ode, age_args(2), orc.SIR_AGE, (2, 1, 1)),
        "sir_age4": (m["age"].sir_ode, age_args(4), orc.SIR_AGE, (4, 1, 1)),
        "sir_age_risk32": (m["risk"].sir_ode, risk_args(3, 2), orc.SIR_AGE_RISK, (3, 2, 1)),
        "seirs_multi_a2s3": (m["multi"].seirs_multi_strain_ode, multi_args(2, 3),
                             orc.SEIRS_MULTISTRAIN, (2, 1, 3)),
    }
    return cases, shapes_for


def random_theta(rng, fam, dims, orc):
    A, R, S = dims
    r0 = rng.uniform(1.2, 4.0, size=max(S, 1))
    inf = rng.uniform(3.0, 10.0, size=max(S, 1))
    lat = rng.uniform(1.5, 5.0, size=max(S, 1))
    wan = rng.uniform(40.0, 200.0, size=max(S, 1))
    beta, gamma, sigma, omega = r0 / inf, 1 / inf, 1 / lat, 1 / wan
    if fam in (orc.SIR_1BIN, orc.SIR_DENSITY, orc.SIR_AGE, orc.SIR_AGE_RISK):
        return np.array([beta[0], gamma[0]])
    if fam == orc.SEIRS_1BIN:
        return np.array([beta[0], gamma[0], sigma[0], omega[0]])
    if fam == orc.SEIRS_SEASONAL:
        return np.array([beta[0], gamma[0], sigma[0], omega[0], rng.uniform(0, 0.4),
                         rng.uniform(0, 2 * np.pi), 365.0])
    return np.concatenate([beta, gamma, sigma, omega])


def random_shared(rng, fam, dims, orc):
    A, R, S = dims
    if fam == orc.SIR_AGE or fam == orc.SEIRS_MULTISTRAIN:
        return rng.uniform(0.1, 1.0, size=(A, A)).ravel()
    if fam == orc.SIR_AGE_RISK:
        ca = rng.uniform(0.1, 1.0, size=(A, A))
        cr = rng.uniform(0.1, 1.0, size=(R, R))
        return np.einsum("ij,kl->ikjl", ca, cr).ravel()
    return np.zeros(1)


def main():
    from oracle import oracle as orc
    from oracle import oracle_np

    m = load_reference_rhs()
    cases, shapes_for = make_cases(m)
    rng = np.random.Generator(np.random.PCG64(20260101))
    out = {}
    for name, (ode, mkargs, fam, dims) in cases.items():
        shapes = shapes_for(fam, *dims)
        sizes = [int(np.prod(s)) for s in shapes]
        n = sum(sizes)
        offs = np.concatenate([[0], np.cumsum(sizes)])

        def unpack(v, shapes=shapes, offs=offs):
            return tuple(J(v[offs[i]:offs[i + 1]]).reshape(shapes[i]) for i in range(len(shapes)))

        K = 16
        ts = rng.uniform(0, 365, size=K)
        ys = rng.uniform(0.0, 100.0, size=(K, n))
        ths = np.stack([random_theta(rng, fam, dims, orc) for _ in range(K)])
        shs = np.stack([random_shared(rng, fam, dims, orc) for _ in range(K)])
        dys = np.empty((K, n))
        for k in range(K):
            res = ode(float(ts[k]), unpack(ys[k]), mkargs(ths[k], shs[k]))
            dys[k] = np.concatenate([np.asarray(c, dtype=np.float64).ravel() for c in res])
        out[f"{name}/family"] = np.array(fam)
        out[f"{name}/dims"] = np.array(dims)
        out[f"{name}/t"], out[f"{name}/y"], out[f"{name}/theta"] = ts, ys, ths
        out[f"{name}/shared"], out[f"{name}/dy"] = shs, dys

        # trajectories: the reference's RHS callable integrated by the numpy-restated solver
        ntraj = 3
        tr_y0, tr_th, tr_sh, tr_ys, tr_stats = [], [], [], [], []
        for k in range(ntraj):
            th = random_theta(rng, fam, dims, orc)
            sh = random_shared(rng, fam, dims, orc)
            if fam == orc.SIR_AGE or fam == orc.SEIRS_MULTISTRAIN:
                A = dims[0]
                c = np.array([[0.7, 0.3], [0.3, 0.7]]) if A == 2 else sh.reshape(A, A) / sh.reshape(A, A).sum(1).max()
                sh = c.ravel()
            y0 = np.zeros(n)
            pop = 1000.0 if n > 4 else 1.0
            w = rng.dirichlet(np.ones(sizes[0]))
            y0[: sizes[0]] = pop * 0.99 * w
            ioff = offs[2] if len(shapes) >= 4 else offs[1]
            isz = sizes[2] if len(shapes) >= 4 else sizes[1]
            y0[ioff: ioff + isz] = pop * 0.01 * rng.dirichlet(np.ones(isz))
            ysol, stats = oracle_np.solve(ode, unpack(y0), mkargs(th, sh), 120)
            tr_y0.append(y0), tr_th.append(th), tr_sh.append(sh)
            tr_ys.append(np.concatenate([c.reshape(c.shape[0], -1) for c in ysol], axis=1))
            tr_stats.append([stats["result"], stats["num_accepted_steps"],
                             stats["num_rejected_steps"], stats["num_steps"]])
        out[f"{name}/traj_y0"], out[f"{name}/traj_theta"] = np.array(tr_y0), np.array(tr_th)
        out[f"{name}/traj_shared"], out[f"{name}/traj_ys"] = np.array(tr_sh), np.array(tr_ys)
        out[f"{name}/traj_stats"] = np.array(tr_stats)
        print(name, "n=", n, "stats", tr_stats)
    path = os.path.join(HERE, "rhs_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
