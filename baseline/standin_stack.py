"""Stand-ins for the third-party stack DynODE imports (jax, diffrax, chex, numpyro, plotting libraries) so that the
REFERENCE's own `dynode` package and `examples/*.py` can be imported -- from where they lie, unmodified -- in an image
that has none of them.  TEST / FIXTURE INFRASTRUCTURE ONLY: nothing under `dynode_b200/` imports this file.

What is real and what is not when `baseline/dump_diffrax_golden.py --standin` runs on top of this:

  real   the reference's `dynode.simulation.simulate` / `build_saveat` (src/dynode/simulation/odes.py:35-198): input
         checks, controller choice, `jump_ts`, the `linspace` save grid, `SubSaveAt` for unsaved compartments;
         the reference's pydantic `SolverParams` (src/dynode/config/params.py:24-67); the reference's RHS callables
         and ODEParams dataclasses (examples/*.py, tests/test_simulation/test_odes.py)
  fake   `diffrax.diffeqsolve` = oracle/oracle_np.py (the numpy RESTATEMENT of diffrax 0.7 -- "parity unpinned"),
         `jax.numpy` = numpy, `jax.vmap` = a Python loop, `chex.dataclass` = dataclasses.dataclass

So a stand-in golden file pins the whole plumbing either side of `diffeqsolve` to the reference's code and proves the
dump script runs end to end; only a run on real diffrax (no `--standin`) pins the solver arithmetic itself.
"""
from __future__ import annotations

import dataclasses
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _AtIndexer:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Setter:
            def set(self, v):
                out = np.array(arr, copy=True).view(JArray)
                out[idx] = v
                return out

        return _Setter()


class JArray(np.ndarray):
    """numpy array with jax's functional `.at[idx].set(v)`."""

    @property
    def at(self):
        return _AtIndexer(self)


def J(x, dtype=None):
    return np.array(x, dtype=dtype or np.float64).view(JArray)


class _AutoStub(types.ModuleType):
    """A module whose every attribute is a fresh dummy class (enough for `from x import Y` and type annotations)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (), {"__init__": lambda self, *a, **k: None, "__module__": self.__name__})
        setattr(self, name, cls)
        return cls


def _stub(name):
    m = _AutoStub(name)
    m.__path__ = []
    sys.modules[name] = m
    return m


# ------------------------------------------------------------------------------------------------ jax
def _tree_index(tree, b):
    if isinstance(tree, dict):
        return {k: _tree_index(v, b) for k, v in tree.items()}
    if isinstance(tree, (tuple, list)):
        return type(tree)(_tree_index(v, b) for v in tree)
    return J(np.asarray(tree)[b])


def _tree_stack(items):
    first = items[0]
    if isinstance(first, dict):
        return {k: _tree_stack([it[k] for it in items]) for k in first}
    if isinstance(first, (tuple, list)):
        return type(first)(_tree_stack([it[i] for it in items]) for i in range(len(first)))
    return np.stack([np.asarray(it) for it in items])


def _tree_len(tree):
    if isinstance(tree, dict):
        return _tree_len(next(iter(tree.values())))
    if isinstance(tree, (tuple, list)):
        return _tree_len(tree[0])
    return np.asarray(tree).shape[0]


def _vmap(fn, in_axes=0):
    def mapped(*args):
        B = _tree_len(args[0])
        return _tree_stack([fn(*[_tree_index(a, b) for a in args]) for b in range(B)])
    return mapped


def _install_jax():
    jnp = types.ModuleType("jax.numpy")
    for name in dir(np):
        if not name.startswith("_"):
            setattr(jnp, name, getattr(np, name))
    jnp.array = lambda x, dtype=None: J(x, dtype)
    jnp.asarray = lambda x, dtype=None: J(x, dtype)
    jnp.zeros_like = lambda x: np.zeros_like(np.asarray(x)).view(JArray)
    jnp.zeros = lambda s, dtype=None: np.zeros(s).view(JArray)
    jnp.linspace = lambda *a, **k: np.linspace(*a, **k).view(JArray)
    jnp.ndarray = np.ndarray
    jax = types.ModuleType("jax")
    jax.__path__ = []
    jax.numpy = jnp
    jax.jit = lambda f=None, **kw: f if f is not None else (lambda g: g)
    jax.vmap = _vmap
    jax.Array = np.ndarray
    jax.__version__ = "standin"
    jax.config = types.SimpleNamespace(update=lambda *a, **k: None)
    typing_mod = types.ModuleType("jax.typing")
    typing_mod.ArrayLike = object
    jax.typing = typing_mod
    random = types.ModuleType("jax.random")
    random.PRNGKey = lambda seed: np.array([0, seed], dtype=np.uint32)
    jax.random = random
    sys.modules.update({"jax": jax, "jax.numpy": jnp, "jax.typing": typing_mod, "jax.random": random})


# ------------------------------------------------------------------------------------------------ chex
def _install_chex():
    chex = types.ModuleType("chex")

    def _dc(cls=None, **kw):
        def wrap(c):
            return dataclasses.dataclass(c)
        return wrap if cls is None else wrap(cls)

    chex.dataclass = _dc
    chex.ArrayDevice = np.ndarray
    sys.modules["chex"] = chex


# ------------------------------------------------------------------------------------------------ diffrax
def _install_diffrax():
    sys.path.insert(0, ROOT)
    from oracle import oracle_np

    dfx = types.ModuleType("diffrax")
    dfx.__version__ = "standin:oracle_np"

    class AbstractSolver:
        pass

    class Tsit5(AbstractSolver):
        pass

    class AbstractStepSizeController:
        pass

    @dataclasses.dataclass
    class PIDController(AbstractStepSizeController):
        rtol: float
        atol: float

    @dataclasses.dataclass
    class ClipStepSizeController(AbstractStepSizeController):
        controller: PIDController
        jump_ts: object = None
        step_ts: object = None

    class ConstantStepSize(AbstractStepSizeController):
        pass

    @dataclasses.dataclass
    class ODETerm:
        vector_field: object

    @dataclasses.dataclass
    class SubSaveAt:
        ts: object = None
        fn: object = None

    @dataclasses.dataclass
    class SaveAt:
        ts: object = None
        subs: object = None

    @dataclasses.dataclass
    class Solution:
        ts: object
        ys: object
        stats: dict
        result: int

    def diffeqsolve(terms, solver, t0, t1, dt0, y0, args=None, *, saveat, stepsize_controller, max_steps=4096, **kw):
        assert isinstance(solver, Tsit5), "the stand-in restates Tsit5 only"
        sub = saveat.subs
        ts = np.asarray(sub.ts if sub is not None else saveat.ts, dtype=np.float64)
        opts = dict(t0=float(t0), max_steps=int(max_steps), save_ts=ts)
        if isinstance(stepsize_controller, ConstantStepSize):
            opts["const_dt"] = float(dt0)
        else:
            pid = stepsize_controller.controller
            opts.update(rtol=pid.rtol, atol=pid.atol)
            if stepsize_controller.jump_ts is not None:
                opts["jump_ts"] = tuple(np.asarray(stepsize_controller.jump_ts, dtype=np.float64))
        y0 = tuple(J(c) for c in y0)
        ys, stats = oracle_np.solve(terms.vector_field, y0, args, float(t1), **opts)
        if stats["result"] != 0:
            # diffrax throw=True: exceeding max_steps raises at run time
            raise RuntimeError("The maximum number of solver steps was reached. Try increasing `max_steps`.")
        if sub is not None:  # SubSaveAt(fn): applied to every saved state
            rows = [sub.fn(ts[k], tuple(J(c[k]) for c in ys), args) for k in range(ts.size)]
            ys = tuple(np.stack([np.asarray(r[i]) for r in rows]) for i in range(len(y0)))
        stats = dict(num_steps=stats["num_steps"], num_accepted_steps=stats["num_accepted_steps"],
                     num_rejected_steps=stats["num_rejected_steps"], max_steps=int(max_steps))
        return Solution(ts=ts.view(JArray), ys=tuple(np.asarray(c).view(JArray) for c in ys), stats=stats, result=0)

    for k, v in dict(locals()).items():
        if k not in ("dfx", "oracle_np"):
            setattr(dfx, k, v)
    sys.modules["diffrax"] = dfx


def install(reference_root="/root/reference"):
    """Registers the stand-ins and puts the reference's `src/` and `examples/` on sys.path."""
    _install_jax()
    _install_chex()
    _install_diffrax()
    for name in ("numpyro", "numpyro.distributions", "numpyro.distributions.transforms", "numpyro.infer",
                 "numpyro.infer.util", "numpyro.infer.svi", "numpyro.infer.hmc", "numpyro.infer.autoguide",
                 "numpyro.optim", "numpyro.handlers", "matplotlib", "matplotlib.pyplot", "matplotlib.colors",
                 "matplotlib.axes", "seaborn", "epiweeks", "arviz", "jaxtyping"):
        _stub(name)
    sys.modules["numpyro"].distributions = sys.modules["numpyro.distributions"]
    sys.modules["numpyro.distributions"].transforms = sys.modules["numpyro.distributions.transforms"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    for sub in ("src", "examples"):
        p = os.path.join(reference_root, sub)
        if p not in sys.path:
            sys.path.insert(0, p)
