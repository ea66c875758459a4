"""Consumes the golden files written by baseline/dump_diffrax_golden.py -- every key -- against any solver.

  tests/golden/diffrax_golden.npz   the reference on real diffrax (absent in this image: PARITY UNPINNED until
                                    someone runs the dump script where DynODE's stack is installed and commits it)
  tests/golden/standin_golden.npz   the reference's own simulate() / RHS / SolverParams driven over
                                    baseline/standin_stack.py, whose `diffeqsolve` is oracle/oracle_np.py: pins
                                    everything either side of the solver arithmetic, not the arithmetic itself

`check(gold, solver)` walks the file; `solver(name, B, **options) -> (ys[B,T,n_saved], stats[B,4])` is the CPU oracle
(tests/test_oracle.py) or the CUDA path through the public API (tests/test_gpu_parity.py).
"""
import os

import numpy as np

from tests.cases import ALL_CASES, make_case

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = {"diffrax": os.path.join(HERE, "golden", "diffrax_golden.npz"),
          "standin": os.path.join(HERE, "golden", "standin_golden.npz")}
OPTION_CASES = ("sir_age2", "seirs_seasonal", "seirs_multi_a2s3")
# what each option of baseline/dump_diffrax_golden.py::OPTIONS means for a solver
OPTIONS = {
    "jump": dict(jump_ts=(30.0, 61.5)),
    "const": dict(const_dt=0.25),
    "step2": dict(save_step=2),
    "step3": dict(save_step=3),
    "step7": dict(save_step=7),
    "sub": dict(sub_save="first_last"),
    "tight": dict(rtol=1e-8, atol=1e-10),
}


def load(kind):
    path = GOLDEN[kind]
    return np.load(path) if os.path.exists(path) else None


def tolerance(gold):
    """Relative tolerance: the north star's 1e-6 against real diffrax; 1e-9 against the stand-in, whose solver is this
    repository's own numpy restatement (two independent statements of one algorithm agree to rounding)."""
    return 1e-6 if str(gold["meta/backend"]) == "diffrax" else 1e-9


def _close(got, ref, rtol, what):
    """|got - ref| <= rtol * |ref| + 1e-3 * rtol * max|ref|: relative, with an absolute floor for compartments that
    decay through zero (1e-9 of the largest value against diffrax -- a thousand times below the solver's own atol)."""
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs golden {ref.shape}"
    scale = np.max(np.abs(ref)) if ref.size else 1.0
    excess = np.abs(got - ref) - (rtol * np.abs(ref) + 1e-3 * rtol * scale)
    assert np.all(excess <= 0), (f"{what}: max rel err "
                                 f"{(np.abs(got - ref) / (np.abs(ref) + 1e-3 * rtol * scale)).max():.3e} > {rtol:g}")


def _stats(st, gold, key, what):
    assert np.all(st[:, 0] == 0), f"{what}: a trajectory did not finish"
    for col, name in ((1, "accepted"), (2, "rejected"), (3, "num_steps")):
        assert np.array_equal(st[:, col], gold[f"{key}/{name}"]), f"{what}: {name} step counts differ from the golden file"


def check(gold, solver, cases=ALL_CASES):
    """Returns the number of (case, option) blocks compared."""
    rtol = tolerance(gold)
    draws = int(gold["meta/draws"])
    n_blocks = 0
    for name in cases:
        ys, st = solver(name, draws)
        _stats(st, gold, name, name)
        nf = gold[f"{name}/ys_full"].shape[0]
        _close(ys[:nf], gold[f"{name}/ys_full"], rtol, f"{name}/ys_full")
        _close(ys[:, gold[f"{name}/rows"]], gold[f"{name}/ys_rows"], rtol, f"{name}/ys_rows")
        t1 = make_case(name, 1)["t1"]
        assert np.array_equal(gold[f"{name}/ts"], np.linspace(0.0, t1, int(t1 // 1) + 1))
        n_blocks += 1
        if name not in OPTION_CASES:
            continue
        for opt, kw in OPTIONS.items():
            key = f"{name}/{opt}"
            ref = gold[f"{key}/ys"]
            ys, st = solver(name, ref.shape[0], **kw)
            _stats(st, gold, key, key)
            _close(ys, ref, rtol, key)
            if "save_step" in kw:  # build_saveat: linspace(0, t1, int(t1 // step) + 1), reference odes.py:177-179
                assert np.array_equal(gold[f"{key}/ts"], np.linspace(0.0, t1, int(t1 // kw["save_step"]) + 1))
            n_blocks += 1
    return n_blocks
