"""ctypes binding of the CPU oracle (oracle/dynode_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED for the solver arithmetic (diffrax is third-party and absent, see the header of
dynode_oracle.cpp); the right-hand sides are pinned by tests/golden/rhs_golden.npz.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (dynode_b200/) never does.
"""

from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdynode_oracle.so")

# family ids (dynode_oracle.cpp enum Family)
SIR_1BIN, SIR_DENSITY, SEIRS_1BIN, SEIRS_SEASONAL, SIR_AGE, SIR_AGE_RISK, SEIRS_MULTISTRAIN, SEIP, SEIPV = range(9)

_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only, no reference sources)."""
    src = os.path.join(_HERE, "dynode_oracle.cpp")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int32)
        L.oracle_solve.restype = ctypes.c_int
        L.oracle_solve.argtypes = [
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64,
            dp, ctypes.c_int64, dp, ctypes.c_int64, dp,
            ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
            ctypes.c_int64, ctypes.c_double,
            dp, ctypes.c_int, ip, ctypes.c_int, ctypes.c_int, ip, dp, dp, dp, ip, ctypes.c_int,
            dp, ctypes.c_int,
        ]
        L.oracle_rhs.restype = ctypes.c_int
        L.oracle_rhs.argtypes = [ctypes.c_int] * 4 + [ctypes.c_double, dp, dp, dp, dp]
        L.oracle_state_size.restype = ctypes.c_int
        L.oracle_state_size.argtypes = [ctypes.c_int] * 4
        L.oracle_theta_size.restype = ctypes.c_int
        L.oracle_theta_size.argtypes = [ctypes.c_int] * 4
        L.oracle_dense_weights.restype = None
        L.oracle_dense_weights.argtypes = [ctypes.c_double, dp]
        L.oracle_tableau.restype = None
        L.oracle_tableau.argtypes = [dp, dp, dp]
        L.oracle_poisson_incidence.restype = ctypes.c_int
        L.oracle_poisson_incidence.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, dp, dp, dp, dp, dp]
        L.oracle_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _dp(a: Optional[np.ndarray]):
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a: Optional[np.ndarray]):
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def _c(a, dtype=np.float64) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=dtype))


def seipv_dims(A: int, W: int, K: int, V: int, NK: int):
    """dims tuple of FAM_SEIPV: the vaccination tiers and spline knots ride in the upper bytes of the third int
    (dynode_oracle.cpp::make_dims)."""
    assert 0 < K < 256 and 0 < V < 256 and 0 <= NK < 256
    return (A, W, K | (V << 8) | (NK << 16))


def state_size(family: int, dims=(1, 1, 1)) -> int:
    return lib().oracle_state_size(family, *dims)


def theta_size(family: int, dims=(1, 1, 1)) -> int:
    return lib().oracle_theta_size(family, *dims)


def saveat_ts(start: float, stop, step=1) -> np.ndarray:
    """build_saveat's time grid (reference src/dynode/simulation/odes.py:177-179)."""
    if step <= 0:
        step = 1
    return np.linspace(start, stop, int(stop // step) + 1)


def rhs(family: int, dims, t: float, y, theta, shared=None) -> np.ndarray:
    y = _c(y).ravel()
    theta = _c(theta).ravel()
    sh = _c(shared).ravel() if shared is not None else np.zeros(1)
    out = np.empty_like(y)
    rc = lib().oracle_rhs(family, *dims, float(t), _dp(y), _dp(theta), _dp(sh), _dp(out))
    if rc != 0:
        raise ValueError(f"oracle_rhs failed rc={rc}")
    return out


def solve(family: int, dims, y0, theta, shared=None, *, t1, t0: float = 0.0, rtol: float = 1e-5,
          atol: float = 1e-6, max_steps: int = 10**6, const_dt: float = 0.0,
          save_ts: Optional[np.ndarray] = None, save_idx: Optional[Sequence[int]] = None,
          wrt: Sequence[int] = (), dy0: Optional[np.ndarray] = None, nthreads: int = 0,
          jump_ts: Sequence[float] = ()):
    """Batched solve.  y0: (B, n) or (n,) shared; theta: (B, P) or (P,) shared.

    Returns (ys[B,T,n_saved], dys[B,T,n_saved,n_wrt] or None, stats[B,4]) with
    stats = (result, num_accepted, num_rejected, num_steps).
    """
    n = state_size(family, dims)
    nt = theta_size(family, dims)
    y0 = _c(y0)
    theta = _c(theta)
    B = 1
    if y0.ndim == 2:
        B = max(B, y0.shape[0])
    if theta.ndim == 2:
        B = max(B, theta.shape[0])
    y0_bs = n if (y0.ndim == 2 and y0.shape[0] > 1) else 0
    th_bs = nt if (theta.ndim == 2 and theta.shape[0] > 1) else 0
    assert y0.shape[-1] == n, (y0.shape, n)
    assert theta.shape[-1] == nt, (theta.shape, nt)
    sh = _c(shared).ravel() if shared is not None else np.zeros(1)
    if save_ts is None:
        save_ts = saveat_ts(t0, t1, 1)
    save_ts = _c(save_ts)
    T = save_ts.shape[0]
    sidx = _c(np.arange(n) if save_idx is None else save_idx, np.int32)
    ns = sidx.shape[0]
    wrt_a = _c(list(wrt), np.int32) if len(wrt) else np.zeros(1, np.int32)
    P = len(wrt)
    ys = np.empty((B, T, ns))
    dys = np.empty((B, T, ns, P)) if P else None
    stats = np.zeros((B, 4), np.int32)
    if dy0 is not None:
        dy0 = _c(dy0)
        assert dy0.shape == (B, P, n)
    rc = lib().oracle_solve(family, *dims, B, _dp(y0), y0_bs, _dp(theta), th_bs, _dp(sh),
                            float(t0), float(t1), rtol, atol, int(max_steps), float(const_dt),
                            _dp(save_ts), T, _ip(sidx), ns, P, _ip(wrt_a), _dp(dy0), _dp(ys),
                            _dp(dys), _ip(stats), int(nthreads),
                            _dp(_c(sorted(jump_ts))) if len(jump_ts) else None, len(jump_ts))
    if rc != 0:
        raise ValueError(f"oracle_solve failed rc={rc}")
    return ys, dys, stats


def poisson_incidence(ys: np.ndarray, dys: Optional[np.ndarray], obs: np.ndarray):
    """lp[B], grad[B,P] of sum Poisson(max(diff(ys,axis=1),1e-6)).log_prob(obs)
    (reference examples/sir_infer_parameters.py:30-38)."""
    ys = _c(ys)
    B, T, m = ys.shape
    P = dys.shape[-1] if dys is not None else 0
    obs = _c(obs)
    assert obs.shape == (T - 1, m)
    lp = np.empty(B)
    grad = np.empty((B, P)) if P else None
    lib().oracle_poisson_incidence(B, T, m, P, _dp(ys), _dp(_c(dys)) if P else None, _dp(obs),
                                   _dp(lp), _dp(grad))
    return lp, grad


def dense_weights(theta: float) -> np.ndarray:
    b = np.empty(7)
    lib().oracle_dense_weights(float(theta), _dp(b))
    return b


def tableau():
    c = np.empty(6)
    a = np.empty(21)
    be = np.empty(7)
    lib().oracle_tableau(_dp(c), _dp(a), _dp(be))
    rows = []
    q = 0
    for i in range(6):
        rows.append(a[q:q + i + 1].copy())
        q += i + 1
    return c, rows, be


def num_threads() -> int:
    return lib().oracle_num_threads()
