"""Device-resident entry to the CTA-per-trajectory kernel of the immune-history / waning family
(include/dynode_b200_seip.h; model after reference ode_model.md:15-53,100-118,179-211).

State layout per trajectory: S[A][H][W], E[A][H][K], I[A][H][K], C[A][H][K] flattened in that order, H = 2^K.
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np

from . import _lib, engine


@dataclass(frozen=True)
class SeipModel:
    n_ages: int
    n_strains: int
    n_wane: int

    @property
    def n_hist(self) -> int:
        return 1 << self.n_strains

    @property
    def state_size(self) -> int:
        return self.n_ages * self.n_hist * (self.n_wane + 3 * self.n_strains)

    def compartment_shapes(self) -> Tuple[Tuple[int, ...], ...]:
        A, H, W, K = self.n_ages, self.n_hist, self.n_wane, self.n_strains
        return ((A, H, W), (A, H, K), (A, H, K), (A, H, K))

    def desc(self) -> _lib.SeipDesc:
        return _lib.SeipDesc(self.n_ages, self.n_strains, self.n_wane)

    def check_supported(self) -> None:
        d = self.desc()
        n = _lib.load().dynode_seip_state_size(ctypes.byref(d))
        if n < 0 or n > 1536:
            raise _lib.DynodeError(
                f"unsupported ODE: SEIP dims ages={self.n_ages} strains={self.n_strains} wane={self.n_wane} "
                "(1..4 strains, at most 1536 state values per trajectory); there is no CPU fallback")


def immunity_table(n_strains: int, base_protection, cross_immunity) -> np.ndarray:
    """immunity[j][w][k] = base_protection[w] * max over strains l in history j of cross_immunity[k][l]
    (0 for the naive history): protection against strain k of someone whose last recovery is w waning stages
    old (WaneBin.base_protection, reference config/bins.py:77-89) and who has seen the strains in j
    (strain_interactions, reference config/params.py)."""
    K, H = n_strains, 1 << n_strains
    base = np.asarray(base_protection, dtype=np.float64)
    cross = np.asarray(cross_immunity, dtype=np.float64).reshape(K, K)
    out = np.zeros((H, base.size, K))
    for j in range(1, H):
        seen = [l for l in range(K) if (j >> l) & 1]
        for k in range(K):
            out[j, :, k] = base * max(cross[k][l] for l in seen)
    return out


def solve_ensemble(model: SeipModel, y0, params: Dict[str, object], contact, pop, immunity, opts: engine.SolverOptions,
                   save_ts, out=None, B: Optional[int] = None):
    """One launch, one thread block per trajectory.  params: beta, sigma, gamma [B|1, K], omega [B|1, W].
    Returns (ys[B, T, n], stats[B, 4]); everything stays on the current CUDA device and stream."""
    torch = _lib.require_cuda()
    model.check_supported()
    dev = torch.device("cuda", torch.cuda.current_device())
    n, K, W, A, H = model.state_size, model.n_strains, model.n_wane, model.n_ages, model.n_hist
    f64 = lambda x: torch.as_tensor(x, dtype=torch.float64, device=dev).contiguous()
    y0_t = f64(y0)
    p = {k: f64(params[k]) for k in ("beta", "sigma", "gamma", "omega")}
    if B is None:
        B = max(1, y0_t.numel() // n, *(p[k].numel() // (W if k == "omega" else K) for k in p))
    B = int(B)
    arr = lambda t, row, name: engine._as_array(t, row, B, name)
    c_t, pop_t, imm_t = f64(contact), f64(pop), f64(immunity)
    if c_t.numel() != A * A or pop_t.numel() != A or imm_t.numel() != H * W * K:
        raise ValueError(f"contact must be {A}x{A}, pop {A}, immunity {H}x{W}x{K}")
    save_dt = engine.uniform_save_dt(save_ts, float(opts.t0), float(opts.t1)) if isinstance(save_ts, np.ndarray) else 0.0
    ts_t = f64(save_ts)
    T = int(ts_t.numel())
    ys = out if out is not None else torch.empty((B, T, n), dtype=torch.float64, device=dev)
    stats = torch.empty((B, 4), dtype=torch.int32, device=dev)
    cp = _lib.SeipParams()
    cp.beta, cp.sigma, cp.gamma = arr(p["beta"], K, "beta"), arr(p["sigma"], K, "sigma"), arr(p["gamma"], K, "gamma")
    cp.omega = arr(p["omega"], W, "omega")
    cp.contact, cp.pop, cp.immunity = c_t.data_ptr(), pop_t.data_ptr(), imm_t.data_ptr()
    md, sd = model.desc(), opts.desc(save_dt)
    _lib.check(_lib.load().dynode_seip_solve_f64(
        ctypes.byref(md), ctypes.byref(sd), B, arr(y0_t, n, "y0"), ctypes.byref(cp), ts_t.data_ptr(), T,
        ys.data_ptr(), stats.data_ptr(), ctypes.c_void_p(_lib.current_stream_ptr())))
    return ys, stats
