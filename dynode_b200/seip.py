"""Device-resident entry to the CTA-per-trajectory kernel of the immune-history / waning family
(include/dynode_b200_seip.h; model after reference ode_model.md:15-53,100-118,179-211).

State layout per trajectory: S[A][H][V][W], E[A][H][V][K], I[A][H][V][K], C[A][H][V][K] flattened in that order,
H = 2^K immune histories, V vaccination tiers (V = 1: no vaccination dimension, the layout of the first version).
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
    n_vax: int = 1
    n_knots: int = 0

    @property
    def n_hist(self) -> int:
        return 1 << self.n_strains

    @property
    def state_size(self) -> int:
        return self.n_ages * self.n_hist * self.n_vax * (self.n_wane + 3 * self.n_strains)

    def compartment_shapes(self) -> Tuple[Tuple[int, ...], ...]:
        A, H, W, K, V = self.n_ages, self.n_hist, self.n_wane, self.n_strains, self.n_vax
        if V == 1:
            return ((A, H, W), (A, H, K), (A, H, K), (A, H, K))
        return ((A, H, V, W), (A, H, V, K), (A, H, V, K), (A, H, V, K))

    def desc(self, save_mask: int = 0) -> _lib.SeipDesc:
        return _lib.SeipDesc(self.n_ages, self.n_strains, self.n_wane, self.n_vax, self.n_knots, int(save_mask))

    def compartment_sizes(self) -> Tuple[int, ...]:
        nS = self.n_ages * self.n_hist * self.n_vax * self.n_wane
        nX = self.n_ages * self.n_hist * self.n_vax * self.n_strains
        return (nS, nX, nX, nX)

    def saved_size(self, mask: int) -> int:
        mask = mask or 15
        return sum(sz for c, sz in enumerate(self.compartment_sizes()) if (mask >> c) & 1)

    def check_supported(self) -> None:
        d = self.desc()
        n = _lib.load().dynode_seip_state_size(ctypes.byref(d))
        if n < 0 or n > 1536:
            raise _lib.DynodeError(
                f"unsupported ODE: SEIP dims ages={self.n_ages} strains={self.n_strains} wane={self.n_wane} "
                f"vax={self.n_vax} (1..4 strains, at most 1536 state values per trajectory); there is no CPU fallback")


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


def immunity_table_vax(n_strains: int, base_protection, cross_immunity, vaccine_efficacy) -> np.ndarray:
    """immunity[j][v][w][k] with a vaccination tier: someone with v doses is protected against strain k by
    vaccine_efficacy[v][k] before waning (Strain.vaccine_efficacy, reference config/strains.py:47-55) on top of the
    infection-acquired protection: 1 - (1 - infection)(1 - base_protection[w] * efficacy)  (ode_model.md:193-211)."""
    inf = immunity_table(n_strains, base_protection, cross_immunity)  # [H][W][K]
    base = np.asarray(base_protection, dtype=np.float64)
    eff = np.asarray(vaccine_efficacy, dtype=np.float64)  # [V][K]
    vax = base[None, :, None] * eff[:, None, :]  # [V][W][K]
    return 1.0 - (1.0 - inf[:, None]) * (1.0 - vax[None])


def solve_ensemble(model: SeipModel, y0, params: Dict[str, object], contact, pop, immunity, opts: engine.SolverOptions,
                   save_ts, out=None, B: Optional[int] = None, vaccination=None, introductions=None,
                   season_tau: Optional[float] = None, save_mask: int = 0):
    """One launch, one thread block per trajectory.  params: beta, sigma, gamma [B|1, K], omega [B|1, W].
    vaccination = (base [A,V,4], knots [A,V,NK], coef [A,V,NK]) shared spline tables (reference utils/splines.py);
    introductions = dict(time, scale, pct [B|1, K], ages [K, A]) (reference config/strains.py:59-109);
    season_tau: seasonal reset of the top tier, phi(t) = sin(2 pi (t + tau) / 730)^1000 (ode_model.md:72-75).
    save_mask: bit c = compartment c of (s, e, i, c) is saved (0 = all); `opts.jump_ts` = discontinuity points.
    Returns (ys[B, T, n_saved], stats[B, 4]); everything stays on the current CUDA device and stream."""
    torch = _lib.require_cuda()
    model.check_supported()
    dev = torch.device("cuda", torch.cuda.current_device())
    n, K, W, A, H, V, NK = (model.state_size, model.n_strains, model.n_wane, model.n_ages, model.n_hist, model.n_vax,
                            model.n_knots)
    f64 = lambda x: torch.as_tensor(x, dtype=torch.float64, device=dev).contiguous()
    y0_t = f64(y0)
    p = {k: f64(params[k]) for k in ("beta", "sigma", "gamma", "omega")}
    intro = None
    if introductions is not None:
        intro = {k: f64(introductions[k]) for k in ("time", "scale", "pct")}
        intro["ages"] = f64(introductions["ages"])
        if intro["ages"].numel() != K * A:
            raise ValueError(f"introduction ages must be {K}x{A}")
    if B is None:
        B = max(1, y0_t.numel() // n, *(p[k].numel() // (W if k == "omega" else K) for k in p))
        if intro is not None:
            B = max(B, *(intro[k].numel() // K for k in ("time", "scale", "pct")))
    B = int(B)
    arr = lambda t, row, name: engine._as_array(t, row, B, name)
    c_t, pop_t, imm_t = f64(contact), f64(pop), f64(immunity)
    if c_t.numel() != A * A or pop_t.numel() != A or imm_t.numel() != H * V * W * K:
        raise ValueError(f"contact must be {A}x{A}, pop {A}, immunity {H}x{V}x{W}x{K}")
    save_dt = engine.uniform_save_dt(save_ts, float(opts.t0), float(opts.t1)) if isinstance(save_ts, np.ndarray) else 0.0
    ts_t = f64(save_ts)
    T = int(ts_t.numel())
    ys = out if out is not None else torch.empty((B, T, model.saved_size(save_mask)), dtype=torch.float64, device=dev)
    stats = torch.empty((B, 4), dtype=torch.int32, device=dev)
    cp = _lib.SeipParams()
    cp.beta, cp.sigma, cp.gamma = arr(p["beta"], K, "beta"), arr(p["sigma"], K, "sigma"), arr(p["gamma"], K, "gamma")
    cp.omega = arr(p["omega"], W, "omega")
    cp.contact, cp.pop, cp.immunity = c_t.data_ptr(), pop_t.data_ptr(), imm_t.data_ptr()
    keep = []
    if vaccination is not None:
        vb, vk, vc = (f64(x) for x in vaccination)
        if vb.numel() != A * V * 4 or vk.numel() != A * V * NK or vc.numel() != A * V * NK:
            raise ValueError(f"vaccination tables must be base {A}x{V}x4, knots and coef {A}x{V}x{NK}")
        cp.vax_base = vb.data_ptr()
        cp.vax_knots = vk.data_ptr() if NK > 0 else None
        cp.vax_coef = vc.data_ptr() if NK > 0 else None
        keep += [vb, vk, vc]
    if intro is not None:
        cp.intro_time, cp.intro_scale = arr(intro["time"], K, "intro_time"), arr(intro["scale"], K, "intro_scale")
        cp.intro_pct = arr(intro["pct"], K, "intro_pct")
        cp.intro_ages = intro["ages"].data_ptr()
    if season_tau is not None:
        cp.season_tau, cp.season_on = float(season_tau), 1.0
    md, sd = model.desc(save_mask), opts.desc(save_dt)
    _lib.check(_lib.load().dynode_seip_solve_f64(
        ctypes.byref(md), ctypes.byref(sd), B, arr(y0_t, n, "y0"), ctypes.byref(cp), ts_t.data_ptr(), T,
        ys.data_ptr(), stats.data_ptr(), ctypes.c_void_p(_lib.current_stream_ptr())))
    return ys, stats
