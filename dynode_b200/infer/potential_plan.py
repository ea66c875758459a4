"""A whole model evaluation in three launches (include/dynode_b200_ppl.h, DynodePotentialPlan).

numpyro evaluates `potential_energy(model)(z)` and its gradient once per leapfrog (reference
src/dynode/infer/inference.py:149-163).  Composed from tensor operations -- sites, bijectors, `get_odeparams`, the
fused ODE log-likelihood, autograd back to z -- that is ~35 launches here, and at a few thousand chains a NUTS round is
little else.  For the models DynODE writes (priors with constant parameters on scalar sites, rates that are monomials
of the sites, one fused Poisson-incidence likelihood) the evaluation is COMPILED into

    dynode_potential_pre_f64  ->  dynode_poisson_loglik_{grad,adjoint}_f64  ->  dynode_potential_post_f64

`compile_plan` discovers the structure by running the model once on grad-tracking site values (the role JAX tracing
plays for the reference), proves the rate map monomial from its log-derivatives at two random points, and then checks
the compiled evaluation against the composed one on random positions; anything that does not fit, or does not agree,
keeps the composed path.  Both paths run on the device; neither is a CPU fallback.
"""

from __future__ import annotations

import ctypes
import os
import warnings
from typing import List, Optional

import torch

from .. import _lib, engine
from ..simulation import autograd as ag
from . import distributions as dist
from . import ppl

_RECORDER: List[Optional[list]] = [None]


def record_loglik_call(cfg, y0r, theta, lp):
    """Called by simulation.simulate_incidence_loglik: remembers the solve behind a model's likelihood factor."""
    if _RECORDER[0] is not None:
        _RECORDER[0].append(dict(cfg=cfg, y0=y0r, theta=theta, lp=lp))


class _Recording:
    def __enter__(self):
        self.prev, _RECORDER[0] = _RECORDER[0], []
        return _RECORDER[0]

    def __exit__(self, *exc):
        _RECORDER[0] = self.prev
        return False


class PotentialPlan:
    """Compiled potential energy of one ModelDensity."""

    def __init__(self, md, c_plan, cfg, y0, rate_cols, why=""):
        self.md, self.c_plan, self.cfg, self.y0 = md, c_plan, cfg, y0
        self.D, self.K = int(c_plan.n_sites), int(c_plan.n_rates)
        self.rate_cols = rate_cols  # theta columns that carry a gradient (cfg.wrt_cols)
        self._cols = {}

    # ------------------------------------------------------------------ evaluation
    def _colmap(self, use_adjoint: bool):
        """theta column -> column of the gradient array the log-likelihood launch writes."""
        m = self._cols.get(use_adjoint)
        if m is None:
            cols = [-1] * self.K
            if use_adjoint:
                S = self.cfg.model.n_strains
                base = {"beta": 0, "gamma": S, "sigma": 2 * S, "omega": 3 * S, "season_amp": 4 * S,
                        "season_phase": 4 * S + 1}
                for col in self.cfg.wrt_cols:
                    for kind, first, width in self.cfg.layout:
                        if first <= col < first + width:
                            cols[col] = base[kind] + (col - first)
            else:
                for j, col in enumerate(self.cfg.wrt_cols):
                    cols[col] = j
            m = (ctypes.c_int32 * self.K)(*cols)
            self._cols[use_adjoint] = m
        return m

    def launch_key(self, C: int, n_rows: int):
        """What `potential_and_grad` would launch for C rows of which n_rows run (hashable): a sampler that replays a
        captured graph asks this to learn whether the capture is still the right one."""
        return ag.use_adjoint(self.cfg.model, len(self.cfg.wrt_cols), self.cfg.opts(), C, n_rows=n_rows)

    def potential_and_grad(self, Z: torch.Tensor):
        """U [C], dU/dz [C, D]: pre -> fused log-likelihood -> post, nothing else touches the device."""
        cfg, pl = self.cfg, self.cfg.payload
        L = _lib.load()
        C = Z.shape[0]
        dev = Z.device
        if Z.stride(1) != 1 or (C > 1 and Z.stride(0) < self.D):
            Z = Z.contiguous()
        z_stride = Z.stride(0) if C > 1 else self.D  # the stride of a size-1 axis is arbitrary
        stream = ctypes.c_void_p(_lib.current_stream_ptr())
        only = engine.current_row_mask(C)
        only_ptr = only.data_ptr() if only is not None else None
        theta = torch.empty((C, self.K), dtype=torch.float64, device=dev)
        aux = torch.empty((C, 3 * self.D + 1), dtype=torch.float64, device=dev)
        _lib.check(L.dynode_potential_pre_f64(ctypes.byref(self.c_plan), C, Z.data_ptr(), z_stride, theta.data_ptr(),
                                              aux.data_ptr(), only_ptr, stream))
        opts = cfg.opts()
        params = ag._kernel_params(cfg, theta)
        n_dir = len(cfg.wrt_cols)
        use_adj = ag.use_adjoint(cfg.model, n_dir, opts, C)
        wrt = cfg.wrt_ids()
        lp_fb = g_fb = st = None
        if use_adj:
            lp, g, _, st = engine.poisson_loglik_adjoint(cfg.model, self.y0, params, pl.contact, opts, pl.save_ts,
                                                         pl.obs_comp, pl.obs, pl.lp_const, B=C,
                                                         cap=ag.adjoint_capacity(), zero_masked=False)
            # rows past the checkpoint capacity: re-evaluated by forward sensitivities, masked to those rows
            over = (st[:, _lib.STAT_RESULT] == _lib.RESULT_ADJOINT_CAPACITY).view(torch.uint8)
            with engine.only_rows(over):
                lp_fb, g_fb, _ = engine.poisson_loglik_grad(cfg.model, self.y0, params, pl.contact, opts, pl.save_ts,
                                                            pl.obs_comp, pl.obs, pl.lp_const, wrt=wrt, B=C,
                                                            zero_masked=False)
        else:
            lp, g, st = engine.poisson_loglik_grad(cfg.model, self.y0, params, pl.contact, opts, pl.save_ts,
                                                   pl.obs_comp, pl.obs, pl.lp_const, wrt=wrt, B=C, zero_masked=False)
        U = torch.empty((C,), dtype=torch.float64, device=dev)
        dU = torch.empty((C, self.D), dtype=torch.float64, device=dev)
        _lib.check(L.dynode_potential_post_f64(
            ctypes.byref(self.c_plan), C, theta.data_ptr(), aux.data_ptr(), lp.data_ptr(), g.data_ptr(), g.stride(0),
            self._colmap(use_adj), lp_fb.data_ptr() if lp_fb is not None else None,
            g_fb.data_ptr() if g_fb is not None else None, g_fb.stride(0) if g_fb is not None else 0,
            self._colmap(False) if g_fb is not None else None,
            st.data_ptr() if g_fb is not None else None, only_ptr, U.data_ptr(), dU.data_ptr(), stream))
        return U, dU


def _site_descs(md):
    """One DynodeSiteDesc per column of z, or None when a site is outside the fused families."""
    descs = []
    for name, info in md.sites.items():
        fn = info["fn"]
        bij, fam = dist._bijector_spec(fn.support), dist._family_spec(fn)
        if bij is None or fam is None:
            return None, f"site `{name}` has a prior outside the fused families (or non-constant parameters)"
        a, b = info["slice"]
        for _ in range(b - a):
            descs.append(_lib.SiteDesc(bij[0], fam[0], bij[1], bij[2], fam[1], fam[2], fam[3], fam[4], fam[5]))
    return descs, ""


def _trace_rates(md, x: torch.Tensor):
    """Run the model once for one draw with the constrained site values `x` [D] as grad-tracking leaves.
    Returns (recorded log-likelihood call, list of trace messages)."""
    vals = {}
    for name, info in md.sites.items():
        a, b = info["slice"]
        vals[name] = x[a:b].reshape(info["shape"])
    with ppl.substitute(data=vals), ppl.trace() as tr, _Recording() as rec:
        md.model(*md.args, **md.kwargs)
    return rec, tr.trace


def compile_plan(md, n_check: int = 12, rtol: float = 1e-9):
    """PotentialPlan for `md`, or (None, reason)."""
    if md.device.type != "cuda":
        return None, "the compiled evaluation runs on a CUDA device"
    if os.environ.get("DYNODE_B200_PLAN", "1") == "0":
        return None, "disabled by DYNODE_B200_PLAN=0"
    D = md.dim
    if D > _lib.PLAN_MAX_SITES:
        return None, f"{D} latent dimensions exceed DYNODE_PLAN_MAX_SITES"
    descs, why = _site_descs(md)
    if descs is None:
        return None, why
    gen = torch.Generator(device=md.device).manual_seed(20260103)
    z0 = md.init_to_median(1)[0]
    points = []
    for k in range(2):
        z = z0 + 0.3 * torch.randn(D, dtype=torch.float64, device=md.device, generator=gen)
        x = _constrained_vector(md, z).detach().requires_grad_(True)
        with torch.enable_grad():
            rec, trace = _trace_rates(md, x)
            if len(rec) != 1:
                return None, f"the model makes {len(rec)} fused log-likelihood calls (exactly one is compiled)"
            call = rec[0]
            for name, msg in trace.items():
                if msg["type"] == "sample" and name not in md.sites:
                    return None, f"site `{name}` is an observed / unlisted sample site (only a fused likelihood factor is compiled)"
                if msg["type"] == "factor" and msg["value"] is not call["lp"]:
                    return None, f"factor `{name}` is not the fused log-likelihood itself"
            if sum(1 for m in trace.values() if m["type"] == "factor") != 1:
                return None, "the model must add exactly one factor (the fused log-likelihood)"
            cfg, theta = call["cfg"], call["theta"]
            if cfg.y0_grad:
                return None, "the initial state depends on latent sites"
            if theta.shape[0] != 1:
                return None, "per-draw rates must form one row"
            K = theta.shape[1]
            if K > _lib.PLAN_MAX_RATES:
                return None, f"{K} rate columns exceed DYNODE_PLAN_MAX_RATES"
            J = torch.zeros((K, D), dtype=torch.float64, device=md.device)
            for kk in range(K):
                if theta[0, kk].requires_grad:
                    (gk,) = torch.autograd.grad(theta[0, kk], x, retain_graph=True, allow_unused=True)
                    if gk is not None:
                        J[kk] = gk
        th = theta[0].detach()
        if bool((th == 0).any()) or bool((x == 0).any()):
            return None, "a rate or a site is exactly zero at the probe point"
        E = J * x.detach()[None, :] / th[:, None]  # d log theta_k / d log x_j
        points.append((x.detach(), th, E, cfg, call["y0"]))
    (x1, th1, E1, cfg, y0), (x2, th2, E2, cfg2, _) = points
    Er = torch.round(E1)
    if float((E1 - Er).abs().max()) > 1e-9 or float((E2 - Er).abs().max()) > 1e-9 or float(Er.abs().max()) > 1:
        return None, "the rates are not monomials (exponents -1, 0, 1) of the latent sites"
    if cfg.wrt_cols != cfg2.wrt_cols or cfg.layout != cfg2.layout:
        return None, "the model's structure changes with the latent values"
    c1 = th1 / torch.prod(x1[None, :] ** Er, dim=1)
    c2 = th2 / torch.prod(x2[None, :] ** Er, dim=1)
    if float(((c1 - c2).abs() / c1.abs()).max()) > 1e-12:
        return None, "the rates are not monomials of the latent sites (coefficient changes between probe points)"
    for kk in range(th1.shape[0]):  # a column that depends on sites must be differentiated, and vice versa
        if bool(Er[kk].abs().sum() > 0) != (kk in cfg.wrt_cols):
            return None, "rate columns that depend on sites do not match the differentiated columns"
    c_plan = _lib.PotentialPlan()
    c_plan.n_sites, c_plan.n_rates = D, int(th1.shape[0])
    for j, d in enumerate(descs):
        c_plan.site[j] = d
    c_host, e_host = c1.cpu().tolist(), Er.cpu().to(torch.int64).tolist()
    for kk in range(c_plan.n_rates):
        c_plan.rate_c[kk] = c_host[kk]
        for j in range(D):
            c_plan.rate_e[kk][j] = int(e_host[kk][j])
    y0_dev = y0.detach()
    plan = PotentialPlan(md, c_plan, cfg, y0_dev, cfg.wrt_cols)
    # ---- the compiled evaluation must reproduce the composed one
    Z = z0[None, :] + 0.5 * torch.randn(n_check, D, dtype=torch.float64, device=md.device, generator=gen)
    U_c, g_c = md.potential_and_grad_composed(Z)
    U_p, g_p = plan.potential_and_grad(Z)
    fin = torch.isfinite(U_c)
    if not bool((torch.isfinite(U_p) == fin).all()):
        return None, "compiled and composed evaluations disagree on which points are finite"
    scale = g_c[fin].abs().max() if bool(fin.any()) else torch.tensor(1.0)
    ok = torch.allclose(U_p[fin], U_c[fin], rtol=rtol, atol=1e-9) and \
        torch.allclose(g_p[fin], g_c[fin], rtol=1e-7, atol=float(1e-8 * scale))
    if not ok:
        err_u = float(((U_p[fin] - U_c[fin]).abs() / U_c[fin].abs().clamp_min(1e-300)).max())
        err_g = float(((g_p[fin] - g_c[fin]).abs()).max() / scale)
        warnings.warn(f"potential plan rejected: compiled vs composed evaluation differ (U rel {err_u:.2e}, "
                      f"grad {err_g:.2e} of scale); keeping the composed path")
        return None, "compiled evaluation does not reproduce the composed one"
    return plan, ""


def _constrained_vector(md, z: torch.Tensor) -> torch.Tensor:
    """Constrained site values as one [D] vector in z's column order."""
    cols = []
    for name, info in md.sites.items():
        a, b = info["slice"]
        t = dist.biject_to(info["fn"].to(md.device).support)
        cols.append(t(z[a:b].reshape(info["shape"])).reshape(-1))
    return torch.cat(cols)
