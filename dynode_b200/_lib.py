"""ctypes binding of libdynode_b200.so (the C ABI in include/dynode_b200.h).

The product path has NO CPU fallback: if the CUDA library is missing or no CUDA device is present,
every compute entry point raises.  torch is used only for device memory and streams.
"""

from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DYNODE_B200_LIB") or os.path.join(HERE, "libdynode_b200.so")

FLOW_SIR, FLOW_SEIRS, FLOW_SEIRS_C = 0, 1, 2
FLAG_SEASONAL, FLAG_DENSITY_DEP = 1, 2
P_BETA, P_GAMMA, P_SIGMA, P_OMEGA, P_SEASON_AMP, P_SEASON_PHASE = range(6)
STAT_RESULT, STAT_ACCEPTED, STAT_REJECTED, STAT_STEPS = range(4)
RESULT_OK, RESULT_MAX_STEPS, RESULT_ADJOINT_CAPACITY = range(3)

EXPORTED_SYMBOLS = (
    "dynode_version", "dynode_last_error", "dynode_state_size", "dynode_num_compartments",
    "dynode_saved_size", "dynode_is_supported", "dynode_solve_f64", "dynode_solve_sens_f64",
    "dynode_poisson_loglik_grad_f64", "dynode_poisson_loglik_adjoint_f64", "dynode_probe_dfma",
    "dynode_probe_hbm_write",
    "dynode_nuts_round_pre", "dynode_nuts_round_post", "dynode_seip_state_size", "dynode_seip_solve_f64",
    "dynode_bijector_f64", "dynode_bijector_vjp_f64", "dynode_site_logdensity_f64",
    "dynode_site_logdensity_vjp_f64",
    "dynode_host_alloc", "dynode_host_free", "dynode_host_info",
    "dynode_potential_pre_f64", "dynode_potential_post_f64",
)


def wrt_id(kind: int, strain: int = 0) -> int:
    return kind * 16 + strain


class DynodeError(RuntimeError):
    """Raised when the native library rejects a call (invalid argument / unsupported ODE)."""


class ModelDesc(ctypes.Structure):
    _fields_ = [("flow", ctypes.c_int32), ("flags", ctypes.c_int32),
                ("n_groups", ctypes.c_int32), ("n_strains", ctypes.c_int32)]


class SolverDesc(ctypes.Structure):
    _fields_ = [("t0", ctypes.c_double), ("t1", ctypes.c_double), ("rtol", ctypes.c_double),
                ("atol", ctypes.c_double), ("const_dt", ctypes.c_double), ("max_steps", ctypes.c_int64),
                ("save_dt", ctypes.c_double), ("jump_ts", ctypes.c_void_p), ("n_jump", ctypes.c_int32),
                ("only", ctypes.c_void_p)]


class SiteDesc(ctypes.Structure):
    """DynodeSiteDesc of include/dynode_b200_ppl.h."""

    _fields_ = [("bijector", ctypes.c_int32), ("family", ctypes.c_int32), ("a", ctypes.c_double),
                ("b", ctypes.c_double), ("p0", ctypes.c_double), ("p1", ctypes.c_double), ("c", ctypes.c_double),
                ("aff_loc", ctypes.c_double), ("aff_scale", ctypes.c_double)]


PLAN_MAX_SITES, PLAN_MAX_RATES = 16, 32


class PotentialPlan(ctypes.Structure):
    """DynodePotentialPlan of include/dynode_b200_ppl.h."""

    _fields_ = [("n_sites", ctypes.c_int32), ("n_rates", ctypes.c_int32), ("site", SiteDesc * PLAN_MAX_SITES),
                ("rate_c", ctypes.c_double * PLAN_MAX_RATES),
                ("rate_e", (ctypes.c_int8 * PLAN_MAX_SITES) * PLAN_MAX_RATES)]


class Array(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("batch_stride", ctypes.c_int64)]


class Params(ctypes.Structure):
    _fields_ = [("beta", Array), ("gamma", Array), ("sigma", Array), ("omega", Array),
                ("season_amp", Array), ("season_phase", Array), ("season_period", Array),
                ("contact", ctypes.c_void_p)]


class SeipDesc(ctypes.Structure):
    _fields_ = [("n_ages", ctypes.c_int32), ("n_strains", ctypes.c_int32), ("n_wane", ctypes.c_int32),
                ("n_vax", ctypes.c_int32), ("n_knots", ctypes.c_int32), ("save_mask", ctypes.c_uint32)]


class SeipParams(ctypes.Structure):
    _fields_ = [("beta", Array), ("sigma", Array), ("gamma", Array), ("omega", Array),
                ("contact", ctypes.c_void_p), ("pop", ctypes.c_void_p), ("immunity", ctypes.c_void_p),
                ("vax_base", ctypes.c_void_p), ("vax_knots", ctypes.c_void_p), ("vax_coef", ctypes.c_void_p),
                ("intro_time", Array), ("intro_scale", Array), ("intro_pct", Array),
                ("intro_ages", ctypes.c_void_p), ("season_tau", ctypes.c_double), ("season_on", ctypes.c_double)]


NUTS_ADAPT, NUTS_WELFORD, NUTS_SAMPLING, NUTS_END_SLOW, NUTS_END_WARMUP = 1, 2, 4, 8, 16  # DYNODE_NUTS_*
NUTS_MAX_DIM, NUTS_MAX_DEPTH = 16, 12  # DYNODE_NUTS_MAX_DIM / _MAX_DEPTH of include/dynode_b200_nuts.h
_NUTS_PTRS = (
    "z U g eps imm msqrt k nwin active need_tree searching fr_dir fr_last sched sched_n energy0 "
    "zL rL gL zR rR gR zP gP r_sum UP weight sum_acc depth nprop turning diverging "
    "s_n s_right s_turn s_div s_z s_r s_g s_zP s_gP s_rsum s_UP s_w s_acc r_ck rs_ck z_new r_half "
    "da_x da_xavg da_gavg da_t da_prox wf_n wf_mean wf_m2 out_z out_accept out_steps out_div out_energy "
    "out_depth last_accept last_steps n_leap any_active").split()


class NutsState(ctypes.Structure):
    """DynodeNutsState of include/dynode_b200_nuts.h (field order is the header's)."""

    _fields_ = ([("C", ctypes.c_int32), ("D", ctypes.c_int32), ("max_depth", ctypes.c_int32), ("N", ctypes.c_int32),
                 ("n_warmup", ctypes.c_int32), ("dense", ctypes.c_int32),
                 ("target_accept", ctypes.c_double)] + [(n, ctypes.c_void_p) for n in _NUTS_PTRS])


_lib: Optional[ctypes.CDLL] = None


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load the native library (building it in-tree with nvcc if it is absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise DynodeError(f"{LIB_PATH} is missing; run `python -m dynode_b200._build`")
        from . import _build
        _build.build()
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, dbl, u32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_uint32
    MP, SP, PP = ctypes.POINTER(ModelDesc), ctypes.POINTER(SolverDesc), ctypes.POINTER(Params)
    L.dynode_version.restype = ctypes.c_int
    L.dynode_last_error.restype = ctypes.c_char_p
    for f in (L.dynode_state_size, L.dynode_num_compartments, L.dynode_is_supported):
        f.restype = ctypes.c_int
        f.argtypes = [MP]
    L.dynode_saved_size.restype = ctypes.c_int
    L.dynode_saved_size.argtypes = [MP, u32]
    L.dynode_solve_f64.restype = ctypes.c_int
    L.dynode_solve_f64.argtypes = [MP, SP, i64, Array, PP, vp, i32, u32, vp, vp, vp]
    L.dynode_solve_sens_f64.restype = ctypes.c_int
    L.dynode_solve_sens_f64.argtypes = [MP, SP, i64, Array, PP, vp, i32, u32, i32,
                                        ctypes.POINTER(i32), vp, vp, vp, vp, vp]
    L.dynode_poisson_loglik_grad_f64.restype = ctypes.c_int
    L.dynode_poisson_loglik_grad_f64.argtypes = [MP, SP, i64, Array, PP, vp, i32, i32, vp, dbl, i32,
                                                 ctypes.POINTER(i32), vp, vp, vp, vp, vp]
    L.dynode_poisson_loglik_adjoint_f64.restype = ctypes.c_int
    L.dynode_poisson_loglik_adjoint_f64.argtypes = [MP, SP, i64, Array, PP, vp, i32, i32, vp, dbl, vp, vp, vp, vp,
                                                    vp, i32, vp, vp]
    L.dynode_seip_state_size.restype = ctypes.c_int
    L.dynode_seip_state_size.argtypes = [ctypes.POINTER(SeipDesc)]
    L.dynode_seip_solve_f64.restype = ctypes.c_int
    L.dynode_seip_solve_f64.argtypes = [ctypes.POINTER(SeipDesc), SP, i64, Array, ctypes.POINTER(SeipParams), vp, i32,
                                        vp, vp, vp]
    NP = ctypes.POINTER(NutsState)
    L.dynode_nuts_round_pre.restype = ctypes.c_int
    L.dynode_nuts_round_pre.argtypes = [NP, vp, vp, vp]
    L.dynode_nuts_round_post.restype = ctypes.c_int
    L.dynode_nuts_round_post.argtypes = [NP, vp, vp, vp, vp]
    dbl = ctypes.c_double
    L.dynode_bijector_f64.restype = ctypes.c_int
    L.dynode_bijector_f64.argtypes = [i32, i64, vp, dbl, dbl, vp, vp, vp]
    L.dynode_bijector_vjp_f64.restype = ctypes.c_int
    L.dynode_bijector_vjp_f64.argtypes = [i32, i64, vp, dbl, vp, vp, vp, vp]
    SD = ctypes.POINTER(SiteDesc)
    L.dynode_site_logdensity_f64.restype = ctypes.c_int
    L.dynode_site_logdensity_f64.argtypes = [SD, i64, vp, i64, vp, vp, vp]
    L.dynode_site_logdensity_vjp_f64.restype = ctypes.c_int
    L.dynode_site_logdensity_vjp_f64.argtypes = [SD, i64, vp, i64, vp, vp, vp, vp]
    PL = ctypes.POINTER(PotentialPlan)
    L.dynode_potential_pre_f64.restype = ctypes.c_int
    L.dynode_potential_pre_f64.argtypes = [PL, i64, vp, i64, vp, vp, vp, vp]
    L.dynode_potential_post_f64.restype = ctypes.c_int
    L.dynode_potential_post_f64.argtypes = [PL, i64, vp, vp, vp, vp, i64, ctypes.POINTER(i32), vp, vp, i64,
                                            ctypes.POINTER(i32), vp, vp, vp, vp, vp]
    L.dynode_probe_dfma.restype = i64
    L.dynode_probe_dfma.argtypes = [vp, i32, vp]
    L.dynode_probe_hbm_write.restype = ctypes.c_int
    L.dynode_probe_hbm_write.argtypes = [vp, i64, vp]
    L.dynode_host_alloc.restype = ctypes.c_int
    L.dynode_host_alloc.argtypes = [ctypes.c_size_t, u32, i32, ctypes.POINTER(vp)]
    L.dynode_host_free.restype = ctypes.c_int
    L.dynode_host_free.argtypes = [vp, ctypes.c_size_t, u32]
    L.dynode_host_info.restype = i64
    L.dynode_host_info.argtypes = [vp, ctypes.c_size_t]
    _lib = L
    return L


def last_error() -> str:
    return load().dynode_last_error().decode()


def check(rc: int) -> None:
    if rc != 0:
        raise DynodeError(last_error())


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise DynodeError(
            "dynode_b200 needs a CUDA device (B200, sm_100a): the engine has no CPU fallback")
    return torch


def current_stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def i32_array(vals: Sequence[int]):
    arr = (ctypes.c_int32 * max(1, len(vals)))(*vals)
    return arr
