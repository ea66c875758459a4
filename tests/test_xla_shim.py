"""The XLA-FFI shim (dynode_b200/csrc/xla_ffi_shim.cc) must meet a compiler in every environment.

* Where jaxlib is installed (`jax.ffi.include_dir()` resolves) the shim is compiled against the REAL headers and its
  handler symbols are checked -- that is the CI step for a DynODE checkout.
* In this image jaxlib is absent (no network), so that test SKIPS LOUDLY, and the shim is compiled against
  tests/mock_xla instead: a model of the header subset it uses whose XLA_FFI_DEFINE_HANDLER_SYMBOL statically checks
  every handler's parameter list against its Ffi::Bind() chain.  A mock proves the file is valid C++ that agrees with
  include/*.h and with its own bindings; it proves nothing about XLA's runtime.
"""
import os
import subprocess
import tempfile

import pytest

from dynode_b200 import _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "tests", "mock_xla")


def _symbols(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


def test_shim_compiles_against_the_real_jaxlib_headers():
    try:
        import jax.ffi  # noqa: F401
    except ImportError:
        pytest.skip("XLA-FFI SHIM NOT COMPILED AGAINST REAL HEADERS: jaxlib is not installed in this image "
                    "(jax.ffi.include_dir() unavailable); see test_shim_compiles_against_the_mock_headers")
    lib = _build.build_xla_shim()
    assert set(_build.XLA_HANDLERS) <= _symbols(lib)


def test_shim_compiles_against_the_mock_headers():
    with tempfile.TemporaryDirectory() as tmp:
        lib = _build.build_xla_shim(include_dir=MOCK, out=os.path.join(tmp, "libshim_mock.so"))
        syms = _symbols(lib)
    missing = set(_build.XLA_HANDLERS) - syms
    assert not missing, f"handlers not exported: {sorted(missing)}"


def test_the_mock_rejects_a_handler_that_disagrees_with_its_binding():
    """The mock's static check is what gives the compile test its value: a handler whose parameters are out of order
    with respect to its Ffi::Bind() chain must not compile."""
    bad = r'''
#include <cuda_runtime.h>
#include "xla/ffi/api/ffi.h"
namespace ffi = xla::ffi;
using F64 = ffi::Buffer<ffi::F64>;
static ffi::Error Impl(cudaStream_t, F64, double, int32_t, ffi::ResultBuffer<ffi::F64>) { return ffi::Error::Success(); }
XLA_FFI_DEFINE_HANDLER_SYMBOL(Bad, Impl, ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>()
                              .Attr<int32_t>("k").Attr<double>("x").Ret<F64>());
'''
    good = bad.replace('.Attr<int32_t>("k").Attr<double>("x")', '.Attr<double>("x").Attr<int32_t>("k")')
    with tempfile.TemporaryDirectory() as tmp:
        for name, src, ok in (("bad", bad, False), ("good", good, True)):
            path = os.path.join(tmp, name + ".cc")
            open(path, "w").write(src)
            r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", MOCK, "-I", _build._cuda_include(), path],
                               capture_output=True, text=True)
            assert (r.returncode == 0) == ok, r.stderr[-1500:]
            if not ok:
                assert "do not match its Ffi::Bind() chain" in r.stderr


# ---------------------------------------------------------------------------------------------------------------
# The handlers' BODIES, executed: tests/mock_xla/shim_driver.cc wraps device pointers in the mock's Buffer and calls
# SolveImpl / LoglikGradImpl.  Results must be bit-identical to the ctypes binding's (same library underneath), so
# any difference is the shim's own plumbing: the [B][3] season convention, batch-stride detection, jump_ts / row-mask
# trailing buffers, attribute order.
def _driver(tmp):
    import ctypes

    from dynode_b200 import _lib

    _lib.load()
    lib = _build.build_xla_shim(include_dir=MOCK, out=os.path.join(tmp, "libshim_driver.so"),
                                source=os.path.join(MOCK, "shim_driver.cc"))
    L = ctypes.CDLL(lib)
    L.shim_driver_last_error.restype = ctypes.c_char_p
    return L


def _dev(torch, x, B=None, row=None):
    if x is None:
        return None
    t = torch.as_tensor(x, dtype=torch.float64).cuda().contiguous()
    if B is not None and t.numel() == row:  # one shared row: the shim sees a buffer without the batch axis
        t = t.reshape(row)
    return t


def _ptr(t):
    return None if t is None else t.data_ptr()


@pytest.mark.gpu
@pytest.mark.parametrize("name,jumps,masked,const_dt", [
    ("seirs_seasonal", (), False, 0.0),
    ("seirs_multi_a2s3", (30.0, 61.5), True, 0.0),
    ("seirs_multi_a2s3", (30.0, 61.5), False, 0.25),  # constant step: the discontinuity points must be ignored
    ("sir_age_risk32", (), True, 0.0),
])
def test_handler_body_of_the_solve_matches_the_ctypes_binding(tmp_path, name, jumps, masked, const_dt):
    import ctypes

    import numpy as np
    import torch

    from dynode_b200 import _lib, engine
    from dynode_b200.synthetic import make_case

    B, t1 = 300, 120
    case = make_case(name, B)
    model, prm = case["model"], case["params"]
    G, S, n = model.n_groups, model.n_strains, model.state_size
    ts = np.linspace(0.0, t1, t1 + 1)
    # simulate() drops the discontinuity points in constant-step mode, as the reference does (odes.py:113-131); the
    # shim has to do the same on its own, so the ctypes arm is given none and the shim arm is given both
    opts = engine.SolverOptions(t1=t1, jump_ts=() if const_dt > 0 else jumps, const_dt=const_dt)
    save_mask = model.full_mask() if name != "seirs_multi_a2s3" else 0b10001
    ns = model.saved_size(save_mask)
    only = (torch.rand(B, device="cuda") < 0.4).view(torch.uint8) if masked else None
    if only is not None:
        with engine.only_rows(only):
            ys0, _, st0 = engine.solve_ensemble(model, case["y0"], prm, case["contact"], opts, ts, save_mask, B=B)
    else:
        ys0, _, st0 = engine.solve_ensemble(model, case["y0"], prm, case["contact"], opts, ts, save_mask, B=B)

    L = _driver(str(tmp_path))
    y0 = _dev(torch, case["y0"])
    rates = {k: _dev(torch, prm.get(k)) for k in ("beta", "gamma", "sigma", "omega")}
    season = None
    if model.flags & _lib.FLAG_SEASONAL:
        season = _dev(torch, np.hstack([prm["season_amp"], prm["season_phase"], prm["season_period"]]))
    contact = _dev(torch, case["contact"])
    tsd, jd = _dev(torch, ts), (_dev(torch, np.asarray(jumps)) if jumps else None)
    ys1 = torch.zeros((B, len(ts), ns), dtype=torch.float64, device="cuda")
    st1 = torch.zeros((B, 4), dtype=torch.int32, device="cuda")
    assert y0.numel() in (n, B * n)  # one shared row or [B][n]: the shim tells them apart by the leading dimension
    c = ctypes
    L.shim_driver_solve.argtypes = ([c.c_void_p] + [c.c_int64] * 7 + [c.c_void_p] * 9 + [c.c_int64, c.c_void_p,
                                    c.c_int32, c.c_int32, c.c_int64] + [c.c_double] * 4 + [c.c_int64, c.c_double,
                                    c.c_void_p, c.c_void_p])
    rc = L.shim_driver_solve(_lib.current_stream_ptr(), B, y0.numel() // n, n, S, G, len(ts), ns, _ptr(y0), _ptr(rates["beta"]),
                             _ptr(rates["gamma"]), _ptr(rates["sigma"]), _ptr(rates["omega"]), _ptr(season),
                             _ptr(contact), _ptr(tsd), _ptr(jd), len(jumps), _ptr(only), model.flow, model.flags,
                             save_mask, float(t1), opts.rtol, opts.atol, const_dt, opts.max_steps,
                             engine.uniform_save_dt(ts, 0.0, float(t1)), ys1.data_ptr(), st1.data_ptr())
    assert rc == 0, L.shim_driver_last_error().decode()
    torch.cuda.synchronize()
    assert torch.equal(st1, st0)
    assert torch.equal(ys1, ys0)
    assert int((st0[:, 0] != 0).sum()) == 0 and bool(ys0.abs().sum() > 0)


@pytest.mark.gpu
def test_handler_body_of_the_loglik_gradient_matches_the_ctypes_binding(tmp_path):
    import ctypes

    import numpy as np
    import torch

    from dynode_b200 import _lib, engine
    from dynode_b200.synthetic import make_case

    B, t1 = 257, 100
    case = make_case("sir_age4", B)
    model, prm = case["model"], case["params"]
    G, S, n = model.n_groups, model.n_strains, model.state_size
    ts = np.linspace(0.0, t1, t1 + 1)
    obs_comp = model.n_compartments - 1
    m = model.compartment_sizes()[obs_comp]
    obs = torch.rand(t1, m, dtype=torch.float64, device="cuda") + 0.1
    wrt = [_lib.wrt_id(_lib.P_BETA, 0), _lib.wrt_id(_lib.P_GAMMA, 0)]
    opts = engine.SolverOptions(t1=t1)
    lp0, g0, st0 = engine.poisson_loglik_grad(model, case["y0"], prm, case["contact"], opts, ts, obs_comp, obs, 2.5,
                                              wrt=wrt, B=B)
    L = _driver(str(tmp_path))
    y0 = _dev(torch, case["y0"])
    assert y0.numel() == n  # shared initial state: one row, batch stride 0
    beta, gamma = _dev(torch, prm["beta"]), _dev(torch, prm["gamma"])
    contact, tsd = _dev(torch, case["contact"]), _dev(torch, ts)
    lp1 = torch.zeros(B, dtype=torch.float64, device="cuda")
    g1 = torch.zeros((B, 2), dtype=torch.float64, device="cuda")
    st1 = torch.zeros((B, 4), dtype=torch.int32, device="cuda")
    w = (ctypes.c_int32 * 2)(*wrt)
    c = ctypes
    L.shim_driver_loglik_grad.argtypes = ([c.c_void_p] + [c.c_int64] * 7 + [c.c_void_p] * 9 + [c.c_int64, c.c_int32,
                                          c.c_int32, c.c_int32] + [c.c_double] * 4 + [c.c_int64, c.c_double] +
                                          [c.c_void_p] * 3)
    rc = L.shim_driver_loglik_grad(_lib.current_stream_ptr(), B, 1, n, S, G, len(ts), m, y0.data_ptr(), beta.data_ptr(),
                                   gamma.data_ptr(), None, None, contact.data_ptr(), tsd.data_ptr(), obs.data_ptr(),
                                   ctypes.cast(w, c.c_void_p), 2, model.flow, model.flags, obs_comp, 2.5, float(t1),
                                   opts.rtol, opts.atol, opts.max_steps, engine.uniform_save_dt(ts, 0.0, float(t1)),
                                   lp1.data_ptr(), g1.data_ptr(), st1.data_ptr())
    assert rc == 0, L.shim_driver_last_error().decode()
    torch.cuda.synchronize()
    assert torch.equal(st1, st0) and torch.equal(lp1, lp0) and torch.equal(g1, g0)


@pytest.mark.gpu
@pytest.mark.parametrize("A,K,W,V,NK,kw,jumps,save_mask", [
    (3, 2, 3, 3, 2, {}, (20.0, 41.5), 0b1001),                       # every optional buffer present
    (2, 3, 4, 2, 0, dict(intro=False), (), 0b1111),                  # no knots, no introductions
    (4, 2, 3, 1, 0, dict(season=False, intro=False, vaccinate=False), (), 0b1111),  # the plain family
])
def test_handler_body_of_the_immune_history_solve_matches_the_ctypes_binding(tmp_path, A, K, W, V, NK, kw, jumps,
                                                                             save_mask):
    import ctypes

    import numpy as np
    import torch

    from dynode_b200 import _lib, engine, seip
    from tests.cases import make_seipv_case

    B, t1 = 9, 60
    case = make_seipv_case(B, A=A, K=K, W=W, V=V, NK=NK, t1=t1, **kw)
    model = case["model"]
    ts = np.linspace(0.0, t1, t1 + 1)
    opts = engine.SolverOptions(t1=float(t1), jump_ts=jumps)
    ys0, st0 = seip.solve_ensemble(model, case["y0"], case["params"], case["contact"], case["pop"], case["immunity"],
                                   opts, ts, vaccination=case["vaccination"], introductions=case["introductions"],
                                   season_tau=case["season_tau"], save_mask=save_mask)
    L = _driver(str(tmp_path))
    d = lambda x: None if x is None else torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64).cuda()
    prm, vac, intro = case["params"], case["vaccination"], case["introductions"]
    bufs = dict(y0=d(case["y0"]), beta=d(prm["beta"]), sigma=d(prm["sigma"]), gamma=d(prm["gamma"]),
                omega=d(prm["omega"]), contact=d(case["contact"]), pop=d(case["pop"]), imm=d(case["immunity"]),
                vbase=d(vac[0]) if vac else None, vknots=d(vac[1]) if vac and NK else None,
                vcoef=d(vac[2]) if vac and NK else None,
                itime=d(intro["time"]) if intro else None, iscale=d(intro["scale"]) if intro else None,
                ipct=d(intro["pct"]) if intro else None, iages=d(intro["ages"]) if intro else None,
                ts=d(ts), jumps=d(np.asarray(jumps)) if jumps else None)
    ns = ys0.shape[2]
    ys1 = torch.zeros((B, len(ts), ns), dtype=torch.float64, device="cuda")
    st1 = torch.zeros((B, 4), dtype=torch.int32, device="cuda")
    c = ctypes
    L.shim_driver_seip.argtypes = ([c.c_void_p] + [c.c_int64] * 9 + [c.c_void_p] * 17 + [c.c_int64, c.c_int64] +
                                   [c.c_double] * 6 + [c.c_int64, c.c_double, c.c_void_p, c.c_void_p])
    tau = case["season_tau"]
    rc = L.shim_driver_seip(_lib.current_stream_ptr(), B, model.state_size, A, K, W, V, NK, len(ts), ns,
                            *[_ptr(bufs[k]) for k in ("y0", "beta", "sigma", "gamma", "omega", "contact", "pop", "imm",
                                                      "vbase", "vknots", "vcoef", "itime", "iscale", "ipct", "iages",
                                                      "ts", "jumps")],
                            len(jumps), save_mask, 0.0 if tau is None else float(tau), 0.0 if tau is None else 1.0,
                            float(t1), opts.rtol, opts.atol, 0.0, opts.max_steps,
                            engine.uniform_save_dt(ts, 0.0, float(t1)), ys1.data_ptr(), st1.data_ptr())
    assert rc == 0, L.shim_driver_last_error().decode()
    torch.cuda.synchronize()
    assert torch.equal(st1, st0) and torch.equal(ys1, ys0)
    assert int((st0[:, 0] != 0).sum()) == 0 and bool(ys0.abs().sum() > 0)
