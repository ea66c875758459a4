"""The C-ABI shared library loads on a CPU-only box and exports every entry point that
include/dynode_b200.h declares; descriptor helpers and argument validation work without a GPU
(no compute call is made here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADERS = [os.path.join(ROOT, "include", f) for f in sorted(os.listdir(os.path.join(ROOT, "include")))
           if f.endswith(".h")]


def _declared_functions():
    names = []
    for h in HEADERS:
        txt = open(h).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        txt = re.sub(r"//[^\n]*", "", txt)
        for m in re.finditer(r"\b(dynode_[a-z0-9_]+)\s*\(", txt):
            names.append(m.group(1))
    return sorted(set(names))


@pytest.fixture(scope="module")
def lib():
    from dynode_b200 import _lib
    return _lib.load()


def test_header_declares_what_python_binds():
    from dynode_b200 import _lib
    assert set(_declared_functions()) == set(_lib.EXPORTED_SYMBOLS)


@pytest.mark.parametrize("sym", _declared_functions())
def test_symbol_is_exported(lib, sym):
    assert getattr(lib, sym) is not None


def test_version_and_descriptor_helpers(lib):
    from dynode_b200 import _lib
    assert lib.dynode_version() == 100
    for flow, ncomp in ((_lib.FLOW_SIR, 3), (_lib.FLOW_SEIRS, 4), (_lib.FLOW_SEIRS_C, 5)):
        d = _lib.ModelDesc(flow, 0, 2, 3)
        assert lib.dynode_num_compartments(ctypes.byref(d)) == ncomp
        assert lib.dynode_state_size(ctypes.byref(d)) == 2 + (ncomp - 1) * 6
        assert lib.dynode_saved_size(ctypes.byref(d), 0b1) == 2
        assert lib.dynode_saved_size(ctypes.byref(d), 0b110) == 12
        assert lib.dynode_saved_size(ctypes.byref(d), (1 << ncomp) - 1) == 2 + (ncomp - 1) * 6
    bad = _lib.ModelDesc(17, 0, 1, 1)
    assert lib.dynode_state_size(ctypes.byref(bad)) == -1


def test_every_compiled_instance_is_reported_supported(lib):
    from dynode_b200 import _lib
    txt = open(os.path.join(ROOT, "dynode_b200", "csrc", "instances.def")).read()
    rows = re.findall(r"^X\((\d+),\s*(\w+),\s*([\w| ]+),\s*(\d+),\s*(\d+)\)", txt, flags=re.M)
    assert len(rows) >= 9
    flows = {"DYNODE_FLOW_SIR": 0, "DYNODE_FLOW_SEIRS": 1, "DYNODE_FLOW_SEIRS_C": 2}
    flags = {"0": 0, "DYNODE_FLAG_SEASONAL": 1, "DYNODE_FLAG_DENSITY_DEP": 2}
    for _, flow, flag, g, s in rows:
        d = _lib.ModelDesc(flows[flow], flags[flag.strip()], int(g), int(s))
        assert lib.dynode_is_supported(ctypes.byref(d)) == 1, (flow, flag, g, s)


def test_unsupported_model_fails_loudly_without_touching_the_gpu(lib):
    """No CPU fallback: a model outside instances.def is rejected by the launch entry points."""
    from dynode_b200 import _lib
    d = _lib.ModelDesc(_lib.FLOW_SEIRS_C, 0, 5, 7)
    assert lib.dynode_is_supported(ctypes.byref(d)) == 0
    sv = _lib.SolverDesc(0.0, 10.0, 1e-5, 1e-6, 0.0, 100, 0.0, None, 0)
    buf = (ctypes.c_double * 8)()
    ptr = ctypes.cast(buf, ctypes.c_void_p)
    arr = _lib.Array(ptr.value, 0)
    prm = _lib.Params()
    prm.beta = prm.gamma = prm.sigma = arr
    rc = lib.dynode_solve_f64(ctypes.byref(d), ctypes.byref(sv), 1, arr, ctypes.byref(prm), ptr, 2, 0b11111,
                              ptr, ptr, None)
    assert rc != 0
    msg = lib.dynode_last_error().decode()
    assert "unsupported ODE" in msg and "no CPU fallback" in msg


@pytest.mark.parametrize("mutate,needle", [
    (lambda sv, prm: setattr(sv, "t1", -1.0), "t1 must be >= t0"),
    (lambda sv, prm: setattr(sv, "rtol", 0.0), "rtol/atol"),
    (lambda sv, prm: setattr(sv, "max_steps", 0), "max_steps"),
    (lambda sv, prm: setattr(prm, "gamma", __import__("dynode_b200")._lib.Array(None, 0)), "beta/gamma"),
    (lambda sv, prm: setattr(sv, "n_jump", 99), "n_jump"),
    (lambda sv, prm: setattr(sv, "n_jump", 2), "jump_ts is null"),
])
def test_argument_validation_messages(lib, mutate, needle):
    from dynode_b200 import _lib
    d = _lib.ModelDesc(_lib.FLOW_SIR, 0, 1, 1)
    sv = _lib.SolverDesc(0.0, 10.0, 1e-5, 1e-6, 0.0, 100, 0.0, None, 0)
    buf = (ctypes.c_double * 8)()
    ptr = ctypes.cast(buf, ctypes.c_void_p)
    arr = _lib.Array(ptr.value, 0)
    prm = _lib.Params()
    prm.beta = prm.gamma = arr
    mutate(sv, prm)
    rc = lib.dynode_solve_f64(ctypes.byref(d), ctypes.byref(sv), 1, arr, ctypes.byref(prm), ptr, 2, 0b111,
                              ptr, ptr, None)
    assert rc != 0 and needle in lib.dynode_last_error().decode()


def test_empty_ensemble_is_a_no_op(lib):
    from dynode_b200 import _lib
    d = _lib.ModelDesc(_lib.FLOW_SIR, 0, 1, 1)
    sv = _lib.SolverDesc(0.0, 10.0, 1e-5, 1e-6, 0.0, 100, 0.0, None, 0)
    buf = (ctypes.c_double * 8)()
    ptr = ctypes.cast(buf, ctypes.c_void_p)
    arr = _lib.Array(ptr.value, 0)
    prm = _lib.Params()
    prm.beta = prm.gamma = arr
    assert lib.dynode_solve_f64(ctypes.byref(d), ctypes.byref(sv), 0, arr, ctypes.byref(prm), ptr, 2, 0b111,
                                ptr, ptr, None) == 0


def test_host_buffers_without_a_gpu(lib):
    """include/dynode_b200_host.h: the mapping itself (2 MiB aligned, huge-page advised, first-touched in parallel)
    needs no device when DYNODE_HOST_NO_PIN is set; registration with CUDA is exercised by the GPU tests / bench."""
    import numpy as np
    from dynode_b200 import hostmem
    buf = hostmem.HostBuffer(3 * (1 << 20) + 17, hugepages=True, pin=False, threads=3)
    assert buf.ptr % (2 << 20) == 0
    a = buf.array((1 << 18,), np.float64)
    a[:] = np.arange(a.size)
    assert a[-1] == a.size - 1 and buf.huge_bytes() >= -1
    t = buf.tensor((4, 8))
    t.fill_(2.5)
    assert float(t.sum()) == 80.0
    del a, t
    buf.free()
    p = ctypes.c_void_p()
    assert lib.dynode_host_alloc(0, 0, 1, ctypes.byref(p)) != 0 and b"zero bytes" in lib.dynode_last_error()


def test_potential_plan_validation(lib):
    from dynode_b200 import _lib
    plan = _lib.PotentialPlan()
    plan.n_sites, plan.n_rates = 0, 1
    assert lib.dynode_potential_pre_f64(ctypes.byref(plan), 4, 1, 1, 1, 1, None, None) != 0
    assert b"n_sites" in lib.dynode_last_error()
    plan.n_sites, plan.n_rates = 2, 2
    for j in range(2):
        plan.site[j] = _lib.SiteDesc(0, 1, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0)
    plan.rate_e[0][0] = 2
    assert lib.dynode_potential_pre_f64(ctypes.byref(plan), 4, 1, 2, 1, 1, None, None) != 0
    assert b"exponents" in lib.dynode_last_error()
    plan.rate_e[0][0] = 1
    assert lib.dynode_potential_pre_f64(ctypes.byref(plan), 4, 1, 1, 1, 1, None, None) != 0  # z_stride < n_sites
    assert lib.dynode_potential_post_f64(ctypes.byref(plan), 4, 1, 1, 1, 1, 2, None, None, None, 0, None, None, None,
                                         1, 1, None) != 0  # missing column map
