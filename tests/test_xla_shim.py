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
