"""In-tree build of libdynode_b200.so (hand-written sm_100a CUDA + the C ABI).

nvcc cross-compiles without a GPU; one translation unit per flow-family instance
(csrc/instances.def) so the instances build in parallel.  The .so is git-ignored but travels to
the GPU box with the repo snapshot.
"""

from __future__ import annotations

import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libdynode_b200.so")
INCLUDE = os.path.join(HERE, "..", "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def instance_ids():
    txt = open(os.path.join(CSRC, "instances.def")).read()
    return [int(m) for m in re.findall(r"^X\((\d+),", txt, flags=re.M)]


def _sources():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f != "xla_ffi_shim.cc"]
    deps += [os.path.join(INCLUDE, "dynode_b200.h"), os.path.join(INCLUDE, "dynode_b200_nuts.h"),
             os.path.join(INCLUDE, "dynode_b200_seip.h"), os.path.join(INCLUDE, "dynode_b200_ppl.h"),
             os.path.join(INCLUDE, "dynode_b200_host.h")]
    return deps


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _sources())


def build(force: bool = False, verbose: bool = False, ptxas_v: bool = False, extra_flags=(),
          out: str = LIB, build_dir: str = BUILD) -> str:
    """Compile and link.  `extra_flags`/`out`/`build_dir` exist for tuning experiments (variant
    libraries selected at run time with DYNODE_B200_LIB)."""
    if not force and out == LIB and not needs_build():
        return LIB
    BUILD_ = build_dir
    os.makedirs(BUILD_, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    flags = NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if ptxas_v else [])
    for k in instance_ids():
        obj = os.path.join(BUILD_, f"inst_{k}.o")
        jobs.append((obj, [nvcc, *flags, f"-DDYN_INST={k}", "-c", os.path.join(CSRC, "inst.cu"), "-o", obj]))
    for unit in ("capi", "nuts_round", "seip_solver", "ppl_kernels", "host_buffers"):
        obj = os.path.join(BUILD_, f"{unit}.o")
        jobs.append((obj, [nvcc, *flags, "-c", os.path.join(CSRC, f"{unit}.cu"), "-o", obj]))
    # the immune-history kernels with the vaccination terms: same source, one more translation unit per
    # elements-per-thread width (parallel build)
    for ept in (4, 8, 12):
        obj = os.path.join(BUILD_, f"seip_solver_ext{ept}.o")
        jobs.append((obj, [nvcc, *flags, "-DSEIP_EXT_UNIT=1", f"-DSEIP_EPT={ept}", "-c",
                           os.path.join(CSRC, "seip_solver.cu"), "-o", obj]))
    jobs.sort(key=lambda j: 0 if "seip_solver" in j[0] else 1)  # the longest units first

    def run(job):
        obj, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(run, jobs))
    if verbose or ptxas_v:
        for obj, log in results:
            if log.strip():
                print(f"--- {os.path.basename(obj)}\n{log}", file=sys.stderr)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + [o for o, _ in results]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return out


XLA_LIB = os.path.join(HERE, "libdynode_b200_xla.so")
XLA_HANDLERS = ("DynodeSolve", "DynodeSolveSens", "DynodePoissonLoglikGrad", "DynodePoissonLoglikAdjoint",
                "DynodeSeipSolve", "DynodeSiteLogdensity", "DynodeSiteLogdensityVjp")


def _cuda_include() -> str:
    return os.path.join(os.path.dirname(os.path.dirname(os.path.realpath(_nvcc()))), "include") \
        if os.path.isabs(_nvcc()) else "/usr/local/cuda/include"


def build_xla_shim(include_dir: str = None, out: str = XLA_LIB, source: str = None) -> str:
    """Compile csrc/xla_ffi_shim.cc (typed XLA-FFI handlers over the C ABI) into libdynode_b200_xla.so.

    `include_dir` defaults to jaxlib's FFI headers (`jax.ffi.include_dir()`): raises ImportError in images without
    jax (this one).  tests/test_xla_shim.py passes tests/mock_xla instead -- a model of the header subset the shim
    uses that checks every handler signature against its binding -- so the file meets a compiler here too."""
    if include_dir is None:
        import jax.ffi  # noqa: F401  (absent in this image)

        include_dir = jax.ffi.include_dir()
    build()
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unused-function", "-I", include_dir,
           "-I", _cuda_include(), source or os.path.join(CSRC, "xla_ffi_shim.cc"), "-o", out, "-L", HERE,
           "-ldynode_b200", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + HERE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"xla shim build failed:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, ptxas_v="--ptxas-v" in sys.argv))
