"""Build libeulerb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libeulerb200.so")
SOURCES = ["eulerb200.cu"]
HEADERS = ["rhs_kernel.cuh", "strict_face.cuh", "halo_kernels.cuh", "vector_kernels.cuh", "euler_math.cuh", "host_setup.h", os.path.join("..", "..", "include", "eulerb200.h")]
# the STRICT build (csrc/strict_face.cuh): reference operation order, no FMA contraction, IEEE division --
# bit-identical to the reference; a verification artefact (tests/test_gpu_strict.py), never the default
LIB_STRICT = os.path.join(HERE, "libeulerb200_strict.so")
STRICT_FLAGS = ["-DEB_STRICT", "-DEB_FAST_BUILD", "-fmad=false"]

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def needs_build(lib=None):
    lib = lib or LIB
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    jobs = []
    for lib, extra in ((LIB, []), (LIB_STRICT, STRICT_FLAGS)):
        if force or needs_build(lib):
            cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
                  ["-o", lib] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
            jobs.append((lib, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for lib, proc in jobs:          # the two builds run side by side
        out, _ = proc.communicate()
        if proc.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed building %s" % os.path.basename(lib))
        if verbose:
            sys.stderr.write(out)
    return LIB


DRIVER = os.path.join(HERE, "euler3d_b200")


def build_driver(force=False):
    """The native explicit driver (host/euler3d_b200.cpp): plain C++ on top of the C ABI."""
    src = os.path.join(HERE, "host", "euler3d_b200.cpp")
    deps = [src, os.path.join(HERE, "host", "problems.hpp"), os.path.join(HERE, "host", "erk_tables.hpp"), os.path.join(HERE, "host", "erk_stepper.hpp"), os.path.join(HERE, "..", "include", "eulerb200.h"), LIB]
    if not force and os.path.exists(DRIVER) and os.path.getmtime(DRIVER) > max(os.path.getmtime(d) for d in deps):
        return DRIVER
    cmd = ["g++", "-std=c++14", "-O2", "-I", os.path.join(HERE, "..", "include"), "-o", DRIVER, src,
           "-L", HERE, "-leulerb200", "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building euler3d_b200")
    return DRIVER


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
