"""Shared fixtures.  Tests marked ``gpu`` need a B200 (run with ``-m gpu``); everything
else runs on CPU only (``-m "not gpu"``)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (the parity tests proper)")


@pytest.fixture(scope="session")
def pkg():
    entry.build()
    return entry.load_package()


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build_port()
    return oracle


@pytest.fixture(scope="session")
def port(oracle_mod):
    return oracle_mod.Port()


@pytest.fixture(scope="session")
def native_emu_exe(tmp_path_factory):
    """The native driver (host/euler3d_b200.cpp) linked against the CPU emulation of the kernel
    source instead of libeulerb200.so (tests/emu/emu_abi.cpp): test infrastructure for the CPU tier."""
    import subprocess
    here = os.path.join(ROOT, "tests")
    d = tmp_path_factory.mktemp("native_emu")
    objs = []
    for src, flags in (("emu_rhs.cpp", ["-O1", "-ffp-contract=off"]), ("emu_abi.cpp", ["-O1"])):
        obj = str(d / (src + ".o"))
        subprocess.check_call(["g++", "-std=c++14", "-w", "-c"] + flags + ["-o", obj, os.path.join(here, "emu", src)])
        objs.append(obj)
    out = str(d / "euler3d_emu")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-I", os.path.join(ROOT, "include"), "-o", out,
                           os.path.join(ROOT, "sundials-manyvector-demo_b200", "host", "euler3d_b200.cpp")] + objs)
    return out


EPS = 2.220446049250313e-16
FLOOR_CAP = 1e-13      # largest share of a sub-vector's scale that rounding_floor() may forgive (see normwise_errors)


def rounding_floor(parts, gamma, d, ulps=32.0):
    """Absolute rounding floor of the RHS per sub-vector: `ulps` units in the last place of the
    terms the divergence differences, S_f = max|F_f| (1/dx + 1/dy + 1/dz) with the split flux
    bounded by |F_f| <= lambda max|w_f| (+ max|p| for momentum / inside et+p for energy),
    lambda = max(|v|_inf + c).  wdot is a DIFFERENCE of such terms: where the flow is smooth and
    well resolved max|wdot| << S_f, and even the reference recompiled with FMA contraction moves
    by a few ulps of S_f (2e-15 normwise on rough data, 1.6e-10 elementwise)."""
    rho, mx, my, mz, et = [np.asarray(x, dtype=np.float64) for x in parts[:5]]
    p = (gamma - 1.0) * (et - 0.5 * (mx * mx + my * my + mz * mz) / rho)
    with np.errstate(invalid="ignore"):
        c = np.sqrt(np.maximum(gamma * p / rho, 0.0))
    lam = float(np.nanmax(np.maximum(np.maximum(np.abs(mx), np.abs(my)), np.abs(mz)) / np.abs(rho) + c))
    inv = 1.0 / d[0] + 1.0 / d[1] + 1.0 / d[2]
    pm = float(np.abs(p).max())
    mm = float(max(np.abs(mx).max(), np.abs(my).max(), np.abs(mz).max()))
    S = [np.abs(rho).max() * lam, mm * lam + pm, mm * lam + pm, mm * lam + pm, (np.abs(et).max() + pm) * lam]
    if len(parts) > 5 and parts[5] is not None:
        S.append(float(np.abs(parts[5]).max()) * lam)
    return [ulps * EPS * float(x) * inv for x in S]


@pytest.fixture(scope="session")
def port_fma(oracle_mod):
    return oracle_mod.Port(fma=True)


def self_noise(port, port_fma, cfg, parts):
    """Normwise distance between the oracle and the SAME source compiled with FMA contraction,
    per sub-vector: the rounding-noise level of this particular state.  For O(1) states it is
    ~2e-15; where the characteristic transform is ill-conditioned (c^2 << 1 in code units, as in
    the primordial_blast set-up with c^2 ~ 1e-9) it reaches 5e-11, and no implementation other
    than a bit-identical one can be closer to the reference than that."""
    _, a, _ = port.feuler(cfg, parts)
    _, b, _ = port_fma.feuler(cfg, parts)
    return normwise_errors(b, a)


def normwise_errors(got, ref, floor=None):
    """The parity metric (DESIGN.md "Tolerance"): per sub-vector
        max(0, max|got-ref| - min(floor_f, 1e-13 scale_f)) / scale_f ,   required <= 1e-12,
    scale_f = max|ref_f| (the three momentum components share the scale of the momentum VECTOR:
    a component whose exact right-hand side is zero, e.g. mx in Rayleigh-Taylor, holds only
    rounding residue in the reference too), floor_f = rounding_floor() or 0."""
    out = []
    mom = max((float(np.abs(ref[f]).max()) for f in (1, 2, 3)), default=0.0)
    k = 0
    for f, (a, b) in enumerate(zip(got, ref)):
        if b is None:
            continue
        a = np.asarray(a)
        scale = mom if f in (1, 2, 3) else float(np.abs(b).max())
        err = float(np.abs(a - b).max())
        if floor is not None:
            # the floor never moves the bar by more than a tenth of it: at most 1e-13 of the scale is forgiven
            # (where the state is so smooth that rounding of the flux terms exceeds that, a test has to use the
            # reference's own self-noise as its bar, as the blast-state tests do -- not a bigger floor)
            err = max(0.0, err - min(floor[k], FLOOR_CAP * scale))
        out.append(err / scale if scale > 0 else err)
        k += 1
    return out
