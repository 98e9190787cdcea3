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


def normwise_errors(got, ref):
    """The parity metric (DESIGN.md "Tolerance"): per sub-vector max|got-ref| / scale with
    scale = max|ref| of that sub-vector, except that the three momentum components share one
    scale (the max over the momentum VECTOR): a component whose exact right-hand side is zero
    (e.g. mx in the Rayleigh-Taylor set-up) holds only the rounding residue of cancelling
    pressure fluxes in the reference as well, and has no scale of its own."""
    out = []
    mom = max((float(np.abs(ref[f]).max()) for f in (1, 2, 3)), default=0.0)
    for f, (a, b) in enumerate(zip(got, ref)):
        if b is None:
            continue
        a = np.asarray(a)
        scale = mom if f in (1, 2, 3) else float(np.abs(b).max())
        err = float(np.abs(a - b).max())
        out.append(err / scale if scale > 0 else err)
    return out
