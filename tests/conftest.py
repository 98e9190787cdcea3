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
    """max|got-ref| / max|ref| per sub-vector (the tolerance definition of DESIGN.md)."""
    out = []
    for a, b in zip(got, ref):
        if b is None:
            continue
        a = np.asarray(a)
        scale = np.abs(b).max()
        out.append(float(np.abs(a - b).max() / scale) if scale > 0 else float(np.abs(a - b).max()))
    return out
