"""Parity of the CUDA fluid RHS (through the C ABI) with the CPU oracle.

Tolerance (BASELINE.json north_star): per sub-vector
max|gpu - ref| / max|ref| <= 1e-12 in FP64.
"""
import numpy as np
import pytest

from conftest import normwise_errors, rounding_floor
from helpers import gpu_feuler, make_udata, oracle_feuler

pytestmark = pytest.mark.gpu
TOL = 1e-12

P, N, D, R = 0, 1, 2, 3   # BC codes (euler3D.hpp:90-93)

CASES = [
    # (n, nchem, bcs)
    ((16, 12, 10), 0, [P] * 6),
    ((16, 12, 10), 2, [N] * 6),
    ((16, 12, 10), 2, [R] * 6),
    ((40, 9, 11), 0, [P, P, R, R, N, N]),
    ((33, 17, 9), 4, [N, N, P, P, R, R]),
    ((64, 20, 24), 10, [R] * 6),
    ((70, 34, 40), 0, [P] * 6),           # several tiles and interior (fast-path) CTAs
    ((3, 40, 36), 0, [N] * 6),            # thin x: hurricane_yz shape
    ((3, 40, 36), 6, [N] * 6),
    ((200, 3, 3), 0, [N] * 6),            # sod_x shape
    ((3, 3, 50), 2, [P, P, P, P, N, N]),
    ((5, 4, 3), 2, [P] * 6),
]


@pytest.mark.parametrize("n,nchem,bcs", CASES)
def test_feuler_matches_oracle(pkg, oracle_mod, port, n, nchem, bcs):
    u = make_udata(pkg, n, nchem, bcs, forcing=[0, 0, -0.1, 0, 0])
    parts = oracle_mod.random_state(n, nchem, seed=sum(n) + nchem)
    ret, got = gpu_feuler(pkg, u, parts)
    ret_ref, ref, _ = oracle_feuler(port, u, parts)
    assert ret == 0 and ret_ref == 0, u.last_error()
    errs = normwise_errors(got, ref, rounding_floor(parts, u.gamma, (u.dx, u.dy, u.dz)))
    assert max(errs) <= TOL, errs
    u.FreeData()


def test_host_pointer_path(pkg, oracle_mod, port):
    n, nchem, bcs = (24, 20, 40), 2, [P, P, R, R, P, P]
    u = make_udata(pkg, n, nchem, bcs)
    parts = oracle_mod.random_state(n, nchem, seed=5)
    ret, got = gpu_feuler(pkg, u, parts, host=True)
    ret_ref, ref, _ = oracle_feuler(port, u, parts)
    assert ret == 0 and ret_ref == 0, u.last_error()
    assert max(normwise_errors(got, ref, rounding_floor(parts, u.gamma, (u.dx, u.dy, u.dz)))) <= TOL
    u.FreeData()


def test_illegal_state_returns_minus_one(pkg, oracle_mod, port):
    n = (16, 12, 10)
    for field, bits in ((0, 1 | 4), (4, 2 | 4)):
        u = make_udata(pkg, n, 0, [P] * 6)
        parts = oracle_mod.random_state(n, 0, seed=3)
        parts[field][137] = -abs(parts[field][137])
        ret, _ = gpu_feuler(pkg, u, parts)
        ret_ref, _, mask = oracle_feuler(port, u, parts)
        assert ret == -1 and ret_ref == -1
        assert "flag = %d" % mask in u.last_error()
        u.FreeData()
