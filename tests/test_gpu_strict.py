"""The STRICT build of the library (libeulerb200_strict.so: csrc/strict_face.cuh, -DEB_STRICT -fmad=false;
SURVEY.md 8(c) "strict build: same operation order, IEEE div/sqrt") reproduces the oracle -- and through
it the unmodified reference, tests/test_oracle.py -- BIT FOR BIT on the B200: every boundary-condition
type, thin grids, many species, tile / z-segment seams, the split launches.  What the fast build differs
by (<= 1e-12 normwise, test_gpu_parity.py) is therefore rounding of its re-derived arithmetic only.

The library is chosen per process (EULERB200_LIB), so the strict runs happen in a child process."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P, N, D, R = 0, 1, 2, 3

CHILD = r"""
import json, sys
import numpy as np
sys.path.insert(0, %(root)r)
sys.path.insert(0, %(root)r + "/tests")
import torch
import oracle
from __graft_entry__ import load_package
from helpers import gpu_feuler, make_udata, oracle_feuler
pkg = load_package()
port = oracle.Port()
P, N, D, R = 0, 1, 2, 3
cases = [((16, 12, 10), 0, [P] * 6), ((16, 12, 10), 2, [N] * 6), ((16, 12, 10), 3, [R] * 6),
         ((40, 9, 11), 0, [P, P, R, R, N, N]), ((33, 17, 9), 4, [N, N, P, P, R, R]), ((64, 20, 24), 10, [R] * 6),
         ((70, 34, 40), 2, [P] * 6), ((3, 40, 36), 6, [N] * 6), ((200, 3, 3), 0, [N] * 6),
         ((40, 26, 12), 24, [R, R, P, P, N, N])]
out = []
for n, nchem, bcs in cases:
    u = make_udata(pkg, n, nchem, bcs, forcing=[0, 0.25, -0.1, 0, 0.5])
    parts = oracle.random_state(n, nchem, seed=sum(n) + nchem)
    ret, got = gpu_feuler(pkg, u, parts)
    ret_ref, ref, _ = oracle_feuler(port, u, parts)
    same = all((a is None and b is None) or np.array_equal(a, b) for a, b in zip(got, ref))
    worst = max(float(np.abs(a - b).max()) for a, b in zip(got, ref) if a is not None)
    out.append(dict(n=n, nchem=nchem, ret=ret, ret_ref=ret_ref, identical=bool(same), max_abs_diff=worst, err=u.last_error() if ret else ""))
    u.FreeData()
print("RESULT " + json.dumps(out))
"""


@pytest.mark.parametrize("split", ["0", "1"])
def test_strict_build_is_bit_identical_to_the_oracle(pkg, split):
    lib = os.path.join(ROOT, "sundials-manyvector-demo_b200", "libeulerb200_strict.so")
    assert os.path.exists(lib), "strict library not built (build.py)"
    env = dict(os.environ, EULERB200_LIB=lib, EULERB200_SPLIT=split)
    res = subprocess.run([sys.executable, "-c", CHILD % dict(root=ROOT)], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("RESULT ")][-1]
    for case in json.loads(line[len("RESULT "):]):
        assert case["ret"] == 0 and case["ret_ref"] == 0, (case["err"], case["n"])
        assert case["identical"], case
