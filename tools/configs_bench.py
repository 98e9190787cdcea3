#!/usr/bin/env python
"""RHS throughput on the grid shapes BASELINE.json names (device-resident state, RHS only).
   python tools/configs_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from __graft_entry__ import build, load_package  # noqa: E402

build()
pkg = load_package()
CASES = [
    ("sod_x", (200, 3, 3), 0),
    ("linear_advection_x", (256, 256, 256), 0),
    ("linear_advection_y", (256, 256, 256), 0),
    ("linear_advection_z", (256, 256, 256), 0),
    ("rayleigh_taylor", (512, 512, 512), 0),
    ("hurricane_yz", (3, 4096, 4096), 0),
    ("hurricane_yz", (3, 4096, 4096), 6),
]
for problem, n, nchem in CASES:
    u = pkg.EulerData(nchem=nchem)
    u.nx, u.ny, u.nz = n
    pkg.problems.configure(problem, u)
    assert u.SetupDecomp(device=0) == 0
    w = pkg.ManyVector.new(u)
    pkg.problems.initial_conditions(problem, 0.0, w, u)
    wdot = pkg.ManyVector.new(u)
    for _ in range(3):
        assert pkg.fEuler(0.0, w, wdot, u) == 0, u.last_error()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        pkg.fEuler(0.0, w, wdot, u, sync=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    cells = n[0] * n[1] * n[2]
    print("%-20s %-18s NVAR=%2d  %9.3f ms/RHS  %7.3f Gcell/s" % (problem, "x".join(map(str, n)), 5 + nchem, ms, cells / ms / 1e6), flush=True)
    u.FreeData()
    del w, wdot
    torch.cuda.empty_cache()
