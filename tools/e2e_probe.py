#!/usr/bin/env python
"""Tuning aid: host-pointer path (eulerb200_rhs_host) timing vs number of z-slabs."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from __graft_entry__ import build, load_package
import bench
build(); pkg = load_package()
n, nchem = (512, 512, 512), 10
for slabs in (8, 16, 32, 64):
    os.environ["EULERB200_HOST_SLABS"] = str(slabs)
    u = pkg.EulerData(nchem=nchem); u.nx, u.ny, u.nz = n
    u.xlbc = u.xrbc = u.ylbc = u.yrbc = u.zlbc = u.zrbc = pkg.BC_REFLECTING; u.gamma = 5.0 / 3.0
    assert u.SetupDecomp(device=0) == 0
    w = pkg.ManyVector(bench.synth_state(torch, u, 1234, u.gamma))
    hw = pkg.ManyVector([torch.empty(s.shape, dtype=torch.float64, pin_memory=True) for s in w.sub])
    for h, d in zip(hw.sub, w.sub): h.copy_(d)
    del w; torch.cuda.empty_cache()
    hwdot = pkg.ManyVector([torch.empty(s.shape, dtype=torch.float64, pin_memory=True) for s in hw.sub])
    assert pkg.fEuler(0.0, hw, hwdot, u) == 0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): assert pkg.fEuler(0.0, hw, hwdot, u) == 0
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print("slabs=%2d  %7.1f ms  %.3f Gcell/s  (%.1f GB/s each way)" % (slabs, dt * 1e3, 512**3 / dt / 1e9, 16.106 / dt), flush=True)
    u.FreeData(); del hw, hwdot; torch.cuda.empty_cache()
