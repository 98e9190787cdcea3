#!/usr/bin/env python
"""Tuning aid: time combinations of EULERB200_* environment settings on one GPU.
   python tools/tune2.py --n 256 256 256 --nchem 10 --env "PAIR=2 KERNEL=1" "PAIR=1" "NO_AUX=1" """
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from __graft_entry__ import build, load_package  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, nargs=3, default=[256, 256, 256])
ap.add_argument("--nchem", type=int, default=10)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--env", nargs="+", default=[""])
args = ap.parse_args()
build()
pkg = load_package()
for envs in args.env:
    for k in list(os.environ):
        if k.startswith("EULERB200_"):
            del os.environ[k]
    for kv in envs.split():
        k, v = kv.split("=")
        os.environ["EULERB200_" + k] = v
    u = pkg.EulerData(nchem=args.nchem)
    u.nx, u.ny, u.nz = args.n
    u.xlbc = u.xrbc = u.ylbc = u.yrbc = u.zlbc = u.zrbc = pkg.BC_REFLECTING
    u.gamma = 5.0 / 3.0
    assert u.SetupDecomp(device=0) == 0
    w = pkg.ManyVector(bench.synth_state(torch, u, 1234, u.gamma))
    wdot = pkg.ManyVector.new(u)
    for _ in range(3):
        assert pkg.fEuler(0.0, w, wdot, u) == 0, u.last_error()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        pkg.fEuler(0.0, w, wdot, u, sync=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    cells = u.nx * u.ny * u.nz
    chk = sum(float(s.abs().sum()) for s in wdot.sub)
    print("n=%s nchem=%2d [%-32s] %8.3f ms  %6.3f Gcell/s  checksum=%.12e"
          % (args.n, args.nchem, envs, ms, cells / ms / 1e6, chk), flush=True)
    u.FreeData()
    del w, wdot
    torch.cuda.empty_cache()
