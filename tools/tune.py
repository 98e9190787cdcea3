#!/usr/bin/env python
"""Tuning aid: time the RHS kernel variants (EULERB200_VARIANT) on one GPU.
   python tools/tune.py [--n 256 256 256] [--nchem 10 0] [--variants 0 1 2 3] [--steps 5]"""
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
ap.add_argument("--nchem", type=int, nargs="+", default=[10, 0])
ap.add_argument("--variants", type=int, nargs="+", default=[0, 1, 2, 3])
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--pair", type=int, nargs="+", default=[2], help="EULERB200_PAIR values to time")
args = ap.parse_args()
if os.environ.get("EB_TUNE_BUILD"): build()
pkg = load_package()
for nchem in args.nchem:
    for v, pair in [(v, q) for v in args.variants for q in args.pair]:
        os.environ["EULERB200_VARIANT"] = str(v)
        os.environ["EULERB200_PAIR"] = str(pair)
        u = pkg.EulerData(nchem=nchem)
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
        print("n=%s nchem=%2d variant=%d pair=%d  %8.3f ms  %6.3f Gcell/s  checksum=%.12e %.12e"
              % (args.n, nchem, v, pair, ms, cells / ms / 1e6, float(wdot.sub[0].double().abs().sum()),
                 float(wdot.sub[-1].double().abs().sum())), flush=True)
        u.FreeData()
        del w, wdot
        torch.cuda.empty_cache()
