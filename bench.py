#!/usr/bin/env python
"""bench.py -- throughput of the fluid right-hand side (fEuler) in Gcell-RHS/s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload primordial_blast|rayleigh_taylor|hurricane_yz|linear_advection_x]
                  [--scaling weak|strong] [--n NX NY NZ] [--nchem C] [--ic random|problem]
                  [--no-e2e] [--no-cpu-baseline] [--no-parity]

One "step" = one complete fEuler evaluation (halo exchange included when N > 1) on a synthetic
admissible state.  Default workload: BASELINE.json's metric configuration, the fluid_blast /
primordial_blast shape -- 512^3 cells per GPU, nchem = 10 (NVAR = 15), unit cube, all-reflecting
boundaries, gamma = 5/3 (tests/primordial_blast/input_*.txt of the reference), weak scaling: every
GPU owns 512^3 cells of a (512 npx, 512 npy, 512 npz) grid decomposed as the reference's SetupDecomp
would.  The other BASELINE.json grid shapes are --workload choices (their own boundary conditions,
forcing and initial condition); --scaling strong keeps the global grid at --n and decomposes it.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      HBM view of the RHS kernel: algorithmic bytes (16*NVAR per cell) / kernel time
  fp64          FP64-pipe view: reference-as-written flops per cell / kernel time
                against the nominal peak and against a DFMA peak measured in this run
  parity        the timed path's wdot against the CPU oracle on windows of the grid (a corner on the
                low and on the high boundary faces, across a z-segment seam of the kernel, the interior
                and, at N > 1, across the seam where up to 8 ranks meet), outside the timed region
  cpu_baseline  the UNMODIFIED reference fEuler (oracle/_ref) on this box's host cores
  e2e           same metric through the host-pointer C-ABI call (pinned host arrays in/out); at N > 1
                the decomposed host path with its halo exchange
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOL = 1e-12          # north_star: max relative error per field (normwise)
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12      # 37.2: 148 SMs x 64 DFMA lanes x 2 x 1.965 GHz


# FP64 operations per cell-RHS executed by the reference as written (BASELINE.md section 2,
# measured with a counting scalar type): NVAR=5 -> 4918, NVAR=15 -> 10414; linear in NVAR.
def ref_flops_per_cell(nvar):
    return 4918.0 + (10414.0 - 4918.0) * (nvar - 5) / 10.0


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


WORKLOADS = {
    # name: (problem plug-in, default cells per GPU, nchem, what the config string says)
    "primordial_blast": ("primordial_blast", (512, 512, 512), 10,
                         "primordial_blast/fluid_blast shape: %dx%dx%d cells per GPU, nchem=%d (NVAR=%d), "
                         "unit cube, all-reflecting, gamma=5/3"),
    "rayleigh_taylor": ("rayleigh_taylor", (512, 512, 512), 0,
                        "rayleigh_taylor shape: %dx%dx%d cells per GPU, nchem=%d (NVAR=%d), x periodic / y reflecting / "
                        "z Neumann, forcing Gmy=-0.1, gamma=1.4"),
    "hurricane_yz": ("hurricane_yz", (3, 4096, 4096), 0,
                     "hurricane_yz shape: %dx%dx%d cells per GPU, nchem=%d (NVAR=%d), all-Neumann, gamma=2"),
    "linear_advection_x": ("linear_advection_x", (256, 256, 256), 0,
                           "linear_advection_x shape: %dx%dx%d cells per GPU, nchem=%d (NVAR=%d), all-periodic, gamma=1.4"),
}


def workload_name(workload, n, nchem):
    return WORKLOADS[workload][3] % (n[0], n[1], n[2], nchem, 5 + nchem)


def dims_create(nnodes, ndims):
    """MPI_Dims_create as every mainstream MPI answers it (balanced, non-increasing) -- restated here
    so that the reference arm needs nothing of the product package."""
    primes, n, p = [], nnodes, 2
    while p * p <= n:
        while n % p == 0:
            primes.append(p)
            n //= p
        p += 1
    if n > 1:
        primes.append(n)
    bins = [1] * ndims
    for q in sorted(primes, reverse=True):
        bins[bins.index(min(bins))] *= q
    return sorted(bins, reverse=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.out, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark(self):
        """Call right before and right after the timed region: only samples in between count."""
        import datetime
        self.marks = getattr(self, "marks", []) + [datetime.datetime.now()]

    def stop(self):
        import datetime
        res = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return res
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
            self.out.close()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            rows = []
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    ts = datetime.datetime.strptime(p[0], "%Y/%m/%d %H:%M:%S.%f")
                    rows.append((ts, float(p[1]), float(p[2]), p[5:9]))
                except ValueError:
                    continue
            os.unlink(self.path)
            marks = getattr(self, "marks", [])
            inside = [r for r in rows if len(marks) >= 2 and marks[0] <= r[0] <= marks[1]]
            use = inside if inside else rows          # the sampler runs from before the warm-up steps
            if use:
                sm = sorted(r[1] for r in use)
                reasons = set()
                for r in use:
                    for nm, v in zip(names, r[3]):
                        if v.lower().startswith("active"):
                            reasons.add(nm)
                res = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(r[2] for r in use),
                       "reasons": sorted(reasons), "samples": len(use),
                       "window": "timed region" if inside else "warm-up + timed region"}
        except Exception:
            pass
        return res


def cpu_reference_run(nvar, steps, warmup, per_rank=40, max_ranks=None):
    """Time the UNMODIFIED reference fEuler (oracle/_ref, compiled from /root/reference by
    oracle/Makefile) on the host cores: P virtual MPI ranks (threads, real halo exchange
    through the shim), one per core, each owning per_rank^3 cells.  Falls back to the C
    port of the oracle (single core) when oracle/_ref is absent.  Uses nothing of the product."""
    import numpy as np
    import oracle
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    nchem = nvar - 5
    bc = [3] * 6
    gamma = 5.0 / 3.0
    if oracle.have_ref(nvar):
        P = cores if max_ranks is None else min(cores, max_ranks)
        R = oracle.Ref(nvar)
        dims = dims_create(P, 3)                 # what the reference's SetupDecomp will choose
        n = tuple(per_rank * d for d in dims)
        w = oracle.random_state(n, nchem, seed=1234, gamma=gamma)
        ret, _, secs, dec = R.feuler(n, [0, 1] * 3, bc, gamma, w, nprocs=P, nrep=warmup + steps)
        assert ret == 0 and sorted(dec, reverse=True) == dims, (ret, dec, dims)
        t = secs[warmup:]
        cells = n[0] * n[1] * n[2]
        return {"kind": "reference", "cores": P, "cells": cells, "sec_per_step": float(np.mean(t)),
                "value": cells / float(np.mean(t)) / 1e9,
                "sample": "unmodified reference fEuler, %d virtual MPI ranks (threads) %dx%dx%d, %d^3 cells/rank, "
                          "NVAR=%d, reflecting, %d timed evals" % (P, dec[0], dec[1], dec[2], per_rank, nvar, steps)}
    port = oracle.Port()
    n = (per_rank,) * 3
    w = oracle.random_state(n, nchem, seed=1234, gamma=gamma)
    cfg = port.cfg(n, nchem, [1.0 / per_rank] * 3, gamma, bc)
    ts = []
    for it in range(warmup + steps):
        t0 = time.time()
        ret, _, _ = port.feuler(cfg, w)
        ts.append(time.time() - t0)
    t = float(np.mean(ts[warmup:]))
    cells = per_rank ** 3
    return {"kind": "port", "cores": 1, "cells": cells, "sec_per_step": t, "value": cells / t / 1e9,
            "sample": "C port of the oracle, 1 core, %d^3 cells, NVAR=%d, %d timed evals" % (per_rank, nvar, steps)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nvar = 5 + args.nchem
    steps = max(1, min(args.steps, 5))
    warmup = max(1, min(args.warmup, 2))
    r = cpu_reference_run(nvar, steps, warmup)
    line = {
        "impl": "reference", "metric": "fEuler cell-RHS evaluations per second", "value": r["value"],
        "unit": "Gcell/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": r["sec_per_step"] * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, args.n, args.nchem),
                   "sample": "each step is a bounded CPU sample of that workload: " + r["sample"]},
        "cpu_baseline": {"value": r["value"], "unit": "Gcell/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "Gcell/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def synth_state(torch, u, seed, gamma):
    """rho=1+0.5U, v=0.3(U-0.5), p=1+0.5U; tracers U * 10^(-8..+6) per species (the wide
    magnitude range of the primordial species is what exercises WENO's epsilon)."""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    N = u.nxl * u.nyl * u.nzl
    U = lambda n: torch.rand(n, generator=g, device="cuda", dtype=torch.float64)
    rho = 1.0 + 0.5 * U(N)
    vx, vy, vz = (0.3 * (U(N) - 0.5) for _ in range(3))
    p = 1.0 + 0.5 * U(N)
    et = p / (gamma - 1.0) + 0.5 * rho * (vx * vx + vy * vy + vz * vz)
    subs = [rho, rho * vx, rho * vy, rho * vz, et]
    del vx, vy, vz, p
    if u.nchem > 0:
        chem = U(N * u.nchem).view(N, u.nchem)
        scale = torch.tensor([10.0 ** (-8 + (14 * v) // max(1, u.nchem - 1)) for v in range(u.nchem)],
                             device="cuda", dtype=torch.float64)
        chem.mul_(scale)
        subs.append(chem.view(-1))
    return subs


# ------------------------------------------------------------------------------------------------
# parity of the timed path against the CPU oracle, on windows of the (global) grid
# ------------------------------------------------------------------------------------------------

def parity_windows(u, size=24):
    """[(name, lo[3], hi[3])] in global cell coordinates."""
    n = (u.nx, u.ny, u.nz)
    nl = (u.nx // u.npx, u.ny // u.npy, u.nz // u.npz)            # block of rank 0
    s = [min(size, n[d]) for d in range(3)]

    def box(c):      # window of s cells per axis around centre c, clipped into the grid
        lo = [max(0, min(n[d] - s[d], c[d] - s[d] // 2)) for d in range(3)]
        return lo, [lo[d] + s[d] for d in range(3)]

    wins = [("corner on the low boundary faces", [0, 0, 0], s),
            ("corner on the high boundary faces", [n[d] - s[d] for d in range(3)], list(n))]
    # the kernel cuts every box into z-segments (host_setup.h launch_geom: 8 of them at 512^3) and
    # 31 x 11-cell tiles: a window across the first z-segment seam of rank 0's box, and across tile seams
    seam_z = max(s[2] // 2, nl[2] // 8)
    wins.append(("across a z-segment seam and tile seams of the kernel",) + box([nl[0] // 3, nl[1] // 3, seam_z]))
    wins.append(("interior",) + box([nl[0] // 2 + 5, nl[1] // 2 + 3, nl[2] // 2 + 1]))
    if u.nprocs > 1:   # where the blocks of up to 8 ranks meet
        c = [nl[d] if (u.npx, u.npy, u.npz)[d] > 1 else nl[d] // 2 for d in range(3)]
        wins.append(("across the seam where %d ranks meet" % (min(2, u.npx) * min(2, u.npy) * min(2, u.npz)),) + box(c))
    return wins


def parity_check(torch, dist, pkg, u, w, wdot, forcing, rank, world, margin=3):
    """Every rank cuts its share of each window (+ margin) out of w and wdot; rank 0 assembles
    the windows, runs the oracle (and its FMA-contracted build: the reference's own rounding noise
    on this state) and compares.  Returns the `parity` object (rank 0) or None.

    Metric, per sub-vector f: err_f = max over the window |gpu - ref| / scale_f with scale_f the
    maximum of |wdot_f| over the WHOLE grid (the three momenta share the scale of the momentum vector):
    north_star's normwise relative error.  self_noise_f: the same distance between the oracle and the
    oracle compiled with FMA contraction -- on smooth or resting states wdot is a difference of O(1/dx)
    larger terms (or exactly zero analytically) and the reference itself moves by that much under a
    different legal compilation; no implementation that is not bit-identical can be expected closer.
    Bar: err_f <= max(1e-12, 2 x self_noise_f) -- two independent roundings of the same value differ
    in their window maxima by a factor of order one, hence the 2; the ratio is reported."""
    import numpy as np
    gmax = torch.stack([s_.abs().max() for s_ in wdot.sub])
    if world > 1:
        dist.all_reduce(gmax, op=dist.ReduceOp.MAX)
    gmax = [float(x) for x in gmax]
    gscale = [max(gmax[1:4]) if f in (1, 2, 3) else gmax[f] for f in range(len(gmax))]
    gscale = [x if x > 0 else 1.0 for x in gscale]
    n = (u.nx, u.ny, u.nz)
    own_lo = (u.is_, u.js, u.ks)
    own_hi = (u.ie + 1, u.je + 1, u.ke + 1)
    nl = (u.nxl, u.nyl, u.nzl)
    nsub = 5 + (1 if u.nchem > 0 else 0)
    wins = parity_windows(u)
    pieces = []
    for name, lo, hi in wins:
        elo = [max(0, lo[d] - margin) for d in range(3)]
        ehi = [min(n[d], hi[d] + margin) for d in range(3)]
        a = [max(elo[d], own_lo[d]) for d in range(3)]
        b = [min(ehi[d], own_hi[d]) for d in range(3)]
        if any(b[d] <= a[d] for d in range(3)):
            pieces.append(None)
            continue
        sl = tuple(slice(a[d] - own_lo[d], b[d] - own_lo[d]) for d in (2, 1, 0))
        cut = []
        for vec in (w, wdot):
            for f in range(nsub):
                shape = (nl[2], nl[1], nl[0]) + ((u.nchem,) if f == 5 else ())
                cut.append(vec.sub[f].view(shape)[sl].contiguous().cpu().numpy())
        pieces.append((a, b, cut))
    if world > 1:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(pieces, gathered, dst=0)
    else:
        gathered = [pieces]
    if rank != 0:
        return None
    import oracle
    port, port_fma = oracle.Port(), oracle.Port(fma=True)
    bcs = u.bcs
    out = {"tolerance": TOL, "metric": "max over the window |gpu-ref| / max over the grid |wdot|, per sub-vector (the "
           "momenta share the scale of the momentum vector); self_noise = the same distance between the oracle and "
           "the oracle compiled with FMA contraction", "oracle": "oracle/euler_oracle.c, pinned bit-for-bit to the "
           "unmodified reference (tests/test_oracle.py)", "margin_cells": margin,
           "scale_per_subvector": gscale, "windows": []}
    worst, worst_ratio = 0.0, 0.0
    for wi, (name, lo, hi) in enumerate(wins):
        elo = [max(0, lo[d] - margin) for d in range(3)]
        ehi = [min(n[d], hi[d] + margin) for d in range(3)]
        en = [ehi[d] - elo[d] for d in range(3)]
        state = [np.full((en[2], en[1], en[0]) + ((u.nchem,) if f == 5 else ()), np.nan) for f in range(nsub)]
        got = [np.full_like(x, np.nan) for x in state]
        ranks = []
        for r in range(world):
            pc = gathered[r][wi]
            if pc is None:
                continue
            ranks.append(r)
            a, b, cut = pc
            sl = tuple(slice(a[d] - elo[d], b[d] - elo[d]) for d in (2, 1, 0))
            for f in range(nsub):
                state[f][sl] = cut[f]
                got[f][sl] = cut[nsub + f]
        assert not any(np.isnan(x).any() for x in state), "window not covered by the ranks"
        # physical boundary condition where the extended window reaches the domain boundary, anything
        # (Neumann) on its artificial faces: cells within `margin` of those are not compared
        wbc = []
        for d in range(3):
            wbc += [bcs[2 * d] if elo[d] == 0 and bcs[2 * d] != pkg.BC_PERIODIC else pkg.BC_NEUMANN,
                    bcs[2 * d + 1] if ehi[d] == n[d] and bcs[2 * d + 1] != pkg.BC_PERIODIC else pkg.BC_NEUMANN]
        parts = [np.ascontiguousarray(x).ravel() for x in state] + ([None] if nsub == 5 else [])
        cfg = port.cfg(en, u.nchem, (u.dx, u.dy, u.dz), u.gamma, wbc, forcing=forcing)
        ret, ref, _ = port.feuler(cfg, parts)
        _, ref_fma, _ = port_fma.feuler(cfg, parts)
        cmp_sl = tuple(slice(lo[d] - elo[d], hi[d] - elo[d]) for d in (2, 1, 0))
        if any(bcs[2 * d] == pkg.BC_PERIODIC and (elo[d] == 0 or ehi[d] == n[d]) for d in range(3)):
            # a periodic face of the grid is an artificial face of the window: keep the margin there
            cmp_sl = tuple(slice(max(lo[d] - elo[d], margin if bcs[2 * d] == pkg.BC_PERIODIC else 0),
                                 min(hi[d] - elo[d], en[d] - (margin if bcs[2 * d] == pkg.BC_PERIODIC else 0)))
                           for d in (2, 1, 0))
        errs, noise = [], []
        for f in range(nsub):
            shape = (en[2], en[1], en[0]) + ((u.nchem,) if f == 5 else ())
            r3 = ref[f].reshape(shape)[cmp_sl]
            errs.append(float(np.abs(got[f][cmp_sl] - r3).max()) / gscale[f])
            noise.append(float(np.abs(ref_fma[f].reshape(shape)[cmp_sl] - r3).max()) / gscale[f])
        ok = ret == 0 and all(e <= max(TOL, 2.0 * nz) for e, nz in zip(errs, noise))
        worst = max(worst, max(errs))
        worst_ratio = max(worst_ratio, max((e / nz if nz > 0 else (0.0 if e == 0 else float("inf")))
                                           for e, nz in zip(errs, noise) if e > TOL) if any(e > TOL for e in errs) else 0.0)
        out["windows"].append({"name": name, "lo": lo, "hi": hi, "ranks": ranks, "oracle_ret": ret,
                               "err_per_subvector": errs, "self_noise_per_subvector": noise,
                               "pass": bool(ok)})
    out["max_err"] = worst
    out["max_err_over_self_noise_where_above_1e-12"] = worst_ratio
    out["pass"] = all(x["pass"] for x in out["windows"])
    out["bar"] = "err <= max(1e-12, 2 x self_noise) per sub-vector"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="primordial_blast", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--n", type=int, nargs=3, default=None, help="cells per GPU (weak) / of the whole grid (strong)")
    ap.add_argument("--nchem", type=int, default=None)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ic", default=None, choices=["random", "problem", "primordial_blast"],
                    help="synthetic state: seeded random admissible state (SURVEY.md 8(d); default of the "
                         "default workload) or the workload's own initial condition (default of the others)")
    args = ap.parse_args()
    problem, n_default, nchem_default, _ = WORKLOADS[args.workload]
    args.n = list(args.n) if args.n else list(n_default)
    args.nchem = nchem_default if args.nchem is None else args.nchem
    if args.ic == "primordial_blast":
        args.ic = "problem"
    if args.ic is None:
        args.ic = "random" if args.workload == "primordial_blast" else "problem"
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from __graft_entry__ import build, load_package
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the fluid RHS has no CPU fallback")
    torch.cuda.set_device(local_rank)
    try:       # pinned host memory of this rank on the NUMA node of its GPU (the e2e leg)
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        numa = "cpu affinity set to the GPU's NUMA node (nvml)"
    except Exception as ex:    # pragma: no cover
        numa = "cpu affinity not set (%s)" % str(ex)[:60]
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        build()
    if world > 1:
        dist.barrier()
    pkg = load_package()

    nvar = 5 + args.nchem
    u = pkg.EulerData(nchem=args.nchem)
    pkg.problems.configure(problem, u)
    gamma, bcs = u.gamma, u.bcs
    # weak scaling: global grid = per-GPU box x process grid of the reference's SetupDecomp;
    # strong scaling: the global grid is --n
    _, dims, _, _, _ = pkg.dims_and_extents(world, 0, tuple(args.n), bcs)
    if args.scaling == "weak":
        u.nx, u.ny, u.nz = (args.n[0] * dims[0], args.n[1] * dims[1], args.n[2] * dims[2])
    else:
        u.nx, u.ny, u.nz = args.n
    assert u.SetupDecomp(myid=rank, nprocs=world, device=local_rank) == 0
    cells_local = u.nxl * u.nyl * u.nzl
    cells_global = u.nx * u.ny * u.nz

    if args.ic == "problem":
        w = pkg.ManyVector.new(u)
        assert pkg.problems.initial_conditions(problem, 0.0, w, u) == 0
    else:
        w = pkg.ManyVector(synth_state(torch, u, 1234 + rank, gamma))
    wdot = pkg.ManyVector.new(u)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                      # nvidia-smi takes a moment to start: begin before the warm-up
    for _ in range(args.warmup):
        ret = pkg.fEuler(0.0, w, wdot, u)
        assert ret == 0, u.last_error()
    launches0 = u.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    sampler.mark()
    ev[0].record()
    for s in range(args.steps):
        ret = pkg.fEuler(0.0, w, wdot, u)
        ev[s + 1].record()
    barrier()
    sampler.mark()
    clocks = sampler.stop() if rank == 0 else None
    assert ret == 0, u.last_error()
    launches = u.launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    tt = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())
    ms_per_step = total_ms / args.steps
    value = cells_global / (ms_per_step * 1e-3) / 1e9

    # parity of what was just timed (the wdot of the last step), outside the timed region
    parity = None
    if not args.no_parity:
        try:
            parity = parity_check(torch, dist, pkg, u, w, wdot, u.forcing, rank, world)
            sums = torch.stack([s_.sum() for s_ in wdot.sub] + [s_.abs().sum() for s_ in wdot.sub])
            if world > 1:
                dist.all_reduce(sums)
            if parity is not None:
                k = len(wdot.sub)
                parity["sum_wdot_per_subvector"] = [float(x) for x in sums[:k]]
                parity["sum_abs_wdot_per_subvector"] = [float(x) for x in sums[k:]]
                parity["conservation_note"] = ("sum(wdot)/sum|wdot| is ~1e-16 on periodic grids (the face fluxes "
                                               "telescope); with physical boundaries the boundary fluxes remain, as in "
                                               "the reference (high-side ghosts are copies, euler3D.hpp:864)")
        except Exception as ex:      # pragma: no cover
            parity = {"error": str(ex)[:300], "pass": False}

    # kernel-only duration (single launch at N=1; at N>1 the interior launch dominates):
    # time the async call alone, events on the launching stream
    kev0, kev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    kev0.record()
    for _ in range(args.steps):
        pkg.fEuler(0.0, w, wdot, u, sync=False)
    kev1.record()
    torch.cuda.synchronize()
    kernel_ms = kev0.elapsed_time(kev1) / args.steps

    # device-time breakdown of one RHS (CUDA events on the streams the phases run on), untimed extra calls
    timeline = None
    try:
        u.profile(on=True, reset=True)
        for _ in range(3):
            pkg.fEuler(0.0, w, wdot, u, sync=False)
        torch.cuda.synchronize()
        mine = u.profile(on=False, reset=True)
        if world > 1:
            allp = [None] * world if rank == 0 else None
            dist.gather_object(mine, allp, dst=0)
        else:
            allp = [mine]
        if rank == 0:
            timeline = {"unit": "ms per RHS, device time (eulerb200_profile)", "per_rank": allp,
                        "note": "transfer = halo pack + send/receive on the exchange stream (highest priority), "
                                "concurrent with prepass and interior; pack = what of it the launching stream waits for; "
                                "shells = what is left of the boundary shells after the interior launch "
                                "(EULERB200_OVERLAP=2: they run beside it); rhs = whole call on the launching stream"}
    except Exception as ex:      # pragma: no cover
        timeline = {"error": str(ex)[:200]}

    hbm_peak, peak_src = measured_peaks()
    traffic = None          # DRAM bytes per step from the committed ncu capture of this exact workload
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_512cube_nchem10.json")))
        if (u.nxl, u.nyl, u.nzl, args.nchem) == (512, 512, 512, 10):
            traffic = tj["per_step_total_bytes"]
    except Exception:
        pass
    alg_bytes = 16.0 * nvar * cells_local
    achieved_gbs = alg_bytes / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "kernel": "rhs_fused_kernel (+ aux_kernel pre-pass, 2 % of the step)", "kernel_ms": kernel_ms,
                "algorithmic_bytes": alg_bytes,
                "traffic_source": "profiles/traffic_512cube_nchem10.json (ncu --set full; bytes per step = per launch)",
                "algorithmic_bytes_per_cell": 16 * nvar,
                "note": "FP64-pipe bound kernel (see fp64): HBM fraction reported as the contract asks"}
    fp64 = None
    try:
        import ctypes as C
        lib = pkg.load_library()
        ach = ref_flops_per_cell(nvar) * cells_local / (kernel_ms * 1e-3) / 1e12
        fp64 = {"achieved": ach, "unit": "TFLOP/s", "peak_nominal": FP64_NOMINAL_TFLOPS,
                "frac_nominal": ach / FP64_NOMINAL_TFLOPS, "flops_per_cell": ref_flops_per_cell(nvar),
                "note": "flops = reference-as-written count per cell (BASELINE.md), div/sqrt = 1; nominal peak = "
                        "148 SMs x 64 lanes x 2 x 1.965 GHz; probe peak = DFMA micro-benchmark measured in this run"}
        tf = C.c_double(0)
        if lib.eulerb200_fp64_peak(C.byref(tf)) == 0:
            fp64["peak"] = tf.value
            fp64["frac"] = ach / tf.value
    except Exception as e:  # pragma: no cover
        fp64 = {"error": str(e)}

    # end to end through the host-pointer C-ABI call: pinned host arrays in, pinned host arrays out.
    # Host memory is bounded: if 2 x state x ranks-on-this-node would take more than 40 % of
    # MemAvailable, the same call is timed on a grid with fewer z-planes per rank (stated in e2e.sample).
    # At N > 1 it is the decomposed path: every rank uploads its block, the halo exchange runs, the
    # RHS is evaluated and the block's wdot is downloaded.
    e2e = None
    if not args.no_e2e:
        try:
            avail = 64e9
            try:
                for ln in open("/proc/meminfo"):
                    if ln.startswith("MemAvailable"):
                        avail = float(ln.split()[1]) * 1024
            except Exception:
                pass
            per_plane = 2 * 8 * nvar * u.nxl * u.nyl
            nz_e2e = int(min(u.nzl, max(8, (0.4 * avail / max(1, world)) // per_plane)))
            full = nz_e2e == u.nzl
            if full:
                ue, src = u, w
            else:
                ue = pkg.EulerData(nchem=args.nchem)
                pkg.problems.configure(problem, ue)
                ue.nx, ue.ny, ue.nz = u.nx, u.ny, nz_e2e * u.npz
                assert ue.SetupDecomp(myid=rank, nprocs=world, device=local_rank) == 0
                assert (ue.nxl, ue.nyl, ue.nzl) == (u.nxl, u.nyl, nz_e2e), "e2e grid decomposes differently"
                ncell = ue.nxl * ue.nyl * ue.nzl
                src = pkg.ManyVector([s_[:ncell * (1 if f < 5 else args.nchem)] for f, s_ in enumerate(w.sub)])
            hw = pkg.ManyVector([torch.empty(s_.shape, dtype=torch.float64, pin_memory=True) for s_ in src.sub])
            for h, d in zip(hw.sub, src.sub):
                h.copy_(d)
            hwdot = pkg.ManyVector([torch.empty(s_.shape, dtype=torch.float64, pin_memory=True) for s_ in src.sub])
            torch.cuda.synchronize()
            ret = pkg.fEuler(0.0, hw, hwdot, ue)      # warm-up (allocates the staging arrays)
            assert ret == 0, ue.last_error()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                ret = pkg.fEuler(0.0, hw, hwdot, ue)
            barrier()
            dt = (time.perf_counter() - t0) / args.e2e_steps
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
            cells_e2e = ue.nx * ue.ny * ue.nz
            nbytes = 8 * nvar * ue.nxl * ue.nyl * ue.nzl
            e2e = {"value": cells_e2e / dt / 1e9, "unit": "Gcell/s", "ms_per_step": dt * 1e3,
                   "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "steps": args.e2e_steps,
                   "bytes_are": "per rank",
                   "api": "eulerb200_rhs_host (fEuler on host ManyVector), pinned host memory" +
                          (", z-slab pipelined" if world == 1 else ", decomposed with halo exchange (NCCL)"),
                   "numa": numa,
                   "sample": ("the full workload" if full else
                              "bounded by host memory: the same decomposed grid with %d instead of %d z-planes per rank "
                              "(global %dx%dx%d)" % (nz_e2e, u.nzl, ue.nx, ue.ny, ue.nz))}
            if full and world == 1:   # the host path must give the same answer as the device path
                err = max(float((a.cuda() - b).abs().max() / b.abs().max()) for a, b in zip(hwdot.sub, wdot.sub))
                e2e["max_rel_diff_vs_device_path"] = err
            if not full:
                ue.FreeData()
            del hw, hwdot
        except Exception as ex:
            e2e = {"value": None, "unit": "Gcell/s", "error": str(ex)[:200],
                   "h2d_bytes_per_step": 8 * nvar * cells_local, "d2h_bytes_per_step": 8 * nvar * cells_local}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_reference_run(nvar, steps=2, warmup=1)
            cpu = {"value": r["value"], "unit": "Gcell/s", "cores": r["cores"], "kind": r["kind"],
                   "sample": r["sample"]}
        except Exception as ex:  # pragma: no cover
            cpu = {"value": None, "unit": "Gcell/s", "cores": 0, "kind": "reference", "sample": "failed: %s" % ex}

    if rank == 0:
        line = {
            "metric": "fEuler cell-RHS evaluations per second", "value": value, "unit": "Gcell/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic" if args.ic == "random" else "synthetic (%s initial condition)" % problem,
            "config": {"workload": workload_name(args.workload, (u.nxl, u.nyl, u.nzl), args.nchem),
                       "global_grid": [u.nx, u.ny, u.nz], "process_grid": [u.npx, u.npy, u.npz],
                       "l2": "inputs (%.1f GB per GPU) far larger than the 126 MB L2; no flush needed"
                             % (8 * nvar * cells_local / 1e9),
                       "per_step_ms_rank0": [round(x, 3) for x in per_step]},
            "roofline": roofline, "fp64": fp64, "parity": parity, "timeline": timeline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    u.FreeData()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
