#!/usr/bin/env python
"""bench.py -- throughput of the fluid right-hand side (fEuler) in Gcell-RHS/s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--n NX NY NZ] [--nchem C] [--no-e2e] [--no-cpu-baseline]

One "step" = one complete fEuler evaluation (halo exchange included when N > 1) on a
synthetic admissible state.  Default workload: BASELINE.json's metric configuration,
fluid_blast/primordial_blast shape -- 512^3 cells per GPU, nchem = 10 (NVAR = 15), unit
cube, all-reflecting boundaries, gamma = 5/3 (tests/primordial_blast/input_*.txt of the
reference).  Weak scaling: every GPU owns 512^3 cells of a (512*npx, 512*npy, 512*npz)
grid decomposed as the reference's SetupDecomp would.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      HBM view of the RHS kernel: algorithmic bytes (16*NVAR per cell) / kernel time
  fp64          FP64-pipe view: reference-as-written flops per cell / kernel time
                against a DFMA peak measured in this run -- the pipe that actually binds
  cpu_baseline  the UNMODIFIED reference fEuler (oracle/_ref) on this box's host cores
  e2e           same metric through the host-pointer C-ABI call (pinned host arrays in/out)
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
    port of the oracle (single core) when oracle/_ref is absent."""
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
        # largest P' <= P whose balanced factorisation keeps the blocks near-cubic
        R = oracle.Ref(nvar)
        from __graft_entry__ import load_package
        pkg = load_package()
        _, dims, _, _, _ = pkg.dims_and_extents(P, 0, (per_rank * 8,) * 3, bc)
        n = tuple(per_rank * d for d in dims)
        w = oracle.random_state(n, nchem, seed=1234, gamma=gamma)
        ret, _, secs, dec = R.feuler(n, [0, 1] * 3, bc, gamma, w, nprocs=P, nrep=warmup + steps)
        assert ret == 0
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


def workload_name(n, nchem):
    return ("primordial_blast/fluid_blast shape: %dx%dx%d cells per GPU, nchem=%d (NVAR=%d), "
            "unit cube, all-reflecting, gamma=5/3" % (n[0], n[1], n[2], nchem, 5 + nchem))


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
        "ms_per_step": r["sec_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.n, args.nchem),
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, nargs=3, default=[512, 512, 512], help="cells per GPU")
    ap.add_argument("--nchem", type=int, default=10)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ic", default="random", choices=["random", "primordial_blast"],
                    help="synthetic state: seeded random admissible state (default, SURVEY.md 8(d)) or the "
                         "reference's primordial_blast initial condition (needs --nchem 10)")
    args = ap.parse_args()
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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        build()
    if world > 1:
        dist.barrier()
    pkg = load_package()

    nvar = 5 + args.nchem
    gamma = 5.0 / 3.0
    bcs = [pkg.BC_REFLECTING] * 6
    # weak scaling: global grid = per-GPU box x process grid of the reference's SetupDecomp
    _, dims, _, _, _ = pkg.dims_and_extents(world, 0, tuple(args.n), bcs)
    u = pkg.EulerData(nchem=args.nchem)
    u.nx, u.ny, u.nz = (args.n[0] * dims[0], args.n[1] * dims[1], args.n[2] * dims[2])
    u.xlbc, u.xrbc, u.ylbc, u.yrbc, u.zlbc, u.zrbc = bcs
    u.gamma = gamma
    assert u.SetupDecomp(myid=rank, nprocs=world, device=local_rank) == 0
    cells_local = u.nxl * u.nyl * u.nzl
    cells_global = u.nx * u.ny * u.nz

    if args.ic == "primordial_blast":
        pkg.problems.configure("primordial_blast", u)      # units only; BCs and gamma are already these
        w = pkg.ManyVector.new(u)
        assert pkg.problems.initial_conditions("primordial_blast", 0.0, w, u) == 0
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
        if hasattr(lib, "eulerb200_fp64_peak"):
            tf = C.c_double(0)
            lib.eulerb200_fp64_peak.restype = C.c_int
            lib.eulerb200_fp64_peak.argtypes = [C.POINTER(C.c_double)]
            if lib.eulerb200_fp64_peak(C.byref(tf)) == 0:
                ach = ref_flops_per_cell(nvar) * cells_local / (kernel_ms * 1e-3) / 1e12
                fp64 = {"achieved": ach, "peak": tf.value, "unit": "TFLOP/s", "frac": ach / tf.value,
                        "flops_per_cell": ref_flops_per_cell(nvar),
                        "note": "flops = reference-as-written count per cell (BASELINE.md), div/sqrt = 1; "
                                "peak = DFMA micro-benchmark measured in this run"}
    except Exception as e:  # pragma: no cover
        fp64 = {"error": str(e)}

    # end to end through the host-pointer C-ABI call: pinned host arrays in, pinned host arrays out.
    # Host memory is bounded: if 2 x state x ranks-on-this-node would take more than 40 % of
    # MemAvailable, every rank times the same call on a z-slab of its box instead (stated in e2e.sample).
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
            full = nz_e2e == u.nzl and world == 1
            if full:
                ue, src = u, w
            else:
                ue = pkg.EulerData(nchem=args.nchem)
                ue.nx, ue.ny, ue.nz = u.nxl, u.nyl, nz_e2e
                ue.xlbc = ue.xrbc = ue.ylbc = ue.yrbc = ue.zlbc = ue.zrbc = pkg.BC_REFLECTING
                ue.gamma = gamma
                assert ue.SetupDecomp(device=local_rank) == 0
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
            cells_e2e = ue.nxl * ue.nyl * ue.nzl
            nbytes = 8 * nvar * cells_e2e
            e2e = {"value": world * cells_e2e / dt / 1e9, "unit": "Gcell/s", "ms_per_step": dt * 1e3,
                   "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "steps": args.e2e_steps,
                   "api": "eulerb200_rhs_host (fEuler on host ManyVector), pinned host memory, z-slab pipelined",
                   "sample": ("the full workload" if full else
                              "bounded by host memory: every rank runs a %dx%dx%d slab of its box, no halo exchange"
                              % (ue.nxl, ue.nyl, ue.nzl))}
            if full:   # the host path must give the same answer as the device path
                err = max(float((a.cuda() - b).abs().max() / b.abs().max()) for a, b in zip(hwdot.sub, wdot.sub))
                e2e["max_rel_diff_vs_device_path"] = err
            else:
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
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic" if args.ic == "random" else "synthetic (primordial_blast initial condition)",
            "config": {"workload": workload_name((u.nxl, u.nyl, u.nzl), args.nchem),
                       "global_grid": [u.nx, u.ny, u.nz], "process_grid": [u.npx, u.npy, u.npz],
                       "l2": "inputs (%.1f GB per GPU) far larger than the 126 MB L2; no flush needed"
                             % (8 * nvar * cells_local / 1e9),
                       "per_step_ms_rank0": [round(x, 3) for x in per_step]},
            "roofline": roofline, "fp64": fp64, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    u.FreeData()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
