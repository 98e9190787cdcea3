"""The reference's OWN main program on the CPU tier.  oracle/Makefile (target refmain) compiles
euler3D_main.cpp, io.cpp, gopt.cpp and one problem file of the reference -- all unmodified, from
where they lie -- against the SUNDIALS/MPI shim, with oracle/shim/shim_arkstep.cpp standing in for
ARKODE's ARKStep on top of this repository's ERK loop (host/erk_stepper.hpp).  Two flavours:

  refmain_<problem>          the reference fEuler / stability (utilities.cpp)
  refmain_dropin_<problem>   OUR fEuler / stability (host/feuler_dropin.cpp) in their place, running
                             the kernel source through the CPU emulation (tests/emu/emu_abi.cpp)

What this pins, in fixed-step runs (so that the step sequence is the same by construction):
  * the drop-in claim end to end: the reference's main, problem file, diagnostics and I/O code call
    our fEuler unchanged and print the same text as with their own;
  * the native driver (host/euler3d_b200.cpp, here linked against the same emulation): its initial
    conditions, errI/errR diagnostics, statistics table, conservation check and final statistics are
    the reference program's, line for line.
It does not pin ARKODE itself (absent from the image): both sides run the same ERK loop."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INPUTS = os.path.join(ROOT, "inputs")

CASES = {
    "sod_x": ("input_sod.txt", ["--nx=40", "--tf=0.01", "--nout=2", "--fixedstep=1", "--hmax=0.0005"]),
    "linear_advection_y/fixed": ("input_linear_advection.txt", ["--nx=3", "--ny=24", "--nz=3", "--tf=0.05", "--nout=2",
                                                                "--fixedstep=1", "--hmax=0.005", "--showstats=1"]),
    "rayleigh_taylor": ("input_rayleigh_taylor.txt", ["--nx=16", "--ny=48", "--tf=0.02", "--nout=2", "--fixedstep=1",
                                                      "--hmax=0.002", "--showstats=1"]),
    "hurricane_yz": ("input_hurricane.txt", ["--nx=3", "--ny=20", "--nz=20", "--tf=0.002", "--nout=2", "--fixedstep=1",
                                             "--hmax=0.0005", "--showstats=1"]),
    "hurricane_zx": ("input_hurricane.txt", ["--nx=18", "--ny=3", "--nz=22", "--tf=0.002", "--nout=2", "--fixedstep=1",
                                             "--hmax=0.0005", "--showstats=1"]),
    "sod_z": ("input_sod.txt", ["--nx=3", "--ny=3", "--nz=40", "--tf=0.01", "--nout=2", "--fixedstep=1", "--hmax=0.0005"]),
    # six colour-stripe species (the reference's NVAR = 11 build): tracer initial condition, species
    # columns of the statistics table, and the species part of the RHS through the drop-in
    "hurricane_zx_color": ("input_hurricane.txt", ["--nx=14", "--ny=3", "--nz=12", "--tf=0.002", "--nout=2", "--fixedstep=1",
                                                   "--hmax=0.0005", "--showstats=1"]),
    # fixedstep with htrans > 0: adaptive (steps <= hmax) over the initial transient, fixed afterwards
    "linear_advection_y": ("input_linear_advection.txt", ["--nx=3", "--ny=24", "--nz=3", "--tf=0.1", "--nout=2",
                                                          "--fixedstep=1", "--hmax=0.005", "--htrans=0.02", "--showstats=1"]),
    # cosmological unit factors: the totals and the statistics table are printed in CGS units
    "fluid_blast": ("input_fluid_blast.txt", ["--tf=0.2", "--nout=2", "--fixedstep=1", "--hmax=0.05", "--showstats=1"]),
}


@pytest.fixture(scope="module")
def refmain(oracle_mod):
    exes = oracle_mod.build_refmain()
    if not all(p.split("/")[0] in exes and "dropin_" + p.split("/")[0] in exes for p in CASES):
        pytest.skip("needs the reference tree (or prebuilt oracle/_ref/refmain_*)")
    return exes


def report(cmd, cwd):
    """stdout from 'Writing initial batch of outputs' on, without wall-clock lines; conservation drifts
    at round-off level (< 1e-13) are replaced by a token (they are sums over differently ordered terms)."""
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(cwd))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    lines = res.stdout.split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("Writing initial batch"))
    keep = []
    for l in lines[start:]:
        if re.match(r"^Total .* time", l) or "GPU kernel launches" in l or l.startswith("Profiling Results") or not l.strip():
            continue
        m = re.match(r"^(\s+(Mass|Energy) conservation relative change\s+= )(\S+)", l)
        if m and float(m.group(3)) < 1e-13:
            l = m.group(1) + "round-off"
        keep.append(l.rstrip())
    return keep


@pytest.mark.parametrize("problem", sorted(CASES))
def test_reference_main_prints_the_same_with_our_feuler_and_the_native_driver_prints_the_same(refmain, native_emu_exe,
                                                                                             tmp_path, problem):
    infile, args = CASES[problem]
    problem = problem.split("/")[0]
    common = ["-f", os.path.join(INPUTS, infile)] + args
    ref = report([refmain[problem]] + common, tmp_path)
    assert any("errI" in l for l in ref) or problem in ("rayleigh_taylor", "fluid_blast")
    assert sum("Total RHS evals" in l for l in ref) == 1
    dropin = report([refmain["dropin_" + problem]] + common, tmp_path)
    assert dropin == ref, "\n".join(dropin) + "\n--- vs reference fEuler ---\n" + "\n".join(ref)
    native_args = ["--problem=hurricane_zx", "--nchem=6"] if problem == "hurricane_zx_color" else ["--problem=" + problem]
    native = report([native_emu_exe] + common + native_args, tmp_path)
    assert native == ref, "\n".join(native) + "\n--- vs reference main ---\n" + "\n".join(ref)


def test_adaptive_runs_agree_to_the_integration_tolerance(refmain, native_emu_exe, tmp_path):
    """Adaptive Sod run (the reference's input file parameters on 40 cells): same controller on both
    sides, right-hand sides that differ at 1e-15 -- the step counts stay within a few per cent and the
    printed errors agree to the digits shown."""
    common = ["-f", os.path.join(INPUTS, "input_sod.txt"), "--nx=40", "--tf=0.05", "--nout=2"]
    ref = report([refmain["sod_x"]] + common, tmp_path)
    nat = report([native_emu_exe] + common, tmp_path)
    assert [l for l in ref if "err" in l and "=  " in l] == [l for l in nat if "err" in l and "=  " in l]
    steps = [int(re.search(r"steps = (\d+)", next(l for l in t if "Internal solver steps" in l)).group(1)) for t in (ref, nat)]
    assert abs(steps[0] - steps[1]) <= 0.05 * steps[0]


def test_cfl_limited_run_goes_through_the_stability_hook(refmain, native_emu_exe, tmp_path):
    """cfl > 0 makes the reference main register `stability` with the integrator
    (euler3D_main.cpp:271-275); with loose tolerances every step is the CFL step, so the three
    programs -- reference stability, drop-in stability, native driver -- take the same steps."""
    common = ["-f", os.path.join(INPUTS, "input_sod.txt"), "--nx=40", "--tf=0.02", "--nout=2", "--cfl=0.2",
              "--rtol=1e-2", "--atol=1e-2"]
    ref = report([refmain["sod_x"]] + common, tmp_path)
    assert any("Internal solver steps = 7 " in l for l in ref)          # 0.2 * dx / alpha ~ 2.3e-3 ... 3e-3
    assert report([refmain["dropin_sod_x"]] + common, tmp_path) == ref
    assert report([native_emu_exe] + common, tmp_path) == ref
