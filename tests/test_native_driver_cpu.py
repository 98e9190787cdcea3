"""The native explicit driver (host/euler3d_b200.cpp) on the CPU tier: linked against the CPU
emulation of the kernel source (tests/emu/emu_abi.cpp, test infrastructure) instead of
libeulerb200.so, it runs small problems end to end without a GPU -- input parsing, problem
plug-ins, Butcher tables, step loop, diagnostics text, solution files -- and must reproduce the
Python driver driven by the CPU oracle.  The device library itself is exercised by
tests/test_gpu_native_driver.py."""
import math
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import NpVec, OracleVecOps

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
P, N, D, R = 0, 1, 2, 3


@pytest.fixture(scope="module")
def exe(native_emu_exe):
    return native_emu_exe


def run(exe, args, cwd):
    res = subprocess.run([exe] + args, capture_output=True, text=True, timeout=600, cwd=str(cwd))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    return res.stdout


def advection_state(n, axis):
    d = [1.0 / n[0], 1.0 / n[1], 1.0 / n[2]]
    idx = np.arange(n[0] * n[1] * n[2])
    c = [(idx % n[0] + 0.5) * d[0], ((idx // n[0]) % n[1] + 0.5) * d[1], (idx // (n[0] * n[1]) + 0.5) * d[2]]
    rho = 1.0 + 0.1 * np.sin(2 * math.pi * c[axis])
    m = [0.5 * rho if a == axis else np.zeros_like(rho) for a in range(3)]
    et = 1.0 / 0.4 + 0.5 * (m[0] ** 2 + m[1] ** 2 + m[2] ** 2) / rho
    return [rho] + m + [et], d


@pytest.mark.parametrize("sel,kw", [(["--order=4"], dict(order=4)), (["--order=3"], dict(order=3)),
                                    (["--order=0", "--etable=8"], dict(order=0, etable=8)),
                                    (["--order=0", "--etable=6"], dict(order=0, etable=6)),
                                    (["--order=0", "--etable=12"], dict(order=0, etable=12)),
                                    (["--order=6"], dict(order=6)),                       # Verner 8-5-6
                                    (["--order=0", "--etable=11"], dict(order=0, etable=11))])   # Fehlberg 13-7-8: 14-term combinations
def test_fixed_step_runs_equal_the_python_driver_with_the_oracle_rhs(pkg, port, exe, tmp_path, sel, kw):
    """linear_advection_y, 3 x 24 x 3, 20 fixed steps with the chosen Butcher table: the state the
    native driver writes equals the Python driver's (numpy stage arithmetic, oracle fEuler) to 1e-12."""
    n, h, tf = (3, 24, 3), 0.005, 0.1
    out = run(exe, ["--problem=linear_advection_y", "--nx=3", "--ny=24", "--nz=3", "--tf=%g" % tf, "--nout=1",
                    "--fixedstep=1", "--hmax=%g" % h, "--output=1"] + sel, tmp_path)
    nst = int(re.search(r"Internal solver steps = (\d+)", out).group(1))
    nfe = int(re.search(r"Fe = (\d+)", out).group(1))
    parts, d = advection_state(n, 1)
    ops = OracleVecOps(port, None, n, 0, d, 1.4, [P] * 6)
    step = pkg.driver.ERKStep(ops, 0.0, NpVec(parts), pkg.driver.ARKODEParameters(fixedstep=1, hmax=h, **kw))
    assert step.evolve(tf) == (0, tf)
    assert (nst, nfe) == (step.stats()["nst"], step.stats()["nfe"]) and nst == 20
    sol = pkg.problems.read_solution(str(tmp_path / pkg.problems.solution_name(1)))
    assert sol["time"] == pytest.approx(tf) and sol["n"] == n
    for f, name in enumerate(pkg.problems.dataset_names(0)):
        ref = step.w.sub[f]
        assert np.abs(sol[name].ravel() - ref).max() <= 1e-12 * max(np.abs(ref).max(), 1.0), name


def test_adaptive_sod_run_prints_the_reference_diagnostics(pkg, port, exe, tmp_path):
    """sod_x 40 x 3 x 3, adaptive order 4: diagnostics text, counters close to the Python driver's
    (same loop; the kernel arithmetic differs from the oracle's at 1e-15, which an adaptive run on a
    shock can turn into a step more or less), conservation while the waves are inside the domain."""
    out = run(exe, ["--problem=sod_x", "--nx=40", "--ny=3", "--nz=3"] + ["--%s=1" % k for k in
              ("xlbc", "xrbc", "ylbc", "yrbc", "zlbc", "zrbc")] + ["--tf=0.05", "--nout=2", "--rtol=1e-5", "--atol=1e-10"], tmp_path)
    errR = [[float(x) for x in m.split()] for m in re.findall(r"errR =\s+(.*)", out)]
    nst = int(re.search(r"Internal solver steps = (\d+)", out).group(1))
    assert len(errR) == 3 and errR[0][0] == 0.0 and all(e[2] == 0.0 and e[3] == 0.0 for e in errR)
    x = (np.arange(360) % 40 + 0.5) / 40
    rho, p = np.where(x < 0.5, 1.0, 0.125), np.where(x < 0.5, 1.0, 0.1)
    ops = OracleVecOps(port, None, (40, 3, 3), 0, (1 / 40, 1 / 3, 1 / 3), 1.4, [N] * 6)
    step = pkg.driver.ERKStep(ops, 0.0, NpVec([rho, np.zeros(360), np.zeros(360), np.zeros(360), p / 0.4]),
                              pkg.driver.ARKODEParameters(order=4, rtol=1e-5, atol=1e-10))
    for tout in (0.025, 0.05):
        assert step.evolve(tout)[0] == 0
    assert abs(step.stats()["nst"] - nst) <= max(2, 0.05 * nst)
    sol = pkg.problems.exact_riemann(0.05, list((np.arange(40) + 0.5) / 40), 0.5, 1.4)
    err = np.sqrt(np.mean((step.w.sub[0] - np.array([s[0] for s in sol])[np.arange(360) % 40]) ** 2))
    assert errR[-1][0] == pytest.approx(err, rel=3e-3)          # printed with three digits


def test_colour_tracers_solution_files_and_restart(pkg, exe, tmp_path):
    """hurricane_xy with four colour-stripe species (the reference's NVAR = 9 build), fixed step:
    species statistics in the text, solution files at every output, and a restart from the middle
    file that ends in the state of the uninterrupted run (io.cpp:716-1150)."""
    args = ["-f", os.path.join(ROOT, "inputs", "input_hurricane.txt"), "--problem=hurricane_xy", "--nx=12", "--ny=10",
            "--nz=3", "--nchem=4", "--tf=0.004", "--fixedstep=1", "--hmax=0.0005", "--output=1"]
    out = run(exe, args + ["--nout=2"], tmp_path)
    assert "num chemical species: 4" in out and "||c3||" in out
    assert int(re.search(r"Internal solver steps = (\d+)", out).group(1)) == 8
    sols = [pkg.problems.read_solution(str(tmp_path / pkg.problems.solution_name(i))) for i in range(3)]
    assert [s["time"] for s in sols] == [0.0, 0.002, 0.004] and sols[0]["nchem"] == 4 and sols[0]["n"] == (12, 10, 3)
    stripes = sum(sols[0]["Chemical-%03d" % v] for v in range(4))
    assert np.array_equal(stripes, np.ones_like(stripes))            # every cell starts in exactly one stripe
    assert not np.array_equal(sols[2]["Chemical-000"], sols[0]["Chemical-000"])        # and the stripes move
    again = tmp_path / "again"
    again.mkdir()
    (again / pkg.problems.solution_name(1)).write_bytes((tmp_path / pkg.problems.solution_name(1)).read_bytes())
    out = run(exe, args + ["--nout=1", "--restart=1"], again)
    assert "restarting from output-0000001.eb200 at t = 0.002" in out
    re2 = pkg.problems.read_solution(str(again / pkg.problems.solution_name(2)))
    assert re2["time"] == 0.004
    for name in pkg.problems.dataset_names(4):
        assert np.abs(re2[name] - sols[2][name]).max() <= 1e-13 * max(np.abs(sols[2][name]).max(), 1e-300), name
    res = subprocess.run([exe] + args + ["--nout=1", "--restart=1", "--nx=13"], capture_output=True, text=True, cwd=str(again))
    assert res.returncode != 0 and "holds a 12 x 10 x 3 grid" in res.stderr


@pytest.mark.parametrize("plane", ["xy", "yz", "zx"])
def test_hurricane_diagnostics_python_equals_native(pkg, exe, tmp_path, plane):
    """output_diagnostics of the hurricane problems (density and the three momenta against the
    critical-rotation solution, hurricane.cpp:217-334): problems.py evaluated on the state the native
    driver wrote gives the numbers the native driver printed (which tests/test_reference_main_cpu.py
    shows to be the reference program's)."""
    import torch
    n = {"xy": (16, 14, 3), "yz": (3, 16, 14), "zx": (14, 3, 16)}[plane]
    out = run(exe, ["-f", os.path.join(ROOT, "inputs", "input_hurricane.txt"), "--problem=hurricane_" + plane,
                    "--nx=%d" % n[0], "--ny=%d" % n[1], "--nz=%d" % n[2], "--tf=0.002", "--nout=1", "--fixedstep=1",
                    "--hmax=0.0005", "--output=1"], tmp_path)
    printed = [[float(x) for x in m.split()] for m in re.findall(r"err[IR] =\s+(.*)", out)]
    assert len(printed) == 4 and all(len(p) == 4 for p in printed)
    sol = pkg.problems.read_solution(str(tmp_path / pkg.problems.solution_name(1)))
    u = pkg.EulerData()
    u.nx, u.ny, u.nz = n
    pkg.problems.configure("hurricane_" + plane, u)
    u.dx, u.dy, u.dz = (u.xr - u.xl) / n[0], (u.yr - u.yl) / n[1], (u.zr - u.zl) / n[2]
    u.nxl, u.nyl, u.nzl = n
    u.is_ = u.js = u.ks = 0
    names = pkg.problems.dataset_names(0)
    w = pkg.ManyVector([torch.from_numpy(np.ascontiguousarray(sol[nm]).ravel().copy()) for nm in names])
    d = pkg.problems.output_diagnostics("hurricane_" + plane, 0.002, w, u, quiet=True)
    for got, want in ((d["errI"], printed[2]), (d["errR"], printed[3])):
        assert len(got) == 4
        for a, b in zip(got, want):
            assert a == pytest.approx(b, rel=6e-3, abs=1e-12)          # three printed digits


def test_initial_transient_then_fixed_steps_equals_the_python_driver(pkg, port, exe, tmp_path):
    """fixedstep = 1 with htrans > 0 (euler3D_main.cpp:87-88,345-367): adaptive steps (<= hmax) over
    (t0, t0+htrans], then ARKStepSetFixedStep(hmax).  Native driver against driver.py + oracle RHS."""
    n, h, tf, htrans = (3, 24, 3), 0.005, 0.1, 0.02
    out = run(exe, ["--problem=linear_advection_y", "--nx=3", "--ny=24", "--nz=3", "--tf=%g" % tf, "--nout=1",
                    "--fixedstep=1", "--hmax=%g" % h, "--htrans=%g" % htrans, "--output=1", "--rtol=1e-6"], tmp_path)
    nst = int(re.search(r"Internal solver steps = (\d+)", out).group(1))
    parts, d = advection_state(n, 1)
    ops = OracleVecOps(port, None, n, 0, d, 1.4, [P] * 6)
    step = pkg.driver.ERKStep(ops, 0.0, NpVec(parts), pkg.driver.ARKODEParameters(order=4, rtol=1e-6, hmax=h))
    assert step.evolve(htrans) == (0, htrans)
    step.set_fixed_step(h)
    assert step.evolve(tf) == (0, tf)
    assert step.stats()["nst"] == nst and nst >= 4 + 16
    sol = pkg.problems.read_solution(str(tmp_path / pkg.problems.solution_name(1)))
    for f, name in enumerate(pkg.problems.dataset_names(0)):
        assert np.abs(sol[name].ravel() - step.w.sub[f]).max() <= 1e-12 * max(np.abs(step.w.sub[f]).max(), 1.0), name


def test_restart_parameters_file_continues_the_run(pkg, exe, tmp_path):
    """write_parameters (io.cpp:645-714): every solution file comes with a restart_parameters.txt that
    continues the run (`-f restart_parameters.txt`): t0 = the file's time, the remaining outputs, the
    current step, restart = <file number>.  Half a run plus its continuation ends where the whole run ends."""
    base = ["--problem=linear_advection_x", "--nx=24", "--ny=3", "--nz=3", "--fixedstep=1", "--hmax=0.005", "--output=1"]
    whole = tmp_path / "whole"
    halves = tmp_path / "halves"
    whole.mkdir(), halves.mkdir()
    run(exe, base + ["--tf=0.1", "--nout=2"], whole)
    run(exe, base + ["--tf=0.1", "--nout=2", "--mxsteps=5000"], halves)      # same run ...
    par = (halves / "restart_parameters.txt").read_text()
    assert "problem = linear_advection_x" in par and re.search(r"^restart = 2$", par, re.M) and re.search(r"^nout = 0$", par, re.M)
    # ... and one that stops after the first output interval, then continues from its parameter file
    for f in halves.iterdir():
        f.unlink()
    run(exe, base + ["--tf=0.05", "--nout=1"], halves)
    par = (halves / "restart_parameters.txt").read_text()
    assert re.search(r"^restart = 1$", par, re.M) and re.search(r"^t0 = 0\.05", par, re.M)
    out = run(exe, ["-f", "restart_parameters.txt", "--tf=0.1", "--nout=1"], halves)
    assert "restarting from output-0000001.eb200 at t = 0.05" in out
    a = pkg.problems.read_solution(str(whole / pkg.problems.solution_name(2)))
    b = pkg.problems.read_solution(str(halves / pkg.problems.solution_name(2)))
    assert a["time"] == b["time"] == 0.1
    for name in pkg.problems.dataset_names(0):
        assert np.abs(a[name] - b[name]).max() <= 1e-13 * max(np.abs(a[name]).max(), 1e-300), name
