"""The native C++ explicit driver (host/euler3d_b200.cpp -> euler3d_b200) on the reference's
Sod input parameters, on linear advection and on Rayleigh-Taylor: runs to completion, prints
the reference's diagnostics, and -- being the same algorithm as driver.py -- reproduces the
Python driver's step counts and errors."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "sundials-manyvector-demo_b200", "euler3d_b200")


def run(args):
    res = subprocess.run([EXE] + args, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    return res.stdout


def parse(out):
    errR = [[float(x) for x in m.split()] for m in re.findall(r"errR =\s+(.*)", out)]
    nst = int(re.search(r"Internal solver steps = (\d+)", out).group(1))
    nfe = int(re.search(r"Fe = (\d+)", out).group(1))
    netf = int(re.search(r"error test failures = (\d+)", out).group(1))
    drift = [float(x) for x in re.findall(r"relative change\s+= (\S+)", out)]
    return errR, nst, nfe, netf, drift


def test_sod_input_file(pkg):
    out = run(["-f", os.path.join(ROOT, "inputs", "input_sod.txt")])
    errR, nst, nfe, netf, drift = parse(out)
    assert len(errR) == 11                                   # initial output + nout = 10
    assert errR[0][0] == 0.0 and all(e[0] < 2e-2 and e[2] == 0.0 and e[3] == 0.0 for e in errR)
    assert nst > 20 and nfe >= 5 * nst
    assert drift[0] < 1e-12                                  # mass: waves have not reached the ends
    # same loop as driver.py: identical counters
    u = pkg.EulerData()
    u.nx, u.ny, u.nz = 200, 3, 3
    pkg.problems.configure("sod_x", u)
    assert u.SetupDecomp(device=0) == 0
    w = pkg.ManyVector.new(u)
    pkg.problems.initial_conditions("sod_x", 0.0, w, u)
    step = pkg.driver.ERKStep(pkg.driver.TorchVecOps(pkg, u), 0.0, w,
                              pkg.driver.ARKODEParameters(order=4, rtol=1e-5, atol=1e-12, mxsteps=10000))
    for i in range(10):
        assert step.evolve(0.0 + (0.2 / 10) * (i + 1))[0] == 0
    st = step.stats()
    assert abs(st["nst"] - nst) <= 0.1 * nst                 # adaptive on a shock: see test_gpu_driver.py
    d = pkg.problems.output_diagnostics("sod_x", 0.2, step.w, u, quiet=True)
    assert abs(d["errR"][0] - errR[-1][0]) <= 0.05 * errR[-1][0]
    u.FreeData()


def test_linear_advection_and_overrides(pkg):
    out = run(["-f", os.path.join(ROOT, "inputs", "input_linear_advection.txt"), "--nx=48", "--tf=0.25", "--nout=2"])
    errR, nst, nfe, netf, drift = parse(out)
    assert "spatial grid: 48 x 16 x 16" in out and len(errR) == 3
    assert errR[-1][0] < 5e-6 and max(drift) < 1e-13
    out_y = run(["-f", os.path.join(ROOT, "inputs", "input_linear_advection.txt"), "--problem=linear_advection_y",
                 "--nx=16", "--ny=48", "--tf=0.25", "--nout=2"])
    errR_y = parse(out_y)[0]
    assert errR_y[-1][0] == pytest.approx(errR[-1][0], rel=1e-6)      # the reference's x/y/z symmetry check


def test_rayleigh_taylor_fixed_step(pkg):
    out = run(["-f", os.path.join(ROOT, "inputs", "input_rayleigh_taylor.txt"), "--nx=32", "--ny=96", "--tf=0.1",
               "--nout=2", "--fixedstep=1", "--hmax=0.002"])
    _, nst, nfe, netf, drift = parse(out)
    assert nst == 50 and nfe == 250 and netf == 0
    # the reference's high-side ghosts are copies, not mirrors (euler3D.hpp:988), so its reflecting
    # wall at y = +0.75 is not exactly flux-free: mass drifts at the 1e-4 level there too
    assert drift[0] < 1e-3


def _python_fixed_run(pkg, problem, n, nchem, touts, h, order=4):
    """The same fixed-step run through driver.py + problems.py (device-resident ManyVector)."""
    u = pkg.EulerData(nchem=nchem)
    u.nx, u.ny, u.nz = n
    pkg.problems.configure(problem, u)
    assert u.SetupDecomp(device=0) == 0
    w = pkg.ManyVector.new(u)
    pkg.problems.initial_conditions(problem, 0.0, w, u)
    step = pkg.driver.ERKStep(pkg.driver.TorchVecOps(pkg, u), 0.0, w,
                              pkg.driver.ARKODEParameters(order=order, fixedstep=1, hmax=h))
    for t in touts:
        assert step.evolve(t)[0] == 0
    return u, step


def _compare_with_file(pkg, u, w, sol, tol=1e-12):
    import numpy as np
    names = pkg.problems.dataset_names(u.nchem)
    scales = pkg.problems._unit_scales(u)
    for f in range(5):
        ref = w.sub[f].cpu().numpy() * scales[f]
        assert np.abs(sol[names[f]].ravel() - ref).max() <= tol * max(np.abs(ref).max(), 1e-300), names[f]
    if u.nchem:
        chem = w.sub[5].cpu().numpy().reshape(-1, u.nchem)
        for v in range(u.nchem):
            assert np.abs(sol[names[5 + v]].ravel() - chem[:, v]).max() <= tol * max(np.abs(chem[:, v]).max(), 1e-300)


def test_hurricane_colour_tracers_output_and_restart(pkg, tmp_path):
    """nchem > 0 in the native driver (NVAR = 9 build of the reference), solution files and restart."""
    import numpy as np
    args = ["-f", os.path.join(ROOT, "inputs", "input_hurricane.txt"), "--problem=hurricane_xy", "--nx=24", "--ny=20",
            "--nz=3", "--nchem=4", "--tf=0.01", "--fixedstep=1", "--hmax=0.0005", "--output=1"]
    res = subprocess.run([EXE] + args + ["--nout=2"], capture_output=True, text=True, timeout=600, cwd=tmp_path)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "num chemical species: 4" in res.stdout and "||c3||" in res.stdout
    assert int(re.search(r"Internal solver steps = (\d+)", res.stdout).group(1)) == 20
    sols = [pkg.problems.read_solution(str(tmp_path / pkg.problems.solution_name(i))) for i in range(3)]
    assert [s["time"] for s in sols] == [0.0, 0.005, 0.01] and sols[0]["nchem"] == 4 and sols[0]["n"] == (24, 20, 3)
    stripes = sum(sols[0]["Chemical-%03d" % v] for v in range(4))
    assert np.array_equal(stripes, np.ones_like(stripes))            # every cell starts in exactly one stripe
    # the same run through the python driver
    u, step = _python_fixed_run(pkg, "hurricane_xy", (24, 20, 3), 4, [0.005, 0.01], 0.0005)
    _compare_with_file(pkg, u, step.w, sols[2])
    u.FreeData()
    # restart from output 1: one more output interval ends in the same state
    again = tmp_path / "again"
    again.mkdir()
    (again / pkg.problems.solution_name(1)).write_bytes((tmp_path / pkg.problems.solution_name(1)).read_bytes())
    res = subprocess.run([EXE] + args + ["--nout=1", "--restart=1"], capture_output=True, text=True, timeout=600, cwd=again)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "restarting from output-0000001.eb200 at t = 0.005" in res.stdout
    re2 = pkg.problems.read_solution(str(again / pkg.problems.solution_name(2)))
    assert re2["time"] == 0.01
    for name in pkg.problems.dataset_names(4):
        scale = max(np.abs(sols[2][name]).max(), 1e-300)
        assert np.abs(re2[name] - sols[2][name]).max() <= 1e-13 * scale, name
    # a restart file of another grid is refused
    res = subprocess.run([EXE] + args + ["--nout=1", "--restart=1", "--nx=25"], capture_output=True, text=True, cwd=again)
    assert res.returncode != 0 and "holds a 24 x 20 x 3 grid" in res.stderr


def test_fluid_blast_input_file_with_tracers(pkg, tmp_path):
    """The reference's fluid_blast parameters with the ten primordial species as passive tracers
    (NVAR = 15): adaptive run completes; the initial file holds the reference's initial state; the
    run equals the python driver's."""
    import numpy as np
    res = subprocess.run([EXE, "-f", os.path.join(ROOT, "inputs", "input_fluid_blast.txt"), "--nchem=10", "--output=1",
                          "--nout=2"], capture_output=True, text=True, timeout=600, cwd=tmp_path)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    _, nst, nfe, netf, drift = parse(res.stdout)
    assert 2 <= nst <= 200 and nfe >= 5 * nst
    assert drift[0] < 1e-3 and drift[1] < 1e-6            # walls: high-side ghosts are copies (see RT test)
    ic = pkg.problems.read_solution(str(tmp_path / pkg.problems.solution_name(0)))
    u = pkg.EulerData(nchem=10)
    u.nx = u.ny = u.nz = 10
    pkg.problems.configure("fluid_blast", u)
    assert u.SetupDecomp(device=0) == 0
    w = pkg.ManyVector.new(u)
    pkg.problems.initial_conditions("fluid_blast", 0.0, w, u)
    _compare_with_file(pkg, u, w, ic, tol=1e-14)
    # same adaptive loop in python: same step sequence on this smooth problem
    step = pkg.driver.ERKStep(pkg.driver.TorchVecOps(pkg, u), 0.0, w,
                              pkg.driver.ARKODEParameters(order=4, rtol=1e-5, atol=1e-9, safety=0.99, bias=2.0, growth=2.0))
    for t in (0.5, 1.0):
        assert step.evolve(t)[0] == 0
    same = step.stats()["nst"] == nst and step.stats()["nfe"] == nfe
    assert abs(step.stats()["nst"] - nst) <= 1               # (the error norm is summed with atomics)
    _compare_with_file(pkg, u, step.w, pkg.problems.read_solution(str(tmp_path / pkg.problems.solution_name(2))),
                       tol=1e-11 if same else 1e-4)
    u.FreeData()
