"""Solution files and native problem plug-ins on CPU:

* host/problems.hpp (the native driver's initial conditions) against the state the UNMODIFIED
  reference produced (tests/golden/ic_fluid_blast.npz) and against problems.py for every problem;
* the .eb200 solution file: native write -> python read, native read -> write round trip,
  python output_solution -> read_restart round trip, and two ranks (gloo) writing their
  sub-boxes into one file == the single-rank file, byte for byte.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("npc") / "native_problems_check")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-ffp-contract=off", "-o", exe,
                           os.path.join(HERE, "native_problems_check.cpp")])
    return exe


def native_state(harness, tmp_path, problem, n, nchem):
    a, b = str(tmp_path / "a.eb200"), str(tmp_path / "b.eb200")
    out = subprocess.run([harness, problem] + [str(x) for x in n] + [str(nchem), a, b], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return a, b, "analytic=1" in out.stdout


def python_state(pkg, problem, n, nchem, t):
    import torch
    u = pkg.EulerData(nchem=nchem)
    u.nx, u.ny, u.nz = n
    pkg.problems.configure(problem, u)
    u.dx, u.dy, u.dz = (u.xr - u.xl) / n[0], (u.yr - u.yl) / n[1], (u.zr - u.zl) / n[2]
    u.nxl, u.nyl, u.nzl = n
    u.is_ = u.js = u.ks = 0
    N = n[0] * n[1] * n[2]
    w = pkg.ManyVector([torch.zeros(N, dtype=torch.float64) for _ in range(5)] +
                       ([torch.zeros(N * nchem, dtype=torch.float64)] if nchem else []))
    assert pkg.problems.initial_conditions(problem, t, w, u) == 0
    return u, w


def test_native_fluid_blast_state_matches_reference(pkg, harness, tmp_path):
    """std::mt19937_64 clumps + blast of host/problems.hpp == the reference's own initial_conditions()."""
    z = np.load(os.path.join(HERE, "golden", "ic_fluid_blast.npz"))
    n = tuple(int(x) for x in z["n"])
    a, _, analytic = native_state(harness, tmp_path, "fluid_blast", n, 0)
    assert not analytic
    sol = pkg.problems.read_solution(a)
    assert sol["n"] == n and sol["nchem"] == 0 and sol["time"] == 0.125
    assert sol["domain"] == [0.0, 1.0, 0.0, 1.0, 0.0, 1.0]
    u, _ = python_state(pkg, "fluid_blast", n, 0, 0.0)
    scales = [u.DensityUnits, u.MomentumUnits, u.MomentumUnits, u.MomentumUnits, u.EnergyUnits]
    for f, name in enumerate(pkg.problems.FLUID_DATASETS):
        ref = z["w%d" % f] * scales[f]                         # the file holds CGS values (io.cpp:887-891)
        assert np.abs(sol[name].ravel() - ref).max() <= 1e-15 * max(np.abs(ref).max(), 1e-300), name


@pytest.mark.parametrize("problem,n,nchem", [
    ("sod_x", (20, 3, 4), 0), ("sod_z", (3, 4, 20), 2), ("linear_advection_y", (4, 16, 5), 0),
    ("rayleigh_taylor", (8, 24, 3), 0), ("hurricane_yz", (3, 16, 18), 4), ("hurricane_xy", (12, 10, 3), 0),
    ("primordial_blast", (10, 9, 8), 10), ("fluid_blast", (7, 8, 9), 10)])
def test_native_and_python_plugins_agree(pkg, harness, tmp_path, problem, n, nchem):
    a, b, analytic = native_state(harness, tmp_path, problem, n, nchem)
    assert analytic == (problem.startswith("sod") or problem.startswith("linear"))
    sol, back = pkg.problems.read_solution(a), pkg.problems.read_solution(b)
    u, w = python_state(pkg, problem, n, nchem, 0.125)
    scales = pkg.problems._unit_scales(u)
    names = pkg.problems.dataset_names(nchem)
    assert len(names) == 5 + nchem and names[-1] == ("Chemical-%03d" % (nchem - 1) if nchem else "TotalEnergy")
    for f in range(5):
        ref = w.sub[f].numpy() * scales[f]
        tol = 1e-14 * max(np.abs(ref).max(), 1e-300)
        assert np.abs(sol[names[f]].ravel() - ref).max() <= tol, names[f]
        # native read -> write round trip (x*s/s*s may move one ulp)
        assert np.abs(back[names[f]] - sol[names[f]]).max() <= 4e-16 * max(np.abs(ref).max(), 1e-300)
    if nchem:
        chem = w.sub[5].numpy().reshape(-1, nchem)
        for v in range(nchem):
            ref = chem[:, v]
            assert np.abs(sol[names[5 + v]].ravel() - ref).max() <= 1e-14 * max(np.abs(ref).max(), 1e-300), names[5 + v]
            assert np.array_equal(back[names[5 + v]], sol[names[5 + v]])


def test_python_output_and_restart_round_trip(pkg, harness, tmp_path):
    import torch
    n, nchem = (6, 5, 4), 10
    u, w = python_state(pkg, "primordial_blast", n, nchem, 0.0)
    assert pkg.problems.output_solution(0.75, w, u, 3, directory=str(tmp_path)) == 0
    path = tmp_path / pkg.problems.solution_name(3)
    assert path.name == "output-0000003.eb200"
    assert path.stat().st_size == pkg.problems.SOLUTION_HEADER_BYTES + 8 * 15 * 120
    w2 = pkg.ManyVector([torch.zeros_like(s) for s in w.sub])
    ret, t = pkg.problems.read_restart(3, w2, u, directory=str(tmp_path))
    assert ret == 0 and t == 0.75
    for a, b in zip(w.sub, w2.sub):
        assert (a - b).abs().max().item() <= 4e-16 * a.abs().max().item()
    assert torch.equal(w.sub[5], w2.sub[5])                    # tracers are stored unscaled
    # a file for another grid is refused, like the reference's dimension check
    u.nx += 1
    assert pkg.problems.read_restart(3, w2, u, directory=str(tmp_path))[0] == -1
    assert pkg.problems.read_restart(4, w2, u, directory=str(tmp_path))[0] == -1
    with open(tmp_path / "junk.eb200", "wb") as fp:
        fp.write(b"not a solution file")
    with pytest.raises(ValueError):
        pkg.problems.read_solution(str(tmp_path / "junk.eb200"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _writer(rank, world, port_no, n, nchem, outdir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = load_package()
    u = pkg.EulerData(nchem=nchem)
    u.nx, u.ny, u.nz = n
    pkg.problems.configure("hurricane_xy", u)
    rc, dims, coords, ext, nbr = pkg.dims_and_extents(world, rank, n, u.bcs)
    assert rc == 0
    u.myid, u.nprocs = rank, world
    u.is_, u.ie, u.js, u.je, u.ks, u.ke = ext
    u.nxl, u.nyl, u.nzl = ext[1] - ext[0] + 1, ext[3] - ext[2] + 1, ext[5] - ext[4] + 1
    u.dx, u.dy, u.dz = (u.xr - u.xl) / n[0], (u.yr - u.yl) / n[1], (u.zr - u.zl) / n[2]
    N = u.nxl * u.nyl * u.nzl
    w = pkg.ManyVector([torch.zeros(N, dtype=torch.float64) for _ in range(5)] + [torch.zeros(N * nchem, dtype=torch.float64)])
    assert pkg.problems.initial_conditions("hurricane_xy", 0.0, w, u) == 0
    assert pkg.problems.output_solution(0.5, w, u, 1, directory=outdir) == 0
    # and back: every rank restarts its own sub-box from the shared file
    w2 = pkg.ManyVector([torch.zeros_like(s) for s in w.sub])
    ret, t = pkg.problems.read_restart(1, w2, u, directory=outdir)
    assert ret == 0 and t == 0.5 and all(torch.equal(a, b) for a, b in zip(w.sub, w2.sub))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_write_one_file(pkg, tmp_path):
    import torch.multiprocessing as mp
    n, nchem = (12, 10, 6), 3
    multi, single = tmp_path / "multi", tmp_path / "single"
    multi.mkdir(), single.mkdir()
    mp.spawn(_writer, args=(2, _free_port(), n, nchem, str(multi)), nprocs=2, join=True)
    u, w = python_state(pkg, "hurricane_xy", n, nchem, 0.0)
    assert pkg.problems.output_solution(0.5, w, u, 1, directory=str(single)) == 0
    name = pkg.problems.solution_name(1)
    assert (multi / name).read_bytes() == (single / name).read_bytes()
    sol = pkg.problems.read_solution(str(multi / name))
    assert sol["Chemical-000"].shape == (6, 10, 12) and sol["Chemical-000"].max() == 1.0
