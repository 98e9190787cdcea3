"""GPU tier: golden vectors of the reference, ghost-layer semantics of every boundary
type (the digit-encoding check of communication_test_main.cpp extended beyond periodic),
stability, and size-independent properties at BASELINE.json's grid sizes."""
import glob
import os

import numpy as np
import pytest

from conftest import normwise_errors, rounding_floor
from helpers import gpu_feuler, make_udata
from test_oracle import FEULER_FILES, load_case

pytestmark = pytest.mark.gpu
P, N, D, R = 0, 1, 2, 3
TOL = 1e-12


@pytest.mark.parametrize("path", FEULER_FILES, ids=[os.path.basename(p)[7:-4] for p in FEULER_FILES])
def test_cuda_feuler_vs_reference_golden(pkg, path):
    c = load_case(path)
    u = make_udata(pkg, c["n"], c["nchem"], c["bcs"], box=tuple(c["box"]), gamma=c["gamma"], forcing=c["forcing"])
    ret, got = gpu_feuler(pkg, u, c["w"])
    assert ret == 0, u.last_error()
    assert max(normwise_errors(got, c["wdot"], rounding_floor(c["w"], c["gamma"], c["d"]))) <= TOL
    u.FreeData()


def encoded_state(n, nchem):
    """value = 0.sXXXyyyZZZ-style encoding of (field, i, j, k) (communication_test_main.cpp:133-181)."""
    Ncell = n[0] * n[1] * n[2]
    idx = np.arange(Ncell)
    i, j, k = idx % n[0], (idx // n[0]) % n[1], idx // (n[0] * n[1])
    enc = lambda v: 0.001 * (v + 1) + 1e-6 * i + 1e-9 * j + 1e-12 * k
    parts = [enc(v) for v in range(5)]
    parts.append(np.stack([enc(5 + v) for v in range(nchem)], axis=1).ravel() if nchem else None)
    return parts


@pytest.mark.parametrize("bcs", [[P] * 6, [N] * 6, [R] * 6, [D] * 6, [N, N, R, R, P, P], [R, R, D, D, N, N]])
@pytest.mark.parametrize("n,nchem", [((12, 10, 8), 2), ((3, 9, 7), 0), ((5, 3, 4), 4)])
def test_ghost_layers_every_bc_type_exact(pkg, port, bcs, n, nchem):
    """Ghost layers as the kernel resolves them == the reference's receive buffers, exactly
    (pure copies / sign flips): low side mirrors, high side copies, reflecting flips the normal
    momentum, Dirichlet flips everything (euler3D.hpp:797-1166)."""
    import torch
    u = make_udata(pkg, n, nchem, bcs)
    parts = encoded_state(n, nchem)
    w = pkg.ManyVector([torch.from_numpy(p).cuda() for p in parts if p is not None])
    cfg = port.cfg(n, nchem, (u.dx, u.dy, u.dz), 1.4, bcs)
    assert u.ExchangeStart(w) == 0 and u.ExchangeEnd() == 0
    for f in range(6):
        got = u.recv_buffer(w, f).cpu().numpy()
        want = port.fill_ghost(cfg, parts, f)
        assert np.array_equal(got, want), "face %d" % f
    u.FreeData()


def test_stability_matches_oracle(pkg, oracle_mod, port):
    import torch
    n = (40, 24, 20)
    u = make_udata(pkg, n, 0, [P] * 6, box=(0, 0.8, 0, 1.2, 0, 1.5))
    u.cfl = 0.5
    parts = oracle_mod.random_state(n, 0, seed=8)
    w = pkg.ManyVector([torch.from_numpy(p).cuda() for p in parts if p is not None])
    ret, dt = pkg.stability(w, 0.0, u)
    cfg = port.cfg(n, 0, (u.dx, u.dy, u.dz), 1.4, [P] * 6)
    want = port.dt_stab(cfg, 0.5, port.max_wavespeed(cfg, parts))
    assert ret == 0 and dt == pytest.approx(want, rel=1e-14)
    u.FreeData()


def test_dirichlet_is_nan_like_the_reference(pkg, oracle_mod, port):
    """With Dirichlet ghosts rho<0 at the boundary face, SUNRsqrt gives 0 and the reference
    divides by it (utilities.cpp:298-304): non-finite wdot next to the boundary, finite inside.
    The corner is unpinned by the reference's own tests; we only check the same cells are hit."""
    n = (12, 10, 9)
    u = make_udata(pkg, n, 0, [D, D, P, P, P, P])
    parts = oracle_mod.random_state(n, 0, seed=2)
    ret, got = gpu_feuler(pkg, u, parts)
    _, ref, _ = port.feuler(port.cfg(n, 0, (u.dx, u.dy, u.dz), 1.4, u.bcs), parts)
    for a, b in zip(got[:5], ref[:5]):
        assert np.array_equal(np.isfinite(a), np.isfinite(b))
        m = np.isfinite(b)
        assert np.abs(a[m] - b[m]).max() <= TOL * np.abs(b[m]).max()
    u.FreeData()


@pytest.mark.parametrize("n,nchem", [((512, 512, 512), 0), ((256, 256, 256), 10), ((3, 4096, 4096), 0)])
def test_properties_at_baseline_sizes(pkg, n, nchem):
    """Size-independent checks at BASELINE.json's full sizes (where the CPU oracle is too slow):
    (1) a constant state is a fixed point: wdot == forcing exactly (compile_test.cpp:45-49);
    (2) periodic: the divergence telescopes, sum(wdot) ~ 0 (conservation, io.cpp:504-541);
    (3) windowed parity: a 20^3 window cut out with its halo agrees with the oracle."""
    import torch
    import oracle
    bcs = [P] * 6
    u = make_udata(pkg, n, nchem, bcs, forcing=[0, 0, -0.1, 0, 0])
    Ncell = n[0] * n[1] * n[2]
    vals = (1.3, 0.2, -0.1, 0.4, 3.0)
    w = pkg.ManyVector([torch.full((Ncell,), v, dtype=torch.float64, device="cuda") for v in vals] +
                       ([torch.full((Ncell * nchem,), 0.7, dtype=torch.float64, device="cuda")] if nchem else []))
    wdot = pkg.ManyVector.new(u)
    assert pkg.fEuler(0.0, w, wdot, u) == 0
    for f, g in enumerate([0, 0, -0.1, 0, 0]):
        assert bool((wdot.sub[f] == g).all())
    if nchem:
        assert bool((wdot.sub[5] == 0).all())
    # random admissible state
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    U = lambda m: torch.rand(m, generator=g, device="cuda", dtype=torch.float64)
    rho = 1 + 0.5 * U(Ncell); vx, vy, vz = (0.3 * (U(Ncell) - 0.5) for _ in range(3)); p = 1 + 0.5 * U(Ncell)
    subs = [rho, rho * vx, rho * vy, rho * vz, p / 0.4 + 0.5 * rho * (vx * vx + vy * vy + vz * vz)]
    if nchem:
        subs.append(U(Ncell * nchem))
    w = pkg.ManyVector(subs)
    u.FreeData()
    u = make_udata(pkg, n, nchem, bcs)
    assert pkg.fEuler(0.0, w, wdot, u) == 0
    for s in wdot.sub:
        assert abs(float(s.sum())) <= 1e-10 * float(s.abs().sum())
    # windowed parity: oracle on a sub-box with >= 3 cells of margin
    m = [min(26, x) for x in n]
    o = [max(0, (x - mm) // 2) for x, mm in zip(n, m)]
    sl = (slice(o[2], o[2] + m[2]), slice(o[1], o[1] + m[1]), slice(o[0], o[0] + m[0]))
    win = [s.view(n[2], n[1], n[0])[sl].contiguous().view(-1).cpu().numpy() for s in w.sub[:5]]
    win.append(w.sub[5].view(n[2], n[1], n[0], nchem)[sl].contiguous().view(-1).cpu().numpy() if nchem else None)
    port = oracle.Port()
    ret, ref, _ = port.feuler(port.cfg(m, nchem, (u.dx, u.dy, u.dz), 1.4, [N] * 6), win)
    assert ret == 0
    inner = tuple(slice(3, mm - 3) if mm == 26 else slice(0, mm) for mm in reversed(m))
    for f in range(5 + (1 if nchem else 0)):
        shape = (n[2], n[1], n[0]) + ((nchem,) if f == 5 else ())
        mshape = (m[2], m[1], m[0]) + ((nchem,) if f == 5 else ())
        got = wdot.sub[f].view(shape)[sl].cpu().numpy()[inner]
        want = ref[f].reshape(mshape)[inner]
        if n[0] == 3:      # thin x: the window's Neumann x-ghosts differ from the periodic wrap
            continue
        assert np.abs(got - want).max() <= TOL * np.abs(want).max()
    u.FreeData()
