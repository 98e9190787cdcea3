"""The explicit driver loop and problem plug-ins on the CPU:
the ERKStep loop driven by the oracle right-hand side.  Checks what the reference's own
diagnostics would show: small errors against the analytic solutions, conservation to
round-off for periodic problems, and the tabulated Sod star state."""
import math

import numpy as np
import pytest

from helpers import NpVec, OracleVecOps

P, N, D, R = 0, 1, 2, 3


def test_exact_riemann_star_state(pkg):
    """Sod tube (rhoL,pL,rhoR,pR) = (1,1,0.125,0.1), gamma 1.4: p* = 0.30313, u* = 0.92745,
    rho left/right of the contact 0.42632 / 0.26557 (textbook values, Toro table 4.2)."""
    sol = pkg.problems.exact_riemann(0.2, [0.1, 0.6, 0.8, 0.95], 0.5, 1.4)
    assert sol[0] == (1.0, 0.0, 1.0)
    assert sol[1][0] == pytest.approx(0.42632, abs=2e-5) and sol[1][1] == pytest.approx(0.92745, abs=2e-5)
    assert sol[1][2] == pytest.approx(0.30313, abs=2e-5)
    assert sol[2][0] == pytest.approx(0.26557, abs=2e-5) and sol[2][2] == pytest.approx(0.30313, abs=2e-5)
    assert sol[3] == (0.125, 0.0, 0.1)


def advection_state(n, axis, t=0.0):
    d = [1.0 / n[0], 1.0 / n[1], 1.0 / n[2]]
    idx = np.arange(n[0] * n[1] * n[2])
    c = [(idx % n[0] + 0.5) * d[0], ((idx // n[0]) % n[1] + 0.5) * d[1], (idx // (n[0] * n[1]) + 0.5) * d[2]]
    rho = 1.0 + 0.1 * np.sin(2 * math.pi * (c[axis] - 0.5 * t))
    m = [0.5 * rho if a == axis else np.zeros_like(rho) for a in range(3)]
    et = 1.0 / 0.4 + 0.5 * (m[0] ** 2 + m[1] ** 2 + m[2] ** 2) / rho
    return [rho] + m + [et], d


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_linear_advection_with_oracle_rhs(pkg, port, axis):
    """linear_advection_{x,y,z}: errors vs rho = 1 + 0.1 sin(2 pi (s - 0.5 t)) small, mass and
    energy conserved to round-off (periodic), fifth-order spatial convergence visible."""
    errs = []
    for m in (16, 32):
        n = [3, 3, 3]
        n[axis] = m
        parts, d = advection_state(n, axis)
        ops = OracleVecOps(port, None, n, 0, d, 1.4, [P] * 6)
        opts = pkg.driver.ARKODEParameters(order=4, rtol=1e-9, atol=1e-12)
        w = NpVec(parts)
        step = pkg.driver.ERKStep(ops, 0.0, w, opts)
        mass0, en0 = parts[0].sum(), parts[4].sum()
        ret, t = step.evolve(0.25)
        assert ret == 0 and t == 0.25
        # the vector handed to the constructor holds the solution afterwards, as ARKStepEvolve's does
        assert step.w is w and all(a is b for a, b in zip(w.sub, step.w.sub))
        true, _ = advection_state(n, axis, t=0.25)
        errs.append(max(np.abs(step.w.sub[0] - true[0]).max(), 1e-300))
        assert abs(step.w.sub[0].sum() - mass0) <= 1e-13 * mass0
        assert abs(step.w.sub[4].sum() - en0) <= 1e-13 * en0
        st = step.stats()
        assert st["nfe"] >= 5 * st["nst"] and st["nst"] > 2
    assert errs[1] < 2e-5
    assert errs[0] / errs[1] > 12.0          # >= ~4th order observed between 16 and 32 cells


def test_sod_with_oracle_rhs(pkg, port):
    """sod_x at 100x3x3 to t = 0.1: L2 error against the exact Riemann solution is at the
    first-order-at-shocks level the reference prints (errR ~ 1e-2), no illegal states."""
    n = (100, 3, 3)
    d = (0.01, 1.0 / 3, 1.0 / 3)
    idx = np.arange(900)
    x = (idx % 100 + 0.5) * 0.01
    rho = np.where(x < 0.5, 1.0, 0.125)
    p = np.where(x < 0.5, 1.0, 0.1)
    parts = [rho, np.zeros(900), np.zeros(900), np.zeros(900), p / 0.4]
    ops = OracleVecOps(port, None, n, 0, d, 1.4, [N] * 6)
    step = pkg.driver.ERKStep(ops, 0.0, NpVec(parts), pkg.driver.ARKODEParameters(order=4, rtol=1e-5, atol=1e-12))
    ret, t = step.evolve(0.1)
    assert ret == 0
    sol = pkg.problems.exact_riemann(0.1, list((np.arange(100) + 0.5) * 0.01), 0.5, 1.4)
    rho_true = np.array([s[0] for s in sol])[idx % 100]
    errR = np.sqrt(np.mean((step.w.sub[0] - rho_true) ** 2))
    assert errR < 1.5e-2
    assert step.w.sub[0].min() > 0.1 and np.all(step.w.sub[2] == 0) and np.all(step.w.sub[3] == 0)


def test_fixed_step_and_cfl_hook(pkg, port):
    n = (24, 3, 3)
    parts, d = advection_state(n, 0)
    ops = OracleVecOps(port, None, n, 0, d, 1.4, [P] * 6)
    fixed = pkg.driver.ERKStep(ops, 0.0, NpVec([p.copy() for p in parts]),
                               pkg.driver.ARKODEParameters(order=3, fixedstep=1, hmax=0.01))
    assert fixed.evolve(0.1) == (0, 0.1)
    assert fixed.stats()["nst"] == 10 and fixed.stats()["netf"] == 0
    ops.cfl = 0.3
    lim = pkg.driver.ERKStep(ops, 0.0, NpVec([p.copy() for p in parts]),
                             pkg.driver.ARKODEParameters(order=4, rtol=1e-2, atol=1e-2), cfl=0.3)
    assert lim.evolve(0.1)[0] == 0
    dt_stab = ops.stability(NpVec(parts), 0.0)[1]
    assert lim.stats()["nst"] >= int(0.1 / dt_stab)      # the CFL bound, not the tolerance, set the steps


def test_fluid_blast_initial_condition_matches_reference(pkg):
    """problems.py's fluid_blast state (std::mt19937_64 clumps + central blast, restated from
    fluid_blast.cpp:65-267) == the state the UNMODIFIED reference initial_conditions() produced
    (tests/golden/ic_fluid_blast.npz, written by oracle/_ref/fluid_blast_ic)."""
    import os
    import torch
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ic_fluid_blast.npz"))
    n = tuple(int(x) for x in z["n"])
    u = pkg.EulerData()
    u.nx, u.ny, u.nz = n
    pkg.problems.configure("fluid_blast", u)
    u.dx, u.dy, u.dz = 1.0 / n[0], 1.0 / n[1], 1.0 / n[2]          # what SetupDecomp would set on one rank
    u.nxl, u.nyl, u.nzl = n
    u.is_ = u.js = u.ks = 0
    w = pkg.ManyVector([torch.zeros(n[0] * n[1] * n[2], dtype=torch.float64) for _ in range(5)])
    assert pkg.problems.initial_conditions("fluid_blast", 0.0, w, u) == 0
    for f in range(5):
        ref = z["w%d" % f]
        assert np.abs(w.sub[f].numpy() - ref).max() <= 4e-16 * max(np.abs(ref).max(), 1e-300)
    # the generator itself: first draws of std::mt19937_64(5489) are 14514284786278117030, 4620546740167642908
    g = pkg.problems.MT19937_64(5489)
    assert g.next() == 14514284786278117030 and g.next() == 4620546740167642908


class _OdeOps:
    """VecOps on a 2-component numpy vector with y' = (-y0^2 + sin t, y0 y1): enough to
    observe the order of a Runge-Kutta table without the fluid RHS."""

    def new_like(self, w):
        from helpers import NpVec
        return NpVec([np.empty_like(s) for s in w.sub])

    def lincomb(self, out, coefs, vecs):
        acc = sum(c * v.sub[0] for c, v in zip(coefs, vecs))
        out.sub[0][...] = acc

    def wrms(self, x, y, rtol, atol):
        q = x.sub[0] / (rtol * np.abs(y.sub[0]) + atol)
        return float(np.sqrt(np.mean(q * q)))

    def rhs(self, t, w, wdot):
        y = w.sub[0]
        wdot.sub[0][...] = [-y[0] * y[0] + np.sin(t), y[0] * y[1]]
        return 0


def _order_residuals(A, b):
    """Rooted-tree order conditions up to order 5 (17 of them); returns {order: max residual}."""
    s = len(b)
    M = np.zeros((s, s))
    for i, row in enumerate(A):
        M[i, :len(row)] = row
    b = np.asarray(b, dtype=float)
    c = M.sum(axis=1)
    Ac, Ac2, Ac3, AAc, AAc2, AAAc = M @ c, M @ c**2, M @ c**3, M @ (M @ c), M @ (M @ c**2), M @ (M @ (M @ c))
    cond = {1: [(b.sum(), 1.0)], 2: [(b @ c, 1 / 2)], 3: [(b @ c**2, 1 / 3), (b @ Ac, 1 / 6)],
            4: [(b @ c**3, 1 / 4), (b @ (c * Ac), 1 / 8), (b @ Ac2, 1 / 12), (b @ AAc, 1 / 24)],
            5: [(b @ c**4, 1 / 5), (b @ (c**2 * Ac), 1 / 10), (b @ (Ac * Ac), 1 / 20), (b @ (c * Ac2), 1 / 15),
                (b @ Ac3, 1 / 20), (b @ (c * AAc), 1 / 30), (b @ (M @ (c * Ac)), 1 / 40), (b @ AAc2, 1 / 60),
                (b @ AAAc, 1 / 120)]}
    return {p: max(abs(x - y) for x, y in v) for p, v in cond.items()}


@pytest.mark.parametrize("tid", [0, 1, 3, 6, 7, 8, 10, 11, 12])
def test_erk_tables_by_id_satisfy_their_order_conditions(pkg, tid):
    """ARKStepSetTableNum ids (euler3D_main.cpp:211-212): the method has order p, the
    embedding order q -- and not one more."""
    A, b, bhat, p, q = pkg.driver.TABLES_BY_ID[tid]
    res = _order_residuals(A, b)                       # all 17 conditions up to order 5
    assert all(res[o] < 1e-14 for o in range(1, min(p, 5) + 1)), res
    assert p >= 5 or res[p + 1] > 1e-6
    if bhat is not None:
        rh = _order_residuals(A, bhat)
        assert all(rh[o] < 1e-14 for o in range(1, min(q, 5) + 1)), rh
        assert q >= 5 or rh[q + 1] > 1e-6
    if p > 5:                                          # beyond: the quadrature and tall-tree conditions
        M = np.zeros((len(b), len(b)))
        for i, row in enumerate(A):
            M[i, :len(row)] = row
        c, bb = M.sum(axis=1), np.asarray(b)
        assert all(abs(bb @ c ** (o - 1) - 1.0 / o) < 1e-14 for o in range(6, p + 1))
        tall, fact = c.copy(), 1.0
        for o in range(2, p + 1):
            fact *= o
            assert abs(bb @ tall - 1.0 / fact) < 1e-14
            tall = M @ tall
        assert abs(bb @ c ** p - 1.0 / (p + 1)) > 1e-7 or tid == 11      # (Fehlberg 7(8) integrates quadratures exactly)


@pytest.mark.parametrize("tid", [3, 6, 8, 10, 11, 12])
def test_erk_tables_observed_order(pkg, tid):
    from helpers import NpVec
    p = pkg.driver.TABLES_BY_ID[tid][3]
    sols = []
    for h in {6: (0.2, 0.1, 0.05), 8: (0.5, 0.25, 0.125)}.get(p, (0.1, 0.05, 0.025)):
        opts = pkg.driver.ARKODEParameters(order=0, etable=tid, fixedstep=1, hmax=h)
        step = pkg.driver.ERKStep(_OdeOps(), 0.0, NpVec([np.array([0.5, 1.0])]), opts)
        ret, t = step.evolve(1.0)
        assert ret == 0 and t == 1.0
        sols.append(step.w.sub[0].copy())
    opts = pkg.driver.ARKODEParameters(order=0, etable=11, fixedstep=1, hmax=1.0 / 64)      # order 8, 64 steps
    ref = pkg.driver.ERKStep(_OdeOps(), 0.0, NpVec([np.array([0.5, 1.0])]), opts)
    assert ref.evolve(1.0)[0] == 0
    errs = [np.abs(s - ref.w.sub[0]).max() for s in sols]
    rates = [np.log2(errs[i] / errs[i + 1]) for i in range(2)]
    assert min(rates) > p - 0.5, (errs, rates)


def test_order_overrides_etable_and_unknown_ids_are_refused(pkg):
    d = pkg.driver
    assert d.select_table(4, 8) is d.ZONNEVELD           # "order" overrides "etable"
    assert d.select_table(0, 8) is d.DORMAND_PRINCE
    assert d.select_table(0) is d.ZONNEVELD
    with pytest.raises(ValueError):
        d.select_table(0, 13)                             # ARK437L2SA's explicit part: not provided
    with pytest.raises(ValueError):
        d.ERKStep(_OdeOps(), 0.0, None, d.ARKODEParameters(order=0, etable=12))   # no embedding


def test_native_driver_tables_equal_the_python_tables(pkg, tmp_path):
    """host/erk_tables.hpp (native driver) against driver.py: same selection rule, same numbers."""
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "native_tables_check")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-I", os.path.join(here, "..", "sundials-manyvector-demo_b200", "host"),
                           "-o", exe, os.path.join(here, "native_tables_check.cpp")])
    d = pkg.driver
    for order, etable in [(2, -1), (3, -1), (4, -1), (5, -1), (6, -1), (8, -1), (4, 8), (0, -1)] + [(0, t) for t in sorted(d.TABLES_BY_ID)]:
        out = subprocess.run([exe, str(order), str(etable)], capture_output=True, text=True, check=True).stdout.split("\n")
        A, b, bhat, p, q = d.select_table(order, etable)
        s = len(b)
        assert [int(x) for x in out[0].split()] == [s, p, q, 0 if bhat is None else 1]
        M = np.array([[float(x) for x in out[1 + i].split()] for i in range(s)])
        for i in range(s):
            assert list(M[i, :len(A[i])]) == list(A[i]) and not M[i, len(A[i]):].any()
        assert [float(x) for x in out[1 + s].split()] == list(b)
        assert [float(x) for x in out[2 + s].split()] == (list(bhat) if bhat is not None else [0.0] * s)
    for order, etable in [(7, -1), (0, 13), (0, 2)]:
        assert subprocess.run([exe, str(order), str(etable)], capture_output=True, text=True).stdout.strip() == "none"
