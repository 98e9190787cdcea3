"""The oracle is pinned: the C restatement (oracle/euler_oracle.c) must agree BIT FOR BIT
with (a) golden vectors produced by the unmodified reference (tests/golden, see
make_golden.py) and (b) the unmodified reference compiled into oracle/_ref when present."""
import glob
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
P, N, D, R = 0, 1, 2, 3


def load_case(path):
    z = np.load(path)
    nvar = int(z["nvar"])
    n = tuple(int(x) for x in z["n"])
    w = [np.ascontiguousarray(z["w%d" % f]) for f in range(5)] + [np.ascontiguousarray(z["w5"]) if nvar > 5 else None]
    wdot = [z["wdot%d" % f] for f in range(5)] + [z["wdot5"] if nvar > 5 else None]
    box = z["box"]
    d = [(box[1] - box[0]) / n[0], (box[3] - box[2]) / n[1], (box[5] - box[4]) / n[2]]
    return dict(n=n, nvar=nvar, nchem=nvar - 5, bcs=[int(b) for b in z["bcs"]], box=box, d=d,
                gamma=float(z["gamma"]), forcing=[float(x) for x in z["forcing"]], w=w, wdot=wdot)


FEULER_FILES = sorted(glob.glob(os.path.join(GOLD, "feuler_*.npz")))


def test_golden_files_present():
    assert len(FEULER_FILES) >= 7
    for nm in ("face_flux_nvar5.npz", "face_flux_nvar15.npz", "exchange_nvar5.npz", "exchange_nvar7.npz",
               "decomp_tables.npz"):
        assert os.path.exists(os.path.join(GOLD, nm))


@pytest.mark.parametrize("path", FEULER_FILES, ids=[os.path.basename(p)[7:-4] for p in FEULER_FILES])
def test_port_feuler_bit_exact_vs_golden(port, path):
    c = load_case(path)
    cfg = port.cfg(c["n"], c["nchem"], c["d"], c["gamma"], c["bcs"], forcing=c["forcing"])
    ret, got, mask = port.feuler(cfg, c["w"])
    assert ret == 0 and mask == 0
    for a, b in zip(got, c["wdot"]):
        if b is not None:
            assert np.array_equal(a, b)


@pytest.mark.parametrize("nvar", [5, 15])
def test_port_face_flux_bit_exact_vs_golden(port, nvar):
    z = np.load(os.path.join(GOLD, "face_flux_nvar%d.npz" % nvar))
    for s, idir, f in zip(z["stencil"], z["idir"], z["flux"]):
        assert np.array_equal(port.face_flux(s, int(idir), float(z["gamma"])), f)


@pytest.mark.parametrize("nvar", [5, 7])
def test_port_halo_layers_vs_reference_exchange(port, nvar):
    """pack_send of the neighbour == what the reference's ExchangeStart/End delivered."""
    z = np.load(os.path.join(GOLD, "exchange_nvar%d.npz" % nvar))
    n = tuple(int(x) for x in z["n"])
    Ntot = n[0] * n[1] * n[2]
    idx = np.arange(Ntot)
    i, j, k = idx % n[0], (idx // n[0]) % n[1], idx // (n[0] * n[1])
    enc = lambda v: 0.001 * v + 1e-6 * i + 1e-9 * j + 1e-12 * k
    full = [enc(v).reshape(n[2], n[1], n[0]) for v in range(nvar)]

    def block(ext):
        sl = (slice(ext[4], ext[5] + 1), slice(ext[2], ext[3] + 1), slice(ext[0], ext[1] + 1))
        parts = [np.ascontiguousarray(full[v][sl]).ravel() for v in range(5)]
        parts.append(np.ascontiguousarray(np.stack([full[v][sl] for v in range(5, nvar)], axis=-1)).ravel()
                     if nvar > 5 else None)
        nl = (ext[1] - ext[0] + 1, ext[3] - ext[2] + 1, ext[5] - ext[4] + 1)
        return parts, nl

    for nprocs in (2, 8):
        exts = [list(z["p%d_r%d_ext" % (nprocs, r)]) for r in range(nprocs)]
        for r in range(nprocs):
            nbr = list(z["p%d_r%d_nbr" % (nprocs, r)])
            for f in range(6):
                src = nbr[f]
                parts, nl = block(exts[src])
                cfg = port.cfg(nl, nvar - 5, (1, 1, 1), 1.4, [P] * 6)
                sent = port.pack_send(cfg, parts, f ^ 1)       # my W ghost = neighbour's E edge
                assert np.array_equal(sent, z["p%d_r%d_recv%d" % (nprocs, r, f)])


@pytest.mark.parametrize("nvar", [5, 7, 9, 11, 15])
def test_port_bit_exact_vs_live_reference(oracle_mod, port, nvar):
    if not oracle_mod.have_ref(nvar):
        pytest.skip("oracle/_ref not built (no reference tree on this machine)")
    ref = oracle_mod.Ref(nvar)
    n = (11, 7, 9)
    for bcs in ([P] * 6, [N] * 6, [R] * 6, [D, D, N, N, P, P], [N, N, R, R, P, P]):
        w = oracle_mod.random_state(n, nvar - 5, seed=nvar + sum(bcs))
        ret_r, wr, _, _ = ref.feuler(n, (0, 1, 0, 1, 0, 1), bcs, 1.4, w, forcing=[0, 0, -0.1, 0, 0])
        cfg = port.cfg(n, nvar - 5, [1.0 / n[0], 1.0 / n[1], 1.0 / n[2]], 1.4, bcs, forcing=[0, 0, -0.1, 0, 0])
        ret_p, wp, _ = port.feuler(cfg, w)
        assert ret_r == ret_p == 0
        for a, b in zip(wp, wr):
            if b is not None:
                assert np.array_equal(a, b, equal_nan=True)    # Dirichlet gives NaN in both


def test_reference_multirank_equals_single_rank(oracle_mod):
    """Decomposition invariance of the reference itself: virtual ranks 2,4,8."""
    if not oracle_mod.have_ref(7):
        pytest.skip("oracle/_ref not built")
    ref = oracle_mod.Ref(7)
    n = (12, 12, 12)
    w = oracle_mod.random_state(n, 2, seed=9)
    one = ref.feuler(n, (0, 1) * 3, [P, P, R, R, N, N], 1.4, w)
    for nprocs, dims in ((2, (2, 1, 1)), (4, (2, 2, 1)), (8, (2, 2, 2))):
        many = ref.feuler(n, (0, 1) * 3, [P, P, R, R, N, N], 1.4, w, nprocs=nprocs)
        assert many[3] == dims
        for a, b in zip(many[1], one[1]):
            assert np.array_equal(a, b)


def test_port_illegal_state_and_stability(oracle_mod, port):
    n = (8, 6, 5)
    w = oracle_mod.random_state(n, 0, seed=1)
    cfg = port.cfg(n, 0, (0.1, 0.2, 0.3), 1.4, [P] * 6)
    alpha = port.max_wavespeed(cfg, w)
    u = np.abs(w[1] / w[0])
    p = 0.4 * (w[4] - 0.5 * (w[1] ** 2 + w[2] ** 2 + w[3] ** 2) / w[0])
    assert alpha == pytest.approx(np.max(u + np.sqrt(1.4 * p / w[0])), rel=1e-14)
    assert port.dt_stab(cfg, 0.5, alpha) == pytest.approx(0.5 * 0.1 / alpha, rel=1e-15)
    if oracle_mod.have_ref(5):
        ret, dt = oracle_mod.Ref(5).stability(n, (0, 0.8, 0, 1.2, 0, 1.5), [P] * 6, 1.4, 0.5, w)
        assert ret == 0 and dt == port.dt_stab(cfg, 0.5, alpha)
    w[0][17] = -1.0
    ret, _, mask = port.feuler(cfg, w)
    assert ret == -1 and mask & 1


def test_port_properties(oracle_mod, port):
    """Constant state -> wdot = forcing exactly (compile_test.cpp:45-49); periodic -> the
    flux differences telescope (conservation, io.cpp:504-541)."""
    n = (9, 8, 7)
    Ncell = n[0] * n[1] * n[2]
    const = [np.full(Ncell, v) for v in (1.3, 0.2, -0.1, 0.4, 3.0)] + [np.full(Ncell * 2, 0.7)]
    cfg = port.cfg(n, 2, (0.1, 0.1, 0.1), 1.4, [P] * 6, forcing=[0, 0, -0.1, 0, 0])
    ret, wd, _ = port.feuler(cfg, const)
    assert ret == 0
    for f, g in enumerate([0, 0, -0.1, 0, 0]):
        assert np.all(wd[f] == g)
    assert np.all(wd[5] == 0)
    w = oracle_mod.random_state(n, 2, seed=4)
    cfg = port.cfg(n, 2, (0.1, 0.1, 0.1), 1.4, [P] * 6)
    ret, wd, _ = port.feuler(cfg, w)
    for a in wd:
        assert abs(a.sum()) <= 1e-11 * np.abs(a).sum()
