"""CPU-side logic tests of the CUDA kernel SOURCE (rhs_kernel.cuh + host_setup.h compiled
through tests/emu/cuda_emu.h): ghost maps, shared-memory exchange, barrier placement,
sub-box launches and halo-buffer reads, against the oracle.  The emulation is test
infrastructure only -- the product has no CPU path; the real parity tests are the
``-m gpu`` ones.  Tolerance as there: normwise 1e-12."""
import numpy as np
import pytest

from conftest import normwise_errors, rounding_floor

P, N, D, R = 0, 1, 2, 3
NO = -1


@pytest.fixture(scope="module")
def emu(pkg):
    from emu.emu import Emu
    return Emu(pkg)


def nbr_single(bcs):
    return [0 if b == P else NO for b in bcs]


@pytest.mark.parametrize("n,nchem,bcs,threads", [
    ((12, 9, 7), 0, [P] * 6, 256),
    ((12, 9, 7), 2, [N] * 6, 384),
    ((12, 9, 7), 3, [R] * 6, 256),          # odd nchem: scalar tracer path
    ((35, 10, 5), 2, [P, P, R, R, N, N], 128),
    ((3, 20, 17), 2, [N] * 6, 256),          # thin x
    ((40, 3, 3), 0, [N] * 6, 256),           # sod_x shape
    ((7, 6, 26), 4, [R, R, P, P, N, N], 64), # several z-segments
    ((70, 12, 10), 2, [P] * 6, 128),         # interior CTAs: the no-ghost path reading aux arrays
    ((33, 9, 4), 24, [N] * 6, 384),          # NVAR = 29: the tile is flattened to fit shared memory
])
def test_emulated_kernel_matches_oracle(emu, oracle_mod, port, n, nchem, bcs, threads):
    w = oracle_mod.random_state(n, nchem, seed=sum(n))
    d = (1.0 / n[0], 2.0 / n[1], 0.5 / n[2])
    forcing = [0, 0, -0.1, 0, 0]
    ret, got, bits = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, forcing=forcing, threads=threads)
    ret_ref, ref, _ = port.feuler(port.cfg(n, nchem, d, 1.4, bcs, forcing=forcing), w)
    assert ret == 0 and ret_ref == 0 and bits == 0
    floor = rounding_floor(w, 1.4, d)
    assert max(normwise_errors(got, ref, floor)) <= 1e-12
    # same answer with every derived value computed on the fly (no aux arrays)
    ret, got2, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, forcing=forcing, threads=threads, use_aux=0)
    assert ret == 0 and max(normwise_errors(got2, ref, floor)) <= 1e-12
    # and with the instantiation for boundary-heavy launches (AG: boundary tiles read the per-cell
    # arrays for owned points and for ghost points that only differ in the sign of a momentum)
    ret, got3, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, forcing=forcing, threads=threads, aux_in_gen=1)
    assert ret == 0 and max(normwise_errors(got3, ref, floor)) <= 1e-12


@pytest.mark.parametrize("nvar", [5, 15])
def test_face_arithmetic_vs_reference_face_flux_golden(emu, nvar):
    """SURVEY.md 8(a2) face by face: the product's face arithmetic (euler_math.cuh: cell_aux, fluid_face,
    tracer_face -- the re-derived formulation, not face_flux recompiled) against the 60 + 60 faces the unmodified
    reference's face_flux produced (tests/golden/face_flux_nvar*.npz, all three directions), per field relative to
    the largest flux of that field: a flux is a sum of O(1) terms, no divergence has cancelled anything yet, so
    the distance is a few ulps (measured 6.5e-16 at NVAR = 5, 5.8e-15 at NVAR = 15) -- the bar here is 1e-14."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "face_flux_nvar%d.npz" % nvar))
    got = np.array([emu.face_flux(s, int(idir), float(z["gamma"])) for s, idir in zip(z["stencil"], z["idir"])])
    ref = np.array(z["flux"])
    assert sorted(set(int(d) for d in z["idir"])) == [0, 1, 2]
    err = np.abs(got - ref).max(axis=0) / np.abs(ref).max(axis=0)
    assert err.max() <= 1e-14, err


def _golden_ids():
    import os
    from test_oracle import FEULER_FILES
    return FEULER_FILES, [os.path.basename(p)[7:-4] for p in FEULER_FILES]


@pytest.mark.parametrize("path", _golden_ids()[0], ids=_golden_ids()[1])
def test_emulated_kernel_vs_reference_golden(emu, path):
    """The reference's golden fEuler cases through the emulated kernel with the bar of the GPU test
    (tests/test_gpu_golden_and_halo.py::test_cuda_feuler_vs_reference_golden), so that an arithmetic change that
    costs parity margin shows on the CPU tier first.  The emulation rounds as the GPU does except where nvcc
    contracts a product and a sum that g++ -ffp-contract=off leaves apart; the smooth advection case, whose
    right-hand side is rounding noise of the flux terms (the reference's own FMA self-noise there is 0.74e-12
    in e_t), came out at 1.19e-12 on both tiers with the shared-projection variant of fluid_face and stays at
    0.91e-12 with the default."""
    from test_oracle import load_case
    c = load_case(path)
    n = tuple(c["n"])
    ret, got, bits = emu.rhs(n, c["nchem"], c["d"], c["gamma"], c["bcs"], nbr_single(c["bcs"]), 0, c["w"],
                             forcing=list(c["forcing"]), threads=128)
    assert ret == 0 and bits == 0
    assert max(normwise_errors(got, c["wdot"], rounding_floor(c["w"], c["gamma"], c["d"]))) <= 1e-12


@pytest.mark.parametrize("bcs", [[D] * 6, [R, R, D, D, N, N], [D, D, P, P, R, R]])
def test_emulated_boundary_instantiation_with_dirichlet_ghosts(emu, oracle_mod, port, bcs):
    """AG instantiation: Dirichlet ghosts negate rho and e_t, so their 1/rho, p, c cannot come from
    the per-cell arrays; the result must equal the default instantiation wherever that is finite
    (the Dirichlet boundary cells are non-finite in the reference too)."""
    n, nchem = (34, 9, 8), 2
    w = oracle_mod.random_state(n, nchem, seed=4)
    d = (0.1, 0.2, 0.3)
    ret0, base, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, threads=128)
    ret1, got, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, threads=128, aux_in_gen=1)
    assert ret0 == ret1
    for a, b in zip(base, got):
        fin = np.isfinite(a)
        assert np.array_equal(fin, np.isfinite(b))
        assert np.abs(a[fin] - b[fin]).max() <= 1e-12 * np.abs(a[fin]).max()


@pytest.mark.parametrize("n,nchem,bcs,threads", [
    ((70, 12, 10), 2, [P] * 6, 128),          # 4 rows per tile, interior and ghost CTAs
    ((35, 25, 9), 0, [P, P, R, R, N, N], 384),  # 12 rows: the production tile shape, two tiles along y
    ((33, 5, 26), 10, [R] * 6, 64),           # 2 rows (one real + the face-only row), several z-segments
])
def test_emulated_pairwise_row_rendezvous(emu, oracle_mod, port, n, nchem, bcs, threads):
    """EULERB200_PAIR path: neighbouring warp rows meet on named barriers, FY double-buffered.
    The emulator schedules fibres barrier by barrier in alternating order, so a missing or
    mis-paired rendezvous shows up as poisoned (NaN) or stale fluxes.  Bit-identical to the
    CTA-wide-barrier path: only the synchronisation differs."""
    w = oracle_mod.random_state(n, nchem, seed=7 + sum(n))
    d = (1.0 / n[0], 2.0 / n[1], 0.5 / n[2])
    ret, base, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, threads=threads)
    for mode in (1, 2):           # two FY buffers / one buffer and a second rendezvous
        ret2, got, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, threads=threads, pair=mode)
        assert ret == 0 and ret2 == 0
        for a, b in zip(base, got):
            assert (a is None and b is None) or np.array_equal(a, b)
    ret_ref, ref, _ = port.feuler(port.cfg(n, nchem, d, 1.4, bcs), w)
    assert max(normwise_errors(got, ref, rounding_floor(w, 1.4, d))) <= 1e-12
    # thin boxes (rows are not warps) silently keep the CTA-wide barriers
    assert emu.rhs((3, 20, 17), 0, d, 1.4, [N] * 6, nbr_single([N] * 6), 0,
                   oracle_mod.random_state((3, 20, 17), 0, seed=1), threads=256, pair=2)[0] == -77


@pytest.mark.parametrize("n,nchem,bcs,threads", [
    ((12, 9, 7), 2, [N] * 6, 384),
    ((35, 10, 5), 3, [P, P, R, R, N, N], 128),
    ((70, 12, 10), 0, [P] * 6, 128),
])
def test_emulated_forcing_taken_from_wdot(emu, oracle_mod, port, n, nchem, bcs, threads):
    """eulerb200_set_forcing_in_wdot (the GW kernel instantiation): the caller has run an arbitrary
    external_forces hook into wdot (utilities.cpp:28,65) and the kernel computes wdot = wdot - div F.
    fEuler is affine in G and the reference rounds G - div once, so the expected result is exactly
    G + (oracle with zero forcing)."""
    w = oracle_mod.random_state(n, nchem, seed=3 + sum(n))
    d = (1.0 / n[0], 2.0 / n[1], 0.5 / n[2])
    rng = np.random.default_rng(5)
    N_ = n[0] * n[1] * n[2]
    G = [rng.normal(size=N_) for _ in range(5)] + [rng.normal(size=N_ * nchem) if nchem else None]
    ret_ref, ref, _ = port.feuler(port.cfg(n, nchem, d, 1.4, bcs), w)
    want = [None if r is None else g + r for g, r in zip(G, ref)]
    ret, got, bits = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, forcing=[9, 9, 9, 9, 9], threads=threads,
                             g_in_wdot=G)
    assert ret == 0 and ret_ref == 0 and bits == 0
    floor = rounding_floor(w, 1.4, d)
    assert max(normwise_errors(got, want, floor)) <= 1e-12
    # bit for bit the plain kernel's divergence: got - G == plain(zero forcing) up to the one rounding of G - div
    ret0, plain, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, threads=threads)
    for g, a, b in zip(G, got, plain):
        if g is not None:
            assert np.array_equal(a, g + b)


@pytest.mark.parametrize("n,nchem,bcs,threads,pair", [
    ((12, 9, 7), 2, [N] * 6, 384, 0),
    ((35, 25, 9), 10, [P, P, R, R, N, N], 384, 2),   # production tile shape, row rendezvous
    ((70, 12, 10), 4, [P] * 6, 128, 1),              # interior CTAs reading the per-cell arrays
    ((3, 20, 17), 3, [N] * 6, 256, 0),               # thin x, odd nchem (scalar species path)
    ((33, 5, 26), 6, [R] * 6, 64, 0),                # several z-segments
])
def test_emulated_split_launches_equal_the_fused_launch(emu, oracle_mod, port, n, nchem, bcs, threads, pair):
    """EULERB200_SPLIT: the fluid fields and the species evaluated by two launches (PART_FLUID,
    PART_TRACERS).  The species launch rebuilds the face-local alpha and the normal velocities from the
    per-cell arrays with the operations fluid_face uses, so the result is bit-identical to the fused
    launch wherever both read the same per-cell values: without the arrays, and in the boundary-heavy
    instantiation.  (In boundary tiles the default fused instantiation derives p from the sweep-ordered
    momenta, the species launch reads the array: c and alpha can differ in the last bit there.)"""
    w = oracle_mod.random_state(n, nchem, seed=5 + sum(n))
    d = (1.0 / n[0], 2.0 / n[1], 0.5 / n[2])
    forcing = [0.1, 0, -0.1, 0, 0.3]
    ret_ref, ref, _ = port.feuler(port.cfg(n, nchem, d, 1.4, bcs, forcing=forcing), w)
    for extra in (dict(use_aux=0), dict(aux_in_gen=1), dict(aux_in_gen=1, chemT=0), dict()):
        ret0, base, b0 = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, forcing=forcing, threads=threads, pair=pair, **extra)
        ret1, got, b1 = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, forcing=forcing, threads=threads, pair=pair, split=1, **extra)
        assert ret0 == 0 and ret1 == 0 and b0 == b1 == 0
        if extra:
            assert all(np.array_equal(a, b) for a, b in zip(base, got))
        assert all(np.array_equal(a, b) for a, b in zip(base[:5], got[:5]))
        assert max(normwise_errors(got, ref, rounding_floor(w, 1.4, d))) <= 1e-12
    rng = np.random.default_rng(9)
    G = [rng.normal(size=x.size) for x in base]
    _, gw0, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, threads=threads, g_in_wdot=G, use_aux=0)
    _, gw1, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, threads=threads, g_in_wdot=G, split=1, use_aux=0)
    assert all(np.array_equal(a, b) for a, b in zip(gw0, gw1))
    ws = [x.copy() for x in w]
    ws[5].reshape(-1, nchem)[:, -1] += 2.0
    _, s0, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, [x.copy() for x in ws], threads=threads, energy_units=3.0, aux_in_gen=1)
    _, s1, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, [x.copy() for x in ws], threads=threads, energy_units=3.0, split=1, aux_in_gen=1)
    assert all(np.array_equal(a, b) for a, b in zip(s0, s1))
    # an illegal state is reported by the fluid launch
    wb = [x.copy() for x in w]
    wb[0][7] = -1.0
    r0 = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, wb, threads=threads)
    r1 = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, wb, threads=threads, split=1)
    assert r0[0] == r1[0] == -1 and r0[2] == r1[2] != 0


def test_emulated_strict_build_is_bit_identical_to_the_oracle(pkg, oracle_mod, port):
    """-DEB_STRICT (csrc/strict_face.cuh): the reference's face_flux arithmetic operation for operation
    and true divisions in the divergence.  Compiled without FMA contraction the kernel source then
    reproduces the oracle BIT FOR BIT -- stencil resolution, every ghost rule, tile and z-segment seams,
    the shared-memory flux exchange, split launches, sub-boxes -- so whatever the fast build differs by
    is rounding of its re-derived arithmetic and nothing else."""
    from emu.emu import Emu
    emu_s = Emu(pkg, strict=True)
    cases = [((12, 9, 7), 0, [P] * 6, 256, {}), ((12, 9, 7), 2, [N] * 6, 384, {}), ((12, 9, 7), 3, [R] * 6, 256, {}),
             ((35, 10, 5), 2, [P, P, R, R, N, N], 128, dict(pair=1)), ((3, 20, 17), 2, [N] * 6, 256, {}),
             ((40, 3, 3), 0, [N] * 6, 256, {}), ((7, 6, 26), 4, [R, R, P, P, N, N], 64, dict(split=1)),
             ((70, 12, 10), 10, [R] * 6, 384, dict(pair=2)), ((33, 9, 4), 24, [N] * 6, 384, {})]
    for n, nchem, bcs, threads, kw in cases:
        w = oracle_mod.random_state(n, nchem, seed=sum(n))
        d = (1.0 / n[0], 2.0 / n[1], 0.5 / n[2])
        forcing = [0, 0.25, -0.1, 0, 0.5]
        ret, got, bits = emu_s.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, forcing=forcing, threads=threads, **kw)
        ret_ref, ref, _ = port.feuler(port.cfg(n, nchem, d, 1.4, bcs, forcing=forcing), w)
        assert ret == 0 and ret_ref == 0 and bits == 0
        for a, b in zip(got, ref):
            assert (a is None and b is None) or np.array_equal(a, b), (n, nchem, bcs)


def test_emulated_bulk_copy_staging_variant(emu, oracle_mod, port):
    """rhs_fused_kernel<..., STAGE> (EULERB200_STAGE=1, the TMA / bulk-copy A/B variant): interior tiles read the
    species of the current plane from the shared-memory window filled by cp.async.bulk + mbarrier (emulated:
    copy at issue, phase flip on the byte count).  Same bits as the default kernel: only where the values
    are read from differs."""
    for n, nchem, bcs in [((70, 40, 10), 10, [P] * 6), ((100, 26, 20), 2, [R, R, P, P, N, N]), ((40, 30, 9), 4, [N] * 6)]:
        w = oracle_mod.random_state(n, nchem, seed=3 + sum(n))
        d = (1.0 / n[0], 2.0 / n[1], 0.5 / n[2])
        ret0, base, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, threads=384)
        ret1, got, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, threads=384, stage=1)
        assert ret0 == 0 and ret1 == 0
        assert all(np.array_equal(a, b) for a, b in zip(base, got))
        ret_ref, ref, _ = port.feuler(port.cfg(n, nchem, d, 1.4, bcs), w)
        assert max(normwise_errors(got, ref, rounding_floor(w, 1.4, d))) <= 1e-12


def test_emulated_full_width_tiles_variant(pkg, emu, oracle_mod, port):
    """rhs_fused_kernel<..., XC>: tiles own all 32 columns and the x-faces that close a tile on the right come
    from the idle lanes of the top warp through two mbarriers (column written / column consumed).  Each face is
    the same arithmetic on the same stencil whoever computes it, so the result must have the bits of the default
    kernel -- in every synchronisation mode, for tiles cut short by the box in x and in y, boundary tiles of every
    ghost rule, the AG / GW instantiations, split launches, the slow mode, sub-boxes (interior / shell launches of a
    decomposed run) -- and the strict build must still be the oracle bit for bit."""
    from emu.emu import Emu
    cases = [((70, 23, 5), 10, [P] * 6, 384, {}), ((64, 12, 4), 2, [R, R, P, P, N, N], 384, dict(pair=1)),
             ((96, 12, 3), 4, [N] * 6, 384, dict(pair=2)), ((33, 11, 3), 3, [R] * 6, 384, dict(pair=2)),
             ((32, 7, 5), 0, [P, P, N, N, R, R], 128, dict(pair=1)), ((100, 9, 3), 2, [R] * 6, 128, {}),
             ((65, 14, 3), 2, [N, N, R, R, P, P], 384, dict(aux_in_gen=1, pair=2)),
             ((67, 13, 3), 4, [P] * 6, 384, dict(split=1, pair=2)), ((45, 12, 3), 1, [R] * 6, 384, dict(use_aux=0))]
    for n, nchem, bcs, threads, kw in cases:
        w = oracle_mod.random_state(n, nchem, seed=5 + sum(n))
        d = (1.0 / n[0], 2.0 / n[1], 0.5 / n[2])
        forcing = [0, 0.25, -0.1, 0, 0.5]
        ret0, base, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, forcing=forcing, threads=threads, **kw)
        ret1, got, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, forcing=forcing, threads=threads, xc=1, **kw)
        assert ret0 == 0 and ret1 == 0
        assert all((a is None and b is None) or np.array_equal(a, b) for a, b in zip(base, got)), (n, nchem, kw)
    # sub-boxes (the interior / shell launches of a decomposed run), against the same launch without XC
    n, nchem, bcs = (80, 20, 9), 2, [P] * 6
    w = oracle_mod.random_state(n, nchem, seed=77)
    d = (0.1, 0.1, 0.1)
    for lo, hi in [((3, 3, 3), (77, 17, 6)), ((0, 0, 0), (3, 20, 9)), ((3, 0, 0), (77, 3, 9)), ((5, 4, 2), (70, 17, 8))]:
        pair = 2 if hi[0] - lo[0] >= 31 and hi[1] - lo[1] >= 2 else 0      # (rows of thin boxes are not warps)
        r0, base, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, threads=384, pair=pair, lo=lo, hi=hi)
        r1, got, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, threads=384, pair=pair, xc=1, lo=lo, hi=hi)
        assert r0 == 0 and r1 == 0
        assert not np.isnan(base[0].reshape(n[2], n[1], n[0])[lo[2]:hi[2], lo[1]:hi[1], lo[0]:hi[0]]).any()
        sl = (slice(lo[2], hi[2]), slice(lo[1], hi[1]), slice(lo[0], hi[0]))
        for a, b in zip(got[:5], base[:5]):
            assert np.array_equal(a.reshape(n[2], n[1], n[0])[sl], b.reshape(n[2], n[1], n[0])[sl])
        assert np.array_equal(got[5].reshape(n[2], n[1], n[0], nchem)[sl], base[5].reshape(n[2], n[1], n[0], nchem)[sl])
    # hook-assigned forcing and the slow mode
    n, nchem, bcs = (66, 13, 4), 2, [R] * 6
    w = oracle_mod.random_state(n, nchem, seed=8)
    rng = np.random.default_rng(10)
    G = [rng.normal(size=x.size) for x in w]
    _, g0, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, threads=384, g_in_wdot=G, pair=2)
    _, g1, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, threads=384, g_in_wdot=G, pair=2, xc=1)
    assert all(np.array_equal(a, b) for a, b in zip(g0, g1))
    _, s0, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, [x.copy() for x in w], threads=384, energy_units=3.0)
    _, s1, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, [x.copy() for x in w], threads=384, energy_units=3.0, xc=1)
    assert all(np.array_equal(a, b) for a, b in zip(s0, s1))
    # strict build + XC: the oracle's bits
    emu_s = Emu(pkg, strict=True)
    for n, nchem, bcs, kw in [((70, 12, 4), 2, [R] * 6, dict(pair=2)), ((64, 13, 3), 0, [P, P, N, N, R, R], {})]:
        w = oracle_mod.random_state(n, nchem, seed=sum(n))
        d = (1.0 / n[0], 2.0 / n[1], 0.5 / n[2])
        ret, got, bits = emu_s.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, threads=384, xc=1, **kw)
        ret_ref, ref, _ = port.feuler(port.cfg(n, nchem, d, 1.4, bcs), w)
        assert ret == 0 and ret_ref == 0 and bits == 0
        for a, b in zip(got, ref):
            assert (a is None and b is None) or np.array_equal(a, b), (n, nchem, bcs)


def test_emulated_illegal_state_bits(emu, oracle_mod, port):
    n = (10, 8, 6)
    w = oracle_mod.random_state(n, 0, seed=2)
    w[4][33] = -5.0
    ret, _, bits = emu.rhs(n, 0, (0.1, 0.1, 0.1), 1.4, [P] * 6, [0] * 6, 0, w)
    _, _, mask = port.feuler(port.cfg(n, 0, (0.1, 0.1, 0.1), 1.4, [P] * 6), w)
    assert ret == -1 and bits == mask == 6


def test_emulated_two_rank_split_with_halo_buffers_and_subboxes(emu, pkg, oracle_mod, port):
    """The N>1 kernel path without a GPU: split a periodic/reflecting box over 2 ranks the way
    SetupDecomp does, hand each rank the neighbour's packed layers as its halo buffers
    (wire layout of euler3D.hpp:648), evaluate interior and boundary shells as separate
    sub-box launches exactly like eulerb200_rhs_async, and compare with the global oracle."""
    n, nchem = (14, 8, 6), 2
    bcs = [P, P, R, R, N, N]
    d = (0.1, 0.2, 0.3)
    w = oracle_mod.random_state(n, nchem, seed=21)
    ret_ref, ref, _ = port.feuler(port.cfg(n, nchem, d, 1.4, bcs), w)
    W3 = [w[f].reshape(n[2], n[1], n[0]) for f in range(5)] + [w[5].reshape(n[2], n[1], n[0], nchem)]
    R3 = [ref[f].reshape(n[2], n[1], n[0]) for f in range(5)] + [ref[5].reshape(n[2], n[1], n[0], nchem)]
    blocks = []
    for rank in range(2):
        rc, dims, coords, ext, nbr = pkg.dims_and_extents(2, rank, n, bcs)
        assert rc == 0 and dims == [2, 1, 1]
        sl = (slice(ext[4], ext[5] + 1), slice(ext[2], ext[3] + 1), slice(ext[0], ext[1] + 1))
        parts = [np.ascontiguousarray(a[sl]).ravel() for a in W3]
        nl = (ext[1] - ext[0] + 1, ext[3] - ext[2] + 1, ext[5] - ext[4] + 1)
        blocks.append(dict(ext=ext, nbr=nbr, parts=parts, nl=nl, sl=sl))
    for rank, b in enumerate(blocks):
        other = blocks[1 - rank]
        ocfg = port.cfg(other["nl"], nchem, d, 1.4, bcs)
        recv = [None] * 6
        for f in range(6):
            if b["nbr"][f] not in (NO, rank):
                recv[f] = port.pack_send(ocfg, other["parts"], f ^ 1)
        nl = b["nl"]
        out = [np.full(nl[0] * nl[1] * nl[2], np.nan) for _ in range(5)] + [np.full(nl[0] * nl[1] * nl[2] * nchem, np.nan)]
        lo = [3 if b["nbr"][2 * a] not in (NO, rank) else 0 for a in range(3)]
        hi = [nl[a] - (3 if b["nbr"][2 * a + 1] not in (NO, rank) else 0) for a in range(3)]
        boxes = [(lo, hi), ([0, lo[1], lo[2]], [lo[0], hi[1], hi[2]]), ([hi[0], lo[1], lo[2]], [nl[0], hi[1], hi[2]])]
        for blo, bhi in boxes:
            ret, part, bits = emu.rhs(nl, nchem, d, 1.4, bcs, b["nbr"], rank, b["parts"], recv=recv, lo=blo, hi=bhi,
                                      aux_in_gen=1 if blo[0] != lo[0] or bhi[0] != hi[0] else 0)   # shells: AG
            assert ret == 0
            for o, p_ in zip(out, part):
                m = ~np.isnan(p_)
                assert np.all(np.isnan(o[m]))       # shells do not overlap
                o[m] = p_[m]
        want = [np.ascontiguousarray(a[b["sl"]]).ravel() for a in R3]
        assert not any(np.isnan(o).any() for o in out)
        assert max(normwise_errors(out, want)) <= 1e-12


def test_emulated_slow_mode_matches_reference_sequence(emu, oracle_mod, port):
    """fslow on the CPU tier: energy rebuild, fEuler, etdot moved into the last
    species -- the kernel's redirected stores against the reference's sequence around the oracle."""
    n, nchem, eu = (34, 9, 8), 4, 2.5e3
    bcs = [R] * 6
    d = (0.1, 0.2, 0.3)
    w = oracle_mod.random_state(n, nchem, seed=12)
    chem = w[5].reshape(-1, nchem)
    chem[:, -1] = eu * (2.0 + chem[:, -1])
    w[4][:] = -7.0                                     # must be rebuilt, not read
    ref_w = [x.copy() for x in w]
    ref_w[4] = chem[:, -1] * (1.0 / eu) + 0.5 / w[0] * (w[1] ** 2 + w[2] ** 2 + w[3] ** 2)
    ret, got, bits = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, energy_units=eu)
    assert ret == 0 and bits == 0
    assert np.array_equal(w[4], ref_w[4])
    ret_ref, ref, _ = port.feuler(port.cfg(n, nchem, d, 1.4, bcs), ref_w)
    ref[5].reshape(-1, nchem)[:, -1] = ref[4]
    ref[4] = np.zeros_like(ref[4])
    assert np.all(got[4] == 0.0)
    floor = rounding_floor(ref_w, 1.4, d)
    floor[4] = 0.0
    assert max(normwise_errors(got, ref, floor)) <= 1e-12
    # the forcing-from-wdot instantiation takes the same redirected stores: with G = 0 in wdot, same bits
    zeros = [np.zeros_like(x) for x in got]
    ret, got_gw, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, energy_units=eu, g_in_wdot=zeros)
    assert ret == 0 and all(np.array_equal(a, b) for a, b in zip(got, got_gw))
    # and the boundary-heavy instantiation
    ret, got_ag, _ = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, energy_units=eu, aux_in_gen=1)
    assert ret == 0 and max(normwise_errors(got_ag, ref, floor)) <= 1e-12


def test_emulated_random_configurations(emu, oracle_mod, port):
    """Seeded sweep over grid shapes (thin, ragged, multi-tile), species counts, boundary-condition
    mixes, CTA sizes, row-synchronisation modes and kernel instantiations: 150 small cases against the
    oracle, tolerance as everywhere (Dirichlet mixes: same non-finite cells as the oracle, finite cells
    to 1e-9 of the field's scale -- the neighbourhood of a Dirichlet face is ill-conditioned)."""
    rng = np.random.default_rng(20261017)
    for case in range(150):
        n = tuple(int(x) for x in rng.integers(3, [70, 30, 20]))
        nchem = int(rng.integers(0, 6))
        bcs = []
        for _ in range(3):
            k = int(rng.choice([P, N, R, D], p=[0.35, 0.3, 0.3, 0.05]))
            bcs += [P, P] if k == P else [k, int(rng.choice([N, R, D], p=[0.5, 0.45, 0.05]))]
        threads = int(rng.choice([64, 128, 256, 384]))
        kw = dict(forcing=[0, 0.3, -0.1, 0, 0.2], threads=threads, pair=int(rng.integers(0, 3)),
                  aux_in_gen=int(rng.integers(0, 2)), use_aux=int(rng.integers(0, 2)))
        kw["split"] = case % 2            # fluid fields and species in separate launches
        kw["chemT"] = (case // 2) % 2     # species read from the pair-interleaved copy / from the vector itself
        w = oracle_mod.random_state(n, nchem, seed=int(rng.integers(1, 1000)))
        d = (1.0 / n[0], 2.0 / n[1], 0.5 / n[2])
        ret, got, bits = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, **kw)
        if ret == -77:                      # rows are not warps: CTA-wide barriers
            kw["pair"] = 0
            ret, got, bits = emu.rhs(n, nchem, d, 1.4, bcs, nbr_single(bcs), 0, w, **kw)
        ret_ref, ref, _ = port.feuler(port.cfg(n, nchem, d, 1.4, bcs, forcing=kw["forcing"]), w)
        tag = (case, n, nchem, bcs, kw)
        if D in bcs:
            for a, b in zip(got, ref):
                if a is None:
                    continue
                fin = np.isfinite(b)
                assert np.array_equal(fin, np.isfinite(a)), tag
                if fin.any():
                    assert np.abs(a[fin] - b[fin]).max() <= 1e-9 * np.abs(b[fin]).max(), tag
        else:
            assert ret == 0 and ret_ref == 0 and bits == 0, tag
            assert max(normwise_errors(got, ref, rounding_floor(w, 1.4, d))) <= 1e-12, tag


def test_emulated_random_decompositions(emu, pkg, oracle_mod, port):
    """Seeded sweep of the N>1 kernel path without GPUs: 60 random grids / BC mixes split over 2, 4
    or 8 ranks the way SetupDecomp does (splits along every axis, periodic wraps onto other ranks),
    every rank reading the layers its neighbours' pack_face_kernel produced (wire layout of
    euler3D.hpp:648,696,744) as halo buffers and evaluating interior box and the six boundary slabs as separate sub-box launches
    exactly like eulerb200_rhs_async; the assembled result against the single-rank oracle on the
    global grid (the decomposed result is the single-rank result)."""
    rng = np.random.default_rng(8)
    d = (0.1, 0.2, 0.3)
    done = 0
    while done < 60:
        world = int(rng.choice([2, 4, 8]))
        n = tuple(int(x) for x in rng.integers([6, 6, 6], [40, 30, 24]))
        nchem = int(rng.integers(0, 4))
        bcs = []
        for _ in range(3):
            k = int(rng.choice([P, N, R], p=[0.4, 0.3, 0.3]))
            bcs += [P, P] if k == P else [k, int(rng.choice([N, R]))]
        threads, ag = int(rng.choice([64, 128, 256])), int(rng.integers(0, 2))
        layout = [pkg.dims_and_extents(world, r, n, bcs) for r in range(world)]
        if any(rc != 0 for rc, *_ in layout):
            continue                              # some local extent < 3: SetupDecomp refuses
        done += 1
        w = oracle_mod.random_state(n, nchem, seed=int(rng.integers(1, 1000)))
        ret_ref, ref, _ = port.feuler(port.cfg(n, nchem, d, 1.4, bcs), w)
        shape = (n[2], n[1], n[0])
        W3 = [w[f].reshape(shape) for f in range(5)] + ([w[5].reshape(shape + (nchem,))] if nchem else [])
        R3 = [ref[f].reshape(shape) for f in range(5)] + ([ref[5].reshape(shape + (nchem,))] if nchem else [])
        blocks = []
        for rc, dims, coords, ext, nbr in layout:
            sl = (slice(ext[4], ext[5] + 1), slice(ext[2], ext[3] + 1), slice(ext[0], ext[1] + 1))
            parts = [np.ascontiguousarray(a[sl]).ravel() for a in W3] + ([] if nchem else [None])
            blocks.append(dict(nbr=nbr, parts=parts, sl=sl, nl=(ext[1] - ext[0] + 1, ext[3] - ext[2] + 1, ext[5] - ext[4] + 1)))
        for rank, b in enumerate(blocks):
            recv = [None] * 6
            for f in range(6):
                o = b["nbr"][f]
                if o not in (NO, rank):            # what the neighbour's pack_face_kernel sends through its opposite face
                    recv[f] = emu.face("pack", blocks[o]["nl"], nchem, bcs, blocks[o]["nbr"], o, blocks[o]["parts"], f ^ 1)
            nl = b["nl"]
            Nl = nl[0] * nl[1] * nl[2]
            out = [np.full(Nl, np.nan) for _ in range(5)] + ([np.full(Nl * nchem, np.nan)] if nchem else [])
            lo = [3 if b["nbr"][2 * a] not in (NO, rank) else 0 for a in range(3)]
            hi = [nl[a] - (3 if b["nbr"][2 * a + 1] not in (NO, rank) else 0) for a in range(3)]
            if any(hi[a] <= lo[a] for a in range(3)):
                boxes = [([0, 0, 0], list(nl))]      # no interior: one launch after the exchange
            else:
                boxes = [(lo, hi), ([0, 0, 0], [nl[0], nl[1], lo[2]]), ([0, 0, hi[2]], list(nl)),
                         ([0, 0, lo[2]], [nl[0], lo[1], hi[2]]), ([0, hi[1], lo[2]], [nl[0], nl[1], hi[2]]),
                         ([0, lo[1], lo[2]], [lo[0], hi[1], hi[2]]), ([hi[0], lo[1], lo[2]], [nl[0], hi[1], hi[2]])]
            tag = (n, nchem, bcs, world, rank, threads, ag)
            for blo, bhi in boxes:
                if any(bhi[a] <= blo[a] for a in range(3)):
                    continue
                ret, part, bits = emu.rhs(nl, nchem, d, 1.4, bcs, b["nbr"], rank, b["parts"], recv=recv, lo=blo, hi=bhi,
                                          threads=threads, aux_in_gen=ag, split=done % 2, chemT=(done // 2) % 2)
                assert ret == 0, tag
                for o_, p_ in zip(out, part):
                    m = ~np.isnan(p_)
                    assert np.all(np.isnan(o_[m])), tag          # boxes do not overlap
                    o_[m] = p_[m]
            want = [np.ascontiguousarray(a[b["sl"]]).ravel() for a in R3]
            assert not any(np.isnan(o_).any() for o_ in out), tag
            assert max(normwise_errors(out, want)) <= 1e-12, tag


@pytest.mark.parametrize("n,nchem", [((12, 10, 8), 2), ((3, 9, 7), 0), ((5, 3, 4), 4)])
def test_emulated_pack_and_ghost_face_kernels_exact(emu, oracle_mod, port, n, nchem):
    """halo_kernels.cuh on the CPU tier, exact (they only move values): pack_face_kernel against the
    oracle's restatement of ExchangeStart's send buffers (euler3D.hpp:644-786) for all six faces, and
    ghost_face_kernel against its boundary-condition fills (euler3D.hpp:797-1166) for every BC type
    and against a halo slab handed in as receive buffer."""
    w = oracle_mod.random_state(n, nchem, seed=31)
    d = (1.0, 1.0, 1.0)
    for bcs in ([P] * 6, [N] * 6, [R] * 6, [D] * 6, [N, N, R, R, P, P], [R, R, D, D, N, N]):
        cfg = port.cfg(n, nchem, d, 1.4, bcs)
        for f in range(6):
            assert np.array_equal(emu.face("pack", n, nchem, bcs, nbr_single(bcs), 0, w, f), port.pack_send(cfg, w, f))
            got = emu.face("ghost", n, nchem, bcs, nbr_single(bcs), 0, w, f)
            if bcs[f] == P:      # single rank: the periodic ghost layers are the opposite side's send layers
                assert np.array_equal(got, port.pack_send(cfg, w, f ^ 1))
            else:
                assert np.array_equal(got, port.fill_ghost(cfg, w, f))
    # a remote neighbour: the ghost layers are whatever the halo slab holds
    slab = np.arange((5 + nchem) * 3 * n[1] * n[2], dtype=np.float64)
    nbr = [1, NO, NO, NO, NO, NO]
    got = emu.face("ghost", n, nchem, [N] * 6, nbr, 0, w, 0, recv=[slab, None, None, None, None, None])
    assert np.array_equal(got, slab)
