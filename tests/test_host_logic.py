"""Host-side logic and the C-ABI surface (no GPU needed)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
P, N, D, R = 0, 1, 2, 3


def test_library_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "eulerb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(eulerb200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    lib = C.CDLL(pkg.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "libeulerb200.so does not export %s" % name
    assert declared == set(pkg.ABI), "Python ABI table and header disagree"
    assert pkg.load_library().eulerb200_version() == 100


def test_config_struct_layout_matches_header(pkg):
    # 3*8 + 2*4 + 4*8 + 6*4 + 6*4 + 2*4 + 5*8 = 160 bytes, no padding surprises
    assert C.sizeof(pkg.Config) == 160
    assert pkg.Config.dx.offset == 32 and pkg.Config.bc.offset == 64 and pkg.Config.forcing.offset == 120


def test_decomposition_matches_reference_setupdecomp(pkg):
    """eulerb200_decompose vs tables produced by the reference's SetupDecomp
    (tests/golden/decomp_tables.npz; euler3D.hpp:396-574) for the BASELINE.json grid shapes."""
    z = np.load(os.path.join(GOLD, "decomp_tables.npz"))
    checked = 0
    for tag in ("cube", "rt", "hurricane", "sod", "periodic"):
        n = [int(x) for x in z[tag + "_n"]]
        bc = [int(x) for x in z[tag + "_bc"]]
        for nprocs in (1, 2, 4, 8):
            tab = z["%s_p%d" % (tag, nprocs)]
            for rank in range(nprocs):
                rc, dims, coords, ext, nbr = pkg.dims_and_extents(nprocs, rank, n, bc)
                if tab[0][0] == -999:
                    assert rc != 0          # the reference refused this decomposition too
                    continue
                assert rc == 0
                assert ext == [int(x) for x in tab[rank][:6]]
                ref_nbr = [int(x) if int(x) != -2 else -1 for x in tab[rank][6:]]   # -2 = MPI_PROC_NULL in the shim
                assert nbr == ref_nbr
                checked += 1
    assert checked >= 40


def test_process_grids_of_the_baseline_configs(pkg):
    R6 = [R] * 6
    assert pkg.dims_and_extents(1, 0, (512, 512, 512), R6)[1] == [1, 1, 1]
    assert pkg.dims_and_extents(2, 0, (512, 512, 512), R6)[1] == [2, 1, 1]
    assert pkg.dims_and_extents(4, 0, (512, 512, 512), R6)[1] == [2, 2, 1]
    assert pkg.dims_and_extents(8, 0, (512, 512, 512), R6)[1] == [2, 2, 2]
    rc, dims, _, ext, _ = pkg.dims_and_extents(8, 5, (3, 4096, 4096), [N] * 6)      # hurricane_yz
    assert rc == 0 and dims == [1, 4, 2]
    assert (ext[1] - ext[0] + 1, ext[3] - ext[2] + 1, ext[5] - ext[4] + 1) == (3, 1024, 2048)
    assert pkg.dims_and_extents(2, 0, (12, 12, 12), [P, N, P, P, P, P])[0] == 1       # half-periodic axis
    assert pkg.dims_and_extents(8, 0, (4, 4, 4), [P] * 6)[0] == -1                   # local extent < 3


def test_exchange_plan_pairs_up(pkg):
    """Every send of rank a to rank b is met, in order, by a receive of b from a on the
    opposite face with the same length (what the MPI tags guarantee in the reference)."""
    for n, bc, nprocs in (((12, 12, 12), [P] * 6, 2), ((12, 12, 12), [P] * 6, 8), ((16, 12, 20), [R] * 6, 4),
                          ((3, 32, 32), [N] * 6, 8), ((12, 16, 20), [P, P, R, R, N, N], 8)):
        plans, lens = [], []
        for rank in range(nprocs):
            u = pkg.EulerData(nchem=2)
            u.nx, u.ny, u.nz = n
            u.xlbc, u.xrbc, u.ylbc, u.yrbc, u.zlbc, u.zrbc = bc
            rc, dims, coords, ext, nbr = pkg.dims_and_extents(nprocs, rank, n, bc)
            assert rc == 0
            u.myid, u.nprocs = rank, nprocs
            u.nxl, u.nyl, u.nzl = ext[1] - ext[0] + 1, ext[3] - ext[2] + 1, ext[5] - ext[4] + 1
            u.dx = u.dy = u.dz = 1.0
            u.ipW, u.ipE, u.ipS, u.ipN, u.ipB, u.ipF = nbr
            plans.append(u.exchange_plan())
            area = [u.nyl * u.nzl, u.nyl * u.nzl, u.nxl * u.nzl, u.nxl * u.nzl, u.nxl * u.nyl, u.nxl * u.nyl]
            lens.append([7 * 3 * a for a in area])
        for a in range(nprocs):
            for b in range(nprocs):
                sends = [(f, lens[a][f]) for k, f, p in plans[a] if k == "send" and p == b]
                recvs = [(f, lens[b][f]) for k, f, p in plans[b] if k == "recv" and p == a]
                assert len(sends) == len(recvs)
                for (fs, ls), (fr, lr) in zip(sends, recvs):
                    assert fr == fs ^ 1 and ls == lr


def test_create_fails_loudly_without_gpu_or_with_bad_config(pkg):
    import torch
    u = pkg.EulerData()
    u.nx, u.ny, u.nz = 8, 8, 8
    if not torch.cuda.is_available():
        with pytest.raises(pkg.EulerB200Error, match="no usable CUDA device"):
            u.SetupDecomp()
    lib = pkg.load_library()
    cfg = pkg.Config()
    cfg.nxl, cfg.nyl, cfg.nzl = 2, 8, 8          # euler3D.hpp:483-494
    cfg.dx = cfg.dy = cfg.dz = 1.0
    ctx = C.c_void_p()
    assert lib.eulerb200_create(C.byref(cfg), C.byref(ctx)) == -1
    assert b"extents" in lib.eulerb200_last_error(None)
    cfg.nxl = 8
    for f in range(6):
        cfg.bc[f], cfg.nbr[f] = P, -1            # periodic face without a neighbour
    assert lib.eulerb200_create(C.byref(cfg), C.byref(ctx)) == -1


def test_product_does_not_import_the_oracle():
    """The product path may not route through the oracle or any CPU fallback."""
    pdir = os.path.join(ROOT, "sundials-manyvector-demo_b200")
    for dirpath, _, files in os.walk(pdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_boundary_tile_fraction_of_the_baseline_shapes(pkg):
    """launch_box() offers the AG instantiation (EULERB200_KERNEL=1) to launches in which a quarter or
    more of the tiles touch a boundary: not the 512^3 bench shape (its SASS stays the measured one),
    but 256^3, the thin hurricane plane and the 3-cell boundary shells of a decomposed run."""
    import ctypes as C
    from emu.emu import build
    lib = C.CDLL(build())
    lib.emu_boundary_tile_fraction.restype = C.c_double
    L3 = C.c_long * 3

    def frac(lo, hi, n, nchem):
        return lib.emu_boundary_tile_fraction(L3(*lo), L3(*hi), C.c_long(n[0]), C.c_long(n[1]), nchem, 384)

    assert 0.10 < frac((0, 0, 0), (512, 512, 512), (512, 512, 512), 10) < 0.20
    assert 0.25 <= frac((0, 0, 0), (256, 256, 256), (256, 256, 256), 0) < 0.40
    assert frac((0, 0, 0), (3, 4096, 4096), (3, 4096, 4096), 0) == 1.0
    assert frac((0, 3, 3), (3, 509, 509), (512, 512, 512), 10) == 1.0          # x-low shell of an 8-GPU run
    assert frac((3, 3, 3), (509, 509, 509), (512, 512, 512), 10) < 0.20         # its interior box


def test_native_driver_help_and_option_errors_need_no_gpu(pkg):
    """euler3d_b200 --help, and the Butcher-table option checks (order overrides etable,
    euler3D_main.cpp:207-213), happen before the device is touched."""
    import subprocess
    exe = os.path.join(ROOT, "sundials-manyvector-demo_b200", "euler3d_b200")
    out = subprocess.run([exe, "--help"], capture_output=True, text=True)
    assert out.returncode == 0 and "--etable=0|1|3|6|7|8|10|11|12" in out.stdout and "primordial_blast" in out.stdout
    for args, msg in ((["--order=7"], "no explicit Butcher table"), (["--order=0", "--etable=13"], "no explicit Butcher table"),
                      (["--order=0", "--etable=12"], "needs fixedstep = 1")):
        out = subprocess.run([exe] + args, capture_output=True, text=True)
        assert out.returncode == 1 and msg in out.stderr, (args, out.stderr)


def test_interior_and_shell_boxes_cover_the_box_once_and_follow_the_tile_grid():
    """eb::overlap_boxes (host_setup.h), the launches rhs_impl() makes for a rank with remote neighbours: the
    interior box keeps three cells from every remote face, interior + shells cover the box exactly once, and with
    tile-thick shells (default) every cut in x and y lies on a seam of the tile grid a single launch over the box
    would use (32 owned columns / 11 owned rows at 384 threads), so no launch runs 3-of-4-column tiles."""
    from emu.emu import build
    lib = C.CDLL(build())
    L3, I6 = C.c_long * 3, C.c_int * 6
    out = (C.c_long * 64)()

    def boxes(n, remote, nchem=10, xc=1, thick=1):
        rc = lib.emu_overlap_boxes(L3(*n), I6(*remote), nchem, 384, xc, thick, out)
        cnt = out[0]
        return rc, [(tuple(out[1 + 6 * q + d] for d in range(3)), tuple(out[4 + 6 * q + d] for d in range(3))) for q in range(cnt)]

    for n, remote, thick in [((512, 512, 512), (0, 1, 0, 1, 0, 1), 1), ((512, 512, 512), (1, 0, 1, 0, 1, 0), 1),
                             ((512, 512, 512), (1, 1, 1, 1, 1, 1), 1), ((200, 90, 40), (1, 1, 0, 1, 1, 0), 1),
                             ((512, 512, 512), (1, 1, 1, 1, 1, 1), 0), ((40, 24, 20), (1, 1, 0, 0, 0, 0), 1),
                             ((3, 4096, 2048), (0, 0, 1, 1, 1, 1), 1)]:
        rc, bl = boxes(n, remote, thick=thick)
        assert rc == 0 and len(bl) == 1 + sum(remote)
        cover = np.zeros((n[2] // 1, n[1], n[0]), dtype=np.int8) if n[0] * n[1] * n[2] <= 1 << 24 else None
        vol = 0
        for lo, hi in bl:
            assert all(0 <= lo[d] < hi[d] <= n[d] for d in range(3))
            vol += (hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2])
            if cover is not None:
                cover[lo[2]:hi[2], lo[1]:hi[1], lo[0]:hi[0]] += 1
        assert vol == n[0] * n[1] * n[2]
        if cover is not None:
            assert (cover == 1).all()
        else:          # disjoint by construction of the six slabs around box 0: check pairwise
            for a in range(len(bl)):
                for b in range(a + 1, len(bl)):
                    assert any(bl[a][1][d] <= bl[b][0][d] or bl[b][1][d] <= bl[a][0][d] for d in range(3))
        lo, hi = bl[0]
        for d in range(3):
            assert lo[d] >= (3 if remote[2 * d] else 0) and hi[d] <= n[d] - (3 if remote[2 * d + 1] else 0)
    # 512^3, all faces remote, 384 threads with full-width tiles: cuts at multiples of 32 (x) and 11 (y), 3 planes in z
    rc, bl = boxes((512, 512, 512), (1, 1, 1, 1, 1, 1))
    assert bl[0] == ((32, 11, 3), (480, 506, 509))
    rc, bl = boxes((512, 512, 512), (0, 1, 0, 1, 0, 1))
    assert bl[0] == ((0, 0, 0), (480, 506, 509))
    rc, bl = boxes((512, 512, 512), (1, 1, 1, 1, 1, 1), thick=0)
    assert bl[0] == ((3, 3, 3), (509, 509, 509))
    rc, bl = boxes((512, 512, 512), (1, 0, 0, 0, 0, 0), xc=0)        # 31-column tiles
    assert bl[0] == ((31, 0, 0), (512, 512, 512))
