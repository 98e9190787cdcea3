#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, compiled from
/root/reference/src by oracle/Makefile).  Run in the container that has the reference:

    python tests/golden/make_golden.py

The fixtures pin the oracle port (tests/test_oracle.py) and the CUDA path
(tests/test_gpu_parity.py) to outputs of the reference itself.  Inputs are stored with
the outputs, so nothing depends on a random-number generator staying stable.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

P, N, D, R = 0, 1, 2, 3

# (name, nvar, n, bcs, box, gamma, forcing, state kind)
FEULER_CASES = [
    ("periodic_nvar5", 5, (10, 8, 6), [P] * 6, (0, 1, 0, 1, 0, 1), 1.4, None, "random"),
    ("neumann_nvar7", 7, (10, 8, 6), [N] * 6, (0, 1, 0, 2, 0, 0.5), 1.4, None, "random"),
    ("reflecting_nvar15", 15, (9, 7, 6), [R] * 6, (0, 1, 0, 1, 0, 1), 5.0 / 3.0, None, "random"),
    ("rt_mixed_nvar5", 5, (8, 12, 3), [P, P, R, R, N, N], (-0.25, 0.25, -0.75, 0.75, 0, 1), 1.4,
     [0, 0, -0.1, 0, 0], "rayleigh_taylor"),
    ("sod_x_nvar5", 5, (40, 3, 3), [N] * 6, (0, 1, 0, 1, 0, 1), 1.4, None, "sod"),
    ("hurricane_yz_nvar5", 5, (3, 16, 16), [N] * 6, (-1, 1, -1, 1, -1, 1), 2.0, None, "hurricane"),
    ("advection_y_nvar9", 9, (6, 16, 5), [P] * 6, (0, 1, 0, 1, 0, 1), 1.4, None, "advection_y"),
]


def cell_centres(n, box):
    x = box[0] + (np.arange(n[0]) + 0.5) * (box[1] - box[0]) / n[0]
    y = box[2] + (np.arange(n[1]) + 0.5) * (box[3] - box[2]) / n[1]
    z = box[4] + (np.arange(n[2]) + 0.5) * (box[5] - box[4]) / n[2]
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")      # index i + nx*(j + ny*k)
    return X.ravel(), Y.ravel(), Z.ravel()


def make_state(kind, n, nchem, box, gamma, seed):
    """Initial states shaped like the reference's problem files (closed forms restated from
    sod.cpp:50-55,139-149; rayleigh_taylor.cpp:48-53,104-109; hurricane.cpp:58-60,157-169;
    linear_advection.cpp:51-68,117-131)."""
    N = n[0] * n[1] * n[2]
    rng = np.random.default_rng(seed)
    if kind == "random":
        return oracle.random_state(n, nchem, seed=seed, gamma=gamma)
    X, Y, Z = cell_centres(n, box)
    mx = np.zeros(N); my = np.zeros(N); mz = np.zeros(N)
    if kind == "sod":
        rho = np.where(X < 0.5, 1.0, 0.125); p = np.where(X < 0.5, 1.0, 0.1)
    elif kind == "rayleigh_taylor":
        rho = np.where(Y > 0, 2.0, 1.0)
        my = rho * 0.01 * (1 + np.cos(4 * np.pi * X)) * (1 + np.cos(3 * np.pi * Y))
        p = 2.5 - 0.1 * rho * Y
    elif kind == "hurricane":
        rho = np.ones(N); th = np.arctan2(Z, Y)
        my = 10.0 * np.sin(th); mz = -10.0 * np.cos(th); p = np.full(N, 25.0)
    elif kind == "advection_y":
        rho = 1.0 + 0.1 * np.sin(2 * np.pi * Y); my = 0.5 * rho; p = np.ones(N)
    else:
        raise ValueError(kind)
    et = p / (gamma - 1.0) + 0.5 * (mx * mx + my * my + mz * mz) / rho
    parts = [rho, mx, my, mz, et]
    parts.append(rng.random(N * nchem) * 10.0 ** rng.integers(-6, 6, size=N * nchem) if nchem else None)
    return parts


def main():
    if not oracle.build_ref():
        raise SystemExit("the reference tree is not available: cannot regenerate golden vectors")
    out = {}
    for name, nvar, n, bcs, box, gamma, forcing, kind in FEULER_CASES:
        R_ = oracle.Ref(nvar)
        parts = make_state(kind, n, nvar - 5, box, gamma, seed=len(name) * 7 + nvar)
        ret, wdot, _, _ = R_.feuler(n, box, bcs, gamma, parts, forcing=forcing)
        assert ret == 0, name
        rec = {"n": np.array(n), "nvar": nvar, "bcs": np.array(bcs), "box": np.array(box, dtype=float),
               "gamma": gamma, "forcing": np.array(forcing if forcing else [0.0] * 5)}
        for f in range(5):
            rec["w%d" % f] = parts[f]; rec["wdot%d" % f] = wdot[f]
        if nvar > 5:
            rec["w5"] = parts[5]; rec["wdot5"] = wdot[5]
        np.savez_compressed(os.path.join(HERE, "feuler_%s.npz" % name), **rec)
        out[name] = float(max(np.abs(w).max() for w in wdot if w is not None))

    # face_flux known answers (utilities.cpp:270-479), NVAR 5 and 15, all three directions
    rng = np.random.default_rng(2024)
    for nvar in (5, 15):
        R_ = oracle.Ref(nvar)
        sten, outs, dirs = [], [], []
        for t in range(60):
            w = np.zeros((6, nvar))
            w[:, 0] = 1 + 0.5 * rng.random(6)
            w[:, 1:4] = 0.8 * (rng.random((6, 3)) - 0.5)
            p = 0.5 + rng.random(6)
            if t % 3 == 0:        # a jump in the middle of the stencil
                w[3:, 0] *= 0.2; p[3:] *= 0.1
            w[:, 4] = p / 0.4 + 0.5 * (w[:, 1:4] ** 2).sum(1) / w[:, 0]
            w[:, 5:] = rng.random((6, nvar - 5)) * 10.0 ** rng.integers(-20, 9, size=nvar - 5)
            sten.append(w); dirs.append(t % 3); outs.append(R_.face_flux(w, t % 3, 1.4))
        np.savez_compressed(os.path.join(HERE, "face_flux_nvar%d.npz" % nvar), stencil=np.array(sten),
                            idir=np.array(dirs), flux=np.array(outs), gamma=1.4)

    # halo exchange of the reference itself: 12^3 periodic over 2 and 8 virtual ranks, values that
    # encode (field, i, j, k) -- the digit trick of communication_test_main.cpp:133-181
    for nvar in (5, 7):
        R_ = oracle.Ref(nvar)
        n = (12, 12, 12)
        N = 12 ** 3
        idx = np.arange(N)
        i, j, k = idx % 12, (idx // 12) % 12, idx // 144
        enc = lambda v: 0.001 * v + 1e-6 * i + 1e-9 * j + 1e-12 * k
        parts = [enc(v) for v in range(5)]
        parts.append(np.stack([enc(5 + v) for v in range(nvar - 5)], axis=1).ravel() if nvar > 5 else None)
        rec = {"n": np.array(n), "nvar": nvar}
        for nprocs in (2, 8):
            for rank in range(nprocs):
                ext, nbr, recv = R_.exchange(n, [P] * 6, parts, nprocs=nprocs, rank=rank)
                rec["p%d_r%d_ext" % (nprocs, rank)] = np.array(ext)
                rec["p%d_r%d_nbr" % (nprocs, rank)] = np.array(nbr)
                for f in range(6):
                    rec["p%d_r%d_recv%d" % (nprocs, rank, f)] = recv[f]
        np.savez_compressed(os.path.join(HERE, "exchange_nvar%d.npz" % nvar), **rec)

    # SetupDecomp tables (euler3D.hpp:396-574) for the BASELINE.json grid shapes
    rec = {}
    R5 = oracle.Ref(5)
    for tag, n, bcs in (("cube", (24, 24, 24), [R] * 6), ("rt", (16, 48, 3), [P, P, R, R, N, N]),
                        ("hurricane", (3, 32, 32), [N] * 6), ("sod", (48, 3, 3), [N] * 6),
                        ("periodic", (12, 16, 20), [P] * 6)):
        N = n[0] * n[1] * n[2]
        parts = [np.ones(N)] * 5 + [None]
        for nprocs in (1, 2, 4, 8):
            try:
                rows = []
                for rank in range(nprocs):
                    ext, nbr, _ = R5.exchange(n, bcs, parts, nprocs=nprocs, rank=rank)
                    rows.append(ext + nbr)
                rec["%s_p%d" % (tag, nprocs)] = np.array(rows)
            except AssertionError:
                rec["%s_p%d" % (tag, nprocs)] = np.array([[-999]])     # SetupDecomp refused
        rec["%s_n" % tag] = np.array(n); rec["%s_bc" % tag] = np.array(bcs)
    np.savez_compressed(os.path.join(HERE, "decomp_tables.npz"), **rec)
    # fluid_blast initial condition of the reference itself (fluid_blast.cpp:65-267): pins the
    # mt19937_64 clump generator and the formulas restated in problems.py
    import subprocess
    import tempfile
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "fluid_blast"])
    n = (14, 12, 10)
    N = n[0] * n[1] * n[2]
    with tempfile.TemporaryDirectory() as td:
        out_path = os.path.join(td, "ic.bin")
        subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "fluid_blast_ic")] + [str(x) for x in n] + [out_path],
                              stdout=subprocess.DEVNULL)
        flat = np.fromfile(out_path, dtype=np.float64)
    arrs = [flat[f * N:(f + 1) * N].copy() for f in range(5)]
    np.savez_compressed(os.path.join(HERE, "ic_fluid_blast.npz"), n=np.array(n), **{"w%d" % f: arrs[f] for f in range(5)})

    print("golden vectors written:", sorted(os.listdir(HERE)))
    print("max |wdot| per case:", out)


if __name__ == "__main__":
    main()
