"""TEST INFRASTRUCTURE.  Builds tests/emu/libemu.so (the product's kernel SOURCE compiled
for the CPU through cuda_emu.h) and calls it.  Never imported by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "sundials-manyvector-demo_b200", "csrc")
SO = os.path.join(HERE, "libemu.so")
_dp = C.POINTER(C.c_double)


def build(strict=False, flags=None, tag=None):
    """strict: the kernel source with -DEB_STRICT (csrc/strict_face.cuh: the reference's operation order).
    flags / tag: extra compiler flags and the library-name suffix of a variant build (tools/blast_tolerance.py)."""
    so = SO.replace("libemu.so", "libemu_strict.so") if strict else SO
    if tag:
        so = SO.replace("libemu.so", "libemu_%s.so" % tag)
    deps = [os.path.join(HERE, f) for f in ("emu_rhs.cpp", "cuda_emu.h")] + \
           [os.path.join(CSRC, f) for f in ("rhs_kernel.cuh", "strict_face.cuh", "halo_kernels.cuh", "vector_kernels.cuh",
                                            "euler_math.cuh", "host_setup.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-std=c++14", "-O1", "-fPIC", "-shared"] +
                              (flags if flags is not None else ["-ffp-contract=off"]) +
                              (["-DEB_STRICT"] if strict else []) + ["-o", so, os.path.join(HERE, "emu_rhs.cpp")])
    return so


def _ptrs(arrs):
    out = (_dp * 6)()
    for i in range(6):
        a = arrs[i] if arrs is not None and i < len(arrs) else None
        out[i] = a.ctypes.data_as(_dp) if a is not None and a.size else None
    return out


class Emu:
    def __init__(self, pkg, strict=False, flags=None, tag=None):
        self.pkg = pkg
        self.lib = C.CDLL(build(strict, flags, tag))
        self.lib.emu_rhs.restype = C.c_int

    def _config(self, n, nchem, d, gamma, bcs, nbr, rank):
        c = self.pkg.Config()
        c.nxl, c.nyl, c.nzl = n
        c.nchem, c.device = nchem, -1
        c.dx, c.dy, c.dz, c.gamma = d[0], d[1], d[2], gamma
        for f in range(6):
            c.bc[f], c.nbr[f] = bcs[f], nbr[f]
        c.rank, c.nranks = rank, 1
        return c

    def face_flux(self, w1d, idir, gamma):
        """One face through the product's face arithmetic on the reference's face_flux interface
        (utilities.cpp:270): w1d[6][nvar] -> f_face[nvar]."""
        w1d = np.ascontiguousarray(w1d, dtype=np.float64)
        nvar = w1d.shape[1]
        out = np.empty(nvar)
        self.lib.emu_face_flux(w1d.ctypes.data_as(_dp), nvar, int(idir), C.c_double(gamma), out.ctypes.data_as(_dp))
        return out

    def face(self, what, n, nchem, bcs, nbr, rank, w, f, recv=None):
        """what = 'pack': the send buffer of face f (pack_face_kernel); 'ghost': the ghost layers of
        face f in the reference's receive-buffer layout (ghost_face_kernel)."""
        c = self._config(n, nchem, (1.0, 1.0, 1.0), 1.4, bcs, nbr, rank)
        other = (n[1] * n[2], n[0] * n[2], n[0] * n[1])[f // 2]
        out = np.full((5 + nchem) * 3 * other, np.nan)
        self.lib.emu_face.restype = C.c_int
        ret = self.lib.emu_face(C.byref(c), _ptrs(w), _ptrs(recv) if recv is not None else None, f,
                                {"pack": 0, "ghost": 1, "pack_warp": 2}[what], out.ctypes.data_as(_dp))
        assert ret == 0
        return out

    def rhs(self, n, nchem, d, gamma, bcs, nbr, rank, w, forcing=None, recv=None, lo=None, hi=None, threads=256,
            use_aux=1, energy_units=0.0, pair=0, g_in_wdot=None, aux_in_gen=0, split=0, chemT=0, stage=0, xc=0):
        c = self.pkg.Config()
        c.nxl, c.nyl, c.nzl = n
        c.nchem, c.device = nchem, -1
        c.dx, c.dy, c.dz, c.gamma = d[0], d[1], d[2], gamma
        for f in range(6):
            c.bc[f], c.nbr[f] = bcs[f], nbr[f]
        c.rank, c.nranks = rank, 1
        for f in range(5):
            c.forcing[f] = forcing[f] if forcing is not None else 0.0
        N = n[0] * n[1] * n[2]
        out = [np.full(N, np.nan) for _ in range(5)] + [np.full(N * nchem, np.nan) if nchem else None]
        if g_in_wdot is not None:         # the external_forces hook has assigned G into wdot
            out = [None if g is None else np.array(g, dtype=np.float64, copy=True) for g in g_in_wdot]
        bits = C.c_int(0)
        L3 = C.c_long * 3
        ret = self.lib.emu_rhs(C.byref(c), _ptrs(w), _ptrs(out), _ptrs(recv) if recv is not None else None,
                               C.byref(bits), L3(*lo) if lo else None, L3(*hi) if hi else None, threads, use_aux, C.c_double(energy_units), pair, 0 if g_in_wdot is None else 1, aux_in_gen, split, chemT, stage, xc)
        return ret, out, bits.value
