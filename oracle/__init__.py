"""oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU checkers for the CUDA fluid right-hand side:

* ``port``  : ``oracle/euler_oracle.c`` -- our plain-C restatement of the reference
  algorithm (``/root/reference/src/utilities.cpp:17-528``, ``euler3D.hpp:577-1414``),
  compiled on demand with gcc into ``oracle/liboracle.so``.
* ``ref``   : ``oracle/_ref/libref_nvar<N>.so`` -- the UNMODIFIED reference sources
  compiled against ``oracle/shim`` (see ``oracle/Makefile``).  Built in the container
  that has ``/root/reference``; the built files travel to the GPU box.

Parity status: pinned -- ``tests/test_oracle.py`` holds the port bit-for-bit against
``ref`` and against golden vectors under ``tests/golden`` generated from ``ref``.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py``
may import this package.  The product (``sundials-manyvector-demo_b200``) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("EULERB200_REFERENCE", "/root/reference")

BC_PERIODIC, BC_NEUMANN, BC_DIRICHLET, BC_REFLECTING, BC_EXTERNAL = 0, 1, 2, 3, -1
NVARS = (5, 7, 9, 11, 15)          # the reference builds these (src/CMakeLists.txt:29)
FACES = ("W", "E", "S", "N", "B", "F")

_dp = C.POINTER(C.c_double)


def build_port(force=False):
    """Compile the C restatement (gcc only; works on the GPU box too)."""
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "euler_oracle.c")
    so_fma = os.path.join(HERE, "liboracle_fma.so")
    if (force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src)
            or not os.path.exists(so_fma) or os.path.getmtime(so_fma) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    return so


def build_ref(nvars=NVARS):
    """Compile the unmodified reference into oracle/_ref (needs /root/reference)."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        return False
    subprocess.check_call(["make", "-s", "-C", HERE, "ref", "REFERENCE=" + REFERENCE_ROOT,
                           "NVARS=" + " ".join(str(n) for n in nvars)])
    return True


def build_dropin():
    """Link check of the drop-in fEuler against the reference's own EulerData (needs the
    reference tree and the built libeulerb200.so); binaries land in oracle/_ref."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        return False
    subprocess.check_call(["make", "-s", "-C", HERE, "dropin", "REFERENCE=" + REFERENCE_ROOT])
    return True


def build_dropin_emu():
    """The same link check against the CPU emulation of the kernel source (tests/emu/emu_abi.cpp)
    instead of libeulerb200.so, so that the CPU test tier can run the drop-in's host logic next to
    the unmodified reference fEuler.  Returns the binary's path, or None without the reference."""
    exe = os.path.join(REF_DIR, "dropin_check_emu_nvar7")
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        subprocess.check_call(["make", "-s", "-C", HERE, "dropin_emu", "REFERENCE=" + REFERENCE_ROOT])
    return exe if os.path.exists(exe) else None


def build_refmain():
    """The reference's own main program (euler3D_main.cpp, io.cpp, gopt.cpp, one problem file,
    all unmodified) with shim/shim_arkstep.cpp in place of ARKODE: refmain_<problem> with the
    reference fEuler, refmain_dropin_<problem> with OUR drop-in fEuler on the kernel emulation.
    Returns {name: path} of what exists, building first where the reference tree is present."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        subprocess.check_call(["make", "-s", "-j4", "-C", HERE, "refmain", "REFERENCE=" + REFERENCE_ROOT])
    out = {}
    if os.path.isdir(REF_DIR):
        for f in os.listdir(REF_DIR):
            if f.startswith("refmain_"):
                out[f[len("refmain_"):]] = os.path.join(REF_DIR, f)
    return out


def have_ref(nvar=5):
    return os.path.exists(os.path.join(REF_DIR, "libref_nvar%d.so" % nvar))


class _Cfg(C.Structure):
    _fields_ = [("nxl", C.c_long), ("nyl", C.c_long), ("nzl", C.c_long), ("nchem", C.c_int),
                ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double), ("gamma", C.c_double),
                ("bc", C.c_int * 6), ("forcing", C.c_double * 5)]


def _ptrs(arrs):
    """6-entry double* array from a list of 5 (+1) contiguous float64 arrays."""
    out = (_dp * 6)()
    for i in range(6):
        if i < len(arrs) and arrs[i] is not None and arrs[i].size > 0:
            a = arrs[i]
            assert a.dtype == np.float64 and a.flags.c_contiguous
            out[i] = a.ctypes.data_as(_dp)
        else:
            out[i] = None
    return out


def split_state(w, n, nchem):
    """View a flat state (5 fluid fields then chem, as the MPIManyVector orders them)."""
    N = int(n[0]) * int(n[1]) * int(n[2])
    parts = [w[f * N:(f + 1) * N] for f in range(5)]
    parts.append(w[5 * N:5 * N + N * nchem] if nchem > 0 else None)
    return parts


class Port:
    """ctypes face of oracle/euler_oracle.c.  ``fma=True`` loads the FMA-contracted build, which
    is not an oracle but a yardstick: |Port(fma=True) - Port()| is how far the reference's own
    arithmetic moves under a different legal compilation of the same source."""

    def __init__(self, fma=False):
        so = build_port()
        if fma:
            so = os.path.join(HERE, "liboracle_fma.so")
        self.lib = C.CDLL(so)
        L = self.lib
        L.oracle_feuler.restype = C.c_int
        L.oracle_feuler.argtypes = [C.POINTER(_Cfg), _dp * 6, _dp * 6, _dp * 6, C.POINTER(C.c_int)]
        L.oracle_face_flux.restype = None
        L.oracle_face_flux.argtypes = [_dp, C.c_int, C.c_int, C.c_double, _dp]
        L.oracle_face_len.restype = C.c_long
        L.oracle_face_len.argtypes = [C.POINTER(_Cfg), C.c_int]
        L.oracle_pack_send.restype = None
        L.oracle_pack_send.argtypes = [C.POINTER(_Cfg), _dp * 6, C.c_int, _dp]
        L.oracle_fill_ghost.restype = None
        L.oracle_fill_ghost.argtypes = [C.POINTER(_Cfg), _dp * 6, C.c_int, _dp]
        L.oracle_max_wavespeed.restype = C.c_double
        L.oracle_max_wavespeed.argtypes = [C.POINTER(_Cfg), _dp * 6]
        L.oracle_dt_stab.restype = C.c_double
        L.oracle_dt_stab.argtypes = [C.POINTER(_Cfg), C.c_double, C.c_double]

    @staticmethod
    def cfg(n, nchem, d, gamma, bc, forcing=None):
        c = _Cfg()
        c.nxl, c.nyl, c.nzl = int(n[0]), int(n[1]), int(n[2])
        c.nchem = int(nchem)
        c.dx, c.dy, c.dz = float(d[0]), float(d[1]), float(d[2])
        c.gamma = float(gamma)
        for i in range(6):
            c.bc[i] = int(bc[i])
        for i in range(5):
            c.forcing[i] = float(forcing[i]) if forcing is not None else 0.0
        return c

    def feuler(self, cfg, w_parts, ext=None):
        """Returns (retval, wdot_parts, state_mask)."""
        N = cfg.nxl * cfg.nyl * cfg.nzl
        wdot = [np.empty(N) for _ in range(5)]
        wdot.append(np.empty(N * cfg.nchem) if cfg.nchem > 0 else None)
        mask = C.c_int(0)
        ret = self.lib.oracle_feuler(C.byref(cfg), _ptrs(w_parts), _ptrs(wdot),
                                     _ptrs(ext if ext is not None else []), C.byref(mask))
        return ret, wdot, mask.value

    def face_flux(self, w1d, idir, gamma):
        s = np.array(w1d, dtype=np.float64, order="C").copy()
        nvar = s.shape[1]
        out = np.empty(nvar)
        self.lib.oracle_face_flux(s.ctypes.data_as(_dp), nvar, idir, gamma, out.ctypes.data_as(_dp))
        return out

    def face_len(self, cfg, f):
        return int(self.lib.oracle_face_len(C.byref(cfg), f))

    def pack_send(self, cfg, w_parts, f):
        buf = np.empty(self.face_len(cfg, f))
        self.lib.oracle_pack_send(C.byref(cfg), _ptrs(w_parts), f, buf.ctypes.data_as(_dp))
        return buf

    def fill_ghost(self, cfg, w_parts, f):
        buf = np.empty(self.face_len(cfg, f))
        self.lib.oracle_fill_ghost(C.byref(cfg), _ptrs(w_parts), f, buf.ctypes.data_as(_dp))
        return buf

    def max_wavespeed(self, cfg, w_parts):
        return float(self.lib.oracle_max_wavespeed(C.byref(cfg), _ptrs(w_parts)))

    def dt_stab(self, cfg, cfl, alpha):
        return float(self.lib.oracle_dt_stab(C.byref(cfg), cfl, alpha))


class Ref:
    """ctypes face of oracle/_ref/libref_nvar<N>.so (the unmodified reference)."""

    def __init__(self, nvar):
        path = os.path.join(REF_DIR, "libref_nvar%d.so" % nvar)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.nvar = nvar
        self.nchem = nvar - 5
        self.lib = C.CDLL(path)
        L = self.lib
        assert L.refdrv_nvar() == nvar
        L.refdrv_set_forcing.argtypes = [_dp]
        L.refdrv_face_flux.argtypes = [_dp, C.c_int, C.c_double, _dp]
        L.refdrv_feuler.restype = C.c_int
        L.refdrv_feuler.argtypes = [C.c_int, C.POINTER(C.c_long), _dp, C.POINTER(C.c_int), C.c_double,
                                    C.c_double, _dp * 6, _dp * 6, C.c_int, _dp, C.POINTER(C.c_int)]
        L.refdrv_stability.restype = C.c_int
        L.refdrv_stability.argtypes = [C.c_int, C.POINTER(C.c_long), _dp, C.POINTER(C.c_int), C.c_double,
                                       C.c_double, _dp * 6, _dp]
        L.refdrv_exchange.restype = C.c_int
        L.refdrv_exchange.argtypes = [C.c_int, C.POINTER(C.c_long), C.POINTER(C.c_int), _dp * 6,
                                      C.c_int, C.POINTER(C.c_long), C.POINTER(C.c_int), _dp * 6]

    def face_flux(self, w1d, idir, gamma):
        s = np.array(w1d, dtype=np.float64, order="C").copy()
        assert s.shape == (6, self.nvar)
        out = np.empty(self.nvar)
        self.lib.refdrv_face_flux(s.ctypes.data_as(_dp), idir, gamma, out.ctypes.data_as(_dp))
        return out

    def feuler(self, n, box, bc, gamma, w_parts, forcing=None, nprocs=1, nrep=1, t=0.0):
        """Global-grid fEuler over `nprocs` virtual ranks.
        Returns (retval, wdot_parts, seconds[nrep], (npx,npy,npz))."""
        N = int(n[0]) * int(n[1]) * int(n[2])
        g = np.zeros(5)
        if forcing is not None:
            g[:] = forcing
        self.lib.refdrv_set_forcing(g.ctypes.data_as(_dp))
        wdot = [np.zeros(N) for _ in range(5)]
        wdot.append(np.zeros(N * self.nchem) if self.nchem > 0 else None)
        secs = np.zeros(nrep)
        dec = (C.c_int * 3)()
        ret = self.lib.refdrv_feuler(nprocs, (C.c_long * 3)(*[int(x) for x in n]),
                                     np.asarray(box, dtype=np.float64).ctypes.data_as(_dp),
                                     (C.c_int * 6)(*[int(b) for b in bc]), gamma, t,
                                     _ptrs(w_parts), _ptrs(wdot), nrep, secs.ctypes.data_as(_dp), dec)
        return ret, wdot, secs, tuple(dec)

    def stability(self, n, box, bc, gamma, cfl, w_parts, nprocs=1):
        dt = C.c_double(0)
        ret = self.lib.refdrv_stability(nprocs, (C.c_long * 3)(*[int(x) for x in n]),
                                        np.asarray(box, dtype=np.float64).ctypes.data_as(_dp),
                                        (C.c_int * 6)(*[int(b) for b in bc]), gamma, cfl,
                                        _ptrs(w_parts), C.byref(dt))
        return ret, dt.value

    def exchange(self, n, bc, w_parts, nprocs=1, rank=0):
        """Returns (ext[6], nbr[6], recv[6 arrays]) of virtual rank `rank` after
        ExchangeStart/ExchangeEnd.  Buffers are sized for the largest possible block."""
        nmax = [int(x) for x in n]
        sizes = [self.nvar * 3 * nmax[1] * nmax[2], self.nvar * 3 * nmax[0] * nmax[2],
                 self.nvar * 3 * nmax[0] * nmax[1]]
        recv = [np.full(sizes[f // 2], np.nan) for f in range(6)]
        ext = (C.c_long * 6)()
        nbr = (C.c_int * 6)()
        ret = self.lib.refdrv_exchange(nprocs, (C.c_long * 3)(*nmax), (C.c_int * 6)(*[int(b) for b in bc]),
                                       _ptrs(w_parts), rank, ext, nbr, _ptrs(recv))
        assert ret == 0
        ext = list(ext)
        nl = [ext[1] - ext[0] + 1, ext[3] - ext[2] + 1, ext[5] - ext[4] + 1]
        lens = [self.nvar * 3 * nl[1] * nl[2], self.nvar * 3 * nl[0] * nl[2], self.nvar * 3 * nl[0] * nl[1]]
        recv = [recv[f][:lens[f // 2]] for f in range(6)]
        return ext, list(nbr), recv


def random_state(n, nchem, seed=1234, gamma=1.4):
    """Admissible seeded state: rho=1+0.5U, v_i=0.3(U-0.5), p=1+0.5U, tracers U
    (the generator SURVEY.md section 8(d) proposes for throughput runs)."""
    rng = np.random.default_rng(seed)
    N = int(n[0]) * int(n[1]) * int(n[2])
    rho = 1.0 + 0.5 * rng.random(N)
    vx, vy, vz = (0.3 * (rng.random(N) - 0.5) for _ in range(3))
    p = 1.0 + 0.5 * rng.random(N)
    et = p / (gamma - 1.0) + 0.5 * rho * (vx * vx + vy * vy + vz * vz)
    parts = [rho, rho * vx, rho * vy, rho * vz, et]
    parts.append(rng.random(N * nchem) if nchem > 0 else None)
    return parts
