"""euler-b200: B200-native fluid right-hand side of sundials-manyvector-demo.

Host-side mirror (Python) of the reference's interface for this path, over the C ABI in
``include/eulerb200.h`` (``libeulerb200.so``, hand-written sm_100a CUDA):

=====================================  ==================================================
reference (``/root/reference/src``)    here
=====================================  ==================================================
``class EulerData`` euler3D.hpp:177    :class:`EulerData` (same field names)
``EulerData::SetupDecomp`` :396        :meth:`EulerData.SetupDecomp`
``ExchangeStart/ExchangeEnd`` :577     :meth:`EulerData.ExchangeStart` / ``ExchangeEnd``
``N_VMake_MPIManyVector`` 5+1 subvecs  :class:`ManyVector` (5 fluid + 1 chem sub-vectors)
``fEuler`` utilities.cpp:17            :func:`fEuler`
``stability`` utilities.cpp:483        :func:`stability`
``external_forces`` hook               ``EulerData.forcing`` (constant per fluid field), or
                                       ``fEuler(..., external_forces=hook)`` for any hook
=====================================  ==================================================

PyTorch is used for device memory, streams and ``torch.distributed`` only.  There is no
CPU implementation here: importing works anywhere, but creating a context without the
built library or without a CUDA device raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libeulerb200.so")

BC_PERIODIC, BC_NEUMANN, BC_DIRICHLET, BC_REFLECTING = 0, 1, 2, 3
NO_NEIGHBOR = -1
FACES = ("W", "E", "S", "N", "B", "F")

_dp = C.POINTER(C.c_double)
_vp6 = C.c_void_p * 6


class Config(C.Structure):
    """``eulerb200_config`` (include/eulerb200.h)."""
    _fields_ = [("nxl", C.c_int64), ("nyl", C.c_int64), ("nzl", C.c_int64),
                ("nchem", C.c_int32), ("device", C.c_int32),
                ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double), ("gamma", C.c_double),
                ("bc", C.c_int32 * 6), ("nbr", C.c_int32 * 6),
                ("rank", C.c_int32), ("nranks", C.c_int32), ("forcing", C.c_double * 5)]


# every symbol include/eulerb200.h declares: name -> (restype, argtypes)
ABI = {
    "eulerb200_version": (C.c_int, []),
    "eulerb200_decompose": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                                      C.POINTER(C.c_int32)]),
    "eulerb200_exchange_plan": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_int32)]),
    "eulerb200_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "eulerb200_destroy": (C.c_int, [C.c_void_p]),
    "eulerb200_last_error": (C.c_char_p, [C.c_void_p]),
    "eulerb200_comm_unique_id": (C.c_int, [C.c_void_p]),
    "eulerb200_comm_attach": (C.c_int, [C.c_void_p, C.c_void_p]),
    "eulerb200_p2p_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "eulerb200_p2p_attach": (C.c_int, [C.c_void_p, C.c_void_p]),
    "eulerb200_rhs": (C.c_int, [C.c_void_p, C.c_double, _vp6, _vp6, C.c_void_p]),
    "eulerb200_rhs_async": (C.c_int, [C.c_void_p, C.c_double, _vp6, _vp6, C.c_void_p]),
    "eulerb200_state_flag": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    "eulerb200_rhs_slow": (C.c_int, [C.c_void_p, C.c_double, _vp6, _vp6, C.c_double, C.c_void_p]),
    "eulerb200_rhs_host": (C.c_int, [C.c_void_p, C.c_double, _vp6, _vp6]),
    "eulerb200_rhs_any": (C.c_int, [C.c_void_p, C.c_double, _vp6, _vp6, C.c_void_p]),
    "eulerb200_exchange_start": (C.c_int, [C.c_void_p, _vp6, C.c_void_p]),
    "eulerb200_exchange_end": (C.c_int, [C.c_void_p, C.c_void_p]),
    "eulerb200_face_len": (C.c_int64, [C.c_void_p, C.c_int32]),
    "eulerb200_ghost_face": (C.c_int, [C.c_void_p, _vp6, C.c_int32, C.c_void_p, C.c_void_p]),
    "eulerb200_stability": (C.c_int, [C.c_void_p, _vp6, C.c_double, C.POINTER(C.c_double), C.c_void_p]),
    "eulerb200_stability_any": (C.c_int, [C.c_void_p, _vp6, C.c_double, C.POINTER(C.c_double), C.c_void_p]),
    "eulerb200_vec_lincomb": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_void_p),
                                       C.c_void_p, C.c_int64, C.c_void_p]),
    "eulerb200_vec_wrms_accum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                          C.c_int64, C.c_void_p, C.c_void_p]),
    "eulerb200_vec_wrms": (C.c_int, [C.c_void_p, _vp6, _vp6, C.c_double, C.c_double, C.c_int64,
                                    C.POINTER(C.c_double), C.c_void_p]),
    "eulerb200_device_alloc": (C.c_void_p, [C.c_int64]),
    "eulerb200_device_free": (None, [C.c_void_p]),
    "eulerb200_managed_alloc": (C.c_void_p, [C.c_int64]),
    "eulerb200_synchronize": (C.c_int, [C.c_void_p]),
    "eulerb200_copy_to_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "eulerb200_copy_to_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "eulerb200_launch_count": (C.c_int64, [C.c_void_p]),
    "eulerb200_set_forcing_in_wdot": (C.c_int, [C.c_void_p, C.c_int32]),
    "eulerb200_profile": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_double)]),
    "eulerb200_fp64_peak": (C.c_int, [C.POINTER(C.c_double)]),
}

_lib = None


def load_library():
    """dlopen libeulerb200.so and bind the ABI.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("EULERB200_LIB", LIB_PATH)      # tuning: another build of the same library
    if not os.path.exists(path):
        raise RuntimeError(
            "%s is missing: build it with __graft_entry__.build() (nvcc, sm_100a). "
            "There is no CPU fallback for the fluid RHS." % path)
    lib = C.CDLL(path)
    for name, (res, args) in ABI.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class EulerB200Error(RuntimeError):
    pass


def dims_and_extents(nprocs, rank, n, bc):
    """``EulerData::SetupDecomp`` arithmetic (euler3D.hpp:416-567) through the C ABI.
    Returns (ret, dims, coords, ext[is,ie,js,je,ks,ke], nbr[ipW..ipF])."""
    lib = load_library()
    dims = (C.c_int32 * 3)()
    coords = (C.c_int32 * 3)()
    ext = (C.c_int64 * 6)()
    nbr = (C.c_int32 * 6)()
    ret = lib.eulerb200_decompose(nprocs, rank, (C.c_int64 * 3)(*[int(x) for x in n]),
                                  (C.c_int32 * 6)(*[int(b) for b in bc]), dims, coords, ext, nbr)
    return ret, list(dims), list(coords), list(ext), list(nbr)


class ManyVector:
    """The MPIManyVector composition of the reference drivers (euler3D_main.cpp:148-175):
    five fluid sub-vectors rho,mx,my,mz,et of N doubles each and, if nchem > 0, one
    chemistry sub-vector of N*nchem doubles (species fastest).  Sub-vectors are CUDA
    float64 tensors (or pinned/pageable host tensors for the *_host path)."""

    def __init__(self, subvecs):
        self.sub = list(subvecs)

    @classmethod
    def new(cls, udata, device="cuda", pin=False):
        import torch
        N = udata.nxl * udata.nyl * udata.nzl
        kw = dict(dtype=torch.float64, device=device)
        if pin and str(device) == "cpu":
            kw["pin_memory"] = True
        subs = [torch.zeros(N, **kw) for _ in range(5)]
        if udata.nchem > 0:
            subs.append(torch.zeros(N * udata.nchem, **kw))
        return cls(subs)

    def N_VGetSubvectorArrayPointer(self, i):
        """N_VGetSubvectorArrayPointer_MPIManyVector (utilities.cpp:31-58)."""
        return self.sub[i].data_ptr() if i < len(self.sub) else None

    def pointers(self):
        out = _vp6()
        for i in range(6):
            out[i] = self.sub[i].data_ptr() if i < len(self.sub) and self.sub[i] is not None else None
        return out

    @property
    def is_cuda(self):
        return self.sub[0].is_cuda


class EulerData:
    """Mirror of ``class EulerData`` (euler3D.hpp:177-1436), hot-path part.

    Usage follows the reference drivers: set ``nx,ny,nz``, the box, the six BC codes,
    ``gamma`` (and ``nchem``, which the reference fixes at compile time through NVAR),
    then ``SetupDecomp()``; afterwards pass the object as ``user_data`` to :func:`fEuler`.
    """

    def __init__(self, nchem=0):
        self.nx = self.ny = self.nz = 3
        self.xl, self.xr, self.yl, self.yr, self.zl, self.zr = 0.0, 1.0, 0.0, 1.0, 0.0, 1.0
        self.xlbc = self.xrbc = self.ylbc = self.yrbc = self.zlbc = self.zrbc = BC_PERIODIC
        self.gamma = 1.4
        self.cfl = 0.0
        self.nchem = int(nchem)
        self.forcing = [0.0] * 5
        self.myid, self.nprocs = 0, 1
        self.npx = self.npy = self.npz = -1
        self.is_ = self.ie = self.js = self.je = self.ks = self.ke = -1
        self.nxl = self.nyl = self.nzl = -1
        self.dx = self.dy = self.dz = 0.0
        self.ipW = self.ipE = self.ipS = self.ipN = self.ipB = self.ipF = NO_NEIGHBOR
        self.device = -1
        self._ctx = None

    # ---- set-up --------------------------------------------------------------------
    @property
    def bcs(self):
        return [self.xlbc, self.xrbc, self.ylbc, self.yrbc, self.zlbc, self.zrbc]

    @property
    def nbrs(self):
        return [self.ipW, self.ipE, self.ipS, self.ipN, self.ipB, self.ipF]

    def SetupDecomp(self, myid=0, nprocs=1, device=None, process_group=None):
        """euler3D.hpp:396-574.  Returns 0 on success (reference convention).
        ``process_group`` (torch.distributed) is only used to hand the NCCL id around."""
        if self._ctx is not None:
            return 1          # "parallel decomposition already set up"
        self.myid, self.nprocs = int(myid), int(nprocs)
        self.dx = (self.xr - self.xl) / self.nx
        self.dy = (self.yr - self.yl) / self.ny
        self.dz = (self.zr - self.zl) / self.nz
        ret, dims, coords, ext, nbr = dims_and_extents(self.nprocs, self.myid,
                                                       (self.nx, self.ny, self.nz), self.bcs)
        if ret != 0:
            return ret
        self.npx, self.npy, self.npz = dims
        self.is_, self.ie, self.js, self.je, self.ks, self.ke = ext
        self.nxl, self.nyl, self.nzl = ext[1] - ext[0] + 1, ext[3] - ext[2] + 1, ext[5] - ext[4] + 1
        self.ipW, self.ipE, self.ipS, self.ipN, self.ipB, self.ipF = nbr
        if device is not None:
            self.device = int(device)
        self._create()
        if self.nprocs > 1:
            self._attach_comm(process_group)
        return 0

    def _config(self):
        c = Config()
        c.nxl, c.nyl, c.nzl = self.nxl, self.nyl, self.nzl
        c.nchem, c.device = self.nchem, self.device
        c.dx, c.dy, c.dz, c.gamma = self.dx, self.dy, self.dz, self.gamma
        for f in range(6):
            c.bc[f] = self.bcs[f]
            c.nbr[f] = self.nbrs[f]
        c.rank, c.nranks = self.myid, self.nprocs
        for f in range(5):
            c.forcing[f] = float(self.forcing[f])
        return c

    def exchange_plan(self):
        """[(kind, face, peer)] in issue order; kind 'send' / 'recv' (eulerb200_exchange_plan)."""
        ops = (C.c_int32 * 36)()
        cfg = self._config()
        n = load_library().eulerb200_exchange_plan(C.byref(cfg), ops)
        return [("send" if ops[3 * q] == 0 else "recv", int(ops[3 * q + 1]), int(ops[3 * q + 2])) for q in range(n)]

    def _create(self):
        lib = load_library()
        ctx = C.c_void_p()
        cfg = self._config()
        ret = lib.eulerb200_create(C.byref(cfg), C.byref(ctx))
        if ret != 0:
            raise EulerB200Error("eulerb200_create failed (%d): %s"
                                 % (ret, lib.eulerb200_last_error(None).decode()))
        self._ctx = ctx

    def _attach_comm(self, process_group=None):
        import torch
        import torch.distributed as dist
        lib = load_library()
        idbuf = (C.c_char * 128)()
        if self.myid == 0:
            self._check(lib.eulerb200_comm_unique_id(idbuf), None)
        t = torch.frombuffer(bytearray(bytes(idbuf)), dtype=torch.uint8).clone()
        if dist.get_backend(process_group) == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0, group=process_group)
        raw = bytes(t.cpu().numpy().tobytes())
        self._check(lib.eulerb200_comm_attach(self._ctx, C.create_string_buffer(raw, 128)), self._ctx)
        # halo transport: "nccl" (default: the one validated at 8 GPUs, profiles/README.md) or the
        # peer-store transport over CUDA IPC with EULERB200_HALO=p2p; every rank must agree
        self.halo_transport = "nccl"
        if os.environ.get("EULERB200_HALO", "nccl") == "p2p":
            blob = (C.c_char * 256)()
            ok = lib.eulerb200_p2p_export(self._ctx, blob) == 0
            mine = torch.frombuffer(bytearray(bytes(blob)), dtype=torch.uint8).clone()
            if dist.get_backend(process_group) == "nccl":
                mine = mine.cuda()
            allb = [torch.empty_like(mine) for _ in range(self.nprocs)]
            dist.all_gather(allb, mine, group=process_group)
            rawall = b"".join(bytes(x.cpu().numpy().tobytes()) for x in allb)
            ok = ok and lib.eulerb200_p2p_attach(self._ctx, C.create_string_buffer(rawall, len(rawall))) == 0
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32)
            if dist.get_backend(process_group) == "nccl":
                flag = flag.cuda()
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=process_group)
            if int(flag.item()) == 1:
                self.halo_transport = "p2p"
            else:
                raise EulerB200Error("peer-store halo transport unavailable on some rank (%s); "
                                     "set EULERB200_HALO=nccl" % lib.eulerb200_last_error(self._ctx).decode())

    def _check(self, ret, ctx):
        if ret != 0:
            raise EulerB200Error("eulerb200 call failed (%d): %s"
                                 % (ret, load_library().eulerb200_last_error(ctx).decode()))

    def FreeData(self):
        """euler3D.hpp:304-377"""
        if self._ctx is not None:
            load_library().eulerb200_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.FreeData()
        except Exception:
            pass

    # ---- halo exchange -------------------------------------------------------------
    @staticmethod
    def _stream():
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def ExchangeStart(self, w):
        """euler3D.hpp:577-1169"""
        return load_library().eulerb200_exchange_start(self._ctx, w.pointers(), self._stream())

    def ExchangeEnd(self):
        """euler3D.hpp:1172-1191"""
        return load_library().eulerb200_exchange_end(self._ctx, self._stream())

    def recv_buffer(self, w, face):
        """Ghost layers of a face in the layout of Wrecv..Frecv (euler3D.hpp:257-262)."""
        import torch
        lib = load_library()
        n = lib.eulerb200_face_len(self._ctx, face)
        out = torch.empty(n, dtype=torch.float64, device=w.sub[0].device)
        self._check(lib.eulerb200_ghost_face(self._ctx, w.pointers(), face, C.c_void_p(out.data_ptr()),
                                             self._stream()), self._ctx)
        return out

    # ---- scalar helpers of the class -----------------------------------------------
    def eos(self, rho, mx, my, mz, et):
        """euler3D.hpp:1383-1388"""
        return (self.gamma - 1.0) * (et - (mx * mx + my * my + mz * mz) * 0.5 / rho)

    def eos_inv(self, rho, mx, my, mz, pr):
        """euler3D.hpp:1393-1398"""
        return pr / (self.gamma - 1.0) + (mx * mx + my * my + mz * mz) * 0.5 / rho

    def launch_count(self):
        return int(load_library().eulerb200_launch_count(self._ctx))

    def profile(self, on=True, reset=False):
        """Device-time profile of the RHS calls since the last reset (eulerb200_profile): dict of
        average milliseconds per call, the counterpart of the reference's Profile slots."""
        out = (C.c_double * 8)()
        self._check(load_library().eulerb200_profile(self._ctx, 1 if on else 0, 1 if reset else 0, out), self._ctx)
        keys = ("rhs", "prepass", "pack", "transfer", "interior", "halo_wait", "shells", "calls")
        return dict(zip(keys, [float(x) for x in out]))

    def last_error(self):
        return load_library().eulerb200_last_error(self._ctx).decode()


def fEuler(t, w, wdot, user_data, sync=True, external_forces=None):
    """``int fEuler(realtype t, N_Vector w, N_Vector wdot, void* user_data)``
    (utilities.cpp:17-253).  Returns 0, or -1 on an illegal state / failure like the
    reference.  Device vectors run on the current CUDA stream; host vectors go through
    the staged host path.  ``sync=False`` skips the legal-state read-back (no host sync).

    ``external_forces(t, G, user_data) -> int`` is the reference's link-time hook
    (euler3D.hpp:1454) for forcing that is not a per-field constant: as at utilities.cpp:28,65
    ``wdot`` is zeroed, the hook ASSIGNS G into it, and the kernel subtracts the flux
    divergence from what it finds there (``user_data.forcing`` is then ignored)."""
    lib = load_library()
    u = user_data
    if u._ctx is None:
        raise EulerB200Error("EulerData.SetupDecomp() has not been called")
    if bool(external_forces) != getattr(u, "_forcing_in_wdot", False):
        if lib.eulerb200_set_forcing_in_wdot(u._ctx, 1 if external_forces else 0) != 0:
            return -1
        u._forcing_in_wdot = bool(external_forces)
    if external_forces:
        for sub in wdot.sub:
            sub.zero_()
        if external_forces(float(t), wdot, u) != 0:
            return -1
    if w.is_cuda:
        fn = lib.eulerb200_rhs if sync else lib.eulerb200_rhs_async
        return fn(u._ctx, float(t), w.pointers(), wdot.pointers(), u._stream())
    return lib.eulerb200_rhs_host(u._ctx, float(t), w.pointers(), wdot.pointers())


def fslow(t, w, wdot, user_data):
    """``fslow`` (multirate_chem_hydro_main.cpp:996-1083) / ``fexpl`` (imex_chem_hydro_main.cpp:
    910-1000) without the Dengo scaling: total energy rebuilt from the gas energy species,
    fEuler, ``chemdot[nchem-1] = etdot``, ``etdot = 0`` -- one fused call, no host copies of
    the chemistry vector.  ``user_data.EnergyUnits`` as in euler3D.hpp:233.  Mutates ``w``'s et."""
    lib = load_library()
    u = user_data
    return lib.eulerb200_rhs_slow(u._ctx, float(t), w.pointers(), wdot.pointers(),
                                  float(getattr(u, "EnergyUnits", 1.0)), u._stream())


def stability(w, t, user_data):
    """``int stability(N_Vector w, realtype t, realtype* dt_stab, void* user_data)``
    (utilities.cpp:483-528).  Returns (retval, dt_stab)."""
    lib = load_library()
    u = user_data
    dt = C.c_double(0.0)
    ret = lib.eulerb200_stability(u._ctx, w.pointers(), float(u.cfl), C.byref(dt), u._stream())
    return ret, dt.value


def __getattr__(name):
    """Lazy sub-modules: ``driver`` (explicit time stepping) and
    ``problems`` (initial conditions / diagnostics, 8(f-2), 8(f-4))."""
    if name in ("driver", "problems"):
        import importlib
        return importlib.import_module(__name__ + "." + name)
    raise AttributeError(name)


def check_flag(flag, funcname, opt):
    """utilities.cpp:542-591 (options 1 and 4, the ones the hot path raises)."""
    import sys
    if opt == 1 and flag < 0:
        sys.stderr.write("\nSUNDIALS_ERROR: %s failed with flag = %d\n\n" % (funcname, flag))
        return 1
    if opt == 4 and flag != 0:
        names = {1: "illegal density", 2: "illegal energy", 3: "illegal density & energy",
                 4: "illegal pressure", 5: "illegal density & pressure", 6: "illegal energy & pressure",
                 7: "illegal density, energy & pressure"}
        sys.stderr.write("\nSTATE_ERROR: %s failed with flag = %d  (%s)\n\n"
                         % (funcname, flag, names.get(flag, "")))
        return 1
    return 0
