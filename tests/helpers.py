"""Test helpers: run the CUDA path through the package (which calls the C ABI)."""
import numpy as np


def make_udata(pkg, n, nchem, bcs, box=(0, 1, 0, 1, 0, 1), gamma=1.4, forcing=None, device=0):
    u = pkg.EulerData(nchem=nchem)
    u.nx, u.ny, u.nz = n
    u.xl, u.xr, u.yl, u.yr, u.zl, u.zr = box
    u.xlbc, u.xrbc, u.ylbc, u.yrbc, u.zlbc, u.zrbc = bcs
    u.gamma = gamma
    if forcing is not None:
        u.forcing = list(forcing)
    assert u.SetupDecomp(device=device) == 0
    return u


def gpu_feuler(pkg, u, parts, host=False):
    """parts: list of 5 (+1) numpy arrays.  Returns (ret, list of numpy arrays)."""
    import torch
    if host:
        w = pkg.ManyVector([torch.from_numpy(p) for p in parts if p is not None])
        wdot = pkg.ManyVector.new(u, device="cpu")
    else:
        w = pkg.ManyVector([torch.from_numpy(p).cuda() for p in parts if p is not None])
        wdot = pkg.ManyVector.new(u)
        for s in wdot.sub:
            s.fill_(float("nan"))
    ret = pkg.fEuler(0.0, w, wdot, u)
    torch.cuda.synchronize()
    out = [s.cpu().numpy() for s in wdot.sub]
    if len(out) == 5:
        out.append(None)
    return ret, out


def oracle_feuler(port, u, parts, n=None):
    cfg = port.cfg(n or (u.nx, u.ny, u.nz), u.nchem, (u.dx, u.dy, u.dz), u.gamma, u.bcs, forcing=u.forcing)
    return port.feuler(cfg, parts)


class NpVec:
    """ManyVector look-alike on numpy arrays (test infrastructure for the driver tests)."""

    def __init__(self, subs):
        self.sub = list(subs)


class OracleVecOps:
    """The driver's VecOps interface with numpy arithmetic and the CPU ORACLE as right-hand
    side: the same ERKStep loop can then be run without a GPU, and its trajectory compared
    with the one the CUDA path produces.  Test infrastructure only."""

    def __init__(self, port, udata_like, n, nchem, d, gamma, bcs, forcing=None):
        self.port = port
        self.cfg = port.cfg(n, nchem, d, gamma, bcs, forcing=forcing)
        self.nglobal = n[0] * n[1] * n[2] * (5 + nchem)
        self.cfl = 0.0

    def new_like(self, w):
        return NpVec([np.empty_like(s) for s in w.sub])

    def lincomb(self, out, coefs, vecs):
        for f in range(len(out.sub)):
            acc = coefs[0] * vecs[0].sub[f]
            for c, v in zip(coefs[1:], vecs[1:]):
                acc = acc + c * v.sub[f]
            out.sub[f][...] = acc

    def wrms(self, x, y, rtol, atol):
        tot = 0.0
        for a, b in zip(x.sub, y.sub):
            q = a / (rtol * np.abs(b) + atol)
            tot += float(np.dot(q, q))
        return float(np.sqrt(tot / self.nglobal))

    def rhs(self, t, w, wdot):
        parts = list(w.sub) + ([None] if len(w.sub) == 5 else [])
        ret, out, _ = self.port.feuler(self.cfg, parts)
        for f in range(len(w.sub)):
            wdot.sub[f][...] = out[f]
        return ret

    def stability(self, w, t):
        parts = list(w.sub) + ([None] if len(w.sub) == 5 else [])
        return 0, self.port.dt_stab(self.cfg, self.cfl, self.port.max_wavespeed(self.cfg, parts))
