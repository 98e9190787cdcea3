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
