"""Problem plug-ins: the reference links one problem file per
executable, each providing ``initial_conditions``, ``external_forces`` and
``output_diagnostics`` (euler3D.hpp:1448-1457).  Here a problem is a name:

    sod_x | sod_y | sod_z                     sod.cpp
    linear_advection_x | _y | _z              linear_advection.cpp
    rayleigh_taylor                           rayleigh_taylor.cpp
    hurricane_xy | hurricane_yz | hurricane_zx    hurricane.cpp
    fluid_blast (nchem = 0) | primordial_blast (nchem = 10, fluid + tracer values only)
                                              fluid_blast.cpp, primordial_blast.cpp:64-309

plus the two diagnostics of io.cpp every driver prints: ``check_conservation`` (:504-541)
and ``print_stats`` (:552-636), and ``output_solution`` / ``read_restart`` (:716-1150) on a flat
file with the reference's dataset order (no HDF5 in this image; the same ``.eb200`` files are
written and read by the native driver, host/problems.hpp).  States are built and reduced on the device (torch is the
array plumbing here; none of this is on the timed path).  The exact Riemann solution used
by the Sod diagnostics (sod.cpp:214-379) is evaluated on the host per x-location.
"""
import math

BC_PERIODIC, BC_NEUMANN, BC_DIRICHLET, BC_REFLECTING = 0, 1, 2, 3


def _coords(torch, u, device):
    f64 = dict(dtype=torch.float64, device=device)
    x = (torch.arange(u.nxl, **f64) + (u.is_ + 0.5)) * u.dx + u.xl
    y = (torch.arange(u.nyl, **f64) + (u.js + 0.5)) * u.dy + u.yl
    z = (torch.arange(u.nzl, **f64) + (u.ks + 0.5)) * u.dz + u.zl
    Z, Y, X = torch.meshgrid(z, y, x, indexing="ij")       # flat index i + nxl*(j + nyl*k)
    return X.reshape(-1), Y.reshape(-1), Z.reshape(-1)


class MT19937_64:
    """std::mt19937_64 (the clump generator of fluid_blast.cpp:102) -- numpy only has the 32-bit
    twister.  uniform(a, b) follows libstdc++'s uniform_real_distribution<double>:
    a + (b - a) * double(x) / 2^64 with one 64-bit draw per value."""
    N, M = 312, 156
    MASK = (1 << 64) - 1

    def __init__(self, seed):
        self.mt = [0] * self.N
        self.mt[0] = seed & self.MASK
        for i in range(1, self.N):
            self.mt[i] = (6364136223846793005 * (self.mt[i - 1] ^ (self.mt[i - 1] >> 62)) + i) & self.MASK
        self.idx = self.N

    def next(self):
        if self.idx >= self.N:
            mt, N, M = self.mt, self.N, self.M
            for i in range(N):
                x = (mt[i] & 0xFFFFFFFF80000000) | (mt[(i + 1) % N] & 0x7FFFFFFF)
                xa = x >> 1
                if x & 1:
                    xa ^= 0xB5026F5AA96619E9
                mt[i] = mt[(i + M) % N] ^ xa
            self.idx = 0
        y = self.mt[self.idx]
        self.idx += 1
        y ^= (y >> 29) & 0x5555555555555555
        y ^= (y << 17) & 0x71D67FFFEDA60000
        y ^= (y << 37) & 0xFFF7EEE000000000
        y ^= y >> 43
        return y & self.MASK

    def uniform(self, a, b):
        r = float(self.next()) / 18446744073709551616.0
        if r >= 1.0:
            r = 1.0 - 2.0 ** -53
        return r * (b - a) + a


def blast_clumps(u, max_strength):
    """10*nprocs Gaussian clumps: centre, radius (cells) and strength (fluid_blast.cpp:98-128,
    primordial_blast.cpp:102-123; MAX_CLUMP_STRENGTH is 10 resp. 5)."""
    gen = MT19937_64(u.nprocs)
    out = []
    for _ in range(10 * u.nprocs):
        cx, cy, cz = gen.uniform(u.xl, u.xr), gen.uniform(u.yl, u.yr), gen.uniform(u.zl, u.zr)
        out.append((cx, cy, cz, gen.uniform(3.0, 6.0), gen.uniform(0.0, max_strength)))
    return out


def configure(problem, u):
    """Domain / boundary conditions / gamma of the shipped input files for `problem`
    (tests/*/input_*.txt); grid sizes are left to the caller."""
    if problem.startswith("sod"):
        u.xlbc = u.xrbc = u.ylbc = u.yrbc = u.zlbc = u.zrbc = BC_NEUMANN
        u.gamma = 1.4
    elif problem.startswith("linear_advection"):
        u.xlbc = u.xrbc = u.ylbc = u.yrbc = u.zlbc = u.zrbc = BC_PERIODIC
        u.gamma = 1.4
    elif problem == "rayleigh_taylor":
        u.xl, u.xr, u.yl, u.yr = -0.25, 0.25, -0.75, 0.75
        u.xlbc = u.xrbc = BC_PERIODIC
        u.ylbc = u.yrbc = BC_REFLECTING
        u.zlbc = u.zrbc = BC_NEUMANN
        u.gamma = 1.4
        u.forcing = [0.0, 0.0, -0.1, 0.0, 0.0]            # rayleigh_taylor.cpp:117-128
    elif problem.startswith("hurricane"):
        u.xl = u.yl = u.zl = -1.0
        u.xr = u.yr = u.zr = 1.0
        u.xlbc = u.xrbc = u.ylbc = u.yrbc = u.zlbc = u.zrbc = BC_NEUMANN
        u.gamma = 2.0
    elif problem in ("fluid_blast", "primordial_blast"):
        # tests/fluid_blast/input_fluid_blast.txt, tests/primordial_blast/input_primordial_blast_mr.txt
        u.xlbc = u.xrbc = u.ylbc = u.yrbc = u.zlbc = u.zrbc = BC_REFLECTING
        u.gamma = 5.0 / 3.0
        u.MassUnits, u.LengthUnits = 3.0e70, 3.0857e30
        u.TimeUnits = 1.0e12 if problem == "fluid_blast" else 1.0e11
        u.DensityUnits = u.MassUnits / u.LengthUnits / u.LengthUnits / u.LengthUnits      # euler3D.hpp:385-393
        u.MomentumUnits = u.MassUnits / u.LengthUnits / u.LengthUnits / u.TimeUnits
        u.EnergyUnits = u.MassUnits / u.LengthUnits / u.TimeUnits / u.TimeUnits
    else:
        raise ValueError("unknown problem %r" % problem)


def _blast_state(problem, w, u, X, Y, Z):
    """fluid_blast.cpp:140-264 / primordial_blast.cpp:180-297: clumpy neutral primordial gas at rest
    plus a hot dense central clump; the ten tracers are the eight number densities, the electron
    density and the gas energy."""
    import torch
    mH, kboltz, Hfrac = 1.67e-24, 1.3806488e-16, 0.76
    m_amu = 1.66053904e-24
    density0 = 1e2 * mH
    # the two problem files differ in three constants (fluid_blast.cpp:50-58, primordial_blast.cpp:50-60)
    fluid = problem == "fluid_blast"
    max_strength, blast_density = (10.0, 10.0) if fluid else (5.0, 5.0)
    blast_temp = 10.0 * 5.0 if fluid else 10.0 * (5.0 - 1.0)
    density = torch.ones_like(X)
    for cx, cy, cz, cr, cs in blast_clumps(u, max_strength):
        cr = cr * u.dx
        rsq = (X - cx).abs() ** 2 + (Y - cy).abs() ** 2 + (Z - cz).abs() ** 2
        density = density + cs * torch.exp(-2.0 * rsq / cr / cr)
    density = density * density0
    cx, cy, cz = u.xl + 0.5 * (u.xr - u.xl), u.yl + 0.5 * (u.yr - u.yl), u.zl + 0.5 * (u.zr - u.zl)
    cr = 0.1 * min(u.xr - u.xl, u.yr - u.yl, u.zr - u.zl)
    rsq = (X - cx).abs() ** 2 + (Y - cy).abs() ** 2 + (Z - cz).abs() ** 2
    bump = torch.exp(-2.0 * rsq / cr / cr)
    density = density + density0 * blast_density * bump
    T = 10.0 + blast_temp * bump
    inside = rsq / cr / cr < 2.0
    tiny, small = 1e-40, 1e-12
    pick = lambda a: torch.where(inside, a * density, 1.0e-3 * density)
    H2I, H2II, HII, HM, HeII, HeIII = pick(tiny), pick(tiny), pick(small), pick(tiny), pick(small), pick(small)
    HeI = (1.0 - Hfrac) * density - HeII - HeIII
    HI = density - (H2I + H2II + HII + HM + HeI + HeII + HeIII)
    wH, wHe = 1.00794 * mH, 4.002602 * mH
    nH2I, nH2II, nHII, nHM = H2I / (2 * wH), H2II / (2 * wH), HII / wH, HM / wH
    nHeII, nHeIII, nHeI, nHI = HeII / wHe, HeIII / wHe, HeI / wHe, HI / wH
    ndens = nH2I + nH2II + nHII + nHM + nHeII + nHeIII + nHeI + nHI
    ge = (kboltz * T * ndens) / (density * (u.gamma - 1.0))
    zero = torch.zeros_like(X)
    for dst, src in zip(w.sub[:5], (density / u.DensityUnits, zero, zero, zero, ge / u.EnergyUnits)):
        dst.copy_(src)
    if u.nchem > 0:
        if u.nchem != 10:
            raise ValueError("primordial_blast carries 10 species")
        de = (nHII + nHeII + 2 * nHeIII - nHM + nH2II) * mH
        chem = torch.stack([nH2I, nH2II, nHI, nHII, nHM, nHeI, nHeII, nHeIII, de / m_amu, ge], dim=1)
        w.sub[5].copy_(chem.reshape(-1))
    return 0


def initial_conditions(problem, t, w, u):
    """``int initial_conditions(const realtype& t, N_Vector w, const EulerData& udata)``"""
    import torch
    dev = w.sub[0].device
    X, Y, Z = _coords(torch, u, dev)
    zero = torch.zeros_like(X)
    mx, my, mz = zero.clone(), zero.clone(), zero.clone()
    if problem in ("fluid_blast", "primordial_blast"):
        return _blast_state(problem, w, u, X, Y, Z)
    if problem.startswith("sod"):                                   # sod.cpp:50-55,120-160
        s = {"x": X, "y": Y, "z": Z}[problem[-1]]
        left = s < 0.5
        one = torch.ones_like(X)
        rho = torch.where(left, one, 0.125 * one)
        p = torch.where(left, one, 0.1 * one)
    elif problem.startswith("linear_advection"):                    # linear_advection.cpp:51-68,117-131
        s = {"x": X, "y": Y, "z": Z}[problem[-1]]
        rho = 1.0 + 0.1 * torch.sin(2.0 * math.pi * (s - 0.5 * t))
        v = 0.5 * rho
        mx, my, mz = (v if problem[-1] == "x" else zero, v if problem[-1] == "y" else zero,
                      v if problem[-1] == "z" else zero)
        p = torch.ones_like(X)
    elif problem == "rayleigh_taylor":                              # rayleigh_taylor.cpp:48-53,104-109
        rho = torch.where(Y > 0.0, 2.0 * torch.ones_like(X), torch.ones_like(X))
        my = rho * 0.01 * (1.0 + torch.cos(4.0 * math.pi * X)) * (1.0 + torch.cos(3.0 * math.pi * Y))
        p = 2.5 - 0.1 * rho * Y
    elif problem.startswith("hurricane"):                           # hurricane.cpp:58-60,120-183
        a, b = {"xy": (X, Y), "zx": (Z, X), "yz": (Y, Z)}[problem[-2:]]
        r = torch.sqrt(a * a + b * b)
        r = torch.where(r == 0.0, torch.full_like(r, 1e-14), r)
        ma, mb = 10.0 * (b / r), -10.0 * (a / r)
        if problem[-2:] == "xy":
            mx, my = ma, mb
        elif problem[-2:] == "zx":
            mz, mx = ma, mb
        else:
            my, mz = ma, mb
        rho = torch.ones_like(X)
        p = torch.full_like(X, 25.0)
        if u.nchem > 0:                                             # colour stripes in angle
            theta = torch.atan2(b, a)
            chem = torch.zeros(X.numel(), u.nchem, dtype=torch.float64, device=dev)
            for v in range(u.nchem):
                lo, hi = -math.pi + v * 2 * math.pi / u.nchem, -math.pi + (v + 1) * 2 * math.pi / u.nchem
                chem[:, v] = ((theta >= lo) & (theta < hi)).to(torch.float64)
            w.sub[5].copy_(chem.reshape(-1))
    else:
        raise ValueError("unknown problem %r" % problem)
    et = p / (u.gamma - 1.0) + (mx * mx + my * my + mz * mz) * 0.5 / rho       # eos_inv, euler3D.hpp:1393
    for dst, src in zip(w.sub[:5], (rho, mx, my, mz, et)):
        dst.copy_(src)
    if u.nchem > 0 and not problem.startswith("hurricane"):
        w.sub[5].zero_()
    return 0


# ---- exact Riemann solution of the Sod tube (sod.cpp:214-379), host side ----------------
def _fsecant(p4, p1, p5, rho1, rho5, g):
    z = p4 / p5 - 1.0
    c1, c5 = math.sqrt(g * p1 / rho1), math.sqrt(g * p5 / rho5)
    fact = (g - 1.0) / (2 * g) * (c5 / c1) * z / math.sqrt(1.0 + (g + 1.0) / (2 * g) * z)
    return p1 * (1.0 - fact) ** (2 * g / (g - 1.0)) - p4


def exact_riemann(t, xs, xI, g, rhoL=1.0, rhoR=0.125, pL=1.0, pR=0.1):
    """rho, u, p at time t for every x in xs (pL > pR branch, the shipped problem)."""
    rho1, p1, rho5, p5 = rhoL, pL, rhoR, pR
    p40, p41 = p1, p5
    f0 = _fsecant(p40, p1, p5, rho1, rho5, g)
    p4 = p41
    for _ in range(50):
        f1 = _fsecant(p41, p1, p5, rho1, rho5, g)
        if f1 == f0:
            break
        p4 = p41 - (p41 - p40) * f1 / (f1 - f0)
        if abs(p4 - p41) / abs(p41) < 1e-14:
            break
        p40, p41, f0 = p41, p4, f1
    z = p4 / p5 - 1.0
    c5 = math.sqrt(g * p5 / rho5)
    gm1, gp1 = g - 1.0, g + 1.0
    fact = math.sqrt(1.0 + 0.5 * gp1 * z / g)
    u4 = c5 * z / (g * fact)
    rho4 = rho5 * (1.0 + 0.5 * gp1 * z / g) / (1.0 + 0.5 * gm1 * z / g)
    wsh = c5 * fact
    p3, u3 = p4, u4
    rho3 = rho1 * (p3 / p1) ** (1.0 / g)
    c1, c3 = math.sqrt(g * p1 / rho1), math.sqrt(g * p3 / rho3)
    xsh, xcd, xft, xhd = xI + wsh * t, xI + u3 * t, xI + (u3 - c3) * t, xI - c1 * t
    out = []
    for x in xs:
        if x < xhd:
            out.append((rho1, 0.0, p1))
        elif x < xft:
            u = 2.0 / gp1 * (c1 + (x - xI) / t)
            f = 1.0 - 0.5 * gm1 * u / c1
            out.append((rho1 * f ** (2.0 / gm1), u, p1 * f ** (2.0 * g / gm1)))
        elif x < xcd:
            out.append((rho3, u3, p3))
        elif x < xsh:
            out.append((rho4, u4, p4))
        else:
            out.append((rho5, 0.0, p5))
    return out


def _true_state(problem, t, u, dev):
    import torch
    X, Y, Z = _coords(torch, u, dev)
    zero = torch.zeros_like(X)
    if problem.startswith("linear_advection"):
        s = {"x": X, "y": Y, "z": Z}[problem[-1]]
        rho = 1.0 + 0.1 * torch.sin(2.0 * math.pi * (s - 0.5 * t))
        v = 0.5 * rho
        m = [v if problem[-1] == a else zero for a in "xyz"]
        et = 1.0 / (u.gamma - 1.0) + (m[0] ** 2 + m[1] ** 2 + m[2] ** 2) * 0.5 / rho
        return [rho] + m + [et]
    if problem.startswith("sod"):
        ax = problem[-1]
        n1 = {"x": u.nxl, "y": u.nyl, "z": u.nzl}[ax]
        o1 = {"x": u.is_, "y": u.js, "z": u.ks}[ax]
        d1 = {"x": u.dx, "y": u.dy, "z": u.dz}[ax]
        l1 = {"x": u.xl, "y": u.yl, "z": u.zl}[ax]
        xs = [(o1 + q + 0.5) * d1 + l1 for q in range(n1)]
        sol = exact_riemann(t, xs, 0.5, u.gamma) if t > 0 else [((1.0, 0.0, 1.0) if x < 0.5 else (0.125, 0.0, 0.1)) for x in xs]
        tab = torch.tensor(sol, dtype=torch.float64, device=dev)              # [n1, 3]
        idx = torch.arange(u.nxl * u.nyl * u.nzl, device=dev)
        q = {"x": idx % u.nxl, "y": (idx // u.nxl) % u.nyl, "z": idx // (u.nxl * u.nyl)}[ax]
        rho, vel, p = tab[q, 0], tab[q, 1], tab[q, 2]
        m = [rho * vel if ax == a else zero for a in "xyz"]
        et = p / (u.gamma - 1.0) + (m[0] ** 2 + m[1] ** 2 + m[2] ** 2) * 0.5 / rho
        return [rho] + m + [et]
    if problem.startswith("hurricane"):
        # critical-rotation solution (hurricane.cpp:236-303): density and the three momenta only.
        # In the rotation plane (a, b): inside r < 2 t sqrt(p0') a paraboloid rho = r^2 / (8 A t^2)
        # with m = rho ((a+b), (b-a)) / (2t); outside rho0 and the velocity of a parcel that
        # started on the v0 circle.  p0' = A gamma rho0^(gamma-1), A = 25, rho0 = 1.
        pl = problem[-2:]
        a, b = {"xy": (X, Y), "zx": (Z, X), "yz": (Y, Z)}[pl]
        A, rho0 = 25.0, 1.0
        p0p = A * u.gamma * rho0 ** (u.gamma - 1.0)
        r = torch.sqrt(a * a + b * b)
        r = torch.where(r == 0, torch.full_like(r, 1e-14), r)
        ca, sb = a / r, b / r
        inside = r < 2.0 * t * math.sqrt(p0p)
        tt = t if t > 0 else 1.0                                   # (inside is empty at t = 0)
        rho_in = r * r / (8.0 * A * tt * tt)
        swirl = math.sqrt(2.0 * p0p) * torch.sqrt(torch.clamp(r * r - 2.0 * t * t * p0p, min=0.0))
        rho = torch.where(inside, rho_in, torch.full_like(r, rho0))
        ma = torch.where(inside, rho_in * (a + b) / (2.0 * tt), rho0 * (2.0 * t * p0p * ca + swirl * sb) / r)
        mb = torch.where(inside, rho_in * (b - a) / (2.0 * tt), rho0 * (2.0 * t * p0p * sb - swirl * ca) / r)
        m = {"xy": [ma, mb, zero], "zx": [mb, zero, ma], "yz": [zero, ma, mb]}[pl]
        return [rho] + m
    return None


def _allreduce(t, op, u, group):
    if u.nprocs > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=getattr(dist.ReduceOp, op), group=group)
    return t


def output_diagnostics(problem, t, w, u, group=None, quiet=False):
    """``output_diagnostics``: errI (max) and errR (RMS) against the analytic solution -- the five
    fluid fields for linear advection and Sod (linear_advection.cpp:168-240, sod.cpp:383-467),
    density and the three momenta for the hurricane problems (hurricane.cpp:217-334); None for
    problems whose reference hook prints nothing."""
    import torch
    true = _true_state(problem, t, u, w.sub[0].device)
    if true is None:
        return None
    errI = torch.stack([(a - b).abs().max() for a, b in zip(true, w.sub[:5])])
    errR = torch.stack([((a - b) ** 2).sum() for a, b in zip(true, w.sub[:5])])
    errI = _allreduce(errI, "MAX", u, group).cpu().tolist()
    errR = torch.sqrt(_allreduce(errR, "SUM", u, group) / (u.nx * u.ny * u.nz)).cpu().tolist()
    if u.myid == 0 and not quiet:
        print("     errI = " + "  ".join("%9.2e" % e for e in errI))
        print("     errR = " + "  ".join("%9.2e" % e for e in errR))
    return {"errI": errI, "errR": errR}


class Conservation:
    """``check_conservation`` (io.cpp:504-541): total mass and energy, then relative drift."""

    def __init__(self):
        self.saved = None

    def __call__(self, t, w, u, group=None, quiet=False):
        import torch
        scales = _unit_scales(u)                                  # CGS totals, io.cpp:522-524
        vol = u.dx * u.dy * u.dz * float(getattr(u, "LengthUnits", 1.0)) ** 3
        tot = torch.stack([w.sub[0].sum() * (vol * scales[0]), w.sub[4].sum() * (vol * scales[4])])
        tot = _allreduce(tot, "SUM", u, group).cpu().tolist()
        if self.saved is None:
            self.saved = tot
            if u.myid == 0 and not quiet:
                print("   Total mass   = %.16e\n   Total energy = %.16e" % tuple(tot))
            return {"mass": tot[0], "energy": tot[1]}
        drift = [abs(tot[i] - self.saved[i]) / self.saved[i] for i in range(2)]
        if u.myid == 0 and not quiet:
            print("   Mass conservation relative change   = %7.2e" % drift[0])
            print("   Energy conservation relative change = %7.2e" % drift[1])
        return {"mass": tot[0], "energy": tot[1], "mass_drift": drift[0], "energy_drift": drift[1]}


def print_stats(t, w, u, nst, group=None, quiet=False):
    """``print_stats`` (io.cpp:552-636): RMS of every field over the global grid."""
    import torch
    sq = [(s * s).sum() for s in w.sub[:5]]
    if u.nchem > 0:
        sq += list((w.sub[5].view(-1, u.nchem) ** 2).sum(0))
    tot = _allreduce(torch.stack(sq), "SUM", u, group)
    rms = torch.sqrt(tot / (u.nx * u.ny * u.nz)).cpu().tolist()
    if u.myid == 0 and not quiet:
        print("  %9.1e " % t + " ".join("%9.1e" % r for r in rms) + "  %6d" % nst)
    return rms


# ---- solution files (output_solution / read_restart, io.cpp:716-1150) ---------------------
# output-<iout>.eb200, little-endian:  char[8] "EB200OUT" | int32 version = 1 | int32 nchem |
# int64 nx, ny, nz | double time | double domain[6] = zl, zr, yl, yr, xl, xr (io.cpp:827-829) |
# 5 + nchem datasets of nx*ny*nz doubles, x fastest, in the reference's order and under the
# reference's names; the fluid fields are stored in CGS (scaled by the unit factors, io.cpp:887-891).
SOLUTION_MAGIC = b"EB200OUT"
SOLUTION_HEADER_BYTES = 8 + 4 + 4 + 3 * 8 + 8 + 6 * 8
FLUID_DATASETS = ("Density", "x-Momentum", "y-Momentum", "z-Momentum", "TotalEnergy")


def solution_name(iout):
    return "output-%07i.eb200" % iout


def dataset_names(nchem):
    return list(FLUID_DATASETS) + ["Chemical-%03d" % v for v in range(nchem)]


def _unit_scales(u):
    d, m, e = (float(getattr(u, k, 1.0)) for k in ("DensityUnits", "MomentumUnits", "EnergyUnits"))
    return [d, m, m, m, e]


def _barrier(u, group):
    if u.nprocs > 1:
        import torch.distributed as dist
        dist.barrier(group=group)


def output_solution(t, w, u, iout, directory=".", group=None):
    """``output_solution``: every rank stores its sub-box of every field into the one shared file
    (the reference does the same through an MPI-IO hyperslab, io.cpp:872-884).  Returns 0 / -1."""
    import os
    import struct
    import numpy as np
    path = os.path.join(directory, solution_name(iout))
    N = u.nx * u.ny * u.nz
    nds = 5 + u.nchem
    try:
        if u.myid == 0:
            with open(path, "wb") as fp:
                fp.write(SOLUTION_MAGIC + struct.pack("<iiqqqd6d", 1, u.nchem, u.nx, u.ny, u.nz, float(t),
                                                      u.zl, u.zr, u.yl, u.yr, u.xl, u.xr))
                fp.truncate(SOLUTION_HEADER_BYTES + 8 * N * nds)
        _barrier(u, group)
        data = np.memmap(path, dtype="<f8", mode="r+", offset=SOLUTION_HEADER_BYTES, shape=(nds, u.nz, u.ny, u.nx))
        box = (slice(u.ks, u.ks + u.nzl), slice(u.js, u.js + u.nyl), slice(u.is_, u.is_ + u.nxl))
        for f, scale in enumerate(_unit_scales(u)):
            data[(f,) + box] = (w.sub[f] * scale).cpu().numpy().reshape(u.nzl, u.nyl, u.nxl)
        if u.nchem > 0:
            chem = w.sub[5].cpu().numpy().reshape(u.nzl, u.nyl, u.nxl, u.nchem)
            for v in range(u.nchem):
                data[(5 + v,) + box] = chem[..., v]
        data.flush()
        del data
        _barrier(u, group)
    except OSError:
        return -1
    return 0


def read_solution(path):
    """The whole file as ``{"time", "nchem", "n": (nx, ny, nz), "domain", <dataset name>: array[nz, ny, nx]}``
    (what the reference's plotting scripts pull out of the HDF5 file)."""
    import struct
    import numpy as np
    with open(path, "rb") as fp:
        head = fp.read(SOLUTION_HEADER_BYTES)
    if len(head) != SOLUTION_HEADER_BYTES or head[:8] != SOLUTION_MAGIC:
        raise ValueError("%s is not a solution file" % path)
    version, nchem, nx, ny, nz, t, *dom = struct.unpack("<iiqqqd6d", head[8:])
    if version != 1:
        raise ValueError("%s: unknown solution file version %d" % (path, version))
    data = np.fromfile(path, dtype="<f8", offset=SOLUTION_HEADER_BYTES)
    if data.size != (5 + nchem) * nx * ny * nz:
        raise ValueError("%s is truncated" % path)
    data = data.reshape(5 + nchem, nz, ny, nx)
    out = {"time": t, "nchem": nchem, "n": (nx, ny, nz), "domain": dom}
    out.update({name: data[f] for f, name in enumerate(dataset_names(nchem))})
    return out


def read_restart(iout, w, u, directory="."):
    """``read_restart`` (io.cpp:940-1150): fill this rank's ``w`` from output-<iout>; the file must
    hold the same grid and species count.  Returns (0, t) or (-1, None)."""
    import os
    import torch
    try:
        sol = read_solution(os.path.join(directory, solution_name(iout)))
    except (OSError, ValueError) as e:
        print("read_restart: %s" % e)
        return -1, None
    if sol["n"] != (u.nx, u.ny, u.nz) or sol["nchem"] != u.nchem:
        print("read_restart: file holds %r with %d species, run asks for %r with %d"
              % (sol["n"], sol["nchem"], (u.nx, u.ny, u.nz), u.nchem))
        return -1, None
    box = (slice(u.ks, u.ks + u.nzl), slice(u.js, u.js + u.nyl), slice(u.is_, u.is_ + u.nxl))
    names = dataset_names(u.nchem)
    for f, scale in enumerate(_unit_scales(u)):
        w.sub[f].copy_(torch.from_numpy((sol[names[f]][box] / scale).reshape(-1).copy()))
    if u.nchem > 0:
        import numpy as np
        chem = np.stack([sol[names[5 + v]][box] for v in range(u.nchem)], axis=-1)
        w.sub[5].copy_(torch.from_numpy(chem.reshape(-1).copy()))
    return 0, sol["time"]
