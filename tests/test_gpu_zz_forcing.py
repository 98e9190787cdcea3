"""Late additions of round 1 (written after its last GPU session, hence sorted last): (1) the
external_forces hook in full generality on the device: forcing that
depends on position and time is run into wdot before the evaluation, as the reference's fEuler
does (utilities.cpp:28,65), and the kernel computes wdot = wdot - div F(w)
(eulerb200_set_forcing_in_wdot).  fEuler is affine in G and the reference rounds G - div once, so
the expected result is exactly G + (oracle with zero forcing); (2) the AG instantiation of the
fused kernel for boundary-heavy launches (EULERB200_KERNEL=1).  Tolerance 1e-12 normwise."""
import os
import subprocess

import numpy as np
import pytest

from conftest import normwise_errors, rounding_floor
from helpers import make_udata, oracle_feuler

pytestmark = pytest.mark.gpu
P, N, D, R = 0, 1, 2, 3
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("host", [False, True])
@pytest.mark.parametrize("n,nchem,bcs", [((40, 24, 20), 2, [P, P, R, R, N, N]), ((33, 12, 40), 0, [R] * 6)])
def test_hook_assigned_forcing_device_and_host_vectors(pkg, oracle_mod, port, n, nchem, bcs, host):
    import torch
    u = make_udata(pkg, n, nchem, bcs, forcing=[7.0, 7.0, 7.0, 7.0, 7.0])     # must be ignored
    parts = oracle_mod.random_state(n, nchem, seed=17)
    rng = np.random.default_rng(3)
    G = [rng.normal(size=p.size) for p in parts if p is not None]
    dev = "cpu" if host else "cuda"
    calls = []

    def hook(t, Gvec, udata):
        calls.append(t)
        for sub, g in zip(Gvec.sub, G):
            assert float(sub.abs().max()) == 0.0          # zeroed first, utilities.cpp:28
            sub.copy_(torch.from_numpy(g + 0.25 * t).to(sub.device))
        return 0

    w = pkg.ManyVector([torch.from_numpy(p).to(dev) for p in parts if p is not None])
    wdot = pkg.ManyVector.new(u, device=dev)
    ret = pkg.fEuler(0.5, w, wdot, u, external_forces=hook)
    torch.cuda.synchronize()
    assert ret == 0 and calls == [0.5], u.last_error()
    u.forcing = [0.0] * 5
    ret_ref, ref, _ = oracle_feuler(port, u, parts)
    want = [g + 0.125 + r for g, r in zip(G, [r for r in ref if r is not None])]
    got = [s.cpu().numpy() for s in wdot.sub]
    floor = rounding_floor(parts, u.gamma, (u.dx, u.dy, u.dz))
    assert max(normwise_errors(got, want, floor)) <= 1e-12
    # and back to the constants of the config when no hook is given
    ret = pkg.fEuler(0.5, w, wdot, u)
    torch.cuda.synchronize()
    u.forcing = [7.0] * 5
    ret_ref, ref7, _ = oracle_feuler(port, u, parts)
    got = [s.cpu().numpy() for s in wdot.sub]
    assert ret == 0 and max(normwise_errors(got, [r for r in ref7 if r is not None], floor)) <= 1e-12
    u.FreeData()


@pytest.mark.parametrize("nvar", [5, 7])
def test_dropin_with_a_varying_hook_against_the_reference(pkg, nvar):
    """oracle/_ref/dropin_check_nvar<N> with a position- and time-dependent external_forces: the
    reference fEuler and the drop-in both run the same hook; host N_Vectors, so the hook's G travels
    to the device with the state (eulerb200_rhs_host)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "dropin_check_nvar%d" % nvar)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin_check_nvar%d not built (needs the reference tree)" % nvar)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=dict(os.environ, EB_DROPIN_VARYING="1"))
    print(res.stdout, res.stderr)
    assert res.returncode == 0 and "DROPIN_CHECK PASS" in res.stdout, res.stdout + res.stderr
    assert res.stdout.count("run before every evaluation") == 5


@pytest.mark.parametrize("n,nchem,bcs", [((16, 12, 10), 2, [R] * 6), ((40, 9, 11), 0, [P, P, R, R, N, N]),
                                         ((64, 20, 24), 10, [R] * 6), ((3, 40, 36), 6, [N] * 6),
                                         ((200, 3, 3), 0, [N] * 6), ((70, 34, 40), 3, [P] * 6)])
def test_boundary_heavy_instantiation_matches_oracle(pkg, oracle_mod, port, monkeypatch, n, nchem, bcs):
    """EULERB200_KERNEL=1 (opt-in until timed on a B200): launches in which a quarter or more of the
    tiles touch a boundary run the AG instantiation, where boundary tiles read the per-cell 1/rho, p, c,
    sqrt(rho) arrays for owned points and for ghost points that only differ in the sign of a momentum
    instead of re-deriving them for all six stencil points.  Same tolerance as the default kernel."""
    from helpers import gpu_feuler
    monkeypatch.setenv("EULERB200_KERNEL", "1")
    u = make_udata(pkg, n, nchem, bcs, forcing=[0, 0, -0.1, 0, 0])
    parts = oracle_mod.random_state(n, nchem, seed=sum(n) + nchem)
    ret, got = gpu_feuler(pkg, u, parts)
    ret_ref, ref, _ = oracle_feuler(port, u, parts)
    assert ret == 0 and ret_ref == 0, u.last_error()
    assert max(normwise_errors(got, ref, rounding_floor(parts, u.gamma, (u.dx, u.dy, u.dz)))) <= 1e-12
    u.FreeData()
