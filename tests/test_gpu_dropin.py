"""Drop-in boundary, end to end: an executable that links the reference's own EulerData /
SetupDecomp / N_Vector composition (compiled from /root/reference, unmodified), the
reference fEuler under a different name, and OUR fEuler/stability with the reference
signatures (host/feuler_dropin.cpp over libeulerb200.so), and calls both on the same
vectors.  Built by `make -C oracle dropin` where the reference tree exists; the binary
travels to the GPU box inside oracle/_ref."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nvar", [5, 7])
def test_reference_eulerdata_drives_our_feuler(pkg, nvar):
    exe = os.path.join(ROOT, "oracle", "_ref", "dropin_check_nvar%d" % nvar)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin_check_nvar%d not built (needs the reference tree)" % nvar)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(res.stdout, res.stderr)
    assert res.returncode == 0 and "DROPIN_CHECK PASS" in res.stdout, res.stdout + res.stderr
