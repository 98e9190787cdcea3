"""The drop-in fEuler / stability (host/feuler_dropin.cpp) on the CPU tier: linked with the
reference's own EulerData, SetupDecomp and N_Vectors and with the UNMODIFIED reference fEuler
(renamed), but against the CPU emulation of the kernel source instead of libeulerb200.so
(tests/emu/emu_abi.cpp, test infrastructure).  Checks the host logic of the drop-in -- sub-vector
plumbing, the probe of the external_forces hook, the per-call hook for forcing that is not a
per-field constant -- on the same vectors as the reference, tolerance 1e-12 normwise as
everywhere.  The device library itself is checked by tests/test_gpu_dropin.py."""
import os
import subprocess

import pytest


@pytest.fixture(scope="module")
def exe(oracle_mod):
    path = oracle_mod.build_dropin_emu()
    if path is None:
        pytest.skip("needs the reference tree (or a prebuilt oracle/_ref/dropin_check_emu_nvar7)")
    return path


def test_dropin_against_reference_feuler_constant_hooks(exe):
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DROPIN_CHECK PASS" in out.stdout, out.stdout + out.stderr
    assert out.stdout.count(" ok") == 5 and "Gmy=-0.1" in out.stdout      # incl. the Rayleigh-Taylor forcing


def test_dropin_runs_a_position_and_time_dependent_hook_before_every_evaluation(exe):
    env = dict(os.environ, EB_DROPIN_VARYING="1")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and "DROPIN_CHECK PASS" in out.stdout, out.stdout + out.stderr
    assert out.stdout.count("run before every evaluation") == 5
