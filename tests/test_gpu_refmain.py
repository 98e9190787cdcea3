"""The reference's UNMODIFIED main program (euler3D_main.cpp + io.cpp + gopt.cpp + problem file) on the
B200 with DEVICE-RESIDENT vectors: oracle/_ref/refmain_gpu_<problem> links it with our fEuler / stability
(host/feuler_dropin.cpp -> libeulerb200.so) and with the managed-memory flavour of the N_Vector stand-in
(oracle/shim/shim_core.h, -DSHIM_MANAGED_VECTORS: the N_VNew_Serial vectors of euler3D_main.cpp:150-171
live in CUDA managed memory, as the reference's own device builds do for the chemistry vector), the
integrator's vector operations running on the device (shim_arkstep.cpp -> eulerb200_vec_lincomb /
_wrms_accum).  In fixed-step runs it must print what the reference main prints with the reference's own
fEuler on the CPU (refmain_<problem>), line for line -- VERDICT r1 "device-resident vectors under a
reference driver" (SURVEY.md 8(b) data access, 8(f-1))."""
import os

import pytest

from test_reference_main_cpu import CASES, INPUTS, report

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


@pytest.mark.parametrize("problem", ["sod_x", "rayleigh_taylor", "hurricane_zx_color"])
def test_reference_main_with_device_resident_vectors_prints_what_the_cpu_reference_prints(tmp_path, problem):
    gpu, cpu = os.path.join(REF, "refmain_gpu_" + problem), os.path.join(REF, "refmain_" + problem)
    if not (os.path.exists(gpu) and os.path.exists(cpu)):
        pytest.skip("needs the prebuilt oracle/_ref/refmain_* programs (built where /root/reference exists)")
    infile, args = CASES[problem]
    common = ["-f", os.path.join(INPUTS, infile)] + args
    ref = report([cpu] + common, tmp_path)
    got = report([gpu] + common, tmp_path)
    assert got == ref, "\n".join(got) + "\n--- vs the reference main with its own fEuler on the CPU ---\n" + "\n".join(ref)


def test_profile_slots_of_the_reference_are_fed_from_cuda_events(tmp_path):
    """EULERB200_PROFILE=1: the drop-in adds device times to the reference's PR_PACKDATA / PR_MPI /
    PR_FACEFLUX slots (profiler.hpp), so the end-of-run profile table of the reference main has them."""
    import subprocess
    gpu = os.path.join(REF, "refmain_gpu_rayleigh_taylor")
    if not os.path.exists(gpu):
        pytest.skip("needs the prebuilt oracle/_ref/refmain_gpu_* programs")
    infile, args = CASES["rayleigh_taylor"]
    res = subprocess.run([gpu, "-f", os.path.join(INPUTS, infile)] + args, capture_output=True, text=True, timeout=600,
                         cwd=str(tmp_path), env=dict(os.environ, EULERB200_PROFILE="1"))
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    import re
    tail = res.stdout[res.stdout.find("Profiling Results"):]
    times = {m.group(1): float(m.group(2)) for m in re.finditer(r"Total (\S+) time = \s*(\S+)", tail)}
    assert times.get("flux", 0.0) > 0.0 and times.get("pack", 0.0) > 0.0, tail       # fed from CUDA events
    assert times["flux"] <= times["RHS"] * 1.05 + 1e-3, tail                          # device time within the wall time
