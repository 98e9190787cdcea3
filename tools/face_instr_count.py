#!/usr/bin/env python
"""Tuning aid (no GPU needed): FP64-pipe instructions of one fluid_face and one tracer_face, counted on probe
kernels (tools/face_instr_probe.cu) cross-compiled for sm_100a -- the currency of an FP64-issue-bound kernel.
   python tools/face_instr_count.py [-DEB_PROJECT_SHARED] [-DEB_RCP_TWO_NEWTON -DEB_NO_FOLD_HALF] ..."""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
flags = sys.argv[1:]
with tempfile.TemporaryDirectory() as tmp:
    cubin = os.path.join(tmp, "probe.cubin")
    subprocess.check_call(["nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-I" + os.path.join(ROOT, "sundials-manyvector-demo_b200", "csrc")] + flags +
                          ["-cubin", "-o", cubin, os.path.join(ROOT, "tools", "face_instr_probe.cu")])
    sass = subprocess.check_output(["cuobjdump", "-sass", cubin], text=True)
cur, cnt = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    if "Function :" in line:
        cur = "fluid_face" if "ffk" in line else "tracer_face"
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(\S+)", line)
    if m and cur:
        op = m.group(1).rstrip(";")
        if op.startswith(("DFMA", "DMUL", "DADD", "DSETP")):
            cnt[cur][op.split(".")[0]] += 1
        elif op.startswith("MUFU"):
            cnt[cur][op] += 1
print("flags:", " ".join(flags) or "(default build)")
for f in ("tracer_face", "fluid_face"):
    c = cnt[f]
    print("%-12s FP64-pipe instructions %4d   %s" % (f, sum(v for k, v in c.items() if not k.startswith("MUFU")), dict(c)))
