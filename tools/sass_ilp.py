#!/usr/bin/env python
"""Tuning aid: histogram of producer->consumer distances (counted in FP64-pipe instructions) for
the DFMA/DMUL/DADD stream of a kernel's SASS.  A DFMA result is usable ~4 FP64 issue slots (8
cycles) later on B200 (tools/fp64_probe.cu), so with W warps per scheduler a distance below ~4/W
stalls the pipe.   cuobjdump -sass lib.so | python tools/sass_ilp.py <kernel-name-substring>"""
import collections
import re
import sys

want = sys.argv[1] if len(sys.argv) > 1 else "rhs_fused"
ins, on = [], False
for l in sys.stdin:
    if "Function :" in l:
        on = want in l
        continue
    if on:
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append(m.group(2).strip())


def regs(tok):
    return [int(m.group(1)) for m in re.finditer(r"\bR(\d+)\b", tok)]


hist = collections.Counter()
lastw, nfp = {}, 0
for t in ins:
    t2 = re.sub(r"^@!?U?P\d+\s+", "", t)
    op = t2.split()[0]
    parts = [p.strip() for p in t2[len(op):].split(",")]
    if op.split(".")[0] in ("DFMA", "DMUL", "DADD"):
        d = 99
        for p in parts[1:]:
            for r in regs(p):
                for rr in (r, r + 1):
                    if rr in lastw:
                        d = min(d, nfp - lastw[rr])
        hist[min(d, 8)] += 1
        nfp += 1
        for r in regs(parts[0]):
            lastw[r] = lastw[r + 1] = nfp
    elif parts and parts[0].startswith("R"):
        for r in regs(parts[0]):
            lastw.pop(r, None)
            lastw.pop(r + 1, None)
tot = sum(hist.values())
print("%d instructions, %d on the FP64 pipe" % (len(ins), tot))
for k in sorted(hist):
    print("  distance %s%d: %5.1f%%" % (">=" if k == 8 else "  ", k, 100.0 * hist[k] / tot))
