#!/usr/bin/env python
"""Tuning aid: instruction mix of the innermost loops of a kernel's SASS that contain at least
--min-ldg128 LDG.E.128 (the species-pair loops of rhs_fused_kernel): FP64-pipe instructions against
everything else per iteration.   cuobjdump -sass <binary> | python tools/sass_loops.py <kernel-substring>"""
import collections
import re
import sys

want = sys.argv[1]
minld = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cur, funcs = None, {}
for l in sys.stdin:
    if "Function :" in l:
        cur = l.split("Function :")[1].strip()
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m and cur is not None:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if want not in name:
        continue
    addr2i = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\s+(?:`\()?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr2i:
                loops.append((addr2i[tgt], i))
    print(name, "instructions:", len(ins))
    for s, e in loops:
        body = [t for _, t in ins[s:e + 1]]
        if sum("LDG.E.128" in t for t in body) < minld or e - s > 1500:
            continue
        key = lambda t: ("IMAD.MOV" if "IMAD.MOV" in t else re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0])
        cnt = collections.Counter(key(t) for t in body)
        fp = sum(cnt[k] for k in ("DFMA", "DMUL", "DADD"))
        rest = {k: v for k, v in cnt.most_common() if k not in ("DFMA", "DMUL", "DADD")}
        print("  loop @%d len=%d fp64=%d other=%d %s" % (s, e - s + 1, fp, e - s + 1 - fp, dict(list(rest.items())[:12])))
