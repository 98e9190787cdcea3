#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` output: registers / spills per rhs_fused_kernel instantiation.
   nvcc ... -Xptxas -v ... 2>&1 | python tools/ptxas_summary.py"""
import re
import sys

txt = sys.stdin.read()
pat = re.compile(r"Compiling entry function '(\S+)' for.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
                 r"(\d+) bytes spill loads\n.*?Used (\d+) registers", re.S)
for m in pat.finditer(txt):
    name = m.group(1)
    k = re.search(r"rhs_fused_kernelILi(\d+)ELi(\d+)ELb(\d)ELb(\d)ELi(\d)E", name)
    if k:
        name = "rhs_fused<T=%s,B=%s,GW=%s,AG=%s,PART=%s>" % k.groups()
    print("%-46s regs=%3s stack=%4s spill_st=%5s spill_ld=%5s" % (name[:46], m.group(5), m.group(2), m.group(3), m.group(4)))
