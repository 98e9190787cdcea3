#!/usr/bin/env python
"""Tuning aid: digest of an `ncu --page source --csv` export (one or more kernels): stall mix,
executed instructions per opcode, the hottest global loads.
   python tools/ncu_source.py export.csv [kernel-index] [cells]"""
import io
import sys

import pandas as pd

path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cells = float(sys.argv[3]) if len(sys.argv) > 3 else 134217728.0
lines = open(path).read().splitlines()
idx = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')] + [len(lines)]
print(lines[idx[which]][:140])
df = pd.read_csv(io.StringIO("\n".join(lines[idx[which] + 1:idx[which + 1]])))
df["src"] = df["Source"].str.strip().str.replace(r"^@!?U?P\d+\s+", "", regex=True)
df["op"] = df["src"].str.split().str[0]
tot = df["# Samples"].sum()
print("samples", tot, " SASS instructions", len(df), " executed per cell %.0f" % (df["Instructions Executed"].sum() * 32 / cells))
stalls = [c for c in df.columns if c.startswith("stall_") and "Not Issued" not in c]
print((df[stalls].sum() / tot * 100).sort_values(ascending=False).round(2).head(12).to_string())
g = df.groupby("op").agg(n=("Source", "count"), ex=("Instructions Executed", "sum"), s=("# Samples", "sum"),
                         tags=("L1 Tag Requests Global", "sum"), lsb=("stall_long_sb", "sum")).sort_values("s", ascending=False)
g["per_cell"] = g["ex"] * 32 / cells
g["s%"] = g["s"] / tot * 100
g["tags/inst"] = g["tags"] / g["ex"]
print(g.head(28).round(2).to_string())
