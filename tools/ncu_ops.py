#!/usr/bin/env python3
"""Executed-instruction histogram by opcode from the source page of an .ncu-rep, per unit of work.
usage: python tools/ncu_ops.py file.ncu-rep [units]   (units: e.g. vertex-warps = vertex-instances / 64)"""
import csv, io, re, subprocess, sys
from collections import Counter
rep = sys.argv[1]; units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = [i for i, r in enumerate(rows) if "Address" in r and "# Samples" in r][0]
hdr = rows[h]; data = [r for r in rows[h + 1:] if len(r) == len(hdr)]
isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
c = Counter()
for r in data:
    op = re.sub(r'^@!?U?P\d+\s+', '', r[isrc].strip()).split()[0].split('.')[0]
    c[op] += int(r[iex] or 0)
tot = sum(c.values())
print("total %d, %.1f per unit" % (tot, tot / units))
print("  ".join("%s %.1f" % (op, n / units) for op, n in c.most_common(36)))
