#!/usr/bin/env python3
"""Stage times of one raster flush, from an ncu launch list (gpu__time_duration.sum per launch, cold-cache and serialised:
read the SHARES).  usage: ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file x.csv python ...;
python tools/raster_stages.py x.csv [first_kernel_regex]"""
import csv, re, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = OrderedDict()
tot = 0.0
for r in rows[1:]:
    if r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ik]).replace("ogrecu::", "").replace("(anonymous namespace)::", "")
    us = float(r[iv].replace(",", "")) / (1000.0 if r[hdr.index("Metric Unit")] in ("nsecond", "ns") else 1.0)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us; tot += us
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-60s x%-4d %10.1f us  %5.1f%%" % (k[:60], n, us, 100 * us / tot))
print("total %.1f us" % tot)
