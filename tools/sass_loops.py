#!/usr/bin/env python3
"""List the loops (backward branches) of a SASS dump with their opcode histograms, or print one loop body.
usage: cuobjdump -sass -fun NAME file.o > x.sass; python tools/sass_loops.py x.sass [start_hex end_hex]"""
import re, sys
from collections import Counter
ins = []
for l in open(sys.argv[1]):
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
def op(t):
    return re.sub(r'^@!?U?P\d+\s+', '', t).split()[0].split('.')[0]
if len(sys.argv) >= 4:
    lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
    for a, t in ins:
        if lo <= a <= hi:
            print("%05x  %s" % (a, t))
else:
    for a, t in ins:
        m = re.search(r'BRA.*0x([0-9a-f]+)', t)
        if m and int(m.group(1), 16) < a:
            tgt = int(m.group(1), 16)
            body = [x for x in ins if tgt <= x[0] <= a]
            if len(body) > 60:
                print(hex(tgt), hex(a), len(body), Counter(op(x[1]) for x in body).most_common(14))
