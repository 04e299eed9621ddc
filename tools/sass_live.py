#!/usr/bin/env python3
"""Approximate register liveness over one function of an `nvdisasm --print-line-info` dump: live general registers at every
instruction of an address range, and for the point of highest pressure which source line defined each live register.
Development aid for register-bound kernels (skin_lane.cu); operand widths are inferred from opcode suffixes.
usage: nvdisasm --print-line-info x.cubin > x.txt; python tools/sass_live.py x.txt FUNCTION_SUBSTRING [lo_hex hi_hex] [--at hex]"""
import re
import sys
from collections import defaultdict

path, fun = sys.argv[1], sys.argv[2]
args = [a for a in sys.argv[3:] if not a.startswith("--")]
at = None
if "--at" in sys.argv:
    at = int(sys.argv[sys.argv.index("--at") + 1], 16)
    args = [a for a in args if a != sys.argv[sys.argv.index("--at") + 1]]
lo = int(args[0], 16) if len(args) >= 2 else 0
hi = int(args[1], 16) if len(args) >= 2 else 1 << 30

ins = []            # (addr, text, line)
labels = {}
infun = False
line_no = 0
for l in open(path):
    if l.startswith("//---") and ".text." in l:
        infun = fun in l
        continue
    if not infun:
        continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
    if m:
        line_no = "%s:%s" % (m.group(1)[:12], m.group(2))
        continue
    m = re.match(r'\s*(\.L_x_\d+):', l)
    if m:
        labels[m.group(1)] = len(ins)
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip(), line_no))

NODEST = ("ST", "STS", "STG", "STL", "BRA", "BSSY", "BSYNC", "EXIT", "ISETP", "FSETP", "PLOP3", "WARPSYNC", "BAR", "MEMBAR", "FENCE",
          "UTMASTG", "RED", "NOP", "CALL", "RET", "DEPBAR", "UTMACMDFLUSH", "ERRBAR", "R2UR", "SYNCS", "UBLKCP", "ATOMS", "NANOSLEEP",
          "BREAK", "UTMAREDG", "CCTL", "LDGDEPBAR", "UTMALDG", "UTMAPF", "BMOV", "YIELD", "ENDCOLLECTIVE", "UISETP", "ULOP3", "UIADD3", "UMOV",
          "ULEA", "UIMAD", "USHF", "S2UR", "USEL", "UPLOP3", "LDCU", "VOTEU", "UFLO", "UPOPC", "REDUX")


def regs(tok, width):
    m = re.search(r'\bR(\d+)', tok)
    if not m:
        return []
    r = int(m.group(1))
    w = width
    if ".64" in tok and width < 2:
        w = 2
    return list(range(r, r + w))


def parse(text):
    t = re.sub(r'^@!?U?P\d+\s+', '', text)
    pred = t is not text and t != text
    parts = t.split(None, 1)
    opf = parts[0]
    op = opf.split('.')[0]
    ops = [o.strip() for o in re.split(r',\s*(?![^\[]*\])', parts[1])] if len(parts) > 1 else []
    vec = 4 if ".128" in opf else 2 if (".64" in opf or op in ("FFMA2", "FADD2", "FMUL2") or ".WIDE" in opf) else 1
    defs, uses = [], []
    k = 0
    if op not in NODEST and ops:
        # predicate destinations come first (LOP3.LUT P1, RZ, ...; IADD3 R, P0, P1 has them after)
        while k < len(ops) and re.match(r'^U?P\d+$|^PT$', ops[k]):
            k += 1
        if k < len(ops) and not ops[k].startswith("["):
            dw = vec if op not in ("IMAD", "IADD3") or ".WIDE" in opf else 1
            if op in ("LDS", "LDG", "LDL", "LD", "LDC", "FFMA2", "FADD2", "FMUL2") or ".WIDE" in opf:
                dw = vec
            elif op == "MUFU" or op in ("MOV", "FSEL", "SEL", "LOP3", "SHF", "LEA", "FFMA", "FADD", "FMUL", "S2R", "VOTE", "SHFL", "I2F", "F2I", "VIMNMX", "HFMA2"):
                dw = 1
            defs = regs(ops[k], dw)
            k += 1
    for o in ops[k:]:
        if o.startswith("["):
            for m in re.finditer(r'R(\d+)(\.64)?', o):
                r = int(m.group(1))
                uses += [r, r + 1] if m.group(2) else [r]
            continue
        if not re.search(r'\bR\d+', o):
            continue
        if op in ("FFMA2", "FADD2", "FMUL2"):
            w = 1 if re.search(r'\.F32(?!x2)', o) else 2
        elif op in ("STS", "STG", "STL", "ST"):
            w = vec
        elif ".WIDE" in opf and o is ops[-1]:
            w = 2
        else:
            w = 1
        uses += regs(o, w)
    tgt = None
    m = re.search(r'`\((\.L_x_\d+)\)', text)
    if m and op in ("BRA",):
        tgt = m.group(1)
    return op, pred, defs, uses, tgt


P = [parse(t) for _, t, _ in ins]
n = len(ins)
succ = [[] for _ in range(n)]
for i, (op, pred, defs, uses, tgt) in enumerate(P):
    if op == "EXIT" and not pred:
        continue
    if op == "BRA" and tgt in labels:
        succ[i].append(labels[tgt])
        if pred or "P" in ins[i][1].split("BRA", 1)[1].split("`")[0]:
            if i + 1 < n:
                succ[i].append(i + 1)
        continue
    if i + 1 < n:
        succ[i].append(i + 1)
live_in = [frozenset()] * n
changed = True
while changed:
    changed = False
    for i in range(n - 1, -1, -1):
        op, pred, defs, uses, tgt = P[i]
        out = set()
        for s in succ[i]:
            out |= live_in[s]
        new = (out - set(defs) if not pred else out) | set(uses)
        new.discard(255)
        if new != live_in[i]:
            live_in[i] = frozenset(new)
            changed = True

best = None
for i in range(n):
    a = ins[i][0]
    if lo <= a <= hi:
        if at is not None:
            if a == at:
                best = i
        elif best is None or len(live_in[i]) > len(live_in[best]):
            best = i
if len(args) >= 2 and at is None and "--trace" in sys.argv:
    for i in range(n):
        if lo <= ins[i][0] <= hi:
            print("%05x %3d %-17s %s" % (ins[i][0], len(live_in[i]), ins[i][2], ins[i][1][:70]))
i = best
print("point %05x: %d live   (%s)" % (ins[i][0], len(live_in[i]), ins[i][1]))
# nearest earlier textual definition of each live register
by_line = defaultdict(list)
for r in sorted(live_in[i]):
    j = i - 1
    while j >= 0 and r not in P[j][2]:
        j -= 1
    if "--loop" in sys.argv:            # definitions inside the address range win (loop-carried values)
        inl = [k for k in range(n) if lo <= ins[k][0] <= hi and r in P[k][2]]
        j = inl[-1] if inl else (j if j >= 0 else -1)
        if not inl:
            by_line[("(invariant) " + (str(ins[j][2]) if j >= 0 else "?"), P[j][0] if j >= 0 else "?")].append(r)
            continue
    by_line[(ins[j][2], P[j][0]) if j >= 0 else ("?", "?")].append(r)
for (ln, op), rs in sorted(by_line.items()):
    print("  %-18s %-8s R%s" % (ln, op, ",".join(map(str, rs))))
