#!/usr/bin/env python3
"""Stall-sample distribution over SASS ranges of an `ncu --page source --csv` export (bucketed, with marker opcodes)."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 60
h = rows[1]
si, ni = h.index("Source"), h.index("# Samples")
stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data = [r for r in rows[2:] if len(r) > ni and r[ni].isdigit()]
srcs = [r[si] for r in data]
# the listing may be duplicated (two copies of the function)
half = len(data) // 2
if srcs[:half] == srcs[half:2 * half]:
    data = data[:half]
tot = sum(int(r[ni]) for r in data)
print("instructions", len(data), "samples", tot)
MARK = ("LDGSTS", "LDTM", "UTCHMMA", "STG", "LDG", "SYNCS", "BAR", "UTCBAR", "RED", "ATOMG", "EXIT", "FENCE", "LDS", "STS", "MUFU", "SHFL")
for b in range(0, len(data), B):
    chunk = data[b:b + B]
    s = sum(int(r[ni]) for r in chunk)
    ops = set()
    for r in chunk:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[si])
        if m and m.group(2).split(".")[0] in MARK:
            ops.add(m.group(2).split(".")[0])
    d = {h[i][6:]: sum(int(r[i]) for r in chunk) for i in stall}
    top = {k: v for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:3]}
    if s > 0.012 * tot:
        print(f"instr {b:5d}-{b + B:5d}: {100 * s / tot:5.1f}%  {sorted(ops)} {top}")
