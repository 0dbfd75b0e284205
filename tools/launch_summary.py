#!/usr/bin/env python3
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel: total us, launches, share."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i
        break
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    agg[r[ki][:100]][0] += 1
    agg[r[ki][:100]][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f"{v[1]:12.1f} us {v[0]:5d} launches {100 * v[1] / tot:5.1f} %  {k}")
print(f"{tot:12.1f} us total")
