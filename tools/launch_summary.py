#!/usr/bin/env python
"""per-kernel mean duration (us) from an `ncu --metrics gpu__time_duration.sum --csv` launch list;
optional 2nd arg: only the last N launches of each kernel (skip warm-up launches)"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    agg.setdefault(r[kn], []).append(v)
last = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tot = 0
for k, v in agg.items():
    if last:
        v = v[-last:]
    m = sum(v) / len(v) / 1000
    print(f"{k[:86]:88s} n={len(v):3d} mean={m:9.1f} us")
