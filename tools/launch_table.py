#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total ms, share."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value"); mu = hdr.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
n = 0
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    n += 1
    if n <= skip: continue
    v = float(r[mv].replace(",", "")); u = r[mu]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    a = agg.setdefault(r[kn], [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.3f | %.1f%% |" % (k[:90], a[0], a[1], 100 * a[1] / tot))
print("total %.3f ms over %d launches" % (tot, sum(a[0] for a in agg.values())))
