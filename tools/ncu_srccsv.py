#!/usr/bin/env python3
"""Hot SASS lines of an exported `ncu --page source --csv` file (gz ok): top lines by stall samples, and sample / instruction shares
per block of consecutive lines.  usage: ncu_srccsv.py file.csv[.gz] [top] [block]"""
import csv, gzip, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
blk = int(sys.argv[3]) if len(sys.argv) > 3 else 50
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = list(csv.reader(f))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) and r[0].startswith("0x")]
S = lambda r: int(r[ci["# Samples"]]); I = lambda r: int(r[ci["Instructions Executed"]]); T = lambda r: int(r[ci["Thread Instructions Executed"]])
tot, toti, tott = sum(map(S, data)), sum(map(I, data)), sum(map(T, data))
print(len(data), "sass lines; samples", tot, "warp-inst", toti, "lanes/inst %.2f" % (tott / max(toti, 1)))
for k in range(0, len(data), blk):
    b = data[k:k + blk]
    s, n, t = sum(map(S, b)), sum(map(I, b)), sum(map(T, b))
    if s * 100 < tot and n * 100 < toti: continue
    print("%5d-%5d samples %5.1f%% inst %5.1f%% lanes %4.1f  %s" % (k, k + len(b), 100 * s / tot, 100 * n / toti, t / max(n, 1), b[0][1].strip()[:50]))
print("-- top lines by samples")
for i in sorted(range(len(data)), key=lambda i: -S(data[i]))[:top]:
    r = data[i]
    print("%5d %5.2f%% inst %9d lanes %4.1f  %s" % (i, 100 * S(r) / tot, I(r), T(r) / max(I(r), 1), r[1].strip()[:90]))
