#!/usr/bin/env python3
"""Per-SASS-region sample/instruction shares from an .ncu-rep source page.
usage: ncu_src.py rep kernel-regex launch-skip [block]"""
import csv, subprocess, sys
rep, kre, skip = sys.argv[1], sys.argv[2], sys.argv[3]
blk = int(sys.argv[4]) if len(sys.argv) > 4 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) and r[0].startswith("0x")]
S = lambda r: int(r[ci["# Samples"]]); I = lambda r: int(r[ci["Instructions Executed"]])
tot, toti = sum(map(S, data)), sum(map(I, data))
print(len(data), "sass lines; samples", tot, "warp-inst", toti)
for k in range(0, len(data), blk):
    b = data[k:k + blk]
    s, n = sum(map(S, b)), sum(map(I, b))
    if s * 200 < tot and n * 200 < toti: continue
    print("%5d-%5d samples %5.1f%% inst %5.1f%%  %s" % (k, k + len(b), 100 * s / tot, 100 * n / toti, b[0][1].strip()[:60]))
if len(sys.argv) > 5:
    lo, hi = map(int, sys.argv[5].split("-"))
    for r in data[lo:hi]:
        print("%6d %9d  %s" % (S(r), I(r), r[1].strip()))
