#!/usr/bin/env python3
"""Where are lanes wasted? Per SASS region: warp-instructions executed, average active threads, share of all issue slots.
usage: ncu_div.py rep kernel-regex launch-skip [block]"""
import csv, subprocess, sys
rep, kre, skip = sys.argv[1], sys.argv[2], sys.argv[3]
blk = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) and r[0].startswith("0x")]
I = lambda r: int(r[ci["Instructions Executed"]]); T = lambda r: int(r[ci["Thread Instructions Executed"]]); S = lambda r: int(r[ci["# Samples"]])
toti, tott, tots = sum(map(I, data)), sum(map(T, data)), sum(map(S, data))
print("warp-inst %d, avg threads %.2f" % (toti, tott / toti))
for k in range(0, len(data), blk):
    b = data[k:k + blk]
    n, t, s = sum(map(I, b)), sum(map(T, b)), sum(map(S, b))
    if n * 100 < toti: continue
    print("%5d-%5d inst %5.1f%% samples %5.1f%% avg threads %5.1f   %s" % (k, k + len(b), 100 * n / toti, 100 * s / max(tots, 1), t / max(n, 1), b[0][1].strip()[:50]))
