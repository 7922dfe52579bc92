#!/usr/bin/env python3
"""Per-source-line shares of warp instructions / thread instructions / stall samples from an .ncu-rep built with -lineinfo and
captured with --import-source on.  usage: ncu_lines.py rep [kernel-regex] [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; kre = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if kre: cmd += ["--kernel-name", "regex:" + kre, "--launch-count", "1"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
path = ""; agg = {}; hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": path = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if not hdr or len(r) < 12 or not r[0]: continue          # cuda lines have a line number; sass rows below them are skipped
    key = (path, int(r[0]), r[1].strip()[:90])
    a = agg.setdefault(key, [0, 0, 0])
    num = lambda v: int(v) if v.lstrip("-").isdigit() else 0
    a[0] += num(r[hdr["Instructions Executed"]]); a[1] += num(r[hdr["Thread Instructions Executed"]]); a[2] += num(r[hdr["# Samples"]])
ti = sum(a[0] for a in agg.values()); tt = sum(a[1] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print("warp-inst %.3g thread-inst %.3g (%.1f lanes) samples %d" % (ti, tt, tt / max(ti, 1), ts))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f lanes %5.1f%% smp  %s:%d  %s" % (100 * a[0] / ti, a[1] / max(a[0], 1), 100 * a[2] / max(ts, 1), k[0], k[1], k[2]))
