#!/bin/bash
# one bench line per BASELINE config shape (+ the harder headline variant), each with parity_check against oracle/_ref
tag=${1:-r02a}
for c in c1 c3 c5 c4r; do
  timeout 900 python bench.py --config $c --no-ext-microbench > gpurun_out/bench_${tag}_$c.json 2> gpurun_out/bench_${tag}_$c.err; echo "$c rc=$?"
  tail -c 300 gpurun_out/bench_${tag}_$c.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${tag}_$c.json")); b=d["breakdown_ms"]
    print("$c", "%.3gM pairs/s e2e %.3gM" % (d["value"]/1e6, d["e2e"]["value"]/1e6), b, d.get("parity_check"), d["counts"], d.get("cpu_baseline",{}).get("value"))
except Exception as e: print("$c: no line", e)
PY
done
