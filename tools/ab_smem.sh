#!/bin/bash
# developer A/B of the seeding kernels' knobs (one bench_breakdown run per setting)
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg bash tools/bench_breakdown.sh 2>&1 | tail -1
done
