#!/bin/bash
# developer A/B of the seeding kernels' knobs (one bench_breakdown run per setting)
for cfg in "BWAFLOW_SMEM_LEGACY=1" "BWAFLOW_TEXT_MODE=0" "BWAFLOW_KMER_K=12" "BWAFLOW_KMER_K=10" "BWAFLOW_KMER_K=0" "BWAFLOW_SMEM_XMIN=8" "BWAFLOW_SMEM_XMIN=16" "$@"; do
  echo "== $cfg"
  env $cfg bash tools/bench_breakdown.sh 2>&1 | tail -1
done
