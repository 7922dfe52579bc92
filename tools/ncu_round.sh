#!/bin/bash
# ncu --set full captures of one production step of the hot path (headline config): the seeding / chaining kernels and the stage-4
# engines separately so each capture stays short.  Run only after the same bench command has exited 0 without ncu.  The reports are
# exported to CSV on the box (raw page; source page of the seeding kernel) and deleted: gpurun_out/ must stay under 64 MiB.
# usage: ncu_round.sh tag
tag=${1:-r02a}
B="python bench.py --steps 1 --warmup 1 --in-flight 1 --no-cpu-baseline --no-ext-microbench --no-sam-stage"
$B > gpurun_out/plain_$tag.json 2> gpurun_out/plain_$tag.err || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.err; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:'pack_reads_kernel|smem_sweep_kernel|smem_fast_p3_kernel|seedpos_dev_kernel|chain_kernel' --launch-skip 10 -c 5 \
  -f -o /tmp/prof_seed_$tag $B > gpurun_out/ncu_seed_$tag.log 2>&1; echo "seed capture rc=$?"
ncu -i /tmp/prof_seed_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_seed_${tag}_raw.csv 2>/dev/null
ncu -i /tmp/prof_seed_$tag.ncu-rep --page source --csv -k regex:smem_sweep_kernel 2>/dev/null | gzip > gpurun_out/ncu_sweep_${tag}_source.csv.gz
ncu -i /tmp/prof_seed_$tag.ncu-rep --page source --csv -k regex:chain_kernel 2>/dev/null | gzip > gpurun_out/ncu_chain_${tag}_source.csv.gz
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:'ext_cols_kernel|c2a_spec_fill|c2a_replay_kernel' --launch-skip 18 -c 9 \
  -f -o /tmp/prof_c2a_$tag $B > gpurun_out/ncu_c2a_$tag.log 2>&1; echo "c2a capture rc=$?"
ncu -i /tmp/prof_c2a_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_c2a_${tag}_raw.csv 2>/dev/null
rm -f /tmp/prof_seed_$tag.ncu-rep /tmp/prof_c2a_$tag.ncu-rep
ls -la gpurun_out/*_$tag*
du -sh gpurun_out
