#!/bin/bash
# ncu --set full captures of one production step of the hot path (headline config), seeding/chaining kernels and stage-4 kernels
# separately so each capture stays short.  Run only after the same bench command has exited 0 without ncu.  usage: ncu_round.sh tag
tag=${1:-r02a}
B="python bench.py --steps 1 --warmup 1 --in-flight 1 --no-cpu-baseline --no-ext-microbench"
$B > gpurun_out/plain_$tag.json 2> gpurun_out/plain_$tag.err || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.err; exit 1; }
timeout 1500 ncu --set full --clock-control none --import-source on \
  -k regex:'pack_reads_kernel|smem_sweep_kernel|smem_fast_p3_kernel|seedpos_dev_kernel|chain_kernel' --launch-skip 10 -c 10 \
  -f -o gpurun_out/prof_seed_$tag $B > gpurun_out/ncu_seed_$tag.log 2>&1; echo "seed capture rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on \
  -k regex:'ext_pair_kernel|ext_kernel|c2a_' --launch-skip 70 -c 36 \
  -f -o gpurun_out/prof_c2a_$tag $B > gpurun_out/ncu_c2a_$tag.log 2>&1; echo "c2a capture rc=$?"
ls -la gpurun_out/*_$tag*
