#!/bin/bash
# one GPU-box visit: parity tests, both bench arms on the headline config, launch list (round-end shape of the driver's own run)
tag=${1:-r02a}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_reference_arm.json 2> gpurun_out/bench_${tag}_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ext-microbench --no-sam-stage > gpurun_out/ncu_$tag.log 2>&1; echo "ncu rc=$?"
cut -c1-1500 gpurun_out/bench_$tag.json
