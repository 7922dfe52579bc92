#!/bin/bash
# developer A/B of the extension engines on dumped mem_chain2aln jobs (tests/perf_ext.py) + the extension parity tests
python -m pytest tests/test_ext_gpu.py tests/test_c2a_gpu.py -x -q 2>&1 | tail -5
for e in pair cols; do
  echo "== engine $e"
  BWAFLOW_EXT_ENGINE=$e python tests/perf_ext.py ${1:-100000} ${2:-4} 2>&1 | tail -4
done
