#!/bin/bash
# Round-2 GPU check: every -m gpu test file in its own process (durations reported), smoke.  Logs: gpurun_out/r2/.
set -u
O=gpurun_out/r2
mkdir -p $O
rm -f $O/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
nproc > $O/nproc.txt; free -g >> $O/nproc.txt
for f in ${TEST_FILES:-kernels mlp render fullsize}; do
  timeout 1500 python -m pytest tests/test_gpu_$f.py -m gpu -q -rA -s --durations=8 --timeout 1200 ${PYTEST_EXTRA:-} > $O/test_$f.log 2>&1
  echo "test_gpu_$f exit $?" >> $O/summary.txt
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "smoke exit $?" >> $O/summary.txt
cat $O/summary.txt
grep -hE "^(FAILED|ERROR)|passed|failed|config 5 bf16|seed [0-9]+:|mean difference|config 1" $O/test_*.log | tail -40
tail -n 2 $O/smoke.log
