#!/bin/bash
# What the driver runs at round end, in the same shape: the whole GPU suite in ONE process, smoke, both bench arms.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout 900 > gpurun_out/driver_tests.log 2>&1
echo "pytest -m gpu exit $?" | tee gpurun_out/driver_summary.txt
tail -n 5 gpurun_out/driver_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/driver_summary.txt
tail -n 2 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
echo "bench reference exit $?" | tee -a gpurun_out/driver_summary.txt
tail -c 1200 gpurun_out/bench_ref.log
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" | tee -a gpurun_out/driver_summary.txt
tail -c 3500 gpurun_out/bench.log
