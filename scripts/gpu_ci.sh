#!/bin/bash
# Full GPU check for one gpurun call: parity tests per file (separate processes so a CUDA fault in one file
# cannot poison the others), smoke, a short bench.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for f in kernels mlp render; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q -rA --timeout 600 > gpurun_out/test_$f.log 2>&1
  echo "test_gpu_$f exit $?" >> gpurun_out/summary.txt
done
NERFAIL_B200_CG=1 timeout 900 python -m pytest tests/test_gpu_mlp.py -m gpu -q -rA --timeout 600 -k fused > gpurun_out/test_mlp_cg1.log 2>&1
echo "test_gpu_mlp (cta_group::1) exit $?" >> gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/summary.txt
if [ "${SKIP_BENCH:-0}" != "1" ]; then
  timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
  echo "bench exit $?" >> gpurun_out/summary.txt
  NERFAIL_B200_CG=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cg1.log 2> gpurun_out/bench_cg1.err
  echo "bench (cta_group::1) exit $?" >> gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
grep -hE "^(FAILED|ERROR)|passed|failed" gpurun_out/test_*.log | tail -30
tail -n 3 gpurun_out/smoke.log
[ -f gpurun_out/bench.log ] && tail -c 2500 gpurun_out/bench.log
[ -f gpurun_out/bench_cg1.log ] && grep -o '"roofline.*' gpurun_out/bench_cg1.log
