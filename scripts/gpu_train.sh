#!/bin/bash
# Training-path check for one gpurun call: wgrad / training parity tests, one profiled retraining step.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlp.py -m gpu -q -rA --timeout 300 -k "wgrad or train" > gpurun_out/test_train.log 2>&1
echo "test train exit $?" > gpurun_out/train_summary.txt
timeout 300 python scripts/profile_train.py 10 > gpurun_out/train_plain.log 2>&1
echo "profile_train exit $?" >> gpurun_out/train_summary.txt
if [ "${NCU:-1}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/train_launches.csv \
    python scripts/profile_train.py 1 > gpurun_out/train_ncu.log 2>&1
  echo "ncu exit $?" >> gpurun_out/train_summary.txt
fi
cat gpurun_out/train_summary.txt
grep -hE "^(FAILED|ERROR)|passed|failed|vs " gpurun_out/test_train.log | tail -20
cat gpurun_out/train_plain.log | tail -5
