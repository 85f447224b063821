#!/bin/bash
# Profiles of the bf16 retraining step: plain timing, ncu launch list of the same command, one --set full capture of each
# tensor-core kernel (fine-network size) after the plain run exited 0.
set -u
mkdir -p gpurun_out
python scripts/profile_train.py 10 > gpurun_out/train_plain.log 2>&1 || exit 1
python scripts/time_train_kernels.py > gpurun_out/train_kernels_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/train_launches.csv \
  python scripts/profile_train.py 1 > gpurun_out/train_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"wgrad_kernel|mlp_train_kernel" -c 21 -o gpurun_out/prof_train -f \
  python scripts/time_train_kernels.py > gpurun_out/prof_train.log 2>&1
cat gpurun_out/train_plain.log gpurun_out/train_kernels_plain.log
