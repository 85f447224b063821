#!/bin/bash
# Overlapped backward (nfb_mlp_bwd): parity test, then a sweep of the producer / consumer SM split at fine- and coarse-network size.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mlp.py -m gpu -q -rA --timeout 120 -k "overlapped" > gpurun_out/test_overlap.log 2>&1
echo "test overlap exit $?" | tee gpurun_out/overlap_summary.txt
tail -n 15 gpurun_out/test_overlap.log
if grep -q "1 passed" gpurun_out/test_overlap.log; then
  timeout 120 python scripts/time_train_kernels.py 2>&1 | tee gpurun_out/overlap_sweep.log
  for g in ${SWEEP:-32 36 38 42 44 48}; do
    ONLY="bwd overlapped" NERFAIL_B200_BWD_PRODUCERS=$g timeout 120 python scripts/time_train_kernels.py 2>&1 | grep -E "producers|overlapped" | tee -a gpurun_out/overlap_sweep.log
  done
  for g in 36 40 44; do
    S=64 ONLY="bwd data,bwd weights,bwd overlapped" NERFAIL_B200_BWD_PRODUCERS=$g timeout 120 python scripts/time_train_kernels.py 2>&1 | tee -a gpurun_out/overlap_sweep.log
  done
  timeout 300 python scripts/profile_train.py 10 2>&1 | tail -2 | tee -a gpurun_out/overlap_sweep.log
  NERFAIL_B200_BWD=serial timeout 300 python scripts/profile_train.py 10 2>&1 | tail -2 | tee -a gpurun_out/overlap_sweep.log
fi
