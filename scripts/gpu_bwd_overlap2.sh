#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/overlap_modes.log
for g in ${SWEEP:-28 34 40}; do
  for m in 1 2 3 0; do
    ONLY="bwd overlapped" NERFAIL_B200_BWD_MODE=$m NERFAIL_B200_BWD_PRODUCERS=$g timeout 120 python scripts/time_train_kernels.py 2>&1 | grep -E "overlapped" | sed "s/^/G=$g mode=$m /" | tee -a gpurun_out/overlap_modes.log
  done
done
ONLY="bwd overlapped" NERFAIL_B200_WGRAD_DBG=1 NERFAIL_B200_BWD_PRODUCERS=34 timeout 120 python scripts/time_train_kernels.py > gpurun_out/overlap_dbg.log 2>&1
tail -n 80 gpurun_out/overlap_dbg.log | head -75
