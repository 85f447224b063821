#!/bin/bash
# the per-GPU share of a strong-scaled retraining step: timing + ncu launch list of ONE graph replay
set -u
O=gpurun_out/r2; mkdir -p $O
python scripts/profile_train_small.py 512 20 > $O/train_small_plain.log 2>&1 && cat $O/train_small_plain.log | tail -2 &&
ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 900 --csv --log-file $O/train_small_launches.csv \
  python scripts/profile_train_small.py 512 1 > $O/train_small_ncu.log 2>&1
echo "launch list exit $?"
python scripts/launch_summary.py $O/train_small_launches.csv 2>/dev/null | tail -40
