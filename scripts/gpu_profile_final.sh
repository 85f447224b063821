#!/bin/bash
# Round-end ncu evidence (each ncu run directly preceded by the identical plain run):
#  (1) launch list of the bench command (render step only: --no-extras keeps it under the capture limit)
#  (2) --set full capture of the fused MLP kernel (production instantiation) at fine-pass size
#  (3) launch list of one bf16 retraining step + --set full captures of the three training kernels
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$B > gpurun_out/bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/bench_ncu.log 2>&1
echo "launch list exit $?"
python scripts/profile_mlp.py 65536 192 > gpurun_out/profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mlp_fused_fwd -s 2 -c 1 -f -o gpurun_out/prof_mlp \
    python scripts/profile_mlp.py 65536 192 > gpurun_out/profile_ncu.log 2>&1
echo "mlp capture exit $?"
python scripts/profile_train.py 10 > gpurun_out/train_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/train_launches.csv \
  python scripts/profile_train.py 1 > gpurun_out/train_ncu.log 2>&1
echo "train launch list exit $?"
ONLY="bwd data,bwd weights" python scripts/time_train_kernels.py > gpurun_out/train_kernels_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"wgrad_kernel|mlp_train_kernel" -s 6 -c 3 -o gpurun_out/prof_train -f \
  env ONLY="bwd data,bwd weights" python scripts/time_train_kernels.py > gpurun_out/prof_train.log 2>&1
echo "train capture exit $?"
cat gpurun_out/profile_plain.log gpurun_out/train_plain.log gpurun_out/train_kernels_plain.log
tail -c 600 gpurun_out/bench_plain.log
