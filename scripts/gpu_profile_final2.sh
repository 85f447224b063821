#!/bin/bash
# Round-end captures of the HBM-bound kernels: grouped wgrad, GaussNet gather / scatter, compositing, hierarchical.
set -u
mkdir -p gpurun_out
ONLY="bwd weights" python scripts/time_train_kernels.py > gpurun_out/wgrad_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 3 -c 1 -f -o gpurun_out/prof_wgrad \
  env ONLY="bwd weights" python scripts/time_train_kernels.py > gpurun_out/prof_wgrad.log 2>&1
echo "wgrad capture exit $?"
python scripts/profile_gauss.py > gpurun_out/gauss_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:"gauss_gather_fwd|gauss_scatter_bwd" -s 8 -c 2 -f -o gpurun_out/prof_gauss \
  python scripts/profile_gauss.py > gpurun_out/gauss_ncu.log 2>&1
echo "gauss capture exit $?"
python scripts/profile_mlp.py 262144 192 > gpurun_out/profile_plain2.log 2>&1 &&
ncu --set full --clock-control none -k regex:'composite_fwd|hierarchical' -c 2 -f -o gpurun_out/prof_hbm \
    python scripts/profile_mlp.py 262144 192 > gpurun_out/profile_ncu2.log 2>&1
echo "hbm capture exit $?"
cat gpurun_out/wgrad_plain.log gpurun_out/gauss_plain.log | tail -8
