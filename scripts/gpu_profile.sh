#!/bin/bash
# ncu evidence for profiles/: (1) launch list of the bench command, (2) full-set capture of the fused MLP kernel,
# (3) full-set capture of the HBM-bound kernels.  Each ncu run is preceded by the identical plain run (&&).
set -u
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
echo "launch list exit $?"
python scripts/profile_mlp.py 65536 192 > gpurun_out/profile_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mlp_fused_fwd -s 2 -c 1 -f -o gpurun_out/prof_mlp \
    python scripts/profile_mlp.py 65536 192 > gpurun_out/profile_ncu.log 2>&1
echo "mlp capture exit $?"
python scripts/profile_mlp.py 65536 192 > gpurun_out/profile_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'composite_fwd|hierarchical' -c 2 -f -o gpurun_out/prof_hbm \
    python scripts/profile_mlp.py 65536 192 > gpurun_out/profile_ncu2.log 2>&1
echo "hbm capture exit $?"
cat gpurun_out/profile_plain.log
ls -la gpurun_out/
