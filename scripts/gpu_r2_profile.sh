#!/bin/bash
# Round-2 ncu evidence; every ncu run is directly preceded by the identical plain run.  Reports land in gpurun_out/r2/ and
# are exported to text under profiles/ by scripts/ncu_export.py in the build container.
set -u
O=gpurun_out/r2; mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$B > $O/prof_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv $B > $O/prof_bench_ncu.log 2>&1
echo "bench launch list exit $?"
python scripts/profile_mlp.py 65536 192 > $O/prof_mlp_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mlp_fused_fwd -s 2 -c 1 -f -o $O/prof_mlp \
    python scripts/profile_mlp.py 65536 192 > $O/prof_mlp_ncu.log 2>&1
echo "mlp capture exit $?"
python scripts/profile_mlp.py 262144 192 > $O/prof_hbm_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'composite_fwd|hierarchical' -c 2 -f -o $O/prof_hbm \
    python scripts/profile_mlp.py 262144 192 > $O/prof_hbm_ncu.log 2>&1
echo "hbm capture exit $?"
python scripts/profile_gauss10.py > $O/prof_gauss_plain.log 2>&1 &&
ncu --set full --metrics lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum \
    --clock-control none -k regex:"gauss_gather_fwd|gauss_scatter_bwd" -s 4 -c 2 -f -o $O/prof_gauss \
    python scripts/profile_gauss10.py clustered > $O/prof_gauss_ncu.log 2>&1
echo "gauss capture exit $?"
python scripts/time_train_kernels.py > $O/prof_train_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"wgrad_kernel|mlp_train_kernel" -s 6 -c 3 -f -o $O/prof_train \
    env ONLY="bwd data,bwd weights" python scripts/time_train_kernels.py > $O/prof_train_ncu.log 2>&1
echo "train capture exit $?"
python scripts/profile_train.py 10 > $O/prof_trainstep_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_train.csv \
    python scripts/profile_train.py 1 > $O/prof_trainstep_ncu.log 2>&1
echo "train launch list exit $?"
cat $O/prof_mlp_plain.log $O/prof_gauss_plain.log $O/prof_train_plain.log $O/prof_trainstep_plain.log | tail -30
tail -c 400 $O/prof_bench_plain.log
