#!/bin/bash
# Round-2 multi-GPU check (gpurun --gpus N): the peer-memory exchange tests, then bench.py at N ranks under torchrun.
set -u
N=${NGPU:-2}
O=gpurun_out/r2
mkdir -p $O
nvidia-smi topo -m > $O/topo_n$N.txt 2>&1
for f in ${TEST_FILES:-multi}; do
  timeout 1200 python -m pytest tests/test_gpu_$f.py -m gpu -q -rA -s --durations=5 --timeout 900 > $O/test_${f}_n$N.log 2>&1
  echo "test_gpu_$f (N=$N) exit $?"
done
if [ "${SKIP_BENCH:-0}" != "1" ]; then
  for n in ${BENCH_N:-$N}; do
    if [ "$n" = "1" ]; then
      timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
        bench.py --gpus $n --steps 5 --warmup 3 > $O/bench_n$n.json 2> $O/bench_n$n.err
    fi
    echo "bench N=$n exit $?"
    tail -c 300 $O/bench_n$n.err
    python - <<PY
import json
try:
    d = json.loads(open("$O/bench_n$n.json").read().strip().splitlines()[-1])
    print("N=$n value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "strong", d.get("strong"))
    ex = d.get("extra", {})
    a = ex.get("attack_iteration", {})
    print("  attack", {k: a.get(k) for k in ("ms_per_iteration", "ms_gather_scatter_only", "ms_exchange_and_update", "ms_per_iteration_nccl_allreduce_then_update")})
    print("  knn_sweep", ex.get("knn_sweep", {}).get("value"), "strict", ex.get("render_strict_chunk_1024", {}).get("value"))
except Exception as e:
    print("no bench line:", e)
PY
  done
fi
grep -hE "^(FAILED|ERROR)|passed|failed" $O/test_*_n$N.log | tail
