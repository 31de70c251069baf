#!/bin/bash
# One gpurun call: chained-step race check, GPU parity tests, bench with and without the
# chained launches.  Usage (repo root, on the GPU box): bash tools/gpu_chain.sh <tag>
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python tools/chain_check.py 100 1024 > $OUT/chain_check_$TAG.log 2>&1
echo "chain_check rc=$?"; tail -4 $OUT/chain_check_$TAG.log
timeout 300 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?"; tail -2 $OUT/pytest_gpu_$TAG.log
for c in 1 0; do
  PB_STEP_CHAIN=$c timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline \
      > $OUT/bench_${TAG}_chain$c.json 2> $OUT/bench_${TAG}_chain$c.err
  echo "bench chain=$c rc=$?"
  python - <<P
import json
try:
    d=json.loads(open("$OUT/bench_${TAG}_chain$c.json").read().strip().splitlines()[-1])
    print("chain=$c", d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
except Exception as e: print("no line", e)
P
done
