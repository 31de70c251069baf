#!/bin/bash
# One gpurun call: parity tests, chained-step race check, bench line, ncu launch list and
# full capture of the convolution kernels.  Usage (from the repo root, on the GPU box):
#   bash tools/gpu_round.sh <tag>
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -3 $OUT/pytest_gpu_$TAG.log
python bench.py --steps 20 --warmup 3 --no-other-configs --table $OUT/table_$TAG.json > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"; cat $OUT/bench_$TAG.json
BENCH_SHORT="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs --min-seconds 0 --max-blocks 1"
$BENCH_SHORT > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
    --log-file $OUT/launches_$TAG.csv $BENCH_SHORT > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
$BENCH_SHORT > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'conv_' -s ${NCU_SKIP:-27} -c ${NCU_COUNT:-9} \
    -f -o $OUT/prof_$TAG $BENCH_SHORT > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT
