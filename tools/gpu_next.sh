#!/bin/bash
# First GPU call of the next round: everything that was written without a GPU.
#   bash tools/gpu_next.sh <tag>            (1 GPU)
#   bash tools/gpu_next.sh <tag> 2          (under gpurun --gpus 2: also the chained DP step)
# 1. opt-in kernels against the oracle (one-warp and cooperative per-sample SGD)
# 2. their rates: circle MLP with PB_TRAIN_WARP=1, per-sample CIFAR with PB_TRAIN_COOP=1
# 4. tools/umma_bf16_probe
# 3. with 2+ GPUs: tools/dp_check.py and the bench with PB_STEP_CHAIN=2 (chained delta pass)
set -u
TAG=${1:-rXX}; N=${2:-1}
OUT=gpurun_out; mkdir -p $OUT
PB_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q \
    -k "cooperative or one_warp" > $OUT/experimental_$TAG.log 2>&1
echo "experimental tests rc=$?"; grep -E "^E  |passed|failed" $OUT/experimental_$TAG.log | cut -c1-220 | tail -15
(cd tests/cpp && PB_EXPERIMENTAL=1 timeout 120 ./mirror_test | tail -3)   # + compute::CudaCommandQueue case
for w in 0 1; do
  PB_TRAIN_WARP=$w timeout 120 python tools/bench_circle.py > $OUT/circle_${TAG}_warp$w.json 2> $OUT/circle_${TAG}_warp$w.err
  echo "circle PB_TRAIN_WARP=$w rc=$?: $(tail -1 $OUT/circle_${TAG}_warp$w.json | cut -c1-160)"
done
for c in 0 1; do
  echo "per-sample CIFAR PB_TRAIN_COOP=$c: $(cd tests/cpp && PB_TRAIN_COOP=$c timeout 200 ./cifar_train --synthetic 4000 --epochs 1 --lr 0.001 --examples 100 --seed 3 2>&1 | grep -E 'Training rate' | tail -1)"
done
# 4. bf16 operand conventions + issue cost of the bf16x3 convolution variant
[ -x tools/umma_bf16_probe ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_bf16_probe tools/umma_bf16_probe.cu
timeout 60 tools/umma_bf16_probe > $OUT/umma_bf16_probe_$TAG.txt 2>&1; echo "bf16 probe rc=$?"; tail -4 $OUT/umma_bf16_probe_$TAG.txt
if [ "$N" -gt 1 ]; then
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
  for c in 1 2; do
    PB_STEP_CHAIN=$c timeout 300 $TR tools/dp_check.py cifar 2>&1 | grep -E "rank 0|rror" | head -3
    PB_STEP_CHAIN=$c timeout 300 $TR bench.py --gpus $N --steps 100 --warmup 5 > $OUT/bench_${TAG}_n${N}_chain$c.json 2> $OUT/bench_${TAG}_n${N}_chain$c.err
    echo "bench n=$N PB_STEP_CHAIN=$c rc=$?: $(python -c "import json,sys; d=json.loads([l for l in open('$OUT/bench_${TAG}_n${N}_chain$c.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])")"
  done
fi
