#!/bin/bash
# ncu captures of the tensor-core GEMM configurations: bash tools/gpu_prof2.sh <tag>
set -u
TAG=${1:-px}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python tools/bench_mlp.py --steps 10 > $OUT/mlp_$TAG.json 2> $OUT/mlp_$TAG.err; echo "mlp rc=$?"
timeout 1200 python tools/bench_yolo.py --steps 2 --warmup 1 > $OUT/yolo_$TAG.json 2> $OUT/yolo_$TAG.err; echo "yolo rc=$?"
CMD="python tools/bench_mlp.py --steps 1 --warmup 3"
ncu --set full --clock-control none --import-source on -k regex:'umma_gemm' -s 12 -c 6 \
    -f -o $OUT/prof_dense_$TAG $CMD > $OUT/ncu_dense_$TAG.log 2>&1; echo "ncu dense rc=$?"
CMD="python tools/bench_yolo.py --steps 1 --warmup 1 --global-batch 32"
ncu --set full --clock-control none --import-source on -k regex:'umma_gemm' -c 11 \
    -f -o $OUT/prof_yolo_$TAG $CMD > $OUT/ncu_yolo_$TAG.log 2>&1; echo "ncu yolo rc=$?"
ls -la $OUT/*.ncu-rep | tail -3
