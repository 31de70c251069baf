#!/bin/bash
set -u
TAG=${1:-yx}
OUT=gpurun_out
mkdir -p $OUT
CMD="python tools/bench_yolo.py --steps 1 --warmup 1 --global-batch 32"
$CMD > $OUT/yolo_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail $OUT/yolo_plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'umma_gemm' -c ${NCU_COUNT:-4} \
    -f -o $OUT/prof_yolo_$TAG $CMD > $OUT/ncu_yolo_$TAG.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/ncu_yolo_$TAG.log | cut -c1-300; ls -la $OUT/prof_yolo_$TAG.ncu-rep
