#!/bin/bash
set -u
TAG=${1:-yx}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python tools/bench_yolo.py --steps 2 --warmup 1 ${YOLO_ARGS:-} > $OUT/yolo_$TAG.json 2> $OUT/yolo_$TAG.err
echo "yolo rc=$?"; cat $OUT/yolo_$TAG.json; tail -5 $OUT/yolo_$TAG.err
