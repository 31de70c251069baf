#!/bin/bash
# ncu --set full of the dense tensor-core GEMM kernels inside the wide-MLP step.
set -u
TAG=${1:-dx}
OUT=gpurun_out
mkdir -p $OUT
CMD="python tools/bench_mlp.py --steps 1 --warmup 3"
$CMD > $OUT/mlp_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail $OUT/mlp_plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'dense_umma' -s 12 -c 6 \
    -f -o $OUT/prof_dense_$TAG $CMD > $OUT/ncu_dense_$TAG.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/ncu_dense_$TAG.log; ls -la $OUT/prof_dense_$TAG.ncu-rep
