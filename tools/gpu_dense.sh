#!/bin/bash
# Dense tensor-core GEMM round: parity tests + wide-MLP bench.  bash tools/gpu_dense.sh <tag>
set -u
TAG=${1:-dx}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_dense_umma.py -x -q > $OUT/pytest_dense_$TAG.log 2>&1
echo "pytest rc=$?"; tail -15 $OUT/pytest_dense_$TAG.log
timeout 600 python tools/bench_mlp.py --steps 10 > $OUT/mlp_$TAG.json 2> $OUT/mlp_$TAG.err
echo "mlp rc=$?"; cat $OUT/mlp_$TAG.json; tail -5 $OUT/mlp_$TAG.err
PB_DISABLE_UMMA=1 timeout 600 python tools/bench_mlp.py --steps 5 > $OUT/mlp_ffma_$TAG.json 2> $OUT/mlp_ffma_$TAG.err
echo "mlp ffma rc=$?"; cat $OUT/mlp_ffma_$TAG.json
