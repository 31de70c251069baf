#!/bin/bash
set -u
TAG=${1:-cx}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_conv_gemm.py tests/test_gpu_dense_umma.py -x -q > $OUT/pytest_conv_$TAG.log 2>&1
echo "pytest rc=$?"; tail -25 $OUT/pytest_conv_$TAG.log
