#!/bin/bash
# tensor-core GEMM round: parity tests of the dense / conv GEMM kernels + MLP and YOLO benches
set -u
TAG=${1:-ax}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_conv_gemm.py tests/test_gpu_dense_umma.py -x -q > $OUT/pytest_gemm_$TAG.log 2>&1
echo "pytest rc=$?"; tail -4 $OUT/pytest_gemm_$TAG.log
timeout 600 python tools/bench_mlp.py --steps 10 > $OUT/mlp_$TAG.json 2> $OUT/mlp_$TAG.err
echo "mlp rc=$?"; tail -3 $OUT/mlp_$TAG.err
timeout 1200 python tools/bench_yolo.py --steps 2 --warmup 1 > $OUT/yolo_$TAG.json 2> $OUT/yolo_$TAG.err
echo "yolo rc=$?"; tail -3 $OUT/yolo_$TAG.err
python - <<'P'
import json,sys
tag=sys.argv[1] if len(sys.argv)>1 else ''
P
