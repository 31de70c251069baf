#!/bin/bash
# Build a variant of libplasticity_b200.so with extra nvcc flags into gpurun_out-free scratch:
#   bash tools/build_variant.sh <name> "<extra nvcc flags>"   -> variants/lib_<name>.so
set -e
NAME=$1; FLAGS=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
D=$ROOT/variants/build_$NAME
mkdir -p $D
cd $ROOT/plasticity_b200/csrc
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
OBJS=""
for f in *.cu; do
  o=$D/${f%.cu}.o
  case $f in
    dense_umma.cu|conv_gemm_umma.cu) $NVCC $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I../../include -I. --expt-relaxed-constexpr $FLAGS -c $f -o $o & ;;
    *) cp build/${f%.cu}.o $o ;;
  esac
  OBJS="$OBJS $o"
done
wait
$NVCC $ARCH -shared -o $ROOT/variants/lib_$NAME.so $OBJS -lcudart -ldl
ls -la $ROOT/variants/lib_$NAME.so
