#!/bin/bash
# variants/lib_prof.so = the product library with conv_wgrad_umma.cu / conv_wt_umma.cu compiled -DPB_WGRAD_PROF / -DPB_WT_PROF
# (per-role cycle counters printed by CTA 0); select it with PB_LIB_PATH.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $ROOT/variants
cd $ROOT/plasticity_b200/csrc
make -j8 >/dev/null
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC \
  -I../../include -I. --expt-relaxed-constexpr -DPB_WGRAD_PROF -c conv_wgrad_umma.cu -o $ROOT/variants/conv_wgrad_umma_prof.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC \
  -I../../include -I. --expt-relaxed-constexpr -DPB_WT_PROF -c conv_wt_umma.cu -o $ROOT/variants/conv_wt_umma_prof.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC \
  -I../../include -I. --expt-relaxed-constexpr -DPB_UMMA_PROF -c conv_umma.cu -o $ROOT/variants/conv_umma_prof.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC \
  -I../../include -I. --expt-relaxed-constexpr -DPB_ROWS_PROF -c conv_rows_umma.cu -o $ROOT/variants/conv_rows_umma_prof.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC \
  -I../../include -I. --expt-relaxed-constexpr -DPB_WROWS_PROF -c conv_wgrad_rows_umma.cu -o $ROOT/variants/conv_wgrad_rows_umma_prof.o
OBJS=$(ls build/*.o | grep -v "conv_wgrad_umma.o\|conv_wt_umma.o\|conv_umma.o\|conv_rows_umma.o\|conv_wgrad_rows_umma.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/variants/lib_prof.so $OBJS $ROOT/variants/conv_wgrad_umma_prof.o $ROOT/variants/conv_wt_umma_prof.o $ROOT/variants/conv_umma_prof.o $ROOT/variants/conv_rows_umma_prof.o $ROOT/variants/conv_wgrad_rows_umma_prof.o -lcudart -ldl
