#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "tensor_core_large or fused_plan or layer_kernels" 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_wg.json 2> $OUT/bench_wg.err; echo rc=$?
PB_WGRAD_MMA=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_wg_off.json 2> $OUT/bench_wg_off.err; echo rc=$?
