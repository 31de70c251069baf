#!/bin/bash
# A/B of GEMM kernel variants: bash tools/gpu_variants.sh <tag> name1 name2 ...
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
for v in base "$@"; do
  if [ $v = base ]; then unset PB_LIB_PATH; else export PB_LIB_PATH=$PWD/variants/lib_$v.so; fi
  timeout 300 python tools/bench_mlp.py --steps 10 > $OUT/mlp_${TAG}_$v.json 2> $OUT/mlp_${TAG}_$v.err || tail -3 $OUT/mlp_${TAG}_$v.err
  timeout 600 python tools/bench_yolo.py --steps 2 --warmup 1 --global-batch 64 > $OUT/yolo_${TAG}_$v.json 2> $OUT/yolo_${TAG}_$v.err || tail -3 $OUT/yolo_${TAG}_$v.err
  python - <<P
import json
for k in ("mlp","yolo"):
    try:
        d=[json.loads(l) for l in open("$OUT/%s_${TAG}_$v.json"%k) if l.startswith("{")][-1]
        rows=d.get("step_kernels") or d.get("layers")
        top=sorted(rows,key=lambda r:-r["ms"])[:6]
        print("$v",k,"ms",round(d["ms_per_step"],3)," ".join("%s=%.3f"%(r["k"].split(":")[0][-14:]+":"+r["k"].split(":")[1][:6],r["ms"]) for r in top))
    except Exception as e: print("$v",k,"failed",e)
P
done
