#!/bin/bash
# data-parallel checks + scaling lines on N GPUs of one box: bash tools/gpu_dp.sh <N> <tag>
set -u
N=${1:-2}; TAG=${2:-dp}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/dp_check.py cifar 2>&1 | grep -E "bit-identical|Error|error" | head -4
timeout 600 $TR tools/dp_check.py mlp 2>&1 | grep -E "bit-identical|Error|error" | head -4
timeout 600 $TR tools/bench_mlp.py --steps 10 > $OUT/mlp_${TAG}_n$N.json 2> $OUT/mlp_${TAG}_n$N.err; echo "mlp overlap rc=$?"
PB_DP_OVERLAP=0 timeout 600 $TR tools/bench_mlp.py --steps 10 > $OUT/mlp_${TAG}_n${N}_noov.json 2> $OUT/mlp_${TAG}_n${N}_noov.err; echo "mlp no-overlap rc=$?"
timeout 900 $TR tools/bench_yolo.py --steps 2 --warmup 1 > $OUT/yolo_${TAG}_n$N.json 2> $OUT/yolo_${TAG}_n$N.err; echo "yolo rc=$?"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err; echo "cifar rc=$?"
python - <<P
import json
for f in ["mlp_${TAG}_n$N", "mlp_${TAG}_n${N}_noov", "yolo_${TAG}_n$N", "bench_${TAG}_n$N"]:
    try:
        d = json.load(open("$OUT/" + f + ".json"))
        print(f, "value", round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 4))
    except Exception as e:
        print(f, "missing", e)
P
