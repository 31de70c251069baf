#!/bin/bash
N=${1:-2}; TAG=${2:-dpc}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/dp_check.py cifar 2>&1 | grep -E "bit-identical|Error|error" | head -3
timeout 600 $TR bench.py --gpus $N --steps 30 --warmup 3 > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err; echo "cifar rc=$?"
PB_DP_BUCKETS=0 timeout 600 $TR bench.py --gpus $N --steps 30 --warmup 3 > $OUT/bench_${TAG}_n${N}_inbox.json 2> $OUT/bench_${TAG}_n${N}_inbox.err; echo "cifar inbox rc=$?"
python - <<P
import json
for f in ["bench_${TAG}_n$N", "bench_${TAG}_n${N}_inbox"]:
    d = json.load(open("$OUT/" + f + ".json"))
    print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"],1))
P
