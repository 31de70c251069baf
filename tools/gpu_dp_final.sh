#!/bin/bash
N=${1:-8}; TAG=${2:-fin}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/dp_check.py cifar 2>&1 | grep -E "rank 0|Error|error" | head -3
timeout 300 $TR bench.py --gpus $N --steps 100 --warmup 3 > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err; echo "cifar rc=$?"
timeout 300 $TR tools/bench_mlp.py --steps 20 > $OUT/mlp_${TAG}_n$N.json 2> $OUT/mlp_${TAG}_n$N.err; echo "mlp rc=$?"
python - <<P
import json
for f in ["bench_${TAG}_n$N", "mlp_${TAG}_n$N"]:
    lines=[l for l in open("$OUT/" + f + ".json") if l.startswith("{")]
    d = json.loads(lines[-1])
    print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d.get("e2e",{}).get("value",0),1))
P
