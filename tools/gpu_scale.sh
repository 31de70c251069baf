#!/bin/bash
# bench.py at N = 1, 2, 4, 8 back to back on one 8-GPU box (gpurun --gpus 8 -- bash tools/gpu_scale.sh TAG)
TAG=${1:-r02e}
OUT=gpurun_out; mkdir -p $OUT
for N in ${NS:-1 2 4 8}; do
  if [ $N = 1 ]; then
    timeout 300 python bench.py --steps 20 --warmup 3 --no-other-configs --no-cpu-baseline > $OUT/scale_${TAG}_n$N.json 2> $OUT/scale_${TAG}_n$N.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N)) \
      bench.py --gpus $N --steps 20 --warmup 3 --no-other-configs > $OUT/scale_${TAG}_n$N.json 2> $OUT/scale_${TAG}_n$N.err
  fi
  echo "N=$N rc=$?"
done
python - <<P
import json
base = None
rows = []
for n in [int(a) for a in "${NS:-1 2 4 8}".split()]:
    try:
        d = json.loads([l for l in open("$OUT/scale_${TAG}_n%d.json" % n) if l.startswith("{")][-1])
    except Exception as e:
        print(n, "no line", e); continue
    if n == 1: base = d["value"]
    rows.append({"n_gpus": n, "value": d["value"], "ms_per_step": d["ms_per_step"],
                 "efficiency": d["value"] / (n * base) if base else None,
                 "e2e": d["e2e"]["value"], "dp_parity": d.get("dp_parity")})
    print(n, round(d["value"]), round(d["ms_per_step"], 4), rows[-1]["efficiency"])
json.dump({"note": "builder-run scaling on one 8-GPU box (gpurun --gpus 8), bench.py at N = 1, 2, 4, 8 back to back; efficiency = value / (N x value at N=1); the driver computes its own from SCALE_r02.json",
           "cifar_cnn_train": rows}, open("$OUT/scaling_${TAG}.json", "w"), indent=1)
P
