#!/usr/bin/env python
"""Summarise gpurun_out/{mlp,yolo}_<tag>.json."""
import json, sys
tag = sys.argv[1]
for kind in ("mlp", "yolo"):
    try:
        d = json.load(open(f"gpurun_out/{kind}_{tag}.json"))
    except Exception as e:
        print(kind, "missing", e); continue
    w = d.get("whole_step") or d.get("whole_pass")
    print(kind, "ms", round(d["ms_per_step"], 3), "value", round(d["value"], 1), "TF/s", round(w["achieved_tflops"], 1), d["clocks"])
    for r in (d.get("step_kernels") or d.get("layers")):
        if r["ms"] > 0.05: print("   ", r["k"][:60].ljust(60), r["ms"], r["tflops"])
