#!/usr/bin/env python
"""CPU baselines of BASELINE.json configs 4 and 5 (a bounded sample,
extrapolated and marked so) with the oracle's plain-C port of the reference arithmetic
(fp64), plus the measured TF32 tensor peak of this GPU when one is present.
One JSON object on stdout."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import pyoracle as po

out = {"host_cores": os.cpu_count()}
rng = np.random.default_rng(0)

# C4: wide MLP, one gradient-sum step over a few samples
Wd, nb = 4096, 4
spec = (f"loss ce\nact {Wd} identity\ndense {Wd} {Wd}\nact {Wd} sigmoid\ndense {Wd} {Wd}\n"
        f"act {Wd} sigmoid\ndense {Wd} 10\nact 10 identity\nsoftmax 10\n")
S = po.RestatedNet(spec)
w = rng.normal(0, 1 / np.sqrt(Wd), S.total_w)
xs = rng.normal(0, 1, (nb, Wd))
es = np.zeros((nb, 10)); es[np.arange(nb), rng.integers(0, 10, nb)] = 1.0
t0 = time.perf_counter()
S.batch_train_mt(xs, es, w, 1e-4)
dt = time.perf_counter() - t0
out["wide_mlp_train"] = {"value": nb / dt, "unit": "samples/s", "kind": "port",
                         "sample": f"one gradient-sum step over {nb} samples (OpenMP over examples, fp64), "
                                   "extrapolated: rate is per sample", "seconds": dt}

# C5: YOLOv1 forward, one image
sys.path.insert(0, ROOT)
from plasticity_b200 import yolov1
import plasticity_b200 as pb
arch = yolov1.YoloArchitecture()
Y = po.RestatedNet(arch.to_spec(pb.MeanSquared))
wy = (rng.standard_normal(Y.total_w) * 0.01)
x = rng.uniform(0, 1, Y.n_in)
t0 = time.perf_counter()
Y.evaluate(x, wy)
dt = time.perf_counter() - t0
out["yolov1_forward"] = {"value": 1.0 / dt, "unit": "images/s", "kind": "port",
                         "sample": "one 448x448x3 image forward (single thread, fp64), extrapolated",
                         "seconds": dt}

try:
    import torch
    if torch.cuda.is_available():
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn(8192, 8192, device="cuda"); b = torch.randn(8192, 8192, device="cuda")
        best = 0.0
        for _ in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
            best = max(best, 2 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        out["tf32_tflops_burst"] = best
        out["tf32_how"] = "torch.matmul fp32 8192^3 with allow_tf32, best of 12, CUDA events"
except Exception as e:  # no torch / no GPU: CPU numbers only
    out["tf32_error"] = str(e)
print(json.dumps(out))
