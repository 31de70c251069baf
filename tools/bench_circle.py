#!/usr/bin/env python
"""BASELINE.json config 1: nnet/circle_test.cc — 2-10-1 sigmoid MLP, MSE, lr 0.3, strict
per-sample SGD on points (x, y) ~ U(-1.25, 1.25)^2 with target [x^2 + y^2 <= 1]
(circle_test.cc:19-38).  Reports samples/s of the device loop (pb_nnet_train_loop) and of
the reference's CPU path (oracle: generated kernels or the plain-C port), plus the
classification accuracy after training."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    import plasticity_b200 as pb
    from plasticity_b200 import _lib as L
    rng = np.random.default_rng(0)
    xs = rng.uniform(-1.25, 1.25, (n, 2))
    es = (xs[:, 0] ** 2 + xs[:, 1] ** 2 <= 1.0).astype(np.float64)[:, None]
    model = pb.Architecture(2).AddDenseLayer(10).AddDenseLayer(1)
    net = pb.Nnet(model, pb.Xavier, pb.MeanSquared, seed=1)
    net.SetLearningParameters(pb.Nnet.LearningParameters(0.3))
    dx, de = net.MakeBuffer(xs.reshape(-1)), net.MakeBuffer(es.reshape(-1))
    net.TrainLoop(dx, de, 1000)  # warm
    net.Finish()
    t0 = time.perf_counter()
    net.TrainLoop(dx, de, n)
    net.Finish()
    dt = time.perf_counter() - t0
    probe = rng.uniform(-1.25, 1.25, (2000, 2))
    out = net.Evaluate(net.MakeBuffer(probe.reshape(-1)), batch=1) if False else None
    ok = 0
    outs = []
    for i in range(0, 2000, 1):
        pass
    # accuracy through a batched evaluate on a fresh network object sharing the weights
    net_b = pb.Nnet(model, pb.NoWeightInit, pb.MeanSquared, max_batch=2000)
    for l in range(len(model.layers)):
        if net.nw[l]:
            net_b.SetLayerWeights(l, net.GetLayerWeights(l))
    o = net_b.Evaluate(net_b.MakeBuffer(probe.reshape(-1)), batch=2000)
    o.MoveToCpu()
    pred = np.asarray(o.cpu) > 0.5
    acc = float((pred == (probe[:, 0] ** 2 + probe[:, 1] ** 2 <= 1.0)).mean())
    line = {"metric": "circle_mlp_per_sample_sgd_samples_per_sec", "value": n / dt, "unit": "samples/s",
            "samples": n, "accuracy_after_training": acc,
            "config": {"workload": "circle_test 2-10-1 sigmoid MLP, MSE, lr 0.3, per-sample SGD"}}
    try:
        from oracle import pyoracle as po
        spec = po.spec_text("circle")
        R = po.RefNet("circle") if po.RefNet.available("circle") else po.RestatedNet(spec)
        w = np.zeros(R.total_w)
        w[:] = rng.normal(0, 0.5, R.total_w)
        m = min(n, 20000)
        t0 = time.perf_counter()
        R.train_loop(xs[:m].copy(), es[:m].copy(), w, 0.3)
        line["cpu_baseline"] = {"value": m / (time.perf_counter() - t0), "unit": "samples/s",
                                "cores": 1, "kind": "reference" if isinstance(R, po.RefNet) else "port",
                                "sample": f"{m} per-sample SGD steps (sequential by construction)"}
    except Exception as e:  # the oracle is optional here
        line["cpu_baseline"] = {"unavailable": str(e)}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
