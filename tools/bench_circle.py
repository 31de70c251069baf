#!/usr/bin/env python
"""BASELINE.json config 1 and the reference's own CIFAR loop: strict per-sample SGD.

* circle: nnet/circle_test.cc — 2-10-1 sigmoid MLP, MSE, lr 0.3, points (x, y) ~
  U(-1.25, 1.25)^2 with target [x^2 + y^2 <= 1] (circle_test.cc:19-38);
* cifar: the loop of nnet/cifar_test.cc:423-428 — one Nnet::Train per sample on the 13-layer
  CNN, lr 1e-4.
Per-sample SGD does not shard (step t+1 needs step t's weights): one GPU, replicas only.
Reports samples/s of the device loop (pb_nnet_train_loop, device-timed) beside the
reference's CPU path (oracle: reference-generated kernels or the plain-C port, one core:
the loop is sequential by construction)."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def _timed_loop(torch, pb, L, net, dx, de, n, device):
    sp = C.c_void_p()
    L.call("pb_context_stream", net.ctx, C.byref(sp))
    stream = torch.cuda.ExternalStream(sp.value, device=torch.device("cuda", device))
    net.TrainLoop(dx, de, min(n, 1000))  # warm
    net.Finish()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = C.c_uint64()
    L.call("pb_context_launch_count", net.ctx, C.byref(l0))
    e0.record(stream)
    net.TrainLoop(dx, de, n)
    e1.record(stream)
    net.Finish()
    l1 = C.c_uint64()
    L.call("pb_context_launch_count", net.ctx, C.byref(l1))
    return e0.elapsed_time(e1) * 1e-3, l1.value - l0.value


def _cpu_loop(name, xs, es, lr, m):
    from oracle import pyoracle as po
    R = po.RefNet(name) if po.RefNet.available(name) else po.RestatedNet(po.spec_text(name))
    rng = np.random.default_rng(3)
    w = np.zeros(R.total_w)
    for l in range(R.n_layers):
        if R.nw[l]:
            w[R.w_off[l]:R.w_off[l + 1]] = rng.normal(0, (1.0 / R.ni[l]) ** 0.5, R.nw[l])
    R.train_loop(xs[:min(m, 50)].astype(np.float64).copy(), es[:min(m, 50)].astype(np.float64).copy(), w, lr)
    t0 = time.perf_counter()
    R.train_loop(xs[:m].astype(np.float64).copy(), es[:m].astype(np.float64).copy(), w, lr)
    dt = time.perf_counter() - t0
    return {"value": m / dt, "unit": "samples/s", "cores": 1,
            "kind": "reference" if isinstance(R, po.RefNet) else "port",
            "sample": f"{m} per-sample SGD steps (sequential by construction), double precision"}


def circle(n=200000, device=0, cpu=True):
    import torch
    import plasticity_b200 as pb
    from plasticity_b200 import _lib as L
    rng = np.random.default_rng(0)
    xs = rng.uniform(-1.25, 1.25, (n, 2))
    es = (xs[:, 0] ** 2 + xs[:, 1] ** 2 <= 1.0).astype(np.float64)[:, None]
    model = pb.Architecture(2).AddDenseLayer(10).AddDenseLayer(1)
    net = pb.Nnet(model, pb.Xavier, pb.MeanSquared, seed=1, device=device)
    net.SetLearningParameters(pb.Nnet.LearningParameters(0.3))
    dx, de = net.MakeBuffer(xs.reshape(-1)), net.MakeBuffer(es.reshape(-1))
    dt, launches = _timed_loop(torch, pb, L, net, dx, de, n, device)
    probe = rng.uniform(-1.25, 1.25, (2000, 2))
    # accuracy through a batched evaluate on a fresh network object holding the same weights
    net_b = pb.Nnet(model, pb.NoWeightInit, pb.MeanSquared, max_batch=2000, device=device)
    for l in range(len(model.layers)):
        if net.nw[l]:
            net_b.SetLayerWeights(l, net.GetLayerWeights(l))
    o = net_b.Evaluate(net_b.MakeBuffer(probe.reshape(-1)), batch=2000)
    o.MoveToCpu()
    pred = np.asarray(o.cpu) > 0.5
    acc = float((pred == (probe[:, 0] ** 2 + probe[:, 1] ** 2 <= 1.0)).mean())
    line = {"metric": "circle_mlp_per_sample_sgd_samples_per_sec", "value": n / dt, "unit": "samples/s",
            "us_per_sample": 1e6 * dt / n, "samples": n, "gpu_launches": launches,
            "accuracy_after_training": acc, "scaling": "replicas only (per-sample SGD does not shard)",
            "bound": "latency (41 parameters, ~10 dependent layer stages per sample)",
            "config": {"workload": "circle_test 2-10-1 sigmoid MLP, MSE, lr 0.3, per-sample SGD"}}
    if cpu:
        try:
            line["cpu_baseline"] = _cpu_loop("circle", xs, es, 0.3, min(n, 20000))
        except Exception as e:  # the oracle is optional here
            line["cpu_baseline"] = {"unavailable": str(e)}
    return line


def cifar(n=2000, device=0, cpu=True):
    import torch
    import plasticity_b200 as pb
    from plasticity_b200 import _lib as L
    import bench
    x, e = bench.synthetic_cifar(np, n, 77)
    net = pb.Nnet(bench.cifar_model(pb), pb.Xavier, pb.CrossEntropy, seed=5, device=device)
    net.SetLearningParameters(pb.Nnet.LearningParameters(1e-4))
    dx, de = net.MakeBuffer(x.reshape(-1)), net.MakeBuffer(e.reshape(-1))
    dt, launches = _timed_loop(torch, pb, L, net, dx, de, n, device)
    line = {"metric": "cifar_cnn_per_sample_sgd_samples_per_sec", "value": n / dt, "unit": "samples/s",
            "us_per_sample": 1e6 * dt / n, "samples": n, "gpu_launches": launches,
            "scaling": "replicas only (per-sample SGD does not shard)",
            "bound": "latency (13 layers, 3 dependent phases each per sample)",
            "config": {"workload": "cifar_test per-sample Train loop (nnet/cifar_test.cc:423-428), lr 1e-4"}}
    if cpu:
        try:
            line["cpu_baseline"] = _cpu_loop("cifar", x, e, 1e-4, min(n, 150))
        except Exception as ex:
            line["cpu_baseline"] = {"unavailable": str(ex)}
    return line


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "circle"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else (200000 if which == "circle" else 2000)
    print(json.dumps(circle(n) if which == "circle" else cifar(n)))


if __name__ == "__main__":
    main()
