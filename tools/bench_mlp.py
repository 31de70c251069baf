#!/usr/bin/env python
"""BASELINE.json config 4: wide dense MLP 4096-4096-4096-10, batch 1024 per GPU,
synthetic N(0,1) inputs, one-hot targets, cross-entropy, lr 1e-4,
gradient-sum minibatch training (Nnet::BatchTrain semantics), data-parallel at N GPUs.

  python tools/bench_mlp.py [--steps K] [--warmup W] [--batch B] [--width 4096]
  torchrun --nproc-per-node N tools/bench_mlp.py ...

One JSON line (rank 0): samples/s, ms/step, algorithmic TFLOP/s (201.57 MFLOP/sample:
forward + input gradient + weight gradient of the three dense layers), the per-kernel
times of one step (CUDA events inside the library) and the dominant kernel's fraction of
(a) the FP32 FFMA peak measured in this run and (b) one third of the TF32 tensor peak
(= half the bf16 figure of MEASURED_PEAKS.json; every product costs three MMAs).
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (ClockSampler, step_profile)


def mlp_model(pb, width, classes=10):
    m = pb.Architecture(width)
    # Architecture(4096).AddDenseLayer(4096).AddDenseLayer(4096)
    #   .AddDenseLayer(10, Identity).AddSoftmaxLayer(10) — AddDenseLayer(n) puts the
    # builder's default Sigmoid behind the layer (architecture.h:63-74)
    m.AddDenseLayer(width)
    m.AddDenseLayer(width)
    m.AddDenseLayer(classes, pb.Identity)
    m.AddSoftmaxLayer(classes)
    return m


def main(argv=None, embedded=False):
    """embedded=True (called from bench.py inside an initialised process group): returns the
    line (rank 0) instead of printing it and leaves the process group alone."""
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--width", type=int, default=4096)
    ap.add_argument("--cpu-baseline", action="store_true", help="time the oracle port on a few samples")
    args = ap.parse_args(argv)
    import numpy as np
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not embedded:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import plasticity_b200 as pb
    from plasticity_b200 import _lib as L

    B, Wd = args.batch, args.width
    ctx = pb.nnet.context(local)
    sp = C.c_void_p()
    L.call("pb_context_stream", ctx, C.byref(sp))
    stream = torch.cuda.ExternalStream(sp.value, device=torch.device("cuda", local))
    net = pb.Nnet(mlp_model(pb, Wd), pb.Xavier, pb.CrossEntropy, max_batch=B, device=local, seed=0)
    net.SetLearningParameters(pb.Nnet.LearningParameters(1e-4))

    comm = None
    if world > 1:
        idbuf = torch.zeros(L.PB_COMM_ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            raw = (C.c_char * L.PB_COMM_ID_BYTES)()
            L.call("pb_comm_unique_id", raw)
            idbuf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
        idbuf = idbuf.cuda()
        dist.broadcast(idbuf, 0)
        comm = C.c_void_p()
        L.call("pb_comm_create", ctx, bytes(idbuf.cpu().numpy().tobytes()), rank, world,
               C.byref(comm))

    NB = 4
    rng = np.random.default_rng(rank)
    x = rng.normal(0, 1, (NB * B, Wd)).astype(np.float32)
    e = np.zeros((NB * B, 10), dtype=np.float32)
    e[np.arange(NB * B), rng.integers(0, 10, NB * B)] = 1.0
    px, pe = C.c_void_p(), C.c_void_p()
    L.call("pb_buffer_alloc", ctx, x.size, C.byref(px))
    L.call("pb_buffer_alloc", ctx, e.size, C.byref(pe))
    L.call("pb_buffer_write_f32", ctx, px, x.ctypes.data, x.size)
    L.call("pb_buffer_write_f32", ctx, pe, e.ctypes.data, e.size)

    def step(i):
        off = (i % NB) * B
        xi = C.c_void_p(px.value + off * Wd * 4)
        ei = C.c_void_p(pe.value + off * 10 * 4)
        if comm is None:
            L.call("pb_nnet_batch_train", net.h, xi, ei, B)
        else:
            L.call("pb_nnet_batch_train_dp", net.h, comm, xi, ei, B)

    def barrier():
        L.call("pb_context_finish", ctx)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(max(args.warmup, 3) + 2 * NB):  # every batch twice: CUDA-graph capture
        step(i)
    barrier()
    sampler = bench.ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop(t0, t1) if sampler else None
    line = None
    if rank == 0:
        macs = Wd * Wd * 2 + Wd * 10
        flop_sample = 3 * 2.0 * macs
        step_ms = ms / args.steps
        tf = C.c_double()
        L.call("pb_measure_fp32_tflops", ctx, C.byref(tf))
        prof = bench.step_profile(L, net, C.c_void_p(px.value), C.c_void_p(pe.value), B, reps=3)
        rows = [{"k": k, "ms": round(v, 4), "tflops": round(bench.layer_flop(net, L, k, B) / (v * 1e-3) / 1e12, 2)}
                for k, v in prof]
        rows.sort(key=lambda r: -r["ms"])
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tf32x3_peak = peaks.get("bf16_tflops", 1590.0) / 2 / 3
        recorded = {}
        try:  # tools/cpu_baselines.py on the GPU box: TF32 matmul peak + CPU rates of this config
            recorded = json.load(open(os.path.join(ROOT, "profiles", "r01_cpu_baselines.json")))
        except Exception:
            pass
        tf32x3_measured = recorded.get("tf32_tflops_burst", 0.0) / 3 or None
        top = rows[0]
        line = {
            "metric": "wide_mlp_train_samples_per_sec", "value": world * B * args.steps / (ms * 1e-3),
            "unit": "samples/s", "n_gpus": world, "steps": args.steps, "ms_per_step": step_ms,
            "scaling": "weak", "dtype": "f32 (3xTF32 tensor-core products, fp32 accumulate)",
            "data": "synthetic",
            "config": {"workload": "wide_mlp_batch_train", "widths": [Wd, Wd, Wd, 10],
                       "batch_per_gpu": B, "lr": 1e-4, "loss": "cross_entropy",
                       "parallelism": f"dp{world}"},
            "clocks": clocks,
            "whole_step": {"achieved_tflops": flop_sample * B * world / (step_ms * 1e-3) / 1e12,
                           "flop_per_sample": flop_sample, "fp32_ffma_peak_tflops": tf.value * world,
                           "tf32x3_peak_tflops": tf32x3_peak * world},
            "roofline": {"bound": "tensor", "kernel": top["k"], "achieved": top["tflops"],
                         "peak": tf32x3_peak, "unit": "TFLOP/s (fp32-equivalent, 3 MMAs/product)",
                         "frac": top["tflops"] / tf32x3_peak,
                         "vs_fp32_ffma_peak": top["tflops"] / tf.value},
            "step_kernels": rows,
        }
        line["roofline"]["peak_tf32_matmul_measured_div3"] = tf32x3_measured
        if tf32x3_measured:
            line["roofline"]["frac_of_measured_tf32"] = top["tflops"] / tf32x3_measured
        if "wide_mlp_train" in recorded:
            cb = dict(recorded["wide_mlp_train"])
            cb["cores"] = recorded.get("host_cores")
            cb["note"] = "recorded by tools/cpu_baselines.py on the GPU box's host (not timed in this run)"
            line["cpu_baseline"] = cb
        if args.cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(np, Wd)
        if not embedded:
            print(json.dumps(line))
    barrier()
    if comm is not None:
        L.call("pb_comm_destroy", comm)
    for p in (px, pe):
        L.call("pb_buffer_free", ctx, p)
    del net
    if world > 1 and not embedded:
        dist.destroy_process_group()
    return (line if rank == 0 else None) if embedded else 0


def cpu_baseline(np, Wd, nb=4):
    """The oracle's plain-C port of the reference arithmetic (fp64, OpenMP over examples) on
    one gradient-sum step over nb samples: a bounded sample, the rate is per sample."""
    from oracle import pyoracle as po
    spec = (f"loss ce\nact {Wd} identity\ndense {Wd} {Wd}\nact {Wd} sigmoid\ndense {Wd} {Wd}\n"
            f"act {Wd} sigmoid\ndense {Wd} 10\nact 10 identity\nsoftmax 10\n")
    S = po.RestatedNet(spec)
    rng = np.random.default_rng(0)
    w = rng.normal(0, 1 / np.sqrt(Wd), S.total_w)
    xs = rng.normal(0, 1, (nb, Wd))
    es = np.zeros((nb, 10))
    es[np.arange(nb), rng.integers(0, 10, nb)] = 1.0
    t0 = time.perf_counter()
    S.batch_train_mt(xs, es, w, 1e-4)
    dt = time.perf_counter() - t0
    return {"value": nb / dt, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"one gradient-sum step over {nb} samples (OpenMP over examples, fp64)"}


if __name__ == "__main__":
    sys.exit(main())
