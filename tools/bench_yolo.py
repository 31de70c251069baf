#!/usr/bin/env python
"""BASELINE.json config 5: YOLOv1 (nnet/yolov1.cc:7-231, repaired per SURVEY 8d) forward
inference, 448x448x3 inputs ~ U[0,1), Xavier weights (device-side, seeded), global batch
256 sharded over the GPUs with no collective (weights replicated, 1.07 GB per GPU).

  python tools/bench_yolo.py [--steps K] [--warmup W] [--global-batch 256]
  torchrun --nproc-per-node N tools/bench_yolo.py ...

One JSON line (rank 0): images/s (whole job), ms per pass, algorithmic TFLOP/s (33.07
GFLOP/image), the per-layer times of one pass (CUDA events inside the library) and the
dominant kernel's fraction of one third of the TF32 tensor peak (= half the bf16 figure of
MEASURED_PEAKS.json; every product costs three MMAs) and of the measured FP32 FFMA peak.
"""
import argparse
import ctypes as C
import json
import os
import re
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main(argv=None, embedded=False):
    """embedded=True (called from bench.py inside an initialised process group): returns the
    line (rank 0) instead of printing it and leaves the process group alone."""
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--global-batch", type=int, default=256)
    ap.add_argument("--width", type=int, default=448)
    ap.add_argument("--cpu-baseline", action="store_true", help="time the oracle port on one image")
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
    from plasticity_b200 import yolov1

    B = args.global_batch // world
    ctx = pb.nnet.context(local)
    sp = C.c_void_p()
    L.call("pb_context_stream", ctx, C.byref(sp))
    stream = torch.cuda.ExternalStream(sp.value, device=torch.device("cuda", local))
    yolo = yolov1.Yolov1(max_batch=B, device=local, seed=0, input_width=args.width)
    net = yolo.net_
    n_in = net.ni[0]
    n_out = net.no[-1]
    # inputs ~ U[0,1), generated on the device in chunks (616 MB at batch 256)
    g = torch.Generator(device="cuda")
    g.manual_seed(100 + rank)
    x = torch.rand((B, n_in), device="cuda", dtype=torch.float32, generator=g)
    out = torch.empty((B, n_out), device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()

    def step():
        L.call("pb_nnet_evaluate", net.h, C.c_void_p(x.data_ptr()), B, C.c_void_p(out.data_ptr()))

    def barrier():
        L.call("pb_context_finish", ctx)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        step()
    barrier()
    sampler = bench.ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop(t0, t1) if sampler else None
    result = None
    if rank == 0:
        layers = net.model_.layers
        macs = 0
        for l in layers:
            d = l.desc
            if d.kind == L.PB_CONVOLUTION:
                macs += l.num_outputs * d.filter_width * d.filter_height * d.filter_depth
            elif d.kind == L.PB_DENSE:
                macs += l.num_inputs * l.num_outputs
        flop_img = 2.0 * macs
        step_ms = ms / args.steps
        tf = C.c_double()
        L.call("pb_measure_fp32_tflops", ctx, C.byref(tf))
        buf = C.create_string_buffer(1 << 16)
        L.call("pb_nnet_profile_evaluate", net.h, C.c_void_p(x.data_ptr()), B, buf, len(buf))
        rows = []
        for line in buf.value.decode().strip().split("\n"):
            k, v = line.split("\t")
            v = float(v)
            m = re.match(r"(?:convolution|dense)_layer_(\d+)", k)
            flop = 0.0
            if m:
                l = layers[int(m.group(1))]
                d = l.desc
                per = (d.filter_width * d.filter_height * d.filter_depth
                       if d.kind == L.PB_CONVOLUTION else l.num_inputs)
                flop = 2.0 * l.num_outputs * per * B
            rows.append({"k": k, "ms": round(v, 4), "tflops": round(flop / (v * 1e-3) / 1e12, 1)})
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
        top = max(rows, key=lambda r: r["ms"])
        line = {
            "metric": "yolov1_forward_images_per_sec", "value": world * B * args.steps / (ms * 1e-3),
            "unit": "images/s", "n_gpus": world, "steps": args.steps, "ms_per_step": step_ms,
            "scaling": "strong", "dtype": "f32 (3xTF32 tensor-core products, fp32 accumulate)",
            "data": "synthetic",
            "config": {"workload": "yolov1_forward", "input": [args.width, args.width, 3],
                       "global_batch": B * world, "batch_per_gpu": B, "params": int(sum(net.nw)),
                       "parallelism": f"batch-sharded x{world}, no collective"},
            "clocks": clocks,
            "whole_pass": {"achieved_tflops": flop_img * B * world / (step_ms * 1e-3) / 1e12,
                           "flop_per_image": flop_img, "fp32_ffma_peak_tflops": tf.value * world,
                           "tf32x3_peak_tflops": tf32x3_peak * world},
            "roofline": {"bound": "tensor", "kernel": top["k"], "achieved": top["tflops"],
                         "peak": tf32x3_peak, "unit": "TFLOP/s (fp32-equivalent, 3 MMAs/product)",
                         "frac": top["tflops"] / tf32x3_peak,
                         "vs_fp32_ffma_peak": top["tflops"] / tf.value},
            "layers": rows,
        }
        line["roofline"]["peak_tf32_matmul_measured_div3"] = tf32x3_measured
        if tf32x3_measured:
            line["roofline"]["frac_of_measured_tf32"] = top["tflops"] / tf32x3_measured
        if "yolov1_forward" in recorded:
            cb = dict(recorded["yolov1_forward"])
            cb["cores"] = recorded.get("host_cores")
            cb["note"] = "recorded by tools/cpu_baselines.py on the GPU box's host (not timed in this run)"
            line["cpu_baseline"] = cb
        if args.cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(np)
        result = line
        if not embedded:
            print(json.dumps(line))
    barrier()
    del yolo, net, x, out
    torch.cuda.empty_cache()
    if world > 1 and not embedded:
        dist.destroy_process_group()
    return result if embedded else 0


def cpu_baseline(np):
    """The oracle's plain-C port of the reference arithmetic (fp64, one thread) on one image."""
    from oracle import pyoracle as po
    import plasticity_b200 as pb
    from plasticity_b200 import yolov1
    Y = po.RestatedNet(yolov1.YoloArchitecture().to_spec(pb.MeanSquared))
    rng = np.random.default_rng(0)
    wy = rng.standard_normal(Y.total_w) * 0.01
    x = rng.uniform(0, 1, Y.n_in)
    t0 = time.perf_counter()
    Y.evaluate(x, wy)
    dt = time.perf_counter() - t0
    return {"value": 1.0 / dt, "unit": "images/s", "cores": 1, "kind": "port",
            "sample": "one 448x448x3 image forward (single thread, fp64)"}


if __name__ == "__main__":
    sys.exit(main())
