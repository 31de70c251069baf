#!/usr/bin/env python
"""YOLOv1 (nnet/yolov1.cc layer list, repaired per SURVEY 8d) gradient-sum TRAINING step,
Yolov1::TrainBatch: forward + input gradients + weight gradients + SGD update of 24
convolutions and 2 dense layers, 3 x 33.07 GFLOP per image.  One GPU, synthetic data.

  python tools/bench_yolo_train.py [--batch 32] [--steps 3] [--width 448]

One JSON line: images/s, ms per step, algorithmic TFLOP/s against the 3xTF32 tensor peak, and
the per-kernel-group times of one step (CUDA events inside the library)."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--width", type=int, default=448)
    args = ap.parse_args(argv)
    import torch
    import plasticity_b200 as pb
    from plasticity_b200 import _lib as L, yolov1
    B = args.batch
    yolo = yolov1.Yolov1(max_batch=B, device=0, seed=0, input_width=args.width)
    yolo.SetParams(learning_rate=1e-6)
    net = yolo.net_
    ctx = net.ctx
    sp = C.c_void_p()
    L.call("pb_context_stream", ctx, C.byref(sp))
    stream = torch.cuda.ExternalStream(sp.value, device=torch.device("cuda", 0))
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    x = torch.rand((B, net.ni[0]), device="cuda", dtype=torch.float32, generator=g)
    e = torch.rand((B, net.no[-1]), device="cuda", dtype=torch.float32, generator=g)
    torch.cuda.synchronize()

    def step():
        L.call("pb_nnet_batch_train", net.h, C.c_void_p(x.data_ptr()), C.c_void_p(e.data_ptr()), B)

    for _ in range(2):
        step()
    L.call("pb_context_finish", ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    L.call("pb_context_finish", ctx)
    ms = e0.elapsed_time(e1) / args.steps
    layers = net.model_.layers
    macs = 0
    for l in layers:
        d = l.desc
        if d.kind == L.PB_CONVOLUTION:
            macs += l.num_outputs * d.filter_width * d.filter_height * d.filter_depth
        elif d.kind == L.PB_DENSE:
            macs += l.num_inputs * l.num_outputs
    flop_img = 3 * 2.0 * macs
    # the input gradient of the first weighted layer has no reader (TrainBatch returns no
    # dE/dInput) and is not computed: not counted either
    first = next(l for l in layers if l.desc.kind in (L.PB_CONVOLUTION, L.PB_DENSE))
    d = first.desc
    flop_img -= 2.0 * (first.num_outputs * d.filter_width * d.filter_height * d.filter_depth
                       if d.kind == L.PB_CONVOLUTION else first.num_inputs * first.num_outputs)
    prof = bench.step_profile(L, net, C.c_void_p(x.data_ptr()), C.c_void_p(e.data_ptr()), B, reps=1)
    rows = sorted(({"k": k, "ms": round(v, 3)} for k, v in prof), key=lambda r: -r["ms"])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops", 1590.0) / 2 / 3
    ach = flop_img * B / (ms * 1e-3) / 1e12
    print(json.dumps({
        "metric": "yolov1_train_images_per_sec", "value": B / (ms * 1e-3), "unit": "images/s",
        "ms_per_step": ms, "n_gpus": 1, "dtype": "f32 (3xTF32 tensor-core products)",
        "config": {"workload": "yolov1_batch_train", "input": [args.width, args.width, 3], "batch": B},
        "whole_step": {"achieved_tflops": ach, "flop_per_image": flop_img, "tf32x3_peak_tflops": peak,
                       "frac": ach / peak},
        "top_kernels": rows[:12], "kernel_groups": len(rows)}))


if __name__ == "__main__":
    main()
