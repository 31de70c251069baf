#!/usr/bin/env python
"""bench.py — CIFAR CNN training throughput of the B200 layer path.

Metric (BASELINE.json): CIFAR CNN train samples/sec, beside the reference's
host-CPU path, with the dominant kernel's fraction of the 3xTF32 tensor roofline.

Workload ("cifar_cnn_batch_train"): the 13-layer network of
nnet/cifar_test.cc:293-350 (22 466 parameters), cross-entropy, lr 1e-4,
synthetic images (pixels ~ U{-128..127}, L2-normalised per image like
cifar_test.cc:66-81, one-hot labels), trained in the gradient-sum minibatch
mode (Nnet::BatchTrain semantics, W <- W - lr * sum_b grad_b(W_old)) with
--batch samples per GPU per step: the only training mode of the path that
shards.  N > 1: one process per GPU, weak scaling, one all-reduce of the
22 466-float delta buffer per step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B]
  python bench.py --impl reference ...     # the reference's CPU path (oracle/_ref)

One JSON line on stdout (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cifar_cnn_train_samples_per_sec"
UNIT = "samples/s"
# fwd + input-grad + weight-grad MACs of the CIFAR net, x2 (23.52e6), minus the
# input gradient of the first convolution (2 x 1.2288e6 MAC): BatchTrain has no reader for
# dE/dInput (nnet.h:617-622 passes a null input_gradients) and the step does not compute it.
TRAIN_FLOP_PER_SAMPLE = 23.52e6 - 2 * 1.2288e6
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 74.4


def cifar_model(pb):
    m = pb.Architecture(32 * 32 * 3)
    m.AddConvolutionLayer(pb.VolumeDimensions(32, 32, 3), pb.FilterParams(5, 5, 3, 1, 2, 16))
    m.AddMaxPoolLayer(pb.VolumeDimensions(32, 32, 16), pb.AreaDimensions(16, 16))
    m.AddConvolutionLayer(pb.VolumeDimensions(16, 16, 16), pb.FilterParams(5, 5, 16, 1, 2, 20))
    m.AddMaxPoolLayer(pb.VolumeDimensions(16, 16, 20), pb.AreaDimensions(8, 8))
    m.AddConvolutionLayer(pb.VolumeDimensions(8, 8, 20), pb.FilterParams(5, 5, 20, 1, 2, 20))
    m.AddMaxPoolLayer(pb.VolumeDimensions(8, 8, 20), pb.AreaDimensions(4, 4))
    m.AddDenseLayer(10, pb.Identity)
    m.AddSoftmaxLayer(10)
    return m


def synthetic_cifar(np, n, seed, records=False):
    rng = np.random.default_rng(seed)
    pix = rng.integers(-128, 128, size=(n, 3072))  # signed char pixels
    px = pix.astype(np.float64)
    norm = np.sqrt((px * px).sum(axis=1, keepdims=True))
    norm[norm == 0] = 1.0
    x = (px / norm).astype(np.float32)
    lab = rng.integers(0, 10, size=n)
    e = np.zeros((n, 10), dtype=np.float32)
    e[np.arange(n), lab] = 1.0
    if records:
        # the same samples as CIFAR-10 binary records: 1 label byte + 3072 pixel bytes
        rec = np.concatenate([lab.astype(np.uint8)[:, None], pix.astype(np.int8).view(np.uint8)],
                             axis=1)
        return x, e, np.ascontiguousarray(rec)
    return x, e


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for r in self.rows if t0 <= r[0] <= t1 + 0.2] or self.rows[-3:]
        for _, line in rows:
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except Exception:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx,
                "samples": len(sm), "reasons": sorted(reasons)}


def reference_rate(np, steps, warmup, sample, threads_note=True):
    """Time the reference's own CPU implementation of the workload: the
    kernels emitted by the reference's generator, host-compiled (oracle/_ref),
    or the plain-C port when that build is absent.  Gradient-sum minibatch,
    examples spread over all host threads."""
    os.environ.setdefault("OMP_WAIT_POLICY", "passive")
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm uses every host core
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    try:  # the OpenMP runtime may already be loaded (it reads the variable once)
        C.CDLL("libgomp.so.1").omp_set_num_threads(cores)
    except OSError:
        pass
    from oracle import pyoracle as po
    if po.RefNet.available("cifar"):
        net, kind = po.RefNet("cifar"), "reference"
    else:
        net, kind = po.RestatedNet(po.spec_text("cifar")), "port"
    x, e = synthetic_cifar(np, sample, 1234)
    x, e = x.astype(np.float64), e.astype(np.float64)
    rng = np.random.default_rng(5)
    w = np.zeros(net.total_w)
    for l in range(net.n_layers):  # Xavier: N(0, 1/num_inputs), nnet/layer.cc:86-100
        if net.nw[l]:
            w[net.w_off[l]:net.w_off[l + 1]] = rng.normal(0, (1.0 / net.ni[l]) ** 0.5, net.nw[l])
    for _ in range(warmup):
        net.batch_train_mt(x, e, w, 1e-4)
    t0 = time.perf_counter()
    for _ in range(steps):
        net.batch_train_mt(x, e, w, 1e-4)
    dt = time.perf_counter() - t0
    return {"value": steps * sample / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{steps} steps x {sample} samples of the same workload "
                      f"(gradient-sum minibatch, OpenMP over examples, double precision)",
            "seconds": dt}


def run_reference(args):
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample = args.ref_sample
    r = reference_rate(np, args.steps, args.warmup, sample)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * r["seconds"] / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cifar_cnn_batch_train", "mode": "gradient_sum_minibatch",
                   "batch_per_gpu": sample, "global_batch": sample, "lr": 1e-4,
                   "loss": "cross_entropy",
                   "network": "nnet/cifar_test.cc:293-350 (13 layers, 22466 params)",
                   "note": "CPU arm: one host, every step = one gradient-sum minibatch of "
                           "batch_per_gpu samples whatever --gpus says"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def step_profile(L, net, xi, ei, batch, reps=5):
    """Per-kernel-group times of the step as it really runs (fused plan), from CUDA
    events recorded inside the library (pb_nnet_profile_step); averaged over reps."""
    buf = C.create_string_buffer(1 << 16)
    acc, order = {}, []
    for _ in range(reps + 1):
        L.call("pb_nnet_profile_step", net.h, xi, ei, batch, buf, len(buf))
        rows = [l.split("\t") for l in buf.value.decode().strip().split("\n") if l]
        if _ == 0:
            continue  # warm
        for label, ms in rows:
            if label not in acc:
                acc[label] = 0.0
                order.append(label)
            acc[label] += float(ms) / reps
    return [(k, acc[k]) for k in order]


def layer_flop(net, L, label, batch):
    """Algorithmic FLOPs (2 x MAC) of one kernel group, from the layer descriptor."""
    import re
    m = re.match(r"(?:convolution|dense)_layer_(\d+)", label)
    if not m or label.startswith("dense_layer") and ".." in label:
        idx = re.match(r"dense_layer_(\d+)\.\.", label)
        if not idx:
            return 0.0
        l = net.model_.layers[int(idx.group(1))]
        return 2.0 * 2 * l.num_inputs * l.num_outputs * batch  # forward + input gradient
    l = net.model_.layers[int(m.group(1))]
    d = l.desc
    if d.kind == L.PB_CONVOLUTION:
        macs = l.num_outputs * d.filter_width * d.filter_height * d.filter_depth
    else:
        macs = l.num_inputs * l.num_outputs
    return 2.0 * macs * batch


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import plasticity_b200 as pb
    from plasticity_b200 import _lib as L

    B = args.batch
    ctx = pb.nnet.context(local)
    sp = C.c_void_p()
    L.call("pb_context_stream", ctx, C.byref(sp))
    stream = torch.cuda.ExternalStream(sp.value, device=torch.device("cuda", local))

    model = cifar_model(pb)
    net = pb.Nnet(model, pb.Xavier, pb.CrossEntropy, max_batch=B, device=local, seed=42)
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
        raw = bytes(idbuf.cpu().numpy().tobytes())
        comm = C.c_void_p()
        L.call("pb_comm_create", ctx, raw, rank, world, C.byref(comm))
        # Same starting weights on every rank (replicas): rank 0's Xavier draw.
        tw = C.c_int64()
        L.call("pb_nnet_total_weights", net.h, C.byref(tw))
        wp = C.c_void_p()
        L.call("pb_nnet_weights_device", net.h, C.byref(wp))
        if rank != 0:
            L.call("pb_buffer_fill", ctx, wp, 0.0, tw.value)
        L.call("pb_comm_allreduce_sum", comm, wp, tw.value)

    # Resident data: NB distinct batches, rotated so consecutive steps never
    # re-read the same inputs; a step's working set (~54k activations x4 B x B
    # samples forward + the same backward) is far larger than the 126 MB L2.
    NB = 8
    x, e, rec = synthetic_cifar(np, NB * B, 1000 + rank, records=True)
    px, pe = C.c_void_p(), C.c_void_p()
    L.call("pb_buffer_alloc", ctx, x.size, C.byref(px))
    L.call("pb_buffer_alloc", ctx, e.size, C.byref(pe))
    L.call("pb_buffer_write_f32", ctx, px, x.ctypes.data, x.size)
    L.call("pb_buffer_write_f32", ctx, pe, e.ctypes.data, e.size)

    def step_resident(i):
        off = (i % NB) * B
        xi = C.c_void_p(px.value + off * 3072 * 4)
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

    def launches():
        n = C.c_uint64()
        L.call("pb_context_launch_count", ctx, C.byref(n))
        return n.value

    def maxreduce(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_blocks(step_fn, first_index, wall=False):
        """The K-step block, bracketed by barrier + synchronize on both sides and timed with
        CUDA events on the library stream (max over ranks), repeated until >= --min-seconds
        of timed work: a 20-step block of this workload is ~6 ms, too short to be a
        measurement on its own.  Returns the per-block times (ms)."""
        out, idx = [], first_index
        total = 0.0
        while not out or (total < args.min_seconds * 1e3 and len(out) < args.max_blocks):
            barrier()
            ea = torch.cuda.Event(enable_timing=True)
            eb = torch.cuda.Event(enable_timing=True)
            ea.record(stream)
            t0 = time.perf_counter()
            for i in range(args.steps):
                step_fn(idx + i)
            eb.record(stream)
            barrier()
            wall_ms = (time.perf_counter() - t0) * 1e3
            dev_ms = ea.elapsed_time(eb)
            blk = maxreduce(max(dev_ms, wall_ms) if wall else dev_ms)
            # every rank takes the same decision: the reduced block time drives the loop
            out.append(blk)
            total += blk
            idx += args.steps
        return out

    def median(v):
        v = sorted(v)
        return v[len(v) // 2] if len(v) % 2 else 0.5 * (v[len(v) // 2 - 1] + v[len(v) // 2])

    # ---------------- device-resident timing ------------------------------
    # Set-up, not warm-up: the library records the step for a given (inputs, targets)
    # pair into a CUDA graph on its second use; touch every resident batch twice so the
    # W warm-up steps and the K timed steps below are all steady-state replays.
    for i in range(2 * NB):
        step_resident(i)
    for i in range(args.warmup):
        step_resident(i)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = launches()
    t_wall0 = time.time()
    blocks = timed_blocks(step_resident, args.warmup)
    t_wall1 = time.time()
    ms = median(blocks)
    n_launch = (launches() - l0) // len(blocks)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    value = world * B * args.steps / (ms * 1e-3)

    # ---------------- end to end through the host API ------------------------
    # Every step: H2D of that step's samples from pinned host memory — as the data set
    # stores them, CIFAR-10 binary records of 1 label byte + 3072 pixel bytes
    # (cifar_test.cc:191-205) — decoding on the device (signed-byte L2 normalisation +
    # one-hot, cifar_test.cc:66-97), BatchTrain, D2H of the step's network outputs
    # (batch x 10 class scores).
    hr, ho = C.c_void_p(), C.c_void_p()
    L.call("pb_host_alloc", rec.nbytes, C.byref(hr))
    L.call("pb_host_alloc", 2 * B * 10 * 4, C.byref(ho))
    C.memmove(hr, rec.ctypes.data, rec.nbytes)

    def step_e2e(i):
        # The user-facing call for host-resident data: pinned host inputs in, the
        # step's network outputs (batch x 10 scores) back to pinned host memory.
        # Asynchronous: the H2D of step i+1 overlaps the compute of step i (two
        # device staging slots inside the library); every step still moves its own
        # inputs and result inside the timed region.
        off = (i % NB) * B
        net.BatchTrainHostRecords(C.c_void_p(hr.value + off * 3073), B,
                                  C.c_void_p(ho.value + (i % 2) * B * 10 * 4), comm=comm)

    def time_e2e(step_fn):
        for i in range(max(min(args.warmup, 3), 4)):  # both staging slots twice (graph capture)
            step_fn(i)
        return median(timed_blocks(step_fn, 0, wall=True))

    ms_e2e = time_e2e(step_e2e)

    # The same end-to-end step with the samples already normalised to fp32 on the host
    # (12 328 B per image over PCIe instead of 3 073): reported beside the headline.
    hx, he = C.c_void_p(), C.c_void_p()
    L.call("pb_host_alloc", x.nbytes, C.byref(hx))
    L.call("pb_host_alloc", e.nbytes, C.byref(he))
    C.memmove(hx, x.ctypes.data, x.nbytes)
    C.memmove(he, e.ctypes.data, e.nbytes)

    def step_e2e_fp32(i):
        off = (i % NB) * B
        net.BatchTrainHost(C.c_void_p(hx.value + off * 3072 * 4),
                           C.c_void_p(he.value + off * 10 * 4), B,
                           C.c_void_p(ho.value + (i % 2) * B * 10 * 4), comm=comm)

    ms_e2e_fp32 = time_e2e(step_e2e_fp32)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)

    # ---------------- data-parallel parity (N > 1), after the timed regions ---------------
    dp_parity = None
    if world > 1:
        tw = C.c_int64()
        L.call("pb_nnet_total_weights", net.h, C.byref(tw))
        wp = C.c_void_p()
        L.call("pb_nnet_weights_device", net.h, C.byref(wp))

        def read_w():
            h = np.zeros(tw.value, dtype=np.float32)
            L.call("pb_buffer_read_f32", ctx, wp, h.ctypes.data, tw.value)
            return h

        barrier()
        w_start = read_w()                       # identical on every rank (checked below)
        w_dp = []
        for i in range(2):
            step_resident(i)                      # data-parallel steps on batches 0 and 1
            barrier()
            w_dp.append(read_w())
        mine = torch.from_numpy(np.concatenate([w_start] + w_dp)).cuda()
        allw = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allw, mine)
        identical = all(torch.equal(allw[0], a) for a in allw)
        # the same two steps on ONE GPU over the global batch (rank 0)
        xg = [torch.empty(2 * B * 3072, dtype=torch.float32, device="cuda") for _ in range(world)]
        eg = [torch.empty(2 * B * 10, dtype=torch.float32, device="cuda") for _ in range(world)]
        dist.all_gather(xg, torch.from_numpy(x[:2 * B].reshape(-1)).cuda())
        dist.all_gather(eg, torch.from_numpy(e[:2 * B].reshape(-1)).cuda())
        if rank == 0:
            one = pb.Nnet(model, pb.NoWeightInit, pb.CrossEntropy, max_batch=world * B, device=local)
            one.SetLearningParameters(pb.Nnet.LearningParameters(1e-4))
            one._sync_to_device()
            wp1 = C.c_void_p()
            L.call("pb_nnet_weights_device", one.h, C.byref(wp1))
            L.call("pb_buffer_write_f32", ctx, wp1, w_start.ctypes.data, tw.value)
            rel = []
            for t in range(2):
                gx = torch.cat([a.view(2, B * 3072)[t] for a in xg]).contiguous()
                ge = torch.cat([a.view(2, B * 10)[t] for a in eg]).contiguous()
                torch.cuda.synchronize()
                L.call("pb_nnet_batch_train", one.h, C.c_void_p(gx.data_ptr()),
                       C.c_void_p(ge.data_ptr()), world * B)
                L.call("pb_context_finish", ctx)
                w_one = np.zeros(tw.value, dtype=np.float32)
                L.call("pb_buffer_read_f32", ctx, wp1, w_one.ctypes.data, tw.value)
                diff = float(np.abs(w_dp[t].astype(np.float64) - w_one).max())
                upd = float(np.abs(w_one.astype(np.float64) - w_start).max())
                rel.append((diff / max(float(np.abs(w_one).max()), 1e-30), diff / max(upd, 1e-30)))
            dp_parity = {"ranks_identical": bool(identical),
                         "max_rel_vs_single_gpu": rel[0][0],
                         "max_diff_over_max_update": rel[0][1],
                         "after_second_step": {"max_rel_vs_single_gpu": rel[1][0],
                                               "max_diff_over_max_update": rel[1][1]},
                         "how": "one data-parallel step vs the same step on ONE GPU over the global "
                                "batch, from the same weights: max|w_dp - w_1| / max|w_1| (and relative "
                                "to the largest weight change of the step) - summation order only; the "
                                "second consecutive step also carries whatever max-pool windows the "
                                "1e-7 weight difference re-routes"}
            del one
        barrier()

    # ---------------- the other BASELINE.json configurations ---------------------------------
    other = None
    if not args.no_other_configs:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        other = {}

        def attempt(name, fn):
            try:
                r = fn()
                if rank == 0 and r is not None:
                    other[name] = r
            except Exception as ex:  # a side measurement must not take the headline down
                if rank == 0:
                    other[name] = {"error": f"{type(ex).__name__}: {ex}"[:300]}

        import bench_mlp
        import bench_yolo
        import bench_circle
        cpu_flag = ["--cpu-baseline"] if (world == 1 and not args.no_cpu_baseline) else []
        attempt("wide_mlp_train", lambda: bench_mlp.main(["--steps", "10", "--warmup", "3"] + cpu_flag,
                                                         embedded=True))
        barrier()
        attempt("yolov1_forward", lambda: bench_yolo.main(["--steps", "3", "--warmup", "2"] + cpu_flag,
                                                          embedded=True))
        barrier()
        if rank == 0:  # per-sample SGD does not shard: one GPU, replicas only
            want_cpu = world == 1 and not args.no_cpu_baseline
            attempt("circle_per_sample", lambda: bench_circle.circle(200000, local, want_cpu))
            attempt("cifar_per_sample", lambda: bench_circle.cifar(2000, local, want_cpu))
        barrier()

    # ---------------- per-kernel table + roofline (rank 0) ---------------------
    line = None
    if rank == 0:
        tf = C.c_double()
        L.call("pb_measure_fp32_tflops", ctx, C.byref(tf))
        xi0, ei0 = C.c_void_p(px.value), C.c_void_p(pe.value)
        prof = step_profile(L, net, xi0, ei0, B)
        rows = [{"kernel": k, "ms": ms, "flop": layer_flop(net, L, k, B)} for k, ms in prof]
        rows.sort(key=lambda r: -r["ms"])
        step_ms = ms / args.steps
        total_prof = sum(r["ms"] for r in rows)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        bf16_peak = peaks.get("bf16_tflops", 1590.0)
        top = rows[0]

        umma_off = os.environ.get("PB_DISABLE_UMMA") == "1"
        wgrad_umma = not umma_off and os.environ.get("PB_WGRAD_UMMA") != "0"

        def roofline_of(r):
            # which pipe the kernel group runs on: tcgen05 (fused conv forward, conv
            # input-gradient, conv weight gradient), FP32 FFMA (dense layers), or none
            # (elementwise: HBM)
            k = r["kernel"]
            is_conv = k.startswith("convolution")
            wgrad = is_conv and k.endswith("weight_update")
            first_fused = k == "convolution_layer_1:activation+pool backward+weight_update"
            wrows = first_fused and os.environ.get("PB_WGRAD_ROWS") != "0"
            first_fwd = k.startswith("convolution_layer_1:evaluate") and \
                os.environ.get("PB_CONV_ROWS") != "0"
            tcgen05 = not umma_off and ("tcgen05" in k or (is_conv and k.endswith(":input_delta")) or
                                        (wgrad and not first_fused and wgrad_umma) or wrows)
            mma_sync = wgrad and not tcgen05 and not umma_off and os.environ.get("PB_WGRAD_MMA") != "0"
            tensor_peak = bf16_peak / 2 / 3
            if wrows:
                pipe = ("tcgen05.mma kind::tf32, bands of image rows in tensor memory, the gradient "
                        "formed from the pooled gradient and staged K-major in shared memory "
                        "(conv_wgrad_rows_umma.cu); the time includes the fixed-order partial reduce + "
                        "SGD update kernel")
            elif mma_sync:
                pipe = ("mma.sync.m16n8k8 tf32 (warp-level; measured peak 278 TFLOP/s of tf32 MMA "
                        "work = 92.7 fp32-equivalent, tools/mma_sync_probe.cu)")
            elif wgrad:
                pipe = ("tcgen05.mma kind::tf32, gradient operand in tensor memory, image operand "
                        "K-major in shared memory (conv_wgrad_umma.cu); the time includes the "
                        "fixed-order partial reduce + SGD update kernel")
            elif first_fwd:
                pipe = ("tcgen05.mma kind::tf32, image rows in tensor memory, fy taps summed over "
                        "tiles in registers (conv_rows_umma.cu)")
            else:
                pipe = ("tcgen05.mma kind::tf32, filters in tensor memory, image operand K-major in "
                        "shared memory, fx taps folded by lane shuffles (conv_wt_umma.cu)")
            if r["flop"] > 0 and not (tcgen05 or mma_sync):
                out = {"bound": "fp32", "peak": tf.value, "unit": "TFLOP/s",
                       "peak_source": "FP32 FFMA micro-benchmark in this run (pb_measure_fp32_tflops; "
                                      f"nominal {NOMINAL_FP32_TFLOPS:.1f}); MEASURED_PEAKS.json has no "
                                      "FP32 entry"}
            elif r["flop"] > 0:
                out = {"bound": "tensor", "peak": tensor_peak, "unit": "TFLOP/s",
                       "peak_source": ("MEASURED_PEAKS.json bf16_tflops (burst)" if peaks else
                                       "fallback 1590") + " / 2 (TF32 runs at half the bf16 rate) / 3 "
                                       "(3xTF32: three MMAs per fp32 product); achieved counts "
                                       "algorithmic fp32 FLOPs",
                       "pipe": pipe}
            else:
                out = {"bound": "hbm", "peak": hbm_peak, "unit": "GB/s",
                       "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"}
            out["kernel"] = r["kernel"]
            out["achieved"] = r["flop"] / (r["ms"] * 1e-3) / 1e12 if r["flop"] > 0 else None
            out["frac"] = out["achieved"] / out["peak"] if out["achieved"] else None
            if out["achieved"]:
                out["vs_fp32_ffma_peak"] = out["achieved"] / tf.value
            out["ms"] = r["ms"]
            out["share_of_step"] = r["ms"] / total_prof
            out["traffic"] = traffic.get(r["kernel"])
            return out

        roof = roofline_of(top)
        top_tc = next((r for r in rows if "tcgen05" in r["kernel"]), None)
        roof_tensor = roofline_of(top_tc) if top_tc else None
        whole = {"achieved_tflops": TRAIN_FLOP_PER_SAMPLE * B * world / (step_ms * 1e-3) / 1e12,
                 "fp32_peak_tflops_measured": tf.value * world,
                 "flop_per_sample": TRAIN_FLOP_PER_SAMPLE}
        whole["frac"] = whole["achieved_tflops"] / whole["fp32_peak_tflops_measured"]
        # every convolution kernel of the step runs on the tensor pipe: the same work against
        # the 3xTF32 peak (bf16 / 2 / 3), the denominator of `roofline`
        whole["frac_of_3xtf32_tensor_peak"] = whole["achieved_tflops"] / (bf16_peak / 2 / 3 * world)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = reference_rate(np, 2, 1, args.ref_sample)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "timing": {"how": "median over blocks of `steps` steps; every block is bracketed by a "
                              "barrier + synchronize on both sides, timed with CUDA events on the "
                              "library stream, max over ranks",
                       "blocks": len(blocks), "block_ms_min": min(blocks), "block_ms_max": max(blocks),
                       "timed_seconds": sum(blocks) * 1e-3},
            "config": {"workload": "cifar_cnn_batch_train", "mode": "gradient_sum_minibatch",
                       "input_gradients": "not requested (BatchTrain's 3-argument overload, "
                                          "nnet.h:617-622): dE/dInput of the first convolution has "
                                          "no reader and is not computed",
                       "batch_per_gpu": B, "global_batch": B * world, "lr": 1e-4,
                       "loss": "cross_entropy",
                       "network": "nnet/cifar_test.cc:293-350 (13 layers, 22466 params)",
                       "parallelism": f"dp{world}",
                       "l2": f"inputs rotate over {NB} resident batches; per-step working set "
                             f"~{2 * 54366 * 4 * B / 1e6:.0f} MB > 126 MB L2"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": B * 3073,
                    "fp32_host_inputs": {"value": world * B * args.steps / (ms_e2e_fp32 * 1e-3),
                                         "h2d_bytes_per_step": (B * 3072 + B * 10) * 4},
                    "d2h_bytes_per_step": B * 10 * 4,
                    "input_format": "CIFAR-10 binary records (1 label byte + 3072 signed pixel "
                                    "bytes per image, nnet/cifar_test.cc:191-205) in pinned host "
                                    "memory; L2 normalisation + one-hot on the device "
                                    "(pb_nnet_batch_train_host_records), bit-identical to the "
                                    "host-side double normalisation of cifar_test.cc:66-81"},
            "gpu_launches": n_launch,
            "roofline": roof,
            "roofline_tensor": roof_tensor,
            "whole_step": whole,
            "cpu_baseline": cpu,
            "dp_parity": dp_parity,
            "other_configs": other,
            "step_kernels": [{"k": r["kernel"], "ms": round(r["ms"], 4)} for r in rows],
        }
        if args.table:
            with open(args.table, "w") as f:
                json.dump({"batch": B, "step_ms": step_ms, "fp32_peak_tflops": tf.value,
                           "rows": rows}, f, indent=1)
    barrier()
    if comm is not None:
        L.call("pb_comm_destroy", comm)
    if world > 1:
        dist.destroy_process_group()
    if line is not None:
        emit(line)
    return 0


def main():
    # Libraries (NCCL's version banner, torchrun notices) print to fd 1; the contract is
    # ONE JSON line on stdout, so everything else is sent to stderr and the line is
    # written to the real stdout at the end.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rc, line = _main()
    if line is not None:
        real_stdout.write(line + "\n")
        real_stdout.flush()
    return rc


def _main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1024, help="samples per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-sample", type=int, default=1024,
                    help="samples per step of the CPU arm (bounded sample of the workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--min-seconds", type=float, default=0.5,
                    help="repeat the K-step block until this much time has been measured")
    ap.add_argument("--max-blocks", type=int, default=400)
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the other BASELINE.json configurations (other_configs)")
    ap.add_argument("--table", default=None, help="write the per-kernel table (JSON) here")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        rc = run_reference(args)
    else:
        rc = run_ours(args)
    return rc, _LINE[0]


_LINE = [None]


def emit(line):
    _LINE[0] = json.dumps(line)


if __name__ == "__main__":
    sys.exit(main())
