"""Time one layer kernel of the CIFAR spec at several batch sizes (CUDA events on the library stream)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import plasticity_b200 as pb
from plasticity_b200 import _lib as L
from oracle import pyoracle as po
import util
spec = po.spec_text("cifar")
S = po.RestatedNet(spec)
model, _ = util.architecture_from_spec(spec)
ctx = pb.nnet.context(0)
sp = C.c_void_p(); L.call("pb_context_stream", ctx, C.byref(sp))
stream = torch.cuda.ExternalStream(sp.value, device=torch.device("cuda", 0))
def dev(n, val=0.01):
    p = C.c_void_p(); L.call("pb_buffer_alloc", ctx, n, C.byref(p)); L.call("pb_buffer_fill", ctx, p, val, n); return p
lr = dev(1, 1e-4)
layers = [int(a) for a in sys.argv[1].split(",")]
kinds = sys.argv[2].split(",")
batches = [int(a) for a in sys.argv[3].split(",")]
for l in layers:
    d = model.layers[l].desc
    ni, no, nw = S.ni[l], S.no[l], S.nw[l]
    for B in batches:
        I, O, G, dI, W, out = dev(ni*B), dev(no*B), dev(no*B), dev(ni*B), dev(nw), dev(nw)
        for kind in kinds:
            def fn():
                if kind == "fwd": L.call("pb_layer_evaluate", ctx, C.byref(d), I, W, O, B)
                elif kind == "bwd": L.call("pb_layer_input_delta", ctx, C.byref(d), I, W, G, dI, B)
                else: L.call("pb_layer_weight_update", ctx, C.byref(d), I, W, G, out, lr, L.PB_WEIGHT_DELTAS, B)
            fn(); fn()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(int(os.environ.get("REPS", "10"))): fn()
            e1.record(stream); e1.synchronize()
            print(f"layer {l} {kind} B={B}: {e0.elapsed_time(e1)/int(os.environ.get("REPS", "10"))*1000:.1f} us", flush=True)
        for p in (I, O, G, dI, W, out): L.call("pb_buffer_free", ctx, p)
