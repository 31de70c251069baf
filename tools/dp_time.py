import ctypes as C, os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, bench, torch
import plasticity_b200 as pb
from plasticity_b200 import _lib as L
B=1024
net=pb.Nnet(bench.cifar_model(pb), pb.Xavier, pb.CrossEntropy, max_batch=B, seed=1)
net.SetLearningParameters(pb.Nnet.LearningParameters(1e-4))
x,e=bench.synthetic_cifar(np,B,3)
bx,be=net.MakeBuffer(x.reshape(-1)),net.MakeBuffer(e.reshape(-1))
bx.MoveToGpu(); be.MoveToGpu()
net._sync_to_device()
def run(fn, n=200):
    for _ in range(10): fn()
    net.Finish()
    t0=time.perf_counter()
    for _ in range(n): fn()
    net.Finish()
    return (time.perf_counter()-t0)/n*1e3
a=run(lambda: L.call("pb_nnet_batch_train", net.h, bx.gpu_buffer(), be.gpu_buffer(), B))
def grads():
    L.call("pb_nnet_batch_gradients", net.h, bx.gpu_buffer(), be.gpu_buffer(), B)
    L.call("pb_nnet_apply_deltas", net.h)
b=run(grads)
print("batch_train %.4f ms   batch_gradients+apply %.4f ms (eager=%s)" % (a, b, os.environ.get("PB_STEP_GRAPHS")))
