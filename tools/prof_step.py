import ctypes as C, os, sys
sys.path.insert(0, "/root/repo")
os.environ["PB_STEP_GRAPHS"]="0"
import numpy as np, bench
import plasticity_b200 as pb
from plasticity_b200 import _lib as L
B=1024
net=pb.Nnet(bench.cifar_model(pb), pb.Xavier, pb.CrossEntropy, max_batch=B, seed=1)
net.SetLearningParameters(pb.Nnet.LearningParameters(1e-4))
x,e=bench.synthetic_cifar(np,B,3)
bx,be=net.MakeBuffer(x.reshape(-1)),net.MakeBuffer(e.reshape(-1))
for _ in range(2): net.BatchTrain(bx,be,B)
net.Finish()
