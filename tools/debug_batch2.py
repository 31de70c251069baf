import os, sys, zlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import plasticity_b200 as pb
from oracle import pyoracle as po
import util
name = "cifar"
spec = po.spec_text(name)
S = po.RestatedNet(spec)
R = po.RefNet(name)
model, loss = util.architecture_from_spec(spec)
rng = np.random.default_rng(zlib.crc32(name.encode()) + 4242)
w0 = util.xavier_like(S, rng)
B = 64
xs = util.f32(rng.normal(0, 1, (3, B, S.n_in)))
es = util.f32(rng.uniform(0, 1, (3, B, S.n_out)))
lr = 0.002
def report(tag, w, w_ref, w_prev):
    line = [tag]
    for l in range(S.n_layers):
        if not S.nw[l]: continue
        a, b = S.w_off[l], S.w_off[l + 1]
        d, dr = w[a:b] - w_prev[a:b], w_ref[a:b] - w_prev[a:b]
        line.append(f"L{l}: {np.abs(d - dr).max() / np.abs(dr).max():.2e} (max|dr|={np.abs(dr).max():.2e})")
    print("  ".join(line), flush=True)
for sub in range(B):
    pass
net = pb.Nnet(model, pb.NoWeightInit, loss, max_batch=B)
net.SetLearningParameters(pb.Nnet.LearningParameters(lr))
util.load_product_weights(net, S, w0)
w_ref = w0.copy()
R.batch_train(xs[0].copy(), es[0].copy(), w_ref, lr)
net.BatchTrain(net.MakeBuffer(xs[0].reshape(-1)), net.MakeBuffer(es[0].reshape(-1)), B)
w = util.read_product_weights(net, S)
report("step0", w, w_ref, w0)
# which samples are responsible? single-sample batches
for b in range(0, B, 8):
    net = pb.Nnet(model, pb.NoWeightInit, loss, max_batch=8)
    net.SetLearningParameters(pb.Nnet.LearningParameters(lr))
    util.load_product_weights(net, S, w0)
    w_ref = w0.copy()
    R.batch_train(xs[0][b:b+8].copy(), es[0][b:b+8].copy(), w_ref, lr)
    net.BatchTrain(net.MakeBuffer(xs[0][b:b+8].reshape(-1)), net.MakeBuffer(es[0][b:b+8].reshape(-1)), 8)
    report(f"samples {b}-{b+7}", util.read_product_weights(net, S), w_ref, w0)
