"""Debug: BatchTrain deltas per layer vs the oracle for several batch sizes and kernel paths."""
import os, sys, zlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import plasticity_b200 as pb
from oracle import pyoracle as po
import util

name = sys.argv[1] if len(sys.argv) > 1 else "cifar"
spec = po.spec_text(name)
S = po.RestatedNet(spec)
R = po.RefNet(name) if po.RefNet.available(name) else S
model, loss = util.architecture_from_spec(spec)
rng = np.random.default_rng(zlib.crc32(name.encode()) + 4242)
w0 = util.xavier_like(S, rng)
lr = 0.002
for B in [int(a) for a in sys.argv[2:]] or [7, 31, 32, 64]:
    xs = util.f32(rng.normal(0, 1, (B, S.n_in)))
    es = util.f32(rng.uniform(0, 1, (B, S.n_out)))
    net = pb.Nnet(model, pb.NoWeightInit, loss, max_batch=B)
    net.SetLearningParameters(pb.Nnet.LearningParameters(lr))
    util.load_product_weights(net, S, w0)
    w_ref = w0.copy()
    R.batch_train(xs.copy(), es.copy(), w_ref, lr)
    net.BatchTrain(net.MakeBuffer(xs.reshape(-1)), net.MakeBuffer(es.reshape(-1)), B)
    w = util.read_product_weights(net, S)
    line = [f"B={B}"]
    for l in range(S.n_layers):
        if not S.nw[l]:
            continue
        a, b = S.w_off[l], S.w_off[l + 1]
        d, dr = w[a:b] - w0[a:b], w_ref[a:b] - w0[a:b]
        line.append(f"L{l}: max|d-dr|/max|dr|={np.abs(d - dr).max() / np.abs(dr).max():.2e} (max|dr|={np.abs(dr).max():.2e})")
    print("  ".join(line), flush=True)
