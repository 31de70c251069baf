"""One library path selected by environment switches, checked against the oracle in its own
process (the switches are read once per process): two fused-plan gradient-sum steps of the
CIFAR network, the chunked variant, a batched Evaluate and a per-sample TrainLoop of the
circle MLP.  Launched by tests/test_gpu_parity.py::test_library_switches."""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import plasticity_b200 as pb
from oracle import pyoracle as po
import util


def main():
    name = "cifar"
    spec = po.spec_text(name)
    S = po.RestatedNet(spec)
    R = po.RefNet(name) if po.RefNet.available(name) else S
    model, loss = util.architecture_from_spec(spec)
    rng = np.random.default_rng(zlib.crc32(b"switches"))
    w0 = util.xavier_like(S, rng)
    B, lr = 64, 0.002
    es = util.f32(rng.uniform(0, 1, (2, B, S.n_out)))
    nets = []
    for max_batch in (B, 40):
        net = pb.Nnet(model, pb.NoWeightInit, loss, max_batch=max_batch)
        net.SetLearningParameters(pb.Nnet.LearningParameters(lr))
        util.load_product_weights(net, S, w0)
        nets.append((max_batch, net))
    w_ref = w0.copy()
    for step in range(2):
        xs = util.draw_with_pool_margin(S, rng, B, w_ref)
        if step == 0:
            out = nets[0][1].Evaluate(nets[0][1].MakeBuffer(xs.reshape(-1)), batch=B)
            out.MoveToCpu()
            want = np.stack([R.layer_output(R.evaluate(x.copy(), w_ref.copy()), S.n_layers - 1) for x in xs])
            util.assert_close(np.asarray(out.cpu).reshape(B, -1), want, "batched Evaluate")
        R.batch_train(xs.copy(), es[step].copy(), w_ref, lr)
        for max_batch, net in nets:
            net.BatchTrain(net.MakeBuffer(xs.reshape(-1)), net.MakeBuffer(es[step].reshape(-1)), B)
            util.assert_close(util.read_product_weights(net, S), w_ref,
                              f"BatchTrain step {step} max_batch={max_batch}")
    # per-sample SGD of a small dense network
    cs = po.spec_text("circle")
    CS = po.RestatedNet(cs)
    cmodel, closs = util.architecture_from_spec(cs)
    cw = util.xavier_like(CS, rng)
    n = 24
    xs = util.f32(rng.normal(0, 1, (n, CS.n_in)))
    ts = util.f32(rng.uniform(0, 1, (n, CS.n_out)))
    cnet = pb.Nnet(cmodel, pb.NoWeightInit, closs)
    cnet.SetLearningParameters(pb.Nnet.LearningParameters(0.05))
    util.load_product_weights(cnet, CS, cw)
    cref = cw.copy()
    CS.train_loop(xs.copy(), ts.copy(), cref, 0.05)
    cnet.TrainLoop(cnet.MakeBuffer(xs.reshape(-1)), cnet.MakeBuffer(ts.reshape(-1)), n)
    util.assert_close(util.read_product_weights(cnet, CS), cref, "circle TrainLoop")
    print("switches ok:", " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("PB_")))


if __name__ == "__main__":
    main()
