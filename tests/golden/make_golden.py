"""Generates tests/golden/*.npz from the reference's own generated kernels
(oracle/_ref/libref_<spec>.so, built by oracle/Makefile from /root/reference).

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
The .npz files are committed; tests on the GPU box (no /root/reference) check
both the plain-C oracle and the CUDA path against them.

Each file holds, for one spec: weights w0, inputs xs[B], targets es[B], and the
reference's results: layer outputs of sample 0, Error of sample 0, input
gradients + weights after 1 and after B per-sample Train steps, weights after
one BatchTrain over all B samples.  fp32-representable inputs, fp64 results.
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle as po  # noqa: E402
import util  # noqa: E402

# savefmt: the reference's conv input-gradient reads GRADIENT out of bounds for
# its 2x2 / pad-0 layer -> no golden from the reference for it.
SPECS = ["circle", "relu3", "simple3", "mazur", "mazur_half", "maxpool443", "conv553", "dense5",
         "conv331", "softmax3", "cifar_grad", "cifar", "mixed"]
B = 4


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name in SPECS:
        R = po.RefNet(name)
        S = po.RestatedNet(po.spec_text(name))
        rng = np.random.default_rng(zlib.crc32(name.encode()) + 17)
        w0 = util.xavier_like(S, rng)
        xs = util.f32(rng.normal(0, 1, (B, R.n_in)))
        if name in ("cifar", "cifar_grad"):
            xs = util.f32(xs / np.linalg.norm(xs, axis=1, keepdims=True))
        es = util.f32(rng.uniform(0, 1, (B, R.n_out)))
        lr = 0.002 if R.n_in > 1000 else 0.05
        lo = R.evaluate(xs[0].copy(), w0.copy())
        err = R.error(R.layer_output(lo, R.n_layers - 1).copy(), es[0].copy())
        w = w0.copy()
        _, ig1 = R.train(xs[0].copy(), es[0].copy(), w, lr, want_input_grads=True)
        w1 = w.copy()
        igs = [ig1]
        for b in range(1, B):
            _, ig = R.train(xs[b].copy(), es[b].copy(), w, lr, want_input_grads=True)
            igs.append(ig)
        wB = w.copy()
        wbatch = w0.copy()
        R.batch_train(xs.copy(), es.copy(), wbatch, lr)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), w0=w0, xs=xs, es=es, lr=lr,
                            layer_out=lo, error=err, input_grads=np.stack(igs), w_after_1=w1,
                            w_after_B=wB, w_after_batch=wbatch)
        print(name, "ok", os.path.getsize(os.path.join(out_dir, name + ".npz")), "bytes")


if __name__ == "__main__":
    main()
