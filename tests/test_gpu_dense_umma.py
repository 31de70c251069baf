"""GPU parity of the 3xTF32 tcgen05 dense GEMM (csrc/dense_umma.cu) — the batched
DenseLayer products of the wide MLP (BASELINE.json config 4) — through the C ABI
(pb_layer_evaluate / pb_layer_input_delta / pb_layer_weight_update), against the
CPU oracle (plain-C restatement of nnet/dense_layer.cc:12-55) on sampled rows and
against its fp64 matrix form on every element.

Shapes are chosen to hit every operand path of the kernel: 16-byte aligned rows
(float4 loads), odd-pitch weight rows (scalar loads), transposed operands, M / N / K
tails, split-K with its fixed-order reduce, and several persistent tiles per CTA.
Tolerance: |rel| <= 1e-4 (north_star), tests/util.py.
"""
import ctypes as C

import numpy as np
import pytest

import plasticity_b200 as pb
from plasticity_b200 import _lib as L
from oracle import pyoracle as po
import util
from test_gpu_parity import Dev

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return pb.nnet.context(0)


def dense_desc(n_in, n_out):
    return L.LayerDesc(kind=L.PB_DENSE, in_=n_in, out=n_out)


# (batch, in, out)
SHAPES = [
    (300, 520, 700),     # A float4 / W scalar; M, N and K tails
    (256, 515, 260),     # in % 4 == 3: activations unaligned -> scalar both
    (129, 1000, 257),    # one row / one column past a tile edge
    (64, 4096, 512),     # few tiles -> split-K forward; K = 64 weight gradient
    (1024, 1024, 1024),  # several tiles per persistent CTA in the weight gradient
    (40, 4096, 10),      # skinny classifier head shapes stay correct (FFMA / skinny paths)
]


@pytest.mark.parametrize("batch,n_in,n_out", SHAPES)
def test_dense_layer_kernels_large(ctx, batch, n_in, n_out):
    rng = np.random.default_rng(batch * 7 + n_in * 3 + n_out)
    S = po.RestatedNet(f"act {n_in} identity\ndense {n_in} {n_out}\n")
    I = util.f32(rng.normal(0, 1, (batch, n_in)))
    G = util.f32(rng.normal(0, 1, (batch, n_out)))
    W = util.f32(rng.normal(0, 1 / np.sqrt(n_in), n_out * (n_in + 1)))
    Wm = W.reshape(n_out, n_in + 1)
    desc = dense_desc(n_in, n_out)
    dI_, dW_, dG = Dev(ctx, I), Dev(ctx, W), Dev(ctx, G)
    dO, dDI = Dev(ctx, np.zeros(batch * n_out)), Dev(ctx, np.zeros(batch * n_in))
    L.call("pb_layer_evaluate", ctx, C.byref(desc), dI_.p, dW_.p, dO.p, batch)
    L.call("pb_layer_input_delta", ctx, C.byref(desc), dI_.p, dW_.p, dG.p, dDI.p, batch)
    O = dO.host().reshape(batch, n_out)
    DI = dDI.host().reshape(batch, n_in)
    # every element against the fp64 matrix form of dense_layer.cc:12-40
    util.assert_close(O, I @ Wm[:, :n_in].T + Wm[:, n_in], "evaluate (matrix form)")
    util.assert_close(DI, G @ Wm[:, :n_in], "input_delta (matrix form)")
    # sampled rows against the oracle itself
    for b in sorted(set([0, batch - 1] + rng.integers(0, batch, 4).tolist())):
        util.assert_close(O[b], S.evaluate_layer(1, I[b].copy(), W), f"evaluate sample {b}")
        util.assert_close(DI[b], S.input_delta(1, I[b].copy(), W, G[b].copy()),
                          f"input_delta sample {b}")
    lr = 0.25
    d_lr = Dev(ctx, [lr])
    grad = np.concatenate([G.T @ I, G.sum(axis=0)[:, None]], axis=1).reshape(-1)
    # oracle cross-check of the matrix form on a small slice of the batch
    nb = min(batch, 3)
    d_small = sum(S.weight_update(1, I[b].copy(), W, G[b].copy(), lr, 1) for b in range(nb))
    g_small = np.concatenate([G[:nb].T @ I[:nb], G[:nb].sum(axis=0)[:, None]], axis=1).reshape(-1)
    np.testing.assert_allclose(d_small, -lr * g_small, rtol=1e-9, atol=1e-12)
    for mode, want in ((L.PB_NEW_WEIGHTS, W - lr * grad), (L.PB_WEIGHT_DELTAS, -lr * grad)):
        dOut = Dev(ctx, np.zeros(W.size))
        L.call("pb_layer_weight_update", ctx, C.byref(desc), dI_.p, dW_.p, dG.p, dOut.p, d_lr.p,
               mode, batch)
        util.assert_close(dOut.host(), want, f"weight mode {mode}")
    dOut = Dev(ctx, np.ones(W.size))
    L.call("pb_layer_weight_update", ctx, C.byref(desc), dI_.p, dW_.p, dG.p, dOut.p, d_lr.p,
           L.PB_ACCUMULATE, batch)
    util.assert_close(dOut.host(), 1.0 - lr * grad, "accumulate")


def test_wide_mlp_batch_train_step(ctx):
    """One gradient-sum step of a scaled-down wide MLP (same layer list as
    BASELINE.json config 4, width 512) against the oracle's BatchTrain."""
    width, B = 512, 96
    spec = (f"loss ce\nact {width} identity\ndense {width} {width}\nact {width} relu\n"
            f"dense {width} {width}\nact {width} relu\ndense {width} 10\nact 10 identity\n"
            f"softmax 10\n")
    S = po.RestatedNet(spec)
    model, loss = util.architecture_from_spec(spec)
    rng = np.random.default_rng(4096)
    w0 = util.xavier_like(S, rng)
    xs = util.f32(rng.normal(0, 1, (B, width)))
    lab = rng.integers(0, 10, B)
    es = np.zeros((B, 10))
    es[np.arange(B), lab] = 1.0
    lr = 1e-3
    net = pb.Nnet(model, pb.NoWeightInit, loss, max_batch=B)
    net.SetLearningParameters(pb.Nnet.LearningParameters(lr))
    util.load_product_weights(net, S, w0)
    w_ref = w0.copy()
    S.batch_train(xs.copy(), es.copy(), w_ref, lr)
    # the deltas themselves (fp32, from the delta buffer) at the tolerance of record, then the
    # step through BatchTrain (update fused into the weight-gradient kernels)
    util.assert_close(util.device_deltas(net, S, xs, es), w_ref - w0, "wide MLP BatchTrain deltas")
    net.BatchTrain(net.MakeBuffer(xs.reshape(-1)), net.MakeBuffer(es.reshape(-1)), B)
    util.assert_close(util.read_product_weights(net, S), w_ref, "wide MLP BatchTrain weights")


def test_wide_classifier_head(ctx):
    """dense 4096 -> 10 + identity + softmax + cross-entropy (the head of BASELINE.json
    config 4: too wide for the fused head kernel's shared-memory budget, so the tensor-core
    dense kernels with split-K, the softmax / error kernels and the skinny weight gradient
    run layer by layer): one gradient-sum step against the oracle."""
    n_in, B = 4096, 72
    spec = f"loss ce\nact {n_in} identity\ndense {n_in} 10\nact 10 identity\nsoftmax 10\n"
    S = po.RestatedNet(spec)
    model, loss = util.architecture_from_spec(spec)
    rng = np.random.default_rng(11)
    w0 = util.xavier_like(S, rng)
    xs = util.f32(rng.normal(0, 1, (B, n_in)))
    es = np.zeros((B, 10))
    es[np.arange(B), rng.integers(0, 10, B)] = 1.0
    lr = 1e-3
    net = pb.Nnet(model, pb.NoWeightInit, loss, max_batch=B)
    net.SetLearningParameters(pb.Nnet.LearningParameters(lr))
    util.load_product_weights(net, S, w0)
    w_ref = w0.copy()
    S.batch_train(xs.copy(), es.copy(), w_ref, lr)
    util.assert_close(util.device_deltas(net, S, xs, es), w_ref - w0, "wide head deltas")
    net.BatchTrain(net.MakeBuffer(xs.reshape(-1)), net.MakeBuffer(es.reshape(-1)), B)
    util.assert_close(util.read_product_weights(net, S), w_ref, "wide head weights")


@pytest.mark.parametrize("n_in,n_out,act,loss", [
    (333, 32, "sigmoid", "mse"),   # every lane owns an output; input tail inside a pass
    (600, 17, "relu", "ce"),       # three passes of 256 inputs, odd output count
    (64, 16, "leaky", "ce"),
    (7, 5, "identity", "mse"),     # fewer inputs than lanes
])
def test_fused_head_shapes(ctx, n_in, n_out, act, loss):
    """dense -> activation -> softmax + loss in the fused head kernel (head.cu: partial sums
    per lane, halving exchange across the warp, gradients broadcast from registers) for
    output counts up to a full warp and input counts off the 32-lane grid: three
    gradient-sum steps against the oracle."""
    B = 70
    spec = (f"loss {loss}\nact {n_in} identity\ndense {n_in} {n_out}\nact {n_out} {act}\n"
            f"softmax {n_out}\n")
    S = po.RestatedNet(spec)
    model, loss_fn = util.architecture_from_spec(spec)
    rng = np.random.default_rng(n_in * 37 + n_out)
    w0 = util.xavier_like(S, rng)
    # softmax behind a bounded activation under MSE has tiny gradients: a learning rate that
    # keeps the updates well above one fp32 ulp of the weights they are added to
    lr = 0.25 if loss == "mse" else 2e-3
    net = pb.Nnet(model, pb.NoWeightInit, loss_fn, max_batch=B)
    net.SetLearningParameters(pb.Nnet.LearningParameters(lr))
    util.load_product_weights(net, S, w0)
    w_ref = w0.copy()
    for step in range(3):
        xs = util.f32(rng.normal(0, 1, (B, n_in)))
        es = np.zeros((B, n_out))
        es[np.arange(B), rng.integers(0, n_out, B)] = 1.0
        w_before = w_ref.copy()
        S.batch_train(xs.copy(), es.copy(), w_ref, lr)
        util.assert_close(util.device_deltas(net, S, xs, es), w_ref - w_before, f"head step {step} deltas")
        net.BatchTrain(net.MakeBuffer(xs.reshape(-1)), net.MakeBuffer(es.reshape(-1)), B)
        got = util.read_product_weights(net, S)
        util.assert_close(got, w_ref, f"head step {step} weights")
        # keep the two trajectories from drifting apart through fp32 rounding
        util.load_product_weights(net, S, w_ref)
