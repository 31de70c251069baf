"""GPU parity of the general 3xTF32 tcgen05 implicit-GEMM convolution
(csrc/conv_gemm_umma.cu) — the forward path of the YOLOv1 configuration (BASELINE.json
config 5: 7x7/2, 3x3, 1x1 and 3x3/2 filters over 3..1024 channels) — through the C ABI,
against the CPU oracle (plain-C restatement of nnet/convolution_layer.cc:36-96) on
sampled samples and against its fp64 sliding-window form on every element.

Shapes are scaled-down YOLOv1 layers plus edge cases: channel counts below a K stage
(RGB), N tails, non-square images, padding 0, stride 2, split-K (few tiles, long K).
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


def conv_ref(I, Wt, W, H, D, fw, fh, stride, pad, nf):
    """fp64 zero-padded cross-correlation, convolution_layer.cc:50-96, all samples."""
    B = I.shape[0]
    x = np.zeros((B, D, H + 2 * pad, W + 2 * pad))
    x[:, :, pad:pad + H, pad:pad + W] = I.reshape(B, D, H, W)
    oh, ow = (H - fh + 2 * pad) // stride + 1, (W - fw + 2 * pad) // stride + 1
    win = np.lib.stride_tricks.sliding_window_view(x, (fh, fw), axis=(2, 3))
    win = win[:, :, ::stride, ::stride][:, :, :oh, :ow]          # B D oh ow fh fw
    Wm = Wt.reshape(nf, D * fh * fw + 1)
    k = Wm[:, :-1].reshape(nf, D, fh, fw)
    out = np.einsum("bdrcyx,fdyx->bfrc", win, k, optimize=True) + Wm[:, -1][None, :, None, None]
    return out.reshape(B, -1)


# (batch, W, H, D, fw, fh, stride, pad, filters)
# batch >= 32: below that the layer entry points use the one-thread-per-output kernels
SHAPES = [
    (32, 64, 64, 3, 7, 7, 2, 3, 64),     # yolov1.cc:11-24 (RGB, 7x7/2), scaled down
    (32, 28, 28, 64, 3, 3, 1, 1, 192),   # yolov1.cc:27-40 (repaired padding 1)
    (32, 28, 28, 192, 1, 1, 1, 0, 128),  # 1x1 reduction
    (32, 14, 14, 64, 3, 3, 2, 1, 272),   # 3x3/2 (yolov1.cc:193-205); N tail past one tile
    (64, 20, 12, 16, 3, 3, 1, 1, 32),    # non-square
    (64, 20, 20, 5, 5, 5, 1, 0, 32),     # 5 channels (padded to 8 per tap), no padding
    (32, 7, 7, 512, 3, 3, 1, 1, 512),    # few tiles, K = 4608: split-K + fixed-order reduce
]


@pytest.mark.parametrize("batch,W,H,D,fw,fh,stride,pad,nf", SHAPES)
def test_conv_gemm_forward(ctx, batch, W, H, D, fw, fh, stride, pad, nf):
    n_in = W * H * D
    S = po.RestatedNet(f"act {n_in} identity\nconv {W} {H} {D} {fw} {fh} {D} {stride} {pad} {nf}\n")
    n_out = S.no[1]
    rng = np.random.default_rng(W * 31 + D * 7 + nf)
    I = util.f32(rng.normal(0, 1, (batch, n_in)))
    Wt = util.f32(rng.normal(0, 1 / np.sqrt(D * fw * fh), nf * (D * fw * fh + 1)))
    desc = L.LayerDesc(kind=L.PB_CONVOLUTION, width=W, height=H, depth=D, filter_width=fw,
                       filter_height=fh, filter_depth=D, stride=stride, padding=pad,
                       num_filters=nf)
    dI_, dW_, dO = Dev(ctx, I), Dev(ctx, Wt), Dev(ctx, np.zeros(batch * n_out))
    L.call("pb_layer_evaluate", ctx, C.byref(desc), dI_.p, dW_.p, dO.p, batch)
    O = dO.host().reshape(batch, n_out)
    util.assert_close(O, conv_ref(I, Wt, W, H, D, fw, fh, stride, pad, nf), "evaluate (window form)")
    for b in sorted({0, batch - 1, int(rng.integers(0, batch))}):
        util.assert_close(O[b], S.evaluate_layer(1, I[b].copy(), Wt), f"evaluate sample {b}")


def conv_bwd_ref(I, G, Wt, W, H, D, fw, fh, stride, pad, nf):
    """fp64 input gradient and weight gradient (sum over the batch) of the zero-padded
    cross-correlation: the adjoints of conv_ref.  (convolution_layer.cc:100-270 for odd filters
    with pad == filter/2 and square or non-square outputs with the bounds the index algebra
    needs, i.e. the oracle's restatement; checked against it sample by sample below.)"""
    B = I.shape[0]
    oh, ow = (H - fh + 2 * pad) // stride + 1, (W - fw + 2 * pad) // stride + 1
    x = np.zeros((B, D, H + 2 * pad, W + 2 * pad))
    x[:, :, pad:pad + H, pad:pad + W] = I.reshape(B, D, H, W)
    g = G.reshape(B, nf, oh, ow)
    k = Wt.reshape(nf, D * fh * fw + 1)[:, :-1].reshape(nf, D, fh, fw)
    win = np.lib.stride_tricks.sliding_window_view(x, (fh, fw), axis=(2, 3))
    win = win[:, :, ::stride, ::stride][:, :, :oh, :ow]
    gw = np.einsum("bfrc,bdrcyx->fdyx", g, win, optimize=True)
    gb = g.sum(axis=(0, 2, 3))
    grad_w = np.concatenate([gw.reshape(nf, -1), gb[:, None]], axis=1).reshape(-1)
    dx = np.zeros_like(x)
    for fy in range(fh):
        for fx in range(fw):
            dx[:, :, fy:fy + stride * oh:stride, fx:fx + stride * ow:stride] += np.einsum(
                "bfrc,fd->bdrc", g, k[:, :, fy, fx], optimize=True)
    return dx[:, :, pad:pad + H, pad:pad + W].reshape(B, -1), grad_w


@pytest.mark.parametrize("batch,W,H,D,fw,fh,stride,pad,nf", SHAPES)
def test_conv_gemm_backward(ctx, batch, W, H, D, fw, fh, stride, pad, nf):
    """Input gradient and weight gradient + SGD update of the same shapes: the tensor-core
    GEMM policies of conv_gemm_umma.cu (transposed gather) and dense_umma.cu (gathered
    operands, K = (sample, pixel)) where their envelope holds, the FFMA kernels elsewhere."""
    n_in = W * H * D
    S = po.RestatedNet(f"act {n_in} identity\nconv {W} {H} {D} {fw} {fh} {D} {stride} {pad} {nf}\n")
    n_out, nw = S.no[1], S.nw[1]
    rng = np.random.default_rng(W * 17 + D * 5 + nf)
    I = util.f32(rng.normal(0, 1, (batch, n_in)))
    G = util.f32(rng.normal(0, 1, (batch, n_out)))
    Wt = util.f32(rng.normal(0, 1 / np.sqrt(D * fw * fh), nw))
    lr = 0.01
    desc = L.LayerDesc(kind=L.PB_CONVOLUTION, width=W, height=H, depth=D, filter_width=fw,
                       filter_height=fh, filter_depth=D, stride=stride, padding=pad,
                       num_filters=nf)
    dI_, dW_, dG = Dev(ctx, I), Dev(ctx, Wt), Dev(ctx, G)
    dDI, d_lr = Dev(ctx, np.zeros(batch * n_in)), Dev(ctx, [lr])
    L.call("pb_layer_input_delta", ctx, C.byref(desc), dI_.p, dW_.p, dG.p, dDI.p, batch)
    DI = dDI.host().reshape(batch, n_in)
    di_ref, gw_ref = conv_bwd_ref(I, G, Wt, W, H, D, fw, fh, stride, pad, nf)
    # the closed form is the restatement's arithmetic: two samples through the oracle itself
    for b in (0, batch - 1):
        util.assert_close(di_ref[b], S.input_delta(1, I[b].copy(), Wt, G[b].copy()),
                          f"closed form vs oracle, input gradient of sample {b}", rtol=1e-9, atol_scale=1e-12)
    _, gw1 = conv_bwd_ref(I[:1], G[:1], Wt, W, H, D, fw, fh, stride, pad, nf)
    util.assert_close(-lr * gw1, S.weight_update(1, I[0].copy(), Wt, G[0].copy(), lr, 1),
                      "closed form vs oracle, weight deltas of sample 0", rtol=1e-9, atol_scale=1e-12)
    util.assert_close(DI, di_ref, "input_delta")
    for mode, want in ((L.PB_NEW_WEIGHTS, Wt - lr * gw_ref), (L.PB_WEIGHT_DELTAS, -lr * gw_ref)):
        dOut = Dev(ctx, np.zeros(nw))
        L.call("pb_layer_weight_update", ctx, C.byref(desc), dI_.p, dW_.p, dG.p, dOut.p, d_lr.p,
               mode, batch)
        util.assert_close(dOut.host(), want, f"weight mode {mode}")
    dOut = Dev(ctx, np.ones(nw))
    L.call("pb_layer_weight_update", ctx, C.byref(desc), dI_.p, dW_.p, dG.p, dOut.p, d_lr.p,
           L.PB_ACCUMULATE, batch)
    util.assert_close(dOut.host(), 1.0 - lr * gw_ref, "accumulate")
    L.call("pb_layer_weight_update", ctx, C.byref(desc), dI_.p, dW_.p, dG.p, dW_.p, d_lr.p,
           L.PB_NEW_WEIGHTS, batch)
    util.assert_close(dW_.host(), Wt - lr * gw_ref, "in place")


def test_conv_activation_fused_in_network_evaluate(ctx):
    """Nnet::Evaluate on a batch: convolution + LeakyReLU run as one kernel, both layer
    outputs are materialised (eval_layer_outputs_), followed by pool + dense."""
    spec = ("act 3072 identity\nconv 32 32 3 3 3 3 1 1 16\nact 16384 leaky\npool 32 32 16 16 16\n"
            "conv 16 16 16 3 3 16 1 1 32\nact 8192 leaky\ndense 8192 24\nact 24 identity\n")
    S = po.RestatedNet(spec)
    model, loss = util.architecture_from_spec(spec)
    rng = np.random.default_rng(5)
    w0 = util.xavier_like(S, rng)
    B = 96
    xs = util.f32(rng.normal(0, 1, (B, S.n_in)))
    net = pb.Nnet(model, pb.NoWeightInit, loss, max_batch=B)
    util.load_product_weights(net, S, w0)
    outs = net.Evaluate(net.MakeBuffer(xs.reshape(-1)), batch=B)
    outs.MoveToCpu()
    got = outs.cpu.reshape(B, -1)
    for b in (0, 17, B - 1):
        lo = S.evaluate(xs[b].copy(), w0.copy())
        util.assert_close(got[b], S.layer_output(lo, S.n_layers - 1), f"network output sample {b}")
        for l in (1, 2, 4, 5):
            util.assert_close(net.LayerOutput(l, batch=B).reshape(B, -1)[b], S.layer_output(lo, l),
                              f"layer {l} output sample {b}")
