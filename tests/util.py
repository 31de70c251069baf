"""Shared helpers for the parity tests: build the product network and the
oracle from ONE spec text, seeded inputs, and the tolerance of record."""
import numpy as np

import plasticity_b200 as pb
from oracle import pyoracle as po

# north_star: outputs, gradients and weights within |rel| <= 1e-4 (fp32 device
# arithmetic against the fp64 oracle).  Elements that cancel to ~0 are judged
# against the tensor's scale: |a-b| <= RTOL*|b| + ATOL_SCALE*max|b|.
RTOL = 1e-4
ATOL_SCALE = 1e-5

ACT = {"identity": pb.Identity, "relu": pb.Relu, "leaky": pb.LeakyRelu, "sigmoid": pb.Sigmoid}


def architecture_from_spec(text):
    """spec text (oracle/ref_gen.cc grammar) -> (pb.Architecture, loss)."""
    model, loss = pb.Architecture(), pb.MeanSquared
    for raw in text.splitlines():
        t = raw.split("#")[0].split()
        if not t:
            continue
        k, a = t[0], t[1:]
        if k == "loss":
            loss = pb.CrossEntropy if a[0] == "ce" else pb.MeanSquared
        elif k == "act":
            model.AddActivationLayer(ACT[a[1]], int(a[0]))
        elif k == "dense":
            model._push(pb._lib.LayerDesc(kind=pb._lib.PB_DENSE, in_=int(a[0]), out=int(a[1])),
                        "dense_layer")
        elif k == "conv":
            v = list(map(int, a))
            model._push(pb._lib.LayerDesc(kind=pb._lib.PB_CONVOLUTION, width=v[0], height=v[1],
                                          depth=v[2], filter_width=v[3], filter_height=v[4],
                                          filter_depth=v[5], stride=v[6], padding=v[7],
                                          num_filters=v[8]), "convolution_layer")
        elif k == "pool":
            v = list(map(int, a))
            model.AddMaxPoolLayer(pb.VolumeDimensions(*v[:3]), pb.AreaDimensions(*v[3:]))
        elif k == "softmax":
            model.AddSoftmaxLayer(int(a[0]))
    return model, loss


def f32(a):
    """Round to fp32 and back: values the device can hold exactly."""
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def xavier_like(net, rng, scale=1.0):
    """Weights ~ N(0, scale^2/fan_in) per layer so activations stay O(1).
    `net` is an oracle RestatedNet (it knows each layer's kind)."""
    w = np.zeros(net.total_w)
    for l in range(net.n_layers):
        if not net.nw[l]:
            continue
        L = net.layers[l]
        fan_in = (L.in_ + 1) if L.kind == po.KINDS["dense"] else (L.fw * L.fh * L.fd + 1)
        w[net.w_off[l]:net.w_off[l + 1]] = rng.normal(0, scale / np.sqrt(fan_in), net.nw[l])
    return f32(w)


def assert_close(actual, expected, what="", rtol=RTOL, atol_scale=ATOL_SCALE):
    actual = np.asarray(actual, dtype=np.float64)
    expected = np.asarray(expected, dtype=np.float64)
    assert actual.shape == expected.shape, (what, actual.shape, expected.shape)
    scale = float(np.max(np.abs(expected))) if expected.size else 0.0
    # 1.2e-38: below the fp32 normal range the device flushes toward zero.
    tol = rtol * np.abs(expected) + atol_scale * scale + 1.2e-38
    err = np.abs(actual - expected)
    bad = err > tol
    if bad.any():
        i = int(np.argmax(err - tol))
        raise AssertionError(
            f"{what}: {int(bad.sum())}/{bad.size} outside |rel|<={rtol} (atol {atol_scale}*max): "
            f"worst idx {i} got {actual.flat[i]!r} want {expected.flat[i]!r} scale {scale:g}")


def load_product_weights(pnet, onet, weights):
    for l in range(onet.n_layers):
        if onet.nw[l]:
            pnet.SetLayerWeights(l, onet.layer_weights(weights, l))


def read_product_weights(pnet, onet):
    w = np.zeros(onet.total_w)
    for l in range(onet.n_layers):
        if onet.nw[l]:
            w[onet.w_off[l]:onet.w_off[l + 1]] = pnet.GetLayerWeights(l)
    return w


def pool_margin(onet, x, weights):
    """Smallest relative gap between the two largest values of any max-pool
    window of the network evaluated (fp64 oracle) at `x`.  Max-pool routing is a
    discontinuous function of its inputs: when two window elements agree to
    ~1e-6 (fp32 / 3xTF32 rounding), the device and the fp64 oracle may route the
    gradient to different elements, which is a property of finite precision and
    not of either implementation.  Parity cases are drawn with a margin."""
    lo = onet.evaluate(np.array(x, dtype=np.float64), np.array(weights, dtype=np.float64))
    worst = np.inf
    for l in range(onet.n_layers):
        L = onet.layers[l]
        if L.kind != po.KINDS["pool"]:
            continue
        v = onet.layer_output(lo, l - 1).reshape(L.D, L.H, L.W)
        gh, gw = L.H // L.oh, L.W // L.ow
        v = v[:, :L.oh * gh, :L.ow * gw].reshape(L.D, L.oh, gh, L.ow, gw)
        win = np.sort(v.transpose(0, 1, 3, 2, 4).reshape(-1, gh * gw), axis=1)
        if win.shape[1] < 2:
            continue
        prev = onet.layers[l - 1] if l > 0 else None
        if prev is not None and prev.kind == po.KINDS["act"] and prev.act == po.ACTS["relu"]:
            # a window that ReLU flattened to zeros routes to every element, but every one of
            # them has derivative 0: harmless.  A window whose maximum is barely positive is
            # not: its sign can differ in fp32.
            live = win[:, -1] > 0
            if np.any(live & (win[:, -1] <= 1e-4)):
                return 0.0
            win = win[live]
            if win.shape[0] == 0:
                continue
        gap = (win[:, -1] - win[:, -2]) / (np.abs(win[:, -1]) + 1e-3)
        worst = min(worst, float(gap.min()))
    return worst


def draw_with_pool_margin(onet, rng, n, weights, margin=2e-5, scale=1.0):
    """n fp32-representable N(0, scale^2) samples whose pool windows all have a
    top-2 gap of at least `margin` (relative) under `weights`."""
    out = []
    tries = 0
    while len(out) < n:
        tries += 1
        if tries > 200 * n + 1000:
            raise RuntimeError("draw_with_pool_margin: no sample with a pool margin in "
                               f"{tries} draws (a network whose windows always tie?)")
        x = f32(rng.normal(0, scale, onet.n_in))
        if pool_margin(onet, x, weights) >= margin:
            out.append(x)
    return np.stack(out)


def device_deltas(pnet, onet, xs, es):
    """-lr * sum_b grad_b as the device forms it (pb_nnet_batch_gradients: the weight-gradient
    kernels in WEIGHT_DELTAS mode), read from the flat delta buffer in the oracle's layout -
    the fp32 deltas themselves, not a difference of rounded weights."""
    import ctypes as C
    from plasticity_b200 import _lib as L
    B = xs.shape[0]
    bx, be = pnet.MakeBuffer(xs.reshape(-1)), pnet.MakeBuffer(es.reshape(-1))
    bx.MoveToGpu()
    be.MoveToGpu()
    pnet._sync_to_device()
    L.call("pb_nnet_batch_gradients", pnet.h, bx.gpu_buffer(), be.gpu_buffer(), B)
    dp = C.c_void_p()
    L.call("pb_nnet_deltas_device", pnet.h, C.byref(dp))
    n_flat = sum((n + 3) & ~3 for n in onet.nw)
    flat = np.zeros(n_flat, dtype=np.float32)
    L.call("pb_buffer_read_f32", pnet.ctx, dp, flat.ctypes.data_as(C.c_void_p), n_flat)
    out = np.zeros(onet.total_w)
    off = 0
    for l in range(onet.n_layers):
        n = onet.nw[l]
        out[onet.w_off[l]:onet.w_off[l + 1]] = flat[off:off + n]
        off += (n + 3) & ~3
    return out
