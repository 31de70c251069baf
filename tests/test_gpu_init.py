"""Device-side weight initialisation: Layer::XavierInitializeWeights draws
every weight of a layer from stats::Normal(0, variance = 1 / num_inputs)
(nnet/layer.cc:86-100, stats/normal.h:9-41) and InitializeWeights(value) fills a constant
(nnet/layer.cc:102-108, Nnet::InitToOne nnet.h:852-856).  The reference seeds its mt19937
from random_device, so its draws are unpinned; here the draw is a counter-based generator on
the device, seeded: reproducible per seed, different across seeds, right distribution."""
import numpy as np
import pytest

import plasticity_b200 as pb

pytestmark = pytest.mark.gpu


def model():
    m = pb.Architecture(24 * 24 * 3)
    m.AddConvolutionLayer(pb.VolumeDimensions(24, 24, 3), pb.FilterParams(5, 5, 3, 1, 2, 32))
    m.AddMaxPoolLayer(pb.VolumeDimensions(24, 24, 32), pb.AreaDimensions(12, 12))
    m.AddDenseLayer(300)
    m.AddDenseLayer(10, pb.Identity)
    m.AddSoftmaxLayer(10)
    return m


def layer_weights(net):
    return {l: np.asarray(net.GetLayerWeights(l)) for l in range(len(net.nw)) if net.nw[l]}


def test_xavier_distribution_and_seed():
    a = layer_weights(pb.Nnet(model(), pb.Xavier, pb.CrossEntropy, seed=7))
    b = layer_weights(pb.Nnet(model(), pb.Xavier, pb.CrossEntropy, seed=7))
    c = layer_weights(pb.Nnet(model(), pb.Xavier, pb.CrossEntropy, seed=8))
    net = pb.Nnet(model(), pb.NoWeightInit, pb.CrossEntropy)
    assert len(a) == 3
    for l, w in a.items():
        n, var = w.size, 1.0 / net.ni[l]          # variance 1 / num_inputs of the LAYER
        assert np.array_equal(w, b[l]), f"layer {l}: same seed, different draw"
        assert not np.array_equal(w, c[l]), f"layer {l}: different seeds, same draw"
        assert np.isfinite(w).all()
        sd = np.sqrt(var)
        assert abs(w.mean()) < 5 * sd / np.sqrt(n), (l, w.mean(), sd)
        # sample variance of n normal draws: relative standard error sqrt(2/n)
        assert abs(w.var() / var - 1) < 6 * np.sqrt(2.0 / n), (l, w.var(), var)
        # shape of the distribution: tail mass and kurtosis of a normal
        z = w / sd
        assert abs((np.abs(z) > 1.959964).mean() - 0.05) < 6 * np.sqrt(0.05 * 0.95 / n)
        assert abs((z ** 4).mean() - 3.0) < 6 * np.sqrt(96.0 / n)
        # no two layers (and no two halves of one layer) share a stream
        h = n // 2
        assert abs(np.corrcoef(w[:h], w[h:2 * h])[0, 1]) < 6 / np.sqrt(h)
    # weights of different layers come from different counters
    w1, w4 = a[min(a)], a[max(a)]
    k = min(w1.size, w4.size)
    assert abs(np.corrcoef(w1[:k], w4[:k])[0, 1]) < 6 / np.sqrt(k)


def test_init_to_one_and_no_weight_init():
    one = layer_weights(pb.Nnet(model(), pb.InitToOne, pb.CrossEntropy))
    for l, w in one.items():
        assert np.array_equal(w, np.ones_like(w)), f"layer {l}"
    # NoWeightInit keeps what the Architecture holds (zeros unless the caller set them)
    m = model()
    m.layers[1].weights[:] = 0.25
    net = pb.Nnet(m, pb.NoWeightInit, pb.CrossEntropy)
    assert np.array_equal(net.GetLayerWeights(1), np.full(net.nw[1], 0.25))
    with pytest.raises(pb._lib.PlasticityError):
        pb.Nnet(model(), 99, pb.CrossEntropy)   # nnet.h:859-863: unknown strategy
