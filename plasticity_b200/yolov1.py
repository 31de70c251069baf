"""nnet::Yolov1 (nnet/yolov1.h:22-31, nnet/yolov1.cc:7-231) over the B200 layer path.

The reference only sketches this class: `yolov1.h` declares `SetParams / Evaluate /
TrainSample` without signatures and `yolov1.cc` is a layer list that does not compile
and whose dimensions do not chain.  The layer list is kept as written with the three
repairs SURVEY.md section 8(d) pins:
  * conv 2 (3x3 over 112x112x64) uses padding 1, not 2 (yolov1.cc:27-40: padding 2
    would give 114x114, the pool behind it expects 112x112);
  * the 1x1 convolution of the second 14x14 pair reads depth 1024 — the output of the
    3x3x1024 before it — not 512 (yolov1.cc:145-172);
  * the model ends at AddDenseLayer(1470) (7*7*30, yolov1.cc:231); the stray
    statements after the `;` (yolov1.cc:232-235) are dropped.
Activations are the builder defaults: LeakyRelu behind every convolution
(architecture.h:77-85), Sigmoid behind AddDenseLayer(n) (architecture.h:63-74).
"""
from typing import Optional

import numpy as np

from . import nnet as N

kInputWidth = 448
kInputSize = kInputWidth * kInputWidth * 3
kOutputSize = 7 * 7 * 30


def YoloArchitecture(input_width: int = kInputWidth, dense_width: int = 4096,
                     outputs: int = kOutputSize) -> N.Architecture:
    """input_width must be a multiple of 64 (448 in the reference); smaller widths give
    the same layer list over smaller images — used by the parity tests."""
    if input_width % 64:
        raise ValueError("Yolov1: input width must be a multiple of 64")
    w = input_width
    V, F, A = N.VolumeDimensions, N.FilterParams, N.AreaDimensions
    m = N.Architecture(w * w * 3)
    m.AddConvolutionLayer(V(w, w, 3), F(7, 7, 3, 2, 3, 64))
    w //= 2
    m.AddMaxPoolLayer(V(w, w, 64), A(w // 2, w // 2))
    w //= 2
    m.AddConvolutionLayer(V(w, w, 64), F(3, 3, 64, 1, 1, 192))  # repaired: padding 1
    m.AddMaxPoolLayer(V(w, w, 192), A(w // 2, w // 2))
    w //= 2
    m.AddConvolutionLayer(V(w, w, 192), F(1, 1, 192, 1, 0, 128))
    m.AddConvolutionLayer(V(w, w, 128), F(3, 3, 128, 1, 1, 256))
    m.AddConvolutionLayer(V(w, w, 256), F(1, 1, 256, 1, 0, 256))
    m.AddConvolutionLayer(V(w, w, 256), F(3, 3, 256, 1, 1, 512))
    m.AddMaxPoolLayer(V(w, w, 512), A(w // 2, w // 2))
    w //= 2
    for _ in range(4):
        m.AddConvolutionLayer(V(w, w, 512), F(1, 1, 512, 1, 0, 256))
        m.AddConvolutionLayer(V(w, w, 256), F(3, 3, 256, 1, 1, 512))
    m.AddConvolutionLayer(V(w, w, 512), F(1, 1, 512, 1, 0, 512))
    m.AddMaxPoolLayer(V(w, w, 512), A(w // 2, w // 2))
    w //= 2
    depth = 512
    for _ in range(2):
        m.AddConvolutionLayer(V(w, w, depth), F(1, 1, depth, 1, 0, 512))  # repaired: depth chains
        m.AddConvolutionLayer(V(w, w, 512), F(3, 3, 512, 1, 1, 1024))
        depth = 1024
    m.AddConvolutionLayer(V(w, w, 1024), F(3, 3, 1024, 1, 1, 1024))
    m.AddConvolutionLayer(V(w, w, 1024), F(3, 3, 1024, 2, 1, 1024))
    w //= 2
    m.AddConvolutionLayer(V(w, w, 1024), F(3, 3, 1024, 1, 1, 1024))
    m.AddConvolutionLayer(V(w, w, 1024), F(3, 3, 1024, 1, 1, 1024))
    m.AddDenseLayer(dense_width)
    m.AddDenseLayer(outputs)
    return m


class Yolov1:
    """The class yolov1.h declares: owns an Nnet over YoloArchitecture()."""

    def __init__(self, max_batch: int = 1, device: int = 0, seed: int = 0,
                 input_width: int = kInputWidth, dense_width: int = 4096,
                 outputs: int = kOutputSize, weight_initialization: int = N.Xavier):
        self.input_width = input_width
        self.net_ = N.Nnet(YoloArchitecture(input_width, dense_width, outputs),
                           weight_initialization, N.MeanSquared, max_batch=max_batch,
                           device=device, seed=seed)

    def SetParams(self, learning_rate: Optional[float] = None,
                  weights: Optional[str] = None) -> bool:
        """Learning rate (Nnet::SetLearningParameters) and / or a weight file in the
        reference's JSON format (Nnet::LoadWeightsFromString)."""
        if learning_rate is not None:
            self.net_.SetLearningParameters(N.Nnet.LearningParameters(learning_rate))
        if weights is not None:
            return self.net_.LoadWeightsFromString(weights)
        return True

    def Evaluate(self, images, batch: int = 1) -> np.ndarray:
        """images: batch x (3 x W x W) planar RGB values; returns batch x 1470."""
        buf = images if isinstance(images, N.ClBuffer) else self.net_.MakeBuffer(
            np.asarray(images, dtype=np.float64).reshape(-1))
        out = self.net_.Evaluate(buf, batch=batch)
        out.MoveToCpu()
        return out.cpu.reshape(batch, -1).copy()

    def TrainSample(self, image, expected) -> None:
        """One per-sample SGD step (Nnet::Train) on the squared error of the 7x7x30 grid."""
        self.net_.Train(self.net_.MakeBuffer(np.asarray(image, dtype=np.float64).reshape(-1)),
                        self.net_.MakeBuffer(np.asarray(expected, dtype=np.float64).reshape(-1)))

    def TrainBatch(self, images, expected, batch: int) -> None:
        """One gradient-sum minibatch step (Nnet::BatchTrain semantics, nnet.h:617-669:
        W <- W - lr * sum over the batch) — the batched counterpart of TrainSample: every
        convolution's input gradient and weight gradient run on the tensor-core GEMM policies
        (conv_gemm_umma.cu, dense_umma.cu)."""
        bx = images if isinstance(images, N.ClBuffer) else self.net_.MakeBuffer(
            np.asarray(images, dtype=np.float64).reshape(-1))
        be = expected if isinstance(expected, N.ClBuffer) else self.net_.MakeBuffer(
            np.asarray(expected, dtype=np.float64).reshape(-1))
        self.net_.BatchTrain(bx, be, batch)
