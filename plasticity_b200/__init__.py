"""plasticity_b200 — B200-native layer forward / back-prop / weight-update path
of jsharf/plasticity behind the reference's own Architecture / Nnet / ClBuffer
interface.

The compute lives in ``lib/libplasticity_b200.so`` (hand-written sm_100a CUDA,
C ABI declared in ``include/plasticity_b200.h``).  This package is the Python
host mirror used by the tests and the bench; the C++ mirror is under ``cpp/``.
There is no CPU path: importing is fine without a GPU, creating a context is not.
"""
from . import _lib
from .nnet import (Architecture, AreaDimensions, ClBuffer, CrossEntropy, FilterParams, Identity,
                   InitToOne, LeakyRelu, MeanSquared, Nnet, NoWeightInit, Relu, Sigmoid,
                   VolumeDimensions, Xavier, DenseSymbolGenerator, ConvSymbolGenerator)

__all__ = ["Architecture", "AreaDimensions", "ClBuffer", "CrossEntropy", "FilterParams",
           "Identity", "InitToOne", "LeakyRelu", "MeanSquared", "Nnet", "NoWeightInit", "Relu",
           "Sigmoid", "VolumeDimensions", "Xavier", "DenseSymbolGenerator",
           "ConvSymbolGenerator", "_lib"]
