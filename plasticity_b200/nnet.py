"""Host-side mirror of the reference's model API over the C ABI.

Names, argument meaning and error behaviour follow the reference so that tests
read like nnet/nnet_test.cc:

* ``Architecture``      nnet/architecture.h:13-122
* ``Nnet``              nnet/nnet.h:33-923
* ``ClBuffer``          compute/cl_buffer.h:42-187 (CPU<->GPU state machine)
* ``DenseSymbolGenerator.WeightNumber`` / ``ConvSymbolGenerator.WeightNumber``
                        nnet/symbol_generator.h:141-156, 339-354

Host values are float64 (the reference's ``double``); the device holds fp32.
"""
from __future__ import annotations

import ctypes as C
import json
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import _lib as L

# symbolic/symbolic_util.h:10-15 — kept as tags the API accepts.
Identity, Relu, LeakyRelu, Sigmoid = (L.PB_IDENTITY, L.PB_RELU, L.PB_LEAKY_RELU, L.PB_SIGMOID)
_ACT_NAMES = {Identity: "identity", Relu: "relu", LeakyRelu: "leaky", Sigmoid: "sigmoid"}
# nnet/layer_dimensions.h:8-11
MeanSquared, CrossEntropy = L.PB_MEAN_SQUARED, L.PB_CROSS_ENTROPY
# nnet/nnet.h:40-48
Xavier, NoWeightInit, InitToOne = 0, 1, 2

kWeightFileFormatVersion = "dev"  # nnet/nnet.h:29


@dataclass
class VolumeDimensions:  # nnet/layer_dimensions.h:38-45
    width: int
    height: int
    depth: int


@dataclass
class FilterParams:  # nnet/layer_dimensions.h:22-36
    width: int
    height: int
    depth: int
    stride: int
    padding: int
    num_filters: int


@dataclass
class AreaDimensions:  # nnet/layer_dimensions.h:47-49
    width: int
    height: int


class DenseSymbolGenerator:
    """Weight layout of a dense layer: row-major [out][in+1], bias last."""

    def __init__(self, num_inputs: int, num_outputs: int):
        self.num_inputs, self.num_outputs = num_inputs, num_outputs

    def WeightNumber(self, node: int, edge: Optional[int] = None) -> int:
        if edge is None:  # bias of `node`
            edge = self.num_inputs
        return node * (self.num_inputs + 1) + edge


class ConvSymbolGenerator:
    """Weight layout of a conv layer: [filter][z][row][col] + bias per filter."""

    def __init__(self, dimensions: VolumeDimensions, params: FilterParams):
        self.d, self.p = dimensions, params

    def WeightNumber(self, filter: int, row: Optional[int] = None, col: int = 0, z: int = 0) -> int:
        size = self.p.width * self.p.height * self.p.depth + 1
        if row is None:  # bias: last slot of the filter block
            return filter * size + size - 1
        return filter * size + self.p.width * self.p.height * z + row * self.p.width + col


class Layer:
    """nnet/layer.h: a layer descriptor plus its host-side weights."""

    def __init__(self, desc: L.LayerDesc, type_name: str, index: int):
        self.desc, self.type_name, self.index = desc, type_name, index
        n = C.c_int64()
        L.call("pb_layer_num_inputs", C.byref(desc), C.byref(n))
        self.num_inputs = n.value
        L.call("pb_layer_num_outputs", C.byref(desc), C.byref(n))
        self.num_outputs = n.value
        L.call("pb_layer_num_weights", C.byref(desc), C.byref(n))
        self.weights = np.zeros(n.value, dtype=np.float64)

    def LayerSuffix(self) -> str:  # nnet/layer.h:99-101
        return f"{self.type_name}_{self.index}"

    def W(self, index: int) -> float:
        if index >= self.weights.size:
            raise IndexError(f"Too large weight index: {index}")  # layer.h:75-78
        return float(self.weights[index])

    def SetW(self, index: int, value: float) -> None:
        if index >= self.weights.size:
            raise IndexError(f"Too large weight index: {index}")
        self.weights[index] = value

    def spec_line(self) -> str:
        d = self.desc
        if d.kind == L.PB_ACTIVATION:
            return f"act {d.n} {_ACT_NAMES[d.activation]}"
        if d.kind == L.PB_DENSE:
            return f"dense {d.in_} {d.out}"
        if d.kind == L.PB_CONVOLUTION:
            return (f"conv {d.width} {d.height} {d.depth} {d.filter_width} {d.filter_height} "
                    f"{d.filter_depth} {d.stride} {d.padding} {d.num_filters}")
        if d.kind == L.PB_MAX_POOL:
            return f"pool {d.width} {d.height} {d.depth} {d.out_width} {d.out_height}"
        return f"softmax {d.n}"


class Architecture:
    """Fluent model builder (nnet/architecture.h)."""

    def __init__(self, input_size: Optional[int] = None):
        self.layers: List[Layer] = []
        if input_size is not None:
            self.AddInputLayer(input_size)

    def _push(self, desc: L.LayerDesc, type_name: str) -> "Architecture":
        self.layers.append(Layer(desc, type_name, len(self.layers)))
        return self

    def AddInputLayer(self, size: int) -> "Architecture":  # architecture.h:45-47
        return self.AddActivationLayer(Identity, size)

    def AddActivationLayer(self, activation_function: int, size: Optional[int] = None):
        if activation_function not in _ACT_NAMES:
            raise L.PlasticityError(
                "ActivationLayer: only Identity/Relu/LeakyRelu/Sigmoid are supported "
                "(no code generator, no fallback)")
        if size is None:
            size = self.layers[-1].num_outputs
        d = L.LayerDesc(kind=L.PB_ACTIVATION, activation=activation_function, n=size)
        return self._push(d, "activation_layer")

    def AddDenseLayer(self, num_outputs: int, activation_function: int = Sigmoid):
        d = L.LayerDesc(kind=L.PB_DENSE, in_=self.layers[-1].num_outputs, out=num_outputs)
        self._push(d, "dense_layer")
        return self.AddActivationLayer(activation_function)  # architecture.h:49-74

    def AddConvolutionLayer(self, dimensions: VolumeDimensions, params: FilterParams,
                            activation_function: int = LeakyRelu):
        if params.depth != dimensions.depth:  # convolution_layer.cc:12-16
            raise L.PlasticityError("Convolution layer input depth != filter depth. Error!")
        d = L.LayerDesc(kind=L.PB_CONVOLUTION, width=dimensions.width, height=dimensions.height,
                        depth=dimensions.depth, filter_width=params.width,
                        filter_height=params.height, filter_depth=params.depth,
                        stride=params.stride, padding=params.padding,
                        num_filters=params.num_filters)
        self._push(d, "convolution_layer")
        return self.AddActivationLayer(activation_function)  # architecture.h:77-85

    def AddMaxPoolLayer(self, input: VolumeDimensions, output: AreaDimensions):
        d = L.LayerDesc(kind=L.PB_MAX_POOL, width=input.width, height=input.height,
                        depth=input.depth, out_width=output.width, out_height=output.height)
        return self._push(d, "max_pool_layer")

    def AddSoftmaxLayer(self, size: int):
        return self._push(L.LayerDesc(kind=L.PB_SOFTMAX, n=size), "softmax_layer")

    def VerifyArchitecture(self) -> bool:  # architecture.h:16-32
        if not self.layers:
            return False
        for a, b in zip(self.layers, self.layers[1:]):
            if a.num_outputs != b.num_inputs:
                return False
        return True

    def input_size(self) -> int:
        return self.layers[0].num_inputs

    def output_size(self) -> int:
        return self.layers[-1].num_outputs

    def to_spec(self, loss: int) -> str:
        """One line per layer in the plain-text layer grammar the tests use."""
        head = "loss ce" if loss == CrossEntropy else "loss mse"
        return "\n".join([head] + [l.spec_line() for l in self.layers]) + "\n"


_ctx_cache = {}


def context(device: int = 0):
    """One pb_context (CUDA stream) per device, shared by the Nnets of a process
    — the reference's SelectDevice() always picks device 0 (nnet.h:188-197)."""
    if device not in _ctx_cache:
        h = C.c_void_p()
        L.call("pb_context_create", device, C.byref(h))
        _ctx_cache[device] = h
    return _ctx_cache[device]


def new_context(device: int = 0):
    """A private pb_context (its own CUDA stream) on `device`: one per rank when several ranks
    of a data-parallel group live in one process (pb_comm_create_local)."""
    h = C.c_void_p()
    L.call("pb_context_create", device, C.byref(h))
    return h


def local_comms(ctxs):
    """pb_comm_create_local: one communicator per context, rank = position."""
    arr = (C.c_void_p * len(ctxs))(*[c.value for c in ctxs])
    out = (C.c_void_p * len(ctxs))()
    L.call("pb_comm_create_local", arr, len(ctxs), out)
    return [C.c_void_p(v) for v in out]


class ClBuffer:
    """compute::ClBuffer: a float64 host vector OR an fp32 device allocation,
    with explicit MoveToGpu / MoveToCpu (blocking, like the reference)."""

    CPU, GPU = 0, 1

    def __init__(self, values=0, ctx=None):
        if isinstance(values, (int, np.integer)):
            self.cpu = np.zeros(int(values), dtype=np.float64)
        else:
            self.cpu = np.array(values, dtype=np.float64).reshape(-1)
        self.state = ClBuffer.CPU
        self.ctx = ctx
        self.dptr = None
        self._n = self.cpu.size

    def size(self) -> int:
        return self._n if self.state == ClBuffer.GPU else self.cpu.size

    def RegisterClBackend(self, ctx) -> None:  # cl_buffer.h:139-144
        self.MoveToCpu()
        self.ctx = ctx

    def MoveToGpu(self) -> None:  # cl_buffer.cc:101-127
        if self.state == ClBuffer.GPU:
            return
        if self.ctx is None:
            raise L.PlasticityError("Error, unexpected nullptr: context_")
        p = C.c_void_p()
        L.call("pb_buffer_alloc", self.ctx, self.cpu.size, C.byref(p))
        if self.cpu.size:
            L.call("pb_buffer_write_f64", self.ctx, p, self.cpu.ctypes.data, self.cpu.size)
        self.dptr, self._n, self.state = p, self.cpu.size, ClBuffer.GPU

    def MoveToCpu(self) -> None:  # cl_buffer.cc:75-99
        if self.state == ClBuffer.CPU:
            return
        self.cpu = np.empty(self._n, dtype=np.float64)
        if self._n:
            L.call("pb_buffer_read_f64", self.ctx, self.dptr, self.cpu.ctypes.data, self._n)
        L.call("pb_buffer_free", self.ctx, self.dptr)
        self.dptr, self.state = None, ClBuffer.CPU

    def gpu_buffer(self):
        if self.state == ClBuffer.CPU:
            raise L.PlasticityError("Requested GPU buffer when in CPU state!")
        return self.dptr

    def resize(self, new_size: int, default_value: float = float("nan")) -> None:
        if new_size == self.size():
            return
        was_gpu = self.state == ClBuffer.GPU
        self.MoveToCpu()
        old = self.cpu
        self.cpu = np.full(new_size, default_value, dtype=np.float64)
        self.cpu[:min(old.size, new_size)] = old[:new_size]
        if was_gpu:
            self.MoveToGpu()

    def DeepClone(self) -> "ClBuffer":  # cl_buffer.h:52-65: device-side copy
        out = ClBuffer(0, self.ctx)
        if self.state == ClBuffer.CPU:
            out.cpu = self.cpu.copy()
            return out
        p = C.c_void_p()
        L.call("pb_buffer_alloc", self.ctx, self._n, C.byref(p))
        L.call("pb_buffer_copy", self.ctx, p, self.dptr, self._n)
        out.dptr, out._n, out.state = p, self._n, ClBuffer.GPU
        return out

    def _host(self) -> np.ndarray:
        if self.state == ClBuffer.GPU:
            raise L.PlasticityError("Error: [] used while buffer is in GPU.")
        return self.cpu

    def __getitem__(self, i):
        return self._host()[i]

    def __setitem__(self, i, v):
        self._host()[i] = v

    def at(self, i):
        return float(self._host()[i])

    def numpy(self) -> np.ndarray:
        """Host copy without changing the buffer's location."""
        if self.state == ClBuffer.CPU:
            return self.cpu.copy()
        out = np.empty(self._n, dtype=np.float64)
        if self._n:
            L.call("pb_buffer_read_f64", self.ctx, self.dptr, out.ctypes.data, self._n)
        return out

    def __del__(self):
        try:
            if self.state == ClBuffer.GPU and self.dptr is not None:
                L.load().pb_buffer_free(self.ctx, self.dptr)
        except Exception:
            pass


class Nnet:
    """nnet::Nnet over pb_nnet_*."""

    class LearningParameters:
        def __init__(self, learning_rate: float, dynamic_learning_rate: bool = False):
            self.learning_rate = learning_rate
            self.dynamic_learning_rate = dynamic_learning_rate

    def __init__(self, model: Architecture, weight_initialization: int = Xavier,
                 loss_function: int = MeanSquared, max_batch: int = 1, device: int = 0,
                 seed: int = 0, ctx=None):
        if not model.VerifyArchitecture():
            raise L.PlasticityError("Invalid dimensions passed to Nnet()")  # nnet.h:55-59
        # The reference copies the Architecture (deep copy of every Layer and its
        # CPU weights, nnet/layer.cc:43-48); later edits of `model` do not leak in.
        self.model_ = Architecture()
        for src in model.layers:
            dst = Layer(src.desc, src.type_name, src.index)
            dst.weights[:] = src.weights
            self.model_.layers.append(dst)
        self.loss = loss_function
        self.ctx = ctx if ctx is not None else context(device)
        self.max_batch = max_batch
        descs = (L.LayerDesc * len(model.layers))(*[l.desc for l in model.layers])
        self.h = C.c_void_p()
        L.call("pb_nnet_create", self.ctx, descs, len(model.layers), loss_function, max_batch,
               C.byref(self.h))
        self.ni = [l.num_inputs for l in model.layers]
        self.no = [l.num_outputs for l in model.layers]
        self.nw = [l.weights.size for l in model.layers]
        self._host_dirty = [True] * len(model.layers)  # host copy newer than device
        self._dev_dirty = [False] * len(model.layers)  # device copy newer than host
        if weight_initialization == Xavier:
            L.call("pb_nnet_init_xavier", self.h, seed)
            self._host_dirty = [False] * len(model.layers)
            self._dev_dirty = [n > 0 for n in self.nw]
        elif weight_initialization == InitToOne:
            L.call("pb_nnet_init_constant", self.h, 1.0)
            self._host_dirty = [False] * len(model.layers)
            self._dev_dirty = [n > 0 for n in self.nw]
        elif weight_initialization != NoWeightInit:
            raise L.PlasticityError("Error: Unknown initialization strategy passed to "
                                    "nnet::Nnet constructor.")

    def __del__(self):
        try:
            if getattr(self, "h", None):
                L.load().pb_nnet_destroy(self.h)
        except Exception:
            pass

    # -- buffers ------------------------------------------------------------
    def MakeBuffer(self, values=0) -> ClBuffer:  # nnet.h:102-122
        return ClBuffer(values, self.ctx)

    def RegisterBuffer(self, buffer: ClBuffer) -> None:  # nnet.h:124-127
        buffer.RegisterClBackend(self.ctx)

    # -- weights ------------------------------------------------------------
    def _sync_to_device(self) -> None:  # LoadWeightsToGpu, nnet.h:490-496
        for i, layer in enumerate(self.model_.layers):
            if self.nw[i] and self._host_dirty[i]:
                L.call("pb_nnet_set_weights_f64", self.h, i, layer.weights.ctypes.data, self.nw[i])
                self._host_dirty[i] = False

    def _sync_to_host(self, i: int) -> None:
        if self.nw[i] and self._dev_dirty[i]:
            w = self.model_.layers[i].weights
            L.call("pb_nnet_get_weights_f64", self.h, i, w.ctypes.data, self.nw[i])
            self._dev_dirty[i] = False

    def model(self) -> Architecture:  # nnet.h:130-135 (weights moved to the CPU)
        for i in range(len(self.model_.layers)):
            self._sync_to_host(i)
        return self.model_

    def layer(self, i: int) -> Layer:
        self._sync_to_host(i)
        return self.model_.layers[i]

    def number_of_layers(self) -> int:
        return len(self.model_.layers)

    def GetWeight(self, layer: int, weight_index: int) -> float:  # nnet.h:199-202
        self._sync_to_host(layer)
        return self.model_.layers[layer].W(weight_index)

    def SetWeight(self, layer: int, weight_index: int, value: float) -> None:
        """`GetWeight(l, i) = v` in the reference (it returns a double&)."""
        self._sync_to_host(layer)
        self.model_.layers[layer].SetW(weight_index, value)
        self._host_dirty[layer] = True

    def SetLayerWeights(self, layer: int, values) -> None:
        w = self.model_.layers[layer].weights
        w[:] = np.asarray(values, dtype=np.float64).reshape(-1)
        self._host_dirty[layer], self._dev_dirty[layer] = True, False

    def GetLayerWeights(self, layer: int) -> np.ndarray:
        self._sync_to_host(layer)
        return self.model_.layers[layer].weights.copy()

    def SetLearningParameters(self, params) -> None:  # nnet.h:357-361
        lr = params.learning_rate if hasattr(params, "learning_rate") else float(params)
        L.call("pb_nnet_set_learning_rate", self.h, lr)

    # -- inference / training ------------------------------------------------
    def _claim(self, buf: ClBuffer) -> None:  # ClaimOwnership, nnet.h:911-920
        if buf.ctx is None or buf.ctx.value != self.ctx.value:
            buf.RegisterClBackend(self.ctx)
        buf.MoveToGpu()

    def Evaluate(self, inputs: ClBuffer, batch: int = 1) -> ClBuffer:  # nnet.h:204-273
        self._claim(inputs)
        self._sync_to_device()
        out = ClBuffer(0, self.ctx)
        p = C.c_void_p()
        L.call("pb_buffer_alloc", self.ctx, self.no[-1] * batch, C.byref(p))
        out.dptr, out._n, out.state = p, self.no[-1] * batch, ClBuffer.GPU
        L.call("pb_nnet_evaluate", self.h, inputs.gpu_buffer(), batch, p)
        L.call("pb_context_finish", self.ctx)  # queue->finish(), nnet.h:270
        return out

    def LayerOutput(self, layer: int, batch: int = 1) -> np.ndarray:
        """eval_layer_outputs_[layer] of the last Evaluate/Train (host copy)."""
        p = C.c_void_p()
        L.call("pb_nnet_layer_output", self.h, layer, C.byref(p))
        out = np.empty(self.no[layer] * batch, dtype=np.float64)
        L.call("pb_buffer_read_f64", self.ctx, p, out.ctypes.data, out.size)
        return out

    def Error(self, actual_output: ClBuffer, expected: ClBuffer) -> float:  # nnet.h:280-314
        self._claim(actual_output)
        self._claim(expected)
        e = C.c_double()
        L.call("pb_nnet_error", self.h, actual_output.gpu_buffer(), expected.gpu_buffer(),
               C.byref(e))
        return e.value

    def Train(self, inputs: ClBuffer, expected: Optional[ClBuffer],
              input_gradients: Optional[ClBuffer] = None,
              initial_backprop_gradients: Optional[ClBuffer] = None) -> None:
        """Nnet::Train, all three overloads (nnet.h:351-355, 369-464, 469-483)."""
        self._claim(inputs)
        if expected is not None:
            self._claim(expected)
        self._sync_to_device()
        ig = None
        if input_gradients is not None:
            # `*input_gradients = *backprop_gradients_` aliases a buffer of
            # max_layer_output_size in the reference (nnet.h:461-463); here the
            # caller's buffer becomes exactly num_inputs long.
            if input_gradients.ctx is None or input_gradients.ctx.value != self.ctx.value:
                input_gradients.RegisterClBackend(self.ctx)
            input_gradients.resize(self.ni[0], 0.0)
            input_gradients.MoveToGpu()
            ig = input_gradients.gpu_buffer()
        init = None
        if initial_backprop_gradients is not None:
            self._claim(initial_backprop_gradients)
            init = initial_backprop_gradients.gpu_buffer()
        L.call("pb_nnet_train", self.h, inputs.gpu_buffer(),
               expected.gpu_buffer() if expected is not None else None, ig, init)
        self._dev_dirty = [n > 0 for n in self.nw]

    def TrainLoop(self, inputs: ClBuffer, expected: ClBuffer, n_samples: int) -> None:
        """The per-sample SGD loop of nnet/cifar_test.cc:423-428 over a resident
        data set (inputs/expected hold n_samples samples, sample-major)."""
        self._claim(inputs)
        self._claim(expected)
        self._sync_to_device()
        L.call("pb_nnet_train_loop", self.h, inputs.gpu_buffer(), expected.gpu_buffer(), n_samples)
        self._dev_dirty = [n > 0 for n in self.nw]

    def BatchTrain(self, inputs: ClBuffer, expected: ClBuffer, batch: int, comm=None) -> None:
        """Nnet::BatchTrain (nnet.h:617-669): W <- W - lr * sum_b grad_b(W_old).
        inputs / expected hold `batch` samples, sample-major."""
        self._claim(inputs)
        self._claim(expected)
        self._sync_to_device()
        if comm is None:
            L.call("pb_nnet_batch_train", self.h, inputs.gpu_buffer(), expected.gpu_buffer(), batch)
        else:
            L.call("pb_nnet_batch_train_dp", self.h, comm, inputs.gpu_buffer(),
                   expected.gpu_buffer(), batch)
        self._dev_dirty = [n > 0 for n in self.nw]

    def BatchTrainHost(self, h_inputs, h_expected, batch: int, h_outputs=None, comm=None) -> None:
        """BatchTrain fed from pinned HOST memory (raw pointers from pb_host_alloc),
        asynchronous and double-buffered (pb_nnet_batch_train_host): the H2D copy of
        the next call overlaps this call's compute; h_outputs (nullable) receives the
        step's pre-update network outputs.  Finish() waits for everything queued."""
        self._sync_to_device()
        L.call("pb_nnet_batch_train_host", self.h, comm, h_inputs, h_expected, batch, h_outputs)
        self._dev_dirty = [n > 0 for n in self.nw]

    def BatchTrainHostRecords(self, h_records, batch: int, h_outputs=None, comm=None) -> None:
        """BatchTrainHost over labelled byte records (1 label byte + input_size pixel
        bytes each, the CIFAR-10 binary format) in pinned host memory; normalisation and
        one-hot encoding happen on the device (cifar_test.cc:66-97)."""
        self._sync_to_device()
        L.call("pb_nnet_batch_train_host_records", self.h, comm, h_records, batch, h_outputs)
        self._dev_dirty = [n > 0 for n in self.nw]

    def Finish(self) -> None:
        L.call("pb_context_finish", self.ctx)

    # -- weight file I/O (nnet.h:671-789) ---------------------------------------
    def WeightsToString(self) -> str:
        layers = []
        for i, layer in enumerate(self.model().layers):
            layers.append({"name": layer.LayerSuffix(), "index": i,
                           "weights": [float(w) for w in layer.weights]})
        try:
            # allow_nan=False: NaN / Infinity are not JSON; a diverged network must not write a
            # file that neither implementation can load back (rapidjson's Writer refuses too)
            return json.dumps({"version": kWeightFileFormatVersion,
                               "number_of_layers": len(layers), "layers": layers},
                              separators=(",", ":"), allow_nan=False)
        except ValueError as exc:
            raise L.PlasticityError("WeightsToString: non-finite weight (the network diverged); "
                                    "refusing to write an unloadable weight file") from exc

    def LoadWeightsFromString(self, weight_string: str) -> bool:
        try:
            d = json.loads(weight_string)
        except ValueError:
            return False
        if not isinstance(d, dict) or not isinstance(d.get("version"), str):
            return False
        if d["version"] != kWeightFileFormatVersion:
            return False
        if not isinstance(d.get("number_of_layers"), int) or \
                d["number_of_layers"] != len(self.model_.layers):
            return False
        layers = d.get("layers")
        if not isinstance(layers, list) or len(layers) != len(self.model_.layers):
            return False
        for i, (entry, layer) in enumerate(zip(layers, self.model_.layers)):
            if not isinstance(entry, dict) or entry.get("name") != layer.LayerSuffix():
                return False
            if entry.get("index") != i:
                return False
            w = entry.get("weights")
            if not isinstance(w, list) or len(w) != layer.weights.size:
                return False
            if any(not isinstance(x, float) for x in w):  # "Weight is not a double!"
                return False
        for i, entry in enumerate(layers):
            if self.nw[i]:
                self.SetLayerWeights(i, entry["weights"])
        return True
