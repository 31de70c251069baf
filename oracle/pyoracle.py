"""TEST INFRASTRUCTURE — ctypes front-end for the two CPU checkers.

* ``RestatedNet``  -> oracle/liboracle.so   (plain-C restatement, plasticity_oracle.c)
* ``RefNet``       -> oracle/_ref/libref_<spec>.so (the reference's own generator
  output compiled for the host by oracle/Makefile; exists only for the specs
  under oracle/specs/)

Both expose the same methods over flat float64 numpy arrays so tests can run
one against the other.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product
package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ACTS = {"identity": 0, "relu": 1, "leaky": 2, "sigmoid": 3}
KINDS = {"act": 0, "dense": 1, "conv": 2, "pool": 3, "softmax": 4}
LOSSES = {"mse": 0, "ce": 1}

_dp = C.POINTER(C.c_double)


def _p(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


class PoLayer(C.Structure):
    _fields_ = [("kind", C.c_int), ("act", C.c_int), ("n", C.c_long),
                ("in_", C.c_long), ("out", C.c_long),
                ("W", C.c_long), ("H", C.c_long), ("D", C.c_long),
                ("fw", C.c_long), ("fh", C.c_long), ("fd", C.c_long),
                ("stride", C.c_long), ("pad", C.c_long), ("nf", C.c_long),
                ("ow", C.c_long), ("oh", C.c_long)]


def parse_spec(text: str) -> Tuple[int, List[PoLayer]]:
    """Spec grammar: see oracle/ref_gen.cc header."""
    loss = 0
    layers: List[PoLayer] = []
    for raw in text.splitlines():
        t = raw.split("#")[0].split()
        if not t:
            continue
        k, a = t[0], t[1:]
        L = PoLayer()
        if k == "loss":
            loss = LOSSES[a[0]]
            continue
        L.kind = KINDS[k]
        if k == "act":
            L.n, L.act = int(a[0]), ACTS[a[1]]
        elif k == "dense":
            L.in_, L.out = int(a[0]), int(a[1])
        elif k == "conv":
            (L.W, L.H, L.D, L.fw, L.fh, L.fd, L.stride, L.pad, L.nf) = map(int, a)
        elif k == "pool":
            (L.W, L.H, L.D, L.ow, L.oh) = map(int, a)
        elif k == "softmax":
            L.n = int(a[0])
        layers.append(L)
    return loss, layers


def build(verbose: bool = False) -> None:
    """Compile liboracle.so (+ _ref/ when /root/reference is present)."""
    r = subprocess.run(["make", "-C", HERE, "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)


class _NetBase:
    n_layers: int
    ni: List[int]
    no: List[int]
    nw: List[int]
    loss: int

    def _finish_shape(self):
        self.w_off = np.concatenate([[0], np.cumsum(self.nw)]).astype(int)
        self.o_off = np.concatenate([[0], np.cumsum(self.no)]).astype(int)
        self.total_w = int(self.w_off[-1])
        self.total_o = int(self.o_off[-1])
        self.n_in = self.ni[0]
        self.n_out = self.no[-1]

    def layer_weights(self, weights: np.ndarray, l: int) -> np.ndarray:
        return weights[self.w_off[l]:self.w_off[l + 1]]

    def layer_output(self, layer_out: np.ndarray, l: int) -> np.ndarray:
        return layer_out[self.o_off[l]:self.o_off[l + 1]]

    def zeros_weights(self) -> np.ndarray:
        return np.zeros(self.total_w, dtype=np.float64)


class RestatedNet(_NetBase):
    """Plain-C restatement (oracle/plasticity_oracle.c)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            path = os.path.join(HERE, "liboracle.so")
            if not os.path.exists(path):
                build()
            L = C.CDLL(path)
            lp = C.POINTER(PoLayer)
            L.po_num_inputs.restype = C.c_long
            L.po_num_outputs.restype = C.c_long
            L.po_num_weights.restype = C.c_long
            for f in (L.po_num_inputs, L.po_num_outputs, L.po_num_weights):
                f.argtypes = [lp]
            L.po_evaluate.argtypes = [lp, _dp, _dp, _dp]
            L.po_input_delta.argtypes = [lp, _dp, _dp, _dp, _dp]
            L.po_weight_update.argtypes = [lp, _dp, _dp, _dp, C.c_double, C.c_int, _dp]
            L.po_error_value.argtypes = [C.c_int, C.c_long, _dp, _dp, _dp]
            L.po_error_gradients.argtypes = [C.c_int, C.c_long, _dp, _dp, _dp]
            L.po_net_evaluate.argtypes = [lp, C.c_int, _dp, _dp, _dp]
            L.po_net_error.argtypes = [lp, C.c_int, C.c_int, _dp, _dp]
            L.po_net_error.restype = C.c_double
            L.po_net_train.argtypes = [lp, C.c_int, C.c_int, _dp, _dp, _dp,
                                       C.c_double, _dp, _dp, _dp]
            L.po_net_train_loop.argtypes = [lp, C.c_int, C.c_int, _dp, _dp, C.c_long,
                                            _dp, C.c_double]
            L.po_net_batch_train.argtypes = [lp, C.c_int, C.c_int, _dp, _dp, C.c_long,
                                             _dp, C.c_double, _dp]
            L.po_net_batch_train_mt.argtypes = [lp, C.c_int, C.c_int, _dp, _dp, C.c_long,
                                                _dp, C.c_double]
            L.po_net_evaluate_batch.argtypes = [lp, C.c_int, _dp, C.c_long, _dp, _dp]
            L.po_set_threads.argtypes = [C.c_int]
            cls._lib = L
        return cls._lib

    def __init__(self, spec_text: str):
        self.loss, layers = parse_spec(spec_text)
        self.n_layers = len(layers)
        self.layers = (PoLayer * self.n_layers)(*layers)
        L = self.lib()
        self.ni = [L.po_num_inputs(C.byref(x)) for x in self.layers]
        self.no = [L.po_num_outputs(C.byref(x)) for x in self.layers]
        self.nw = [L.po_num_weights(C.byref(x)) for x in self.layers]
        self._finish_shape()

    def set_threads(self, n: int):
        self.lib().po_set_threads(n)

    # per-kernel
    def evaluate_layer(self, l, I, W):
        O = np.empty(self.no[l])
        self.lib().po_evaluate(C.byref(self.layers[l]), _p(I), _p(W), _p(O))
        return O

    def input_delta(self, l, I, W, G):
        dI = np.empty(self.ni[l])
        self.lib().po_input_delta(C.byref(self.layers[l]), _p(I), _p(W), _p(G), _p(dI))
        return dI

    def weight_update(self, l, I, W, G, lr, mode):
        out = np.empty(self.nw[l])
        self.lib().po_weight_update(C.byref(self.layers[l]), _p(I), _p(W), _p(G),
                                    lr, mode, _p(out))
        return out

    def error_gradients(self, O, E):
        g = np.empty(self.n_out)
        self.lib().po_error_gradients(self.loss, self.n_out, _p(O), _p(E), _p(g))
        return g

    # network
    def evaluate(self, x, weights):
        lo = np.empty(self.total_o)
        self.lib().po_net_evaluate(self.layers, self.n_layers, _p(x), _p(weights), _p(lo))
        return lo

    def error(self, out, expected):
        return self.lib().po_net_error(self.layers, self.n_layers, self.loss,
                                       _p(out), _p(expected))

    def train(self, x, expected, weights, lr, want_input_grads=False,
              initial_gradients=None):
        """In-place SGD step on `weights`; returns (layer_out, input_grads|None)."""
        lo = np.empty(self.total_o)
        ig = np.empty(self.n_in) if want_input_grads else None
        self.lib().po_net_train(self.layers, self.n_layers, self.loss, _p(x),
                                _p(expected), _p(weights), lr, _p(lo), _p(ig),
                                _p(initial_gradients))
        return lo, ig

    def train_loop(self, xs, expected, weights, lr):
        n = xs.shape[0]
        self.lib().po_net_train_loop(self.layers, self.n_layers, self.loss, _p(xs),
                                     _p(expected), n, _p(weights), lr)

    def batch_train(self, xs, expected, weights, lr):
        sd = np.empty(self.total_w)
        self.lib().po_net_batch_train(self.layers, self.n_layers, self.loss, _p(xs),
                                      _p(expected), xs.shape[0], _p(weights), lr, _p(sd))
        return sd

    def batch_train_mt(self, xs, expected, weights, lr):
        self.lib().po_net_batch_train_mt(self.layers, self.n_layers, self.loss, _p(xs),
                                         _p(expected), xs.shape[0], _p(weights), lr)

    def evaluate_batch(self, xs, weights):
        outs = np.empty((xs.shape[0], self.n_out))
        self.lib().po_net_evaluate_batch(self.layers, self.n_layers, _p(xs),
                                         xs.shape[0], _p(weights), _p(outs))
        return outs


class RefNet(_NetBase):
    """The reference generator's emitted kernels, host-compiled (oracle/_ref)."""

    @staticmethod
    def path(name: str, fp32: bool = False) -> str:
        return os.path.join(HERE, "_ref", f"libref{'32' if fp32 else ''}_{name}.so")

    @classmethod
    def available(cls, name: str, fp32: bool = False) -> bool:
        return os.path.exists(cls.path(name, fp32))

    def __init__(self, name: str, fp32: bool = False):
        """fp32=True: the same generated text evaluated in single precision (ref_shim.h
        REF_FP32); only evaluate() is offered there."""
        if not self.available(name, fp32):
            raise FileNotFoundError(self.path(name, fp32) + " (run `make -C oracle`)")
        L = self.L = C.CDLL(self.path(name, fp32))
        self.fp32 = fp32
        self.name = name
        self.n_layers = L.ref_num_layers()
        self.loss = L.ref_loss()
        L.ref_layer_info.argtypes = [C.c_int] + [C.POINTER(C.c_long)] * 3 + [C.POINTER(C.c_char_p)]
        self.ni, self.no, self.nw, self.names = [], [], [], []
        for l in range(self.n_layers):
            a, b, c, s = C.c_long(), C.c_long(), C.c_long(), C.c_char_p()
            L.ref_layer_info(l, C.byref(a), C.byref(b), C.byref(c), C.byref(s))
            self.ni.append(a.value), self.no.append(b.value), self.nw.append(c.value)
            self.names.append(s.value.decode())
        self._finish_shape()
        L.ref_evaluate.argtypes = [C.c_int, _dp, _dp, _dp]
        L.ref_input_delta.argtypes = [C.c_int, _dp, _dp, _dp, _dp]
        L.ref_new_weights.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp]
        L.ref_weight_deltas.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp]
        L.ref_error_gradients.argtypes = [_dp, _dp, _dp]
        L.ref_net_evaluate.argtypes = [_dp, _dp, _dp]
        L.ref_net_error.argtypes = [_dp, _dp]
        L.ref_net_error.restype = C.c_double
        L.ref_net_train.argtypes = [_dp, _dp, _dp, C.c_double, _dp, _dp, _dp, C.c_int]
        L.ref_net_train_loop.argtypes = [_dp, _dp, C.c_long, _dp, C.c_double, C.c_int]
        L.ref_net_batch_train.argtypes = [_dp, _dp, C.c_long, _dp, C.c_double, _dp]
        L.ref_net_batch_train_mt.argtypes = [_dp, _dp, C.c_long, _dp, C.c_double]

    def set_parallel_threshold(self, n: int):
        C.c_int.in_dll(self.L, "ref_parallel_threshold").value = n

    def evaluate_layer(self, l, I, W):
        O = np.empty(self.no[l])
        W = W if W is not None and W.size else np.zeros(1)
        self.L.ref_evaluate(l, _p(I), _p(W), _p(O))
        return O

    def input_delta(self, l, I, W, G):
        dI = np.empty(self.ni[l])
        W = W if W is not None and W.size else np.zeros(1)
        self.L.ref_input_delta(l, _p(I), _p(W), _p(G), _p(dI))
        return dI

    def weight_update(self, l, I, W, G, lr, mode):
        out = np.empty(self.nw[l])
        lrb = np.array([lr], dtype=np.float64)
        fn = self.L.ref_new_weights if mode == 0 else self.L.ref_weight_deltas
        fn(l, _p(I), _p(W), _p(G), _p(out), _p(lrb))
        return out

    def error_gradients(self, O, E):
        g = np.empty(self.n_out)
        self.L.ref_error_gradients(_p(O), _p(E), _p(g))
        return g

    def evaluate(self, x, weights):
        w = weights if weights.size else np.zeros(1)
        if self.fp32:
            fp = C.POINTER(C.c_float)
            x32 = np.ascontiguousarray(x, dtype=np.float32)
            w32 = np.ascontiguousarray(w, dtype=np.float32)
            lo32 = np.empty(self.total_o, dtype=np.float32)
            self.L.ref_net_evaluate.argtypes = [fp, fp, fp]
            self.L.ref_net_evaluate(x32.ctypes.data_as(fp), w32.ctypes.data_as(fp), lo32.ctypes.data_as(fp))
            return lo32.astype(np.float64)
        lo = np.empty(self.total_o)
        self.L.ref_net_evaluate(_p(x), _p(w), _p(lo))
        return lo

    def error(self, out, expected):
        return self.L.ref_net_error(_p(out), _p(expected))

    def train(self, x, expected, weights, lr, want_input_grads=False,
              initial_gradients=None, forwards=1):
        lo = np.empty(self.total_o)
        ig = np.empty(self.n_in) if want_input_grads else None
        w = weights if weights.size else np.zeros(1)
        self.L.ref_net_train(_p(x), _p(expected), _p(w), lr, _p(lo), _p(ig),
                             _p(initial_gradients), forwards)
        return lo, ig

    def train_loop(self, xs, expected, weights, lr, forwards=1):
        self.L.ref_net_train_loop(_p(xs), _p(expected), xs.shape[0], _p(weights), lr, forwards)

    def batch_train(self, xs, expected, weights, lr):
        sd = np.empty(self.total_w)
        self.L.ref_net_batch_train(_p(xs), _p(expected), xs.shape[0], _p(weights), lr, _p(sd))
        return sd


    def batch_train_mt(self, xs, expected, weights, lr):
        self.L.ref_net_batch_train_mt(_p(xs), _p(expected), xs.shape[0], _p(weights), lr)


def spec_text(name: str) -> str:
    with open(os.path.join(HERE, "specs", name + ".spec")) as f:
        return f.read()
