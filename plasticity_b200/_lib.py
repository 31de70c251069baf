"""ctypes binding of include/plasticity_b200.h.  Fails loudly when the CUDA
library is missing — there is no fallback of any kind."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# PB_LIB_PATH: an alternative build of the same library (kernel A/B experiments)
LIB_PATH = os.environ.get("PB_LIB_PATH") or os.path.join(HERE, "lib", "libplasticity_b200.so")

fp = C.POINTER(C.c_float)
dp = C.POINTER(C.c_double)


class LayerDesc(C.Structure):
    """pb_layer_desc"""
    _fields_ = [("kind", C.c_int32), ("activation", C.c_int32), ("n", C.c_int64),
                ("in_", C.c_int64), ("out", C.c_int64),
                ("width", C.c_int64), ("height", C.c_int64), ("depth", C.c_int64),
                ("filter_width", C.c_int64), ("filter_height", C.c_int64),
                ("filter_depth", C.c_int64),
                ("stride", C.c_int64), ("padding", C.c_int64), ("num_filters", C.c_int64),
                ("out_width", C.c_int64), ("out_height", C.c_int64)]


PB_ACTIVATION, PB_DENSE, PB_CONVOLUTION, PB_MAX_POOL, PB_SOFTMAX = range(5)
PB_IDENTITY, PB_RELU, PB_LEAKY_RELU, PB_SIGMOID = range(4)
PB_MEAN_SQUARED, PB_CROSS_ENTROPY = range(2)
PB_NEW_WEIGHTS, PB_WEIGHT_DELTAS, PB_ACCUMULATE = range(3)
PB_COMM_ID_BYTES = 128

_ld = C.POINTER(LayerDesc)
_vp = C.c_void_p

# name -> argtypes; every function returns int (0 == ok) unless noted.
SIGNATURES = {
    "pb_device_count": [C.POINTER(C.c_int)],
    "pb_context_create": [C.c_int, C.POINTER(_vp)],
    "pb_context_destroy": [_vp],
    "pb_context_finish": [_vp],
    "pb_context_device": [_vp, C.POINTER(C.c_int)],
    "pb_context_stream": [_vp, C.POINTER(_vp)],
    "pb_context_launch_count": [_vp, C.POINTER(C.c_uint64)],
    "pb_buffer_alloc": [_vp, C.c_size_t, C.POINTER(_vp)],
    "pb_buffer_free": [_vp, _vp],
    "pb_buffer_write_f64": [_vp, _vp, _vp, C.c_size_t],
    "pb_buffer_read_f64": [_vp, _vp, _vp, C.c_size_t],
    "pb_buffer_write_f32": [_vp, _vp, _vp, C.c_size_t],
    "pb_buffer_read_f32": [_vp, _vp, _vp, C.c_size_t],
    "pb_buffer_write_f32_async": [_vp, _vp, _vp, C.c_size_t],
    "pb_buffer_read_f32_async": [_vp, _vp, _vp, C.c_size_t],
    "pb_buffer_copy": [_vp, _vp, _vp, C.c_size_t],
    "pb_buffer_fill": [_vp, _vp, C.c_float, C.c_size_t],
    "pb_host_alloc": [C.c_size_t, C.POINTER(_vp)],
    "pb_host_free": [_vp],
    "pb_layer_num_inputs": [_ld, C.POINTER(C.c_int64)],
    "pb_layer_num_outputs": [_ld, C.POINTER(C.c_int64)],
    "pb_layer_num_weights": [_ld, C.POINTER(C.c_int64)],
    "pb_layer_evaluate": [_vp, _ld, _vp, _vp, _vp, C.c_int64],
    "pb_layer_input_delta": [_vp, _ld, _vp, _vp, _vp, _vp, C.c_int64],
    "pb_layer_weight_update": [_vp, _ld, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int64],
    "pb_error_value": [_vp, C.c_int, C.c_int64, _vp, _vp, _vp, C.c_int64],
    "pb_error_gradients": [_vp, C.c_int, C.c_int64, _vp, _vp, _vp, C.c_int64],
    "pb_vector_accumulate": [_vp, _vp, _vp, C.c_int64],
    "pb_nnet_create": [_vp, _ld, C.c_int32, C.c_int32, C.c_int64, C.POINTER(_vp)],
    "pb_nnet_destroy": [_vp],
    "pb_nnet_num_layers": [_vp, C.POINTER(C.c_int32)],
    "pb_nnet_total_weights": [_vp, C.POINTER(C.c_int64)],
    "pb_nnet_set_weights_f64": [_vp, C.c_int32, _vp, C.c_int64],
    "pb_nnet_get_weights_f64": [_vp, C.c_int32, _vp, C.c_int64],
    "pb_nnet_layer_weights_device": [_vp, C.c_int32, C.POINTER(_vp)],
    "pb_nnet_weights_device": [_vp, C.POINTER(_vp)],
    "pb_nnet_init_xavier": [_vp, C.c_uint64],
    "pb_nnet_init_constant": [_vp, C.c_double],
    "pb_nnet_set_learning_rate": [_vp, C.c_double],
    "pb_nnet_evaluate": [_vp, _vp, C.c_int64, _vp],
    "pb_nnet_layer_output": [_vp, C.c_int32, C.POINTER(_vp)],
    "pb_nnet_error": [_vp, _vp, _vp, C.POINTER(C.c_double)],
    "pb_nnet_error_gradients": [_vp, _vp, _vp, _vp, C.c_int64],
    "pb_nnet_train": [_vp, _vp, _vp, _vp, _vp],
    "pb_nnet_train_loop": [_vp, _vp, _vp, C.c_int64],
    "pb_nnet_batch_gradients": [_vp, _vp, _vp, C.c_int64],
    "pb_nnet_deltas_device": [_vp, C.POINTER(_vp)],
    "pb_nnet_apply_deltas": [_vp],
    "pb_nnet_batch_train": [_vp, _vp, _vp, C.c_int64],
    "pb_comm_unique_id": [_vp],
    "pb_comm_create": [_vp, _vp, C.c_int, C.c_int, C.POINTER(_vp)],
    "pb_comm_create_local": [C.POINTER(_vp), C.c_int, C.POINTER(_vp)],
    "pb_comm_destroy": [_vp],
    "pb_comm_allreduce_sum": [_vp, _vp, C.c_int64],
    "pb_nnet_batch_train_dp": [_vp, _vp, _vp, _vp, C.c_int64],
    "pb_nnet_batch_train_host": [_vp, _vp, _vp, _vp, C.c_int64, _vp],
    "pb_records_decode": [_vp, _vp, _vp, _vp, C.c_int64, C.c_int64, C.c_int64],
    "pb_nnet_batch_train_host_records": [_vp, _vp, _vp, C.c_int64, _vp],
    "pb_nnet_profile_step": [_vp, _vp, _vp, C.c_int64, C.c_char_p, C.c_size_t],
    "pb_nnet_profile_evaluate": [_vp, _vp, C.c_int64, C.c_char_p, C.c_size_t],
    "pb_measure_fp32_tflops": [_vp, C.POINTER(C.c_double)],
    "pb_measure_copy_gbs": [_vp, C.c_size_t, C.POINTER(C.c_double)],
}

_lib = None


class PlasticityError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen the CUDA library; raise if it was never built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PlasticityError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). plasticity_b200 has no CPU or eager fallback.")
        lib = C.CDLL(LIB_PATH)
        lib.pb_last_error.restype = C.c_char_p
        lib.pb_abi_version.restype = C.c_int
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.argtypes = args
            fn.restype = C.c_int
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise PlasticityError(load().pb_last_error().decode())


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args))
