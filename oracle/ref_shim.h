// TEST INFRASTRUCTURE — not part of the product path.
//
// Lets the OpenCL C text emitted by the reference's generator compile as host
// C++: address-space qualifiers vanish, get_global_id reads a thread-local
// that REF_NDRANGE sets, i.e. the loop over work-items that
// cl::CommandQueue::enqueueNDRangeKernel (nnet/nnet.h:251-253) would run.
#ifndef ORACLE_REF_SHIM_H_
#define ORACLE_REF_SHIM_H_
#include <cmath>
#include <cstddef>
#include <cstdio>

using std::isinf;  // max_pool_layer.cc:80 emits a bare isinf()

// REF_FP32: the SAME generated text evaluated in single precision (every `double` of the
// emitted kernels becomes `float`; built with -fsingle-precision-constant so the printed
// constants are float too).  Not what the reference runs - it exists to tell which max-pool
// windows an fp32 evaluation in the reference's own summation order routes differently
// from fp64 (tests/test_gpu_parity.py::test_pool_routing_unfiltered_inputs).  The system
// headers above are already parsed; only the generated text and the driver see the macro.
#ifdef REF_FP32
#define double float
#endif

#define global
#define kernel

static thread_local size_t ref_gid__ = 0;
static inline size_t get_global_id(int) { return ref_gid__; }

extern "C" int ref_parallel_threshold;  // defined in ref_driver.cc

// Work-items are independent (each writes exactly one output element), so the
// loop may run in any order; OpenMP over host cores stands in for POCL's
// pthread driver.  Small ranges stay serial: forking threads for 10 items is
// slower than the items.
#define REF_NDRANGE(N, CALL)                                            \
  do {                                                                  \
    const long n__ = (long)(N);                                         \
    if (n__ >= ref_parallel_threshold) {                                \
      _Pragma("omp parallel for schedule(static)") for (long i__ = 0;   \
                                                        i__ < n__;      \
                                                        ++i__) {        \
        ref_gid__ = (size_t)i__;                                        \
        CALL;                                                           \
      }                                                                 \
    } else {                                                            \
      for (long i__ = 0; i__ < n__; ++i__) {                            \
        ref_gid__ = (size_t)i__;                                        \
        CALL;                                                           \
      }                                                                 \
    }                                                                   \
  } while (0)

#endif  // ORACLE_REF_SHIM_H_
