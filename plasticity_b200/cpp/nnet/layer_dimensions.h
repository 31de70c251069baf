// nnet/layer_dimensions.h:8-51 of the reference: same names, same field order.
#ifndef PLASTICITY_B200_LAYER_DIMENSIONS_H_
#define PLASTICITY_B200_LAYER_DIMENSIONS_H_

#include <cstddef>
#include <initializer_list>

namespace nnet {

enum LossFunction {
  MeanSquared = 0,
  CrossEntropy,
};

struct Dimensions {
  size_t num_inputs;
  size_t num_outputs;
};
using LinearDimensions = Dimensions;

struct FilterParams {
  size_t width, height, depth;  // of each filter
  size_t stride;
  size_t padding;               // zeros added to EACH side of the input
  size_t num_filters;
};

struct VolumeDimensions {
  size_t width, height, depth;
};

struct AreaDimensions {
  size_t width, height;
};

// nnet/layer_dimensions.cc:12-54: the NDRange work-group size the reference picks for a
// launch of n work-items — n itself up to 128; otherwise the smallest divisor of n in
// [32, min(n/2, 1024)); otherwise the first of 96, 128, 160, 100, 1024, 512, 256 that
// divides n; otherwise 1.  The CUDA kernels choose their own launch shapes; the value is
// kept because Layer::{eval,weight_train,bp_train}_workgroup_size() are public.
inline size_t CalculateWorkgroupSize(size_t global_num_tasks) {
  if (global_num_tasks <= 128) return global_num_tasks ? global_num_tasks : 1;
  const size_t half = global_num_tasks / 2, limit = half < 1024 ? half : 1024;
  for (size_t s = 32; s < limit; ++s)
    if (global_num_tasks % s == 0) return s;
  for (size_t s : {96u, 128u, 160u, 100u, 1024u, 512u, 256u})
    if (global_num_tasks % s == 0) return s;
  return 1;
}

}  // namespace nnet
#endif
