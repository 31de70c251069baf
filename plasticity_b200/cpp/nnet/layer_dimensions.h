// nnet/layer_dimensions.h:8-51 of the reference: same names, same field order.
#ifndef PLASTICITY_B200_LAYER_DIMENSIONS_H_
#define PLASTICITY_B200_LAYER_DIMENSIONS_H_

#include <cstddef>

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

}  // namespace nnet
#endif
