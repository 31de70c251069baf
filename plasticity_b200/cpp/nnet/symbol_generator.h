// Weight-index helpers of nnet/symbol_generator.h (the layout contract):
//   dense  row-major [out][in+1], bias last          (:58-86, 141-156)
//   conv   [filter][z][row][col] + bias per filter   (:229-249, 339-354)
// Only WeightNumber() survives — the Expression-producing members belonged to
// the code generator.
#ifndef PLASTICITY_B200_SYMBOL_GENERATOR_H_
#define PLASTICITY_B200_SYMBOL_GENERATOR_H_

#include "nnet/layer_dimensions.h"

namespace nnet {

class DenseSymbolGenerator {
 public:
  explicit DenseSymbolGenerator(Dimensions dimensions) : dimensions_(dimensions) {}
  size_t WeightNumber(size_t node_idx, size_t edge_idx) const {
    return node_idx * (dimensions_.num_inputs + 1) + edge_idx;
  }
  size_t WeightNumber(size_t node) const {  // bias of `node`
    return node * (dimensions_.num_inputs + 1) + dimensions_.num_inputs;
  }

 private:
  Dimensions dimensions_;
};

class ConvSymbolGenerator {
 public:
  ConvSymbolGenerator(const VolumeDimensions &dimensions, const FilterParams &params)
      : dimensions_(dimensions), params_(params) {}
  size_t WeightNumber(size_t filter, size_t row, size_t col, size_t z) const {
    const size_t filter_size = params_.width * params_.height * params_.depth + 1;
    return filter * filter_size + params_.width * params_.height * z + row * params_.width + col;
  }
  size_t WeightNumber(size_t filter) const {  // bias: last slot of the filter block
    const size_t filter_size = params_.width * params_.height * params_.depth + 1;
    return filter * filter_size + (filter_size - 1);
  }

 private:
  VolumeDimensions dimensions_;
  FilterParams params_;
};

}  // namespace nnet
#endif
