// nnet::Layer of the B200 host mirror (reference: nnet/layer.h, nnet/layer.cc).
// A layer is a pb_layer_desc (what the reference derives from its LayerImpl,
// nnet/layer_impl.h:24-43) plus its host-side weights in the reference layout.
// The kernel-source generators (GenerateEvaluationKernel ..., layer.cc:140-238)
// have no counterpart: the CUDA kernels are dispatched on desc.kind.
#ifndef PLASTICITY_B200_NNET_LAYER_H_
#define PLASTICITY_B200_NNET_LAYER_H_

#include <cstdlib>
#include <functional>
#include <iostream>
#include <string>

#include "compute/cl_buffer.h"
#include "nnet/layer_dimensions.h"
#include "plasticity_b200.h"
#include "symbolic/expression.h"
#include "symbolic/symbolic_util.h"

namespace nnet {

class Nnet;

class Layer {
 public:
  // nnet/layer_impl.h:21-22
  using ActivationFunctionType =
      std::function<symbolic::Expression(const symbolic::Expression &)>;

  Layer(const Layer &other) = default;  // deep copy incl. CPU weights (layer.cc:43-48)
  Layer(Layer &&other) = default;
  Layer &operator=(const Layer &rhs) = delete;
  Layer &operator=(Layer &&rhs) = delete;
  virtual ~Layer() {}

  static Layer MakeDenseLayer(size_t layer_index, const Dimensions &dimensions) {
    pb_layer_desc d{};
    d.kind = PB_DENSE;
    d.in = (int64_t)dimensions.num_inputs;
    d.out = (int64_t)dimensions.num_outputs;
    return Layer(d, "dense_layer", layer_index);
  }

  static Layer MakeConvolutionLayer(size_t layer_index, const VolumeDimensions &dimensions,
                                    const FilterParams &params) {
    if (params.depth != dimensions.depth) {  // convolution_layer.cc:12-16
      std::cerr << "Convolution layer input depth != filter depth. Error!" << std::endl;
      std::exit(1);
    }
    pb_layer_desc d{};
    d.kind = PB_CONVOLUTION;
    d.width = (int64_t)dimensions.width; d.height = (int64_t)dimensions.height;
    d.depth = (int64_t)dimensions.depth;
    d.filter_width = (int64_t)params.width; d.filter_height = (int64_t)params.height;
    d.filter_depth = (int64_t)params.depth;
    d.stride = (int64_t)params.stride; d.padding = (int64_t)params.padding;
    d.num_filters = (int64_t)params.num_filters;
    return Layer(d, "convolution_layer", layer_index);
  }

  static Layer MakeActivationLayer(size_t layer_index, size_t size,
                                   const ActivationFunctionType &activation_function) {
    // Identify the built-in by calling it on a probe expression.
    const symbolic::Expression r =
        activation_function(symbolic::Expression(symbolic::Expression::kProbe));
    if (r.tag() < symbolic::Expression::kIdentity || r.tag() > symbolic::Expression::kSigmoid) {
      std::cerr << "ActivationLayer: only symbolic::Identity/Relu/LeakyRelu/Sigmoid are "
                   "supported (no code generator, no fallback)" << std::endl;
      std::exit(1);
    }
    pb_layer_desc d{};
    d.kind = PB_ACTIVATION;
    d.activation = (int32_t)r.tag();
    d.n = (int64_t)size;
    return Layer(d, "activation_layer", layer_index);
  }

  static Layer MakeSoftmaxLayer(size_t layer_index, size_t size) {
    pb_layer_desc d{};
    d.kind = PB_SOFTMAX;
    d.n = (int64_t)size;
    return Layer(d, "softmax_layer", layer_index);
  }

  static Layer MakeMaxPoolLayer(size_t layer_index, const VolumeDimensions &input,
                                const AreaDimensions &output) {
    pb_layer_desc d{};
    d.kind = PB_MAX_POOL;
    d.width = (int64_t)input.width; d.height = (int64_t)input.height; d.depth = (int64_t)input.depth;
    d.out_width = (int64_t)output.width; d.out_height = (int64_t)output.height;
    return Layer(d, "max_pool_layer", layer_index);
  }

  double &W(size_t index) {  // layer.h:74-81
    if (index >= weights_.size()) {
      std::cerr << "Too large weight index: " << index << std::endl;
      std::exit(1);
    }
    weights_.MoveToCpu();
    return weights_[index];
  }

  compute::ClBuffer &weight_buffer() { return weights_; }
  const compute::ClBuffer &weight_buffer() const { return weights_; }

  Dimensions GetDimensions() const {
    int64_t a = 0, b = 0;
    pb_layer_num_inputs(&desc_, &a);
    pb_layer_num_outputs(&desc_, &b);
    return Dimensions{(size_t)a, (size_t)b};
  }

  // nnet/layer.h:119-123, layer.cc:28-34: the OpenCL work-group sizes of the layer's
  // three launches (outputs / weights / inputs work-items).
  size_t eval_workgroup_size() const { return CalculateWorkgroupSize(GetDimensions().num_outputs); }
  size_t weight_train_workgroup_size() const { return CalculateWorkgroupSize(weights_.size()); }
  size_t bp_train_workgroup_size() const { return CalculateWorkgroupSize(GetDimensions().num_inputs); }

  std::string LayerSuffix() const { return type_ + "_" + std::to_string(index_); }  // layer.h:99-101
  // Kernel names of the reference's generated program (layer.h:91-114): what a
  // compute::CommandQueue is asked to launch (compute/cuda_command_queue.h parses them).
  const std::string &layer_type() const { return type_; }
  std::string EvaluateKernelName() const { return "evaluate_" + std::to_string(index_); }
  std::string InputGradientKernelName() const { return "input_delta_" + LayerSuffix(); }
  std::string WeightUpdateKernelName() const { return "new_weights_" + LayerSuffix(); }
  std::string WeightGradientKernelName() const { return "weight_delta_" + LayerSuffix(); }
  const pb_layer_desc &desc() const { return desc_; }
  size_t layer_index() const { return index_; }

 private:
  Layer(const pb_layer_desc &desc, const std::string &type, size_t index)
      : desc_(desc), type_(type), index_(index) {
    int64_t nw = 0;
    if (pb_layer_num_weights(&desc_, &nw) != 0) {
      std::cerr << pb_last_error() << std::endl;
      std::exit(1);
    }
    weights_ = compute::ClBuffer(std::vector<double>((size_t)nw, 0.0));
  }

  pb_layer_desc desc_;
  std::string type_;
  size_t index_;
  compute::ClBuffer weights_;
};

}  // namespace nnet
#endif
