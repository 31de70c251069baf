// nnet::Architecture — the fluent model builder of nnet/architecture.h:13-122,
// unchanged in names, defaults and implicit layers: the input layer is an
// Identity activation (:45-47), Dense adds dense + activation (default
// Sigmoid, :49-74), Convolution adds conv + activation (default LeakyRelu,
// :77-85), VerifyArchitecture checks only flat in/out counts (:16-32).
#ifndef PLASTICITY_B200_NNET_ARCHITECTURE_H_
#define PLASTICITY_B200_NNET_ARCHITECTURE_H_

#include <sstream>
#include <string>
#include <vector>

#include "nnet/layer.h"
#include "nnet/layer_dimensions.h"
#include "symbolic/expression.h"
#include "symbolic/symbolic_util.h"

namespace nnet {

struct Architecture {
  std::vector<Layer> layers;

  bool VerifyArchitecture() const {
    for (size_t i = 1; i < layers.size(); ++i) {
      const size_t prev_output = layers[i - 1].GetDimensions().num_outputs;
      const size_t curr_input = layers[i].GetDimensions().num_inputs;
      if (prev_output != curr_input) {
        std::cerr << "Dimension mismatch between output of layer " << i - 1 << " (" << prev_output
                  << ") and input of layer " << i << " (" << curr_input << ")." << std::endl;
        return false;
      }
    }
    return true;
  }

  Architecture(size_t input_size) { AddInputLayer(input_size); }
  Architecture() {}

  using ActivationFunctionType = Layer::ActivationFunctionType;

  Architecture &AddInputLayer(size_t size) { return AddActivationLayer(size, symbolic::Identity); }

  Architecture &AddDenseLayer(size_t num_outputs,
                              const ActivationFunctionType &activation_function) {
    Dimensions dimensions = {layers[layers.size() - 1].GetDimensions().num_outputs, num_outputs};
    layers.push_back(Layer::MakeDenseLayer(layers.size(), dimensions));
    AddActivationLayer(activation_function);
    return *this;
  }

  Architecture &AddDenseLayer(size_t num_outputs) {
    return AddDenseLayer(num_outputs, symbolic::Sigmoid);
  }

  Architecture &AddConvolutionLayer(
      const VolumeDimensions &dimensions, const FilterParams &params,
      const ActivationFunctionType &activation_function = symbolic::LeakyRelu) {
    layers.push_back(Layer::MakeConvolutionLayer(layers.size(), dimensions, params));
    const size_t conv_outputs = layers[layers.size() - 1].GetDimensions().num_outputs;
    AddActivationLayer(conv_outputs, activation_function);
    return *this;
  }

  Architecture &AddSoftmaxLayer(size_t size) {
    layers.push_back(Layer::MakeSoftmaxLayer(layers.size(), size));
    return *this;
  }

  Architecture &AddActivationLayer(size_t size, const ActivationFunctionType &activation_function) {
    layers.push_back(Layer::MakeActivationLayer(layers.size(), size, activation_function));
    return *this;
  }

  Architecture &AddActivationLayer(const ActivationFunctionType &activation_function) {
    return AddActivationLayer(layers[layers.size() - 1].GetDimensions().num_outputs,
                              activation_function);
  }

  Architecture &AddMaxPoolLayer(const VolumeDimensions &input, const AreaDimensions &output) {
    layers.push_back(Layer::MakeMaxPoolLayer(layers.size(), input, output));
    return *this;
  }

  std::string to_string() const {
    std::stringstream buffer;
    buffer << "Architecture{ inputs: " << input_size() << ", outputs: " << output_size() << "}";
    return buffer.str();
  }

  size_t input_size() const { return layers[0].GetDimensions().num_inputs; }
  size_t output_size() const { return layers[layers.size() - 1].GetDimensions().num_outputs; }
};

}  // namespace nnet
#endif
