// nnet::Yolov1 — the class nnet/yolov1.h:22-31 declares (SetParams / Evaluate /
// TrainSample over a private Nnet) with the layer list of nnet/yolov1.cc:7-231.
// The reference's files do not compile (no return types, undefined kInputSize, a
// stray `;` before the last two builder calls) and the listed dimensions do not
// chain; the three repairs of SURVEY.md section 8(d) are applied and marked below.
// Activations are the builder defaults: LeakyRelu behind every convolution
// (architecture.h:77-85), Sigmoid behind AddDenseLayer(n) (architecture.h:63-74).
#ifndef PLASTICITY_B200_NNET_YOLOV1_H_
#define PLASTICITY_B200_NNET_YOLOV1_H_

#include <memory>
#include <string>
#include <vector>

#include "nnet/nnet.h"

namespace nnet {

constexpr size_t kYoloInputWidth = 448;
constexpr size_t kYoloOutputSize = 7 * 7 * 30;  // yolov1.cc:231

// input_width must be a multiple of 64; 448 is the reference's size, smaller widths
// give the same layer list over smaller images (tests).
inline Architecture YoloArchitecture(size_t input_width = kYoloInputWidth,
                                     size_t dense_width = 4096,
                                     size_t outputs = kYoloOutputSize) {
  size_t w = input_width;
  Architecture model(w * w * 3);
  model.AddConvolutionLayer({w, w, 3}, {7, 7, 3, 2, 3, 64});
  w /= 2;
  model.AddMaxPoolLayer({w, w, 64}, {w / 2, w / 2});
  w /= 2;
  model.AddConvolutionLayer({w, w, 64}, {3, 3, 64, 1, 1, 192});  // repaired: padding 1 (was 2)
  model.AddMaxPoolLayer({w, w, 192}, {w / 2, w / 2});
  w /= 2;
  model.AddConvolutionLayer({w, w, 192}, {1, 1, 192, 1, 0, 128})
      .AddConvolutionLayer({w, w, 128}, {3, 3, 128, 1, 1, 256})
      .AddConvolutionLayer({w, w, 256}, {1, 1, 256, 1, 0, 256})
      .AddConvolutionLayer({w, w, 256}, {3, 3, 256, 1, 1, 512})
      .AddMaxPoolLayer({w, w, 512}, {w / 2, w / 2});
  w /= 2;
  for (size_t i = 0; i < 4; ++i) {
    model.AddConvolutionLayer({w, w, 512}, {1, 1, 512, 1, 0, 256})
        .AddConvolutionLayer({w, w, 256}, {3, 3, 256, 1, 1, 512});
  }
  model.AddConvolutionLayer({w, w, 512}, {1, 1, 512, 1, 0, 512})
      .AddMaxPoolLayer({w, w, 512}, {w / 2, w / 2});
  w /= 2;
  size_t depth = 512;
  for (size_t i = 0; i < 2; ++i) {
    // repaired: the second pair reads the 1024 planes the 3x3 before it wrote (was 512)
    model.AddConvolutionLayer({w, w, depth}, {1, 1, depth, 1, 0, 512})
        .AddConvolutionLayer({w, w, 512}, {3, 3, 512, 1, 1, 1024});
    depth = 1024;
  }
  model.AddConvolutionLayer({w, w, 1024}, {3, 3, 1024, 1, 1, 1024})
      .AddConvolutionLayer({w, w, 1024}, {3, 3, 1024, 2, 1, 1024});
  w /= 2;
  model.AddConvolutionLayer({w, w, 1024}, {3, 3, 1024, 1, 1, 1024})
      .AddConvolutionLayer({w, w, 1024}, {3, 3, 1024, 1, 1, 1024})
      .AddDenseLayer(dense_width)
      .AddDenseLayer(outputs);  // repaired: the model ends here (yolov1.cc:232-235 dropped)
  return model;
}

class Yolov1 {
 public:
  explicit Yolov1(size_t max_batch = 1, size_t input_width = kYoloInputWidth,
                  size_t dense_width = 4096, size_t outputs = kYoloOutputSize)
      : net_(YoloArchitecture(input_width, dense_width, outputs), Nnet::Xavier, MeanSquared,
             max_batch) {}

  // Learning rate and, optionally, a weight file in the reference's JSON format.
  bool SetParams(double learning_rate, const std::string &weights = "") {
    net_.SetLearningParameters(Nnet::LearningParameters{learning_rate});
    return weights.empty() || net_.LoadWeightsFromString(weights);
  }

  // images: k * (3 * W * W) planar RGB values -> k * outputs grid values.
  std::unique_ptr<compute::ClBuffer> Evaluate(const std::unique_ptr<compute::ClBuffer> &images) {
    return net_.Evaluate(images);
  }

  // One per-sample SGD step on the squared error of the output grid.
  void TrainSample(std::unique_ptr<compute::ClBuffer> &image,
                   std::unique_ptr<compute::ClBuffer> &expected) {
    net_.Train(image, expected);
  }

  // One gradient-sum minibatch step over `batch` samples stored back to back (the batched
  // counterpart of TrainSample: Nnet::BatchTrain semantics, nnet.h:617-669).
  void TrainBatch(std::unique_ptr<compute::ClBuffer> &images,
                  std::unique_ptr<compute::ClBuffer> &expected, size_t batch) {
    net_.BatchTrain(images, expected, batch);
  }

  Nnet &net() { return net_; }

 private:
  Nnet net_;
};

}  // namespace nnet
#endif
