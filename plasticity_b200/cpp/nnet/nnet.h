// nnet::Nnet over libplasticity_b200 — the reference's runtime class
// (nnet/nnet.h:33-923) with the same public surface, re-written over the CUDA
// C ABI (include/plasticity_b200.h).  What the reference does per call with
// generated OpenCL kernels, cl::Buffer allocations and blocking transfers
//  happens behind pb_nnet_*: device-resident fp32 weights and
// per-layer activations, one CUDA stream, hand-written sm_100a kernels.
//
// Host-visible values stay `double` (nnet.h:27 `typedef double Number`).
// Layer weights follow the reference's CPU<->GPU ownership: GetWeight(),
// layer(), model() bring them to the host (and may modify them); the next
// device operation uploads what changed.
#ifndef PLASTICITY_B200_NNET_NNET_H_
#define PLASTICITY_B200_NNET_NNET_H_

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "compute/cl_buffer.h"
#include "nnet/architecture.h"
#include "nnet/layer.h"
#include "nnet/layer_dimensions.h"
#include "nnet/symbol_generator.h"
#include "plasticity_b200.h"
#include "symbolic/expression.h"
#include "symbolic/symbolic_util.h"

namespace nnet {

typedef double Number;

constexpr const char *kWeightFileFormatVersion = "dev";  // nnet.h:29

class Nnet {
 public:
  struct LearningParameters {
    Number learning_rate;
    bool dynamic_learning_rate = false;
  };

  enum InitStrategy {
    Xavier = 0,
    NoWeightInit,  // testing: weights are set by hand
    InitToOne,
  };

  // nnet.h:52-100.  max_batch (new, defaulted) sizes the device scratch for
  // batched Evaluate / BatchTrain; larger batches run as chunks.
  Nnet(const Architecture &model, InitStrategy weight_initialization = Xavier,
       LossFunction loss_function = MeanSquared, size_t max_batch = 1)
      : model_(model), loss_(loss_function), max_batch_(max_batch) {
    if (!model_.VerifyArchitecture()) {
      std::cerr << "Invalid dimensions passed to Nnet(): " << model.to_string() << std::endl;
      std::exit(1);
    }
    ctx_ = SharedContext();  // SelectDevice(): device 0, nnet.h:188-197
    std::vector<pb_layer_desc> descs;
    for (auto &l : model_.layers) descs.push_back(l.desc());
    PB_CHECK_OK(pb_nnet_create(ctx_, descs.data(), (int32_t)descs.size(),
                               loss_ == CrossEntropy ? PB_CROSS_ENTROPY : PB_MEAN_SQUARED,
                               (int64_t)max_batch_, &net_));
    host_dirty_.assign(model_.layers.size(), true);
    device_newer_.assign(model_.layers.size(), false);
    CalculateInitialWeights(weight_initialization);
  }

  ~Nnet() {
    if (net_) pb_nnet_destroy(net_);
  }

  // One context (CUDA stream) per process: every Nnet and every buffer shares
  // it, so buffers may outlive the network that made them and networks can
  // exchange device buffers without the reference's host bounce.
  static pb_context *SharedContext() {
    static pb_context *ctx = [] {
      pb_context *c = nullptr;
      PB_CHECK_OK(pb_context_create(0, &c));
      return c;
    }();
    return ctx;
  }
  Nnet(const Nnet &) = delete;
  Nnet &operator=(const Nnet &) = delete;

  std::unique_ptr<compute::ClBuffer> MakeBuffer(size_t size) {
    return std::make_unique<compute::ClBuffer>(size, ctx_);
  }
  std::unique_ptr<compute::ClBuffer> MakeBuffer(const std::vector<double> &values) {
    return std::make_unique<compute::ClBuffer>(values, ctx_);
  }
  std::unique_ptr<compute::ClBuffer> MakeBuffer(std::initializer_list<double> values) {
    std::vector<double> vals(values);
    return MakeBuffer(vals);
  }
  std::unique_ptr<compute::ClBuffer> MakeBuffer() { return std::make_unique<compute::ClBuffer>(ctx_); }

  void RegisterBuffer(compute::ClBuffer *buffer) { buffer->RegisterClBackend(ctx_); }

  // Intended mostly for testing or low-level hacks (nnet.h:129-142).
  Architecture &model() {
    for (size_t i = 0; i < model_.layers.size(); ++i) SyncLayerToHost(i);
    return model_;
  }
  Layer &layer(size_t layer) {
    SyncLayerToHost(layer);
    return model_.layers[layer];
  }
  size_t number_of_layers() const { return model_.layers.size(); }

  double &GetWeight(size_t layer, size_t weight_index) {  // nnet.h:199-202
    SyncLayerToHost(layer);
    return model_.layers[layer].weight_buffer()[weight_index];
  }

  std::unique_ptr<compute::ClBuffer> Evaluate(const std::unique_ptr<compute::ClBuffer> &in) {
    std::unique_ptr<std::vector<compute::ClBuffer>> _(nullptr);
    return Evaluate(in, _);
  }

  // (*out_layer_outputs)[i] is a GPU buffer with the outputs of layer i
  // (nnet.h:213-273).  Inputs holding k * input_size values evaluate k samples.
  std::unique_ptr<compute::ClBuffer> Evaluate(
      const std::unique_ptr<compute::ClBuffer> &inputs,
      std::unique_ptr<std::vector<compute::ClBuffer>> &out_layer_outputs) {
    ClaimOwnership(inputs);
    LoadWeightsToGpu();
    const size_t n_in = model_.input_size(), n_out = model_.output_size();
    if (inputs->size() == 0 || inputs->size() % n_in != 0) {
      std::cerr << "Evaluate: input size " << inputs->size() << " is not a multiple of " << n_in
                << std::endl;
      std::exit(1);
    }
    const size_t total = inputs->size() / n_in;
    float *out = nullptr;
    PB_CHECK_OK(pb_buffer_alloc(ctx_, total * n_out, &out));
    for (size_t b0 = 0; b0 < total; b0 += max_batch_) {
      const size_t nb = std::min(max_batch_, total - b0);
      PB_CHECK_OK(pb_nnet_evaluate(net_, inputs->gpu_buffer() + b0 * n_in, (int64_t)nb,
                                   out + b0 * n_out));
    }
    if (out_layer_outputs) {
      out_layer_outputs->clear();
      const size_t nb = std::min(max_batch_, total);
      for (size_t l = 0; l < model_.layers.size(); ++l) {
        const size_t n = model_.layers[l].GetDimensions().num_outputs * nb;
        float *src = nullptr, *dst = nullptr;
        PB_CHECK_OK(pb_nnet_layer_output(net_, (int32_t)l, &src));
        PB_CHECK_OK(pb_buffer_alloc(ctx_, n, &dst));
        PB_CHECK_OK(pb_buffer_copy(ctx_, dst, src, n));
        out_layer_outputs->emplace_back(ctx_, dst, n);
      }
    }
    PB_CHECK_OK(pb_context_finish(ctx_));  // queue->finish(), nnet.h:270
    return std::make_unique<compute::ClBuffer>(ctx_, out, total * n_out);
  }

  double Error(std::unique_ptr<compute::ClBuffer> &actual_output,
               std::unique_ptr<compute::ClBuffer> &expected) {  // nnet.h:280-314
    ClaimOwnership(actual_output);
    ClaimOwnership(expected);
    double e = 0;
    PB_CHECK_OK(pb_nnet_error(net_, actual_output->gpu_buffer(), expected->gpu_buffer(), &e));
    return e;
  }

  void ErrorGradients(const std::unique_ptr<compute::ClBuffer> &actual_output,
                      const std::unique_ptr<compute::ClBuffer> &expected,
                      const std::unique_ptr<compute::ClBuffer> &out_error_gradients) {
    ClaimOwnership(actual_output);  // nnet.h:316-349
    ClaimOwnership(expected);
    ClaimOwnership(out_error_gradients);
    if (out_error_gradients->size() < model_.output_size()) {
      out_error_gradients->MoveToCpu();
      out_error_gradients->resize(model_.output_size());
      out_error_gradients->MoveToGpu();
    }
    PB_CHECK_OK(pb_nnet_error_gradients(net_, actual_output->gpu_buffer(), expected->gpu_buffer(),
                                        out_error_gradients->gpu_buffer(), 1));
  }

  void Train(std::unique_ptr<compute::ClBuffer> &in, std::unique_ptr<compute::ClBuffer> &o) {
    std::unique_ptr<compute::ClBuffer> _(nullptr);
    return Train(in, o, _);
  }

  void SetLearningParameters(const LearningParameters &params) {  // nnet.h:357-361
    PB_CHECK_OK(pb_nnet_set_learning_rate(net_, params.learning_rate));
  }

  // input_gradients (out, may be null): dE/dInput.  initial_backprop_gradients:
  // custom gradient that starts back-prop instead of the loss (nnet.h:363-464).
  void Train(std::unique_ptr<compute::ClBuffer> &in, std::unique_ptr<compute::ClBuffer> &o,
             const std::unique_ptr<compute::ClBuffer> &input_gradients,
             const std::unique_ptr<compute::ClBuffer> &initial_backprop_gradients) {
    ClaimOwnership(in);
    if (o) ClaimOwnership(o);
    ClaimOwnership(initial_backprop_gradients);
    TrainImpl(in, o, input_gradients, initial_backprop_gradients->gpu_buffer());
  }

  void Train(std::unique_ptr<compute::ClBuffer> &in, std::unique_ptr<compute::ClBuffer> &o,
             const std::unique_ptr<compute::ClBuffer> &input_gradients) {  // nnet.h:469-483
    ClaimOwnership(in);
    ClaimOwnership(o);
    TrainImpl(in, o, input_gradients, nullptr);
  }

  // The loop of nnet/cifar_test.cc:423-428 on the device: n per-sample SGD steps
  // over samples stored back to back (CUDA-graph replay per sample).
  void TrainLoop(std::unique_ptr<compute::ClBuffer> &ins, std::unique_ptr<compute::ClBuffer> &outs,
                 size_t n_samples) {
    ClaimOwnership(ins);
    ClaimOwnership(outs);
    LoadWeightsToGpu();
    PB_CHECK_OK(pb_nnet_train_loop(net_, ins->gpu_buffer(), outs->gpu_buffer(), (int64_t)n_samples));
    MarkDeviceNewer();
  }

  // Minibatch step, W <- W - lr * sum_b grad_b(W_old)  (nnet.h:617-669).
  void BatchTrain(std::vector<std::unique_ptr<compute::ClBuffer>> &ins,
                  std::vector<std::unique_ptr<compute::ClBuffer>> &outs,
                  std::set<int> indices_to_train) {
    const size_t n_in = model_.input_size(), n_out = model_.output_size();
    const size_t B = indices_to_train.size();
    if (B == 0) return;
    LoadWeightsToGpu();
    float *x = nullptr, *e = nullptr;
    PB_CHECK_OK(pb_buffer_alloc(ctx_, B * n_in, &x));
    PB_CHECK_OK(pb_buffer_alloc(ctx_, B * n_out, &e));
    size_t k = 0;
    for (int i : indices_to_train) {
      ClaimOwnership(ins[i]);
      ClaimOwnership(outs[i]);
      PB_CHECK_OK(pb_buffer_copy(ctx_, x + k * n_in, ins[i]->gpu_buffer(), n_in));
      PB_CHECK_OK(pb_buffer_copy(ctx_, e + k * n_out, outs[i]->gpu_buffer(), n_out));
      ++k;
    }
    PB_CHECK_OK(pb_nnet_batch_train(net_, x, e, (int64_t)B));
    PB_CHECK_OK(pb_context_finish(ctx_));
    pb_buffer_free(ctx_, x);
    pb_buffer_free(ctx_, e);
    MarkDeviceNewer();
  }

  // Same step over samples already stored back to back on the device.
  void BatchTrain(std::unique_ptr<compute::ClBuffer> &ins, std::unique_ptr<compute::ClBuffer> &outs,
                  size_t batch) {
    ClaimOwnership(ins);
    ClaimOwnership(outs);
    LoadWeightsToGpu();
    PB_CHECK_OK(pb_nnet_batch_train(net_, ins->gpu_buffer(), outs->gpu_buffer(), (int64_t)batch));
    MarkDeviceNewer();
  }

  void LoadWeightsToGpu() {  // nnet.h:490-496: already-resident layers are skipped
    for (size_t l = 0; l < model_.layers.size(); ++l) {
      if (!host_dirty_[l]) continue;
      compute::ClBuffer &w = model_.layers[l].weight_buffer();
      if (w.size()) {
        w.MoveToCpu();
        PB_CHECK_OK(pb_nnet_set_weights_f64(net_, (int32_t)l, &w[0], (int64_t)w.size()));
      }
      host_dirty_[l] = false;
    }
  }

  // JSON weight file, format "dev" (nnet.h:761-789).
  std::string WeightsToString() {
    std::ostringstream out;
    out << "{\"version\":\"" << kWeightFileFormatVersion << "\",\"number_of_layers\":"
        << model_.layers.size() << ",\"layers\":[";
    for (size_t l = 0; l < model_.layers.size(); ++l) {
      SyncLayerToHost(l);
      if (l) out << ",";
      out << "{\"name\":\"" << model_.layers[l].LayerSuffix() << "\",\"index\":" << l
          << ",\"weights\":[";
      compute::ClBuffer &w = model_.layers[l].weight_buffer();
      for (size_t i = 0; i < w.size(); ++i) out << (i ? "," : "") << FormatDouble(w[i]);
      out << "]}";
    }
    out << "]}";
    return out.str();
  }

  // nnet.h:671-759: validates version, layer count, names, indices, sizes.
  bool LoadWeightsFromString(const std::string &weight_string) {
    JsonReader r(weight_string);
    std::string version;
    long n_layers = -1;
    bool have_layers = false;
    std::vector<std::vector<double>> weights(model_.layers.size());
    if (!r.Expect('{')) return Fail("not a JSON object");
    bool first = true;
    while (!r.Peek('}')) {
      if (!first && !r.Expect(',')) return Fail("malformed object");
      first = false;
      std::string key;
      if (!r.String(&key) || !r.Expect(':')) return Fail("malformed key");
      if (key == "version") {
        if (!r.String(&version)) return Fail("Missing version field!");
      } else if (key == "number_of_layers") {
        double v;
        bool is_double;
        if (!r.Number(&v, &is_double) || is_double)
          return Fail("missing number_of_layers (or wrong type)");
        n_layers = (long)v;
      } else if (key == "layers") {
        have_layers = true;
        if (!r.Expect('[')) return Fail("missing layers (or wrong type)");
        size_t index = 0;
        while (!r.Peek(']')) {
          if (index && !r.Expect(',')) return Fail("malformed layers array");
          if (index >= model_.layers.size())
            return Fail("number_of_layers matches, but the provided number of layers was not the same.");
          if (!ReadLayer(&r, index, &weights[index])) return false;
          ++index;
        }
        r.Expect(']');
        if (index != model_.layers.size())
          return Fail("number_of_layers matches, but the provided number of layers was not the same.");
      } else {
        if (!r.SkipValue()) return Fail("malformed value");
      }
    }
    if (version.empty()) return Fail("Missing version field!");
    if (version != kWeightFileFormatVersion)
      return Fail("Weight file format mismatch! " + version + " != (expected) " +
                  kWeightFileFormatVersion);
    if (n_layers < 0) return Fail("missing number_of_layers (or wrong type)");
    if ((size_t)n_layers != model_.layers.size())
      return Fail("Unexpected number of layers, model provided does not match weight file.");
    if (!have_layers) return Fail("missing layers (or wrong type)");
    for (size_t l = 0; l < model_.layers.size(); ++l) {
      for (size_t i = 0; i < weights[l].size(); ++i) model_.layers[l].W(i) = weights[l][i];
      host_dirty_[l] = true;
      device_newer_[l] = false;
    }
    return true;
  }

  pb_context *context() { return ctx_; }
  pb_nnet *handle() { return net_; }

 private:
  void TrainImpl(std::unique_ptr<compute::ClBuffer> &in, std::unique_ptr<compute::ClBuffer> &o,
                 const std::unique_ptr<compute::ClBuffer> &input_gradients,
                 const float *initial_gradients) {
    LoadWeightsToGpu();
    float *ig = nullptr;
    if (input_gradients) {
      if (input_gradients->context() != ctx_) input_gradients->RegisterClBackend(ctx_);
      if (input_gradients->size() != model_.input_size()) {
        input_gradients->MoveToCpu();
        input_gradients->resize(model_.input_size(), 0.0);
      }
      input_gradients->MoveToGpu();
      ig = input_gradients->gpu_buffer();
    }
    PB_CHECK_OK(pb_nnet_train(net_, in->gpu_buffer(), o ? o->gpu_buffer() : nullptr, ig,
                              initial_gradients));
    MarkDeviceNewer();
  }

  // nnet.h:911-920: a buffer made by another Nnet (or by nobody) is bounced
  // through the host into this network's context.
  void ClaimOwnership(const std::unique_ptr<compute::ClBuffer> &buffer) {
    if (buffer->context() != ctx_) buffer->RegisterClBackend(ctx_);
    buffer->MoveToGpu();
  }

  void SyncLayerToHost(size_t l) {
    compute::ClBuffer &w = model_.layers[l].weight_buffer();
    w.MoveToCpu();
    if (device_newer_[l] && w.size())
      PB_CHECK_OK(pb_nnet_get_weights_f64(net_, (int32_t)l, &w[0], (int64_t)w.size()));
    device_newer_[l] = false;
    host_dirty_[l] = true;  // the caller gets a mutable reference
  }

  void MarkDeviceNewer() {
    for (size_t l = 0; l < model_.layers.size(); ++l)
      if (model_.layers[l].weight_buffer().size()) device_newer_[l] = true;
  }

  void CalculateInitialWeights(InitStrategy weight_initialization) {  // nnet.h:792-808
    switch (weight_initialization) {
      case NoWeightInit:
        break;  // host weights of the model (zeros unless set) are uploaded on first use
      case InitToOne:
        PB_CHECK_OK(pb_nnet_init_constant(net_, 1.0));
        host_dirty_.assign(model_.layers.size(), false);
        MarkDeviceNewer();
        break;
      case Xavier:
      default: {
        // Layer::XavierInitializeWeights, layer.cc:86-100: N(0, 1/num_inputs),
        // seeded from random_device like stats/normal.h:11,37.
        // PLASTICITY_SEED=<n> in the environment makes the draw reproducible.
        uint64_t seed;
        if (const char *fixed = std::getenv("PLASTICITY_SEED")) {
          seed = std::strtoull(fixed, nullptr, 10);
        } else {
          std::random_device rd;
          seed = ((uint64_t)rd() << 32) ^ rd();
        }
        PB_CHECK_OK(pb_nnet_init_xavier(net_, seed));
        host_dirty_.assign(model_.layers.size(), false);
        MarkDeviceNewer();
        break;
      }
    }
  }

  static std::string FormatDouble(double v) {
    if (!std::isfinite(v)) {  // not JSON: rapidjson's Writer::Double refuses it as well
      std::cerr << "WeightsToString: non-finite weight (the network diverged); refusing to "
                   "write an unloadable weight file" << std::endl;
      std::exit(1);
    }
    char buf[40];
    std::snprintf(buf, sizeof(buf), "%.17g", v);
    std::string s(buf);
    // rapidjson's Writer::Double always writes a double-typed token.
    if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
    return s;
  }

  static bool Fail(const std::string &msg) {
    std::cerr << msg << std::endl;
    return false;
  }

  // Minimal reader for the weight-file grammar (objects, arrays, strings, numbers).
  class JsonReader {
   public:
    explicit JsonReader(const std::string &s) : s_(s) {}
    void Ws() { while (p_ < s_.size() && std::isspace((unsigned char)s_[p_])) ++p_; }
    bool Peek(char c) { Ws(); return p_ < s_.size() ? s_[p_] == c : true; }
    bool Expect(char c) { Ws(); if (p_ < s_.size() && s_[p_] == c) { ++p_; return true; } return false; }
    bool String(std::string *out) {
      Ws();
      if (p_ >= s_.size() || s_[p_] != '"') return false;
      out->clear();
      for (++p_; p_ < s_.size() && s_[p_] != '"'; ++p_) {
        if (s_[p_] == '\\' && p_ + 1 < s_.size()) ++p_;
        out->push_back(s_[p_]);
      }
      return Expect('"');
    }
    bool Number(double *v, bool *is_double) {
      Ws();
      const char *b = s_.c_str() + p_;
      char *e = nullptr;
      *v = std::strtod(b, &e);
      if (e == b) return false;
      *is_double = std::string(b, (size_t)(e - b)).find_first_of(".eE") != std::string::npos;
      p_ += (size_t)(e - b);
      return true;
    }
    bool SkipValue() {
      Ws();
      if (p_ >= s_.size()) return false;
      const char c = s_[p_];
      std::string tmp;
      double v; bool d;
      if (c == '"') return String(&tmp);
      if (c == '{' || c == '[') {
        const char close = c == '{' ? '}' : ']';
        ++p_;
        bool first = true;
        while (!Peek(close)) {
          if (!first && !Expect(',')) return false;
          first = false;
          if (c == '{' && (!String(&tmp) || !Expect(':'))) return false;
          if (!SkipValue()) return false;
        }
        return Expect(close);
      }
      if (s_.compare(p_, 4, "true") == 0 || s_.compare(p_, 4, "null") == 0) { p_ += 4; return true; }
      if (s_.compare(p_, 5, "false") == 0) { p_ += 5; return true; }
      return Number(&v, &d);
    }
   private:
    const std::string &s_;
    size_t p_ = 0;
  };

  bool ReadLayer(JsonReader *r, size_t index, std::vector<double> *weights) {
    bool have_name = false, have_index = false, have_weights = false;
    if (!r->Expect('{')) return Fail("malformed layer");
    bool first = true;
    while (!r->Peek('}')) {
      if (!first && !r->Expect(',')) return Fail("malformed layer");
      first = false;
      std::string key;
      if (!r->String(&key) || !r->Expect(':')) return Fail("malformed layer key");
      if (key == "name") {
        std::string name;
        if (!r->String(&name)) return Fail("Missing name");
        const std::string expected_name = model_.layers[index].LayerSuffix();
        if (name != expected_name)
          return Fail("Model name mismatch. " + name + " != (expected) " + expected_name);
        have_name = true;
      } else if (key == "index") {
        double v; bool d;
        if (!r->Number(&v, &d)) return Fail("Missing layer index");
        if ((size_t)v != index)
          return Fail("Layer index mismatch: " + std::to_string((long)v) + " != (expected) " +
                      std::to_string(index));
        have_index = true;
      } else if (key == "weights") {
        if (!r->Expect('[')) return Fail("Layer missing weights (or wrong type).");
        while (!r->Peek(']')) {
          if (!weights->empty() && !r->Expect(',')) return Fail("malformed weights");
          double v; bool d;
          if (!r->Number(&v, &d)) return Fail("malformed weight");
          if (!d) return Fail("Weight is not a double!");
          weights->push_back(v);
        }
        r->Expect(']');
        if (weights->size() != model_.layers[index].weight_buffer().size())
          return Fail("layer size mismatch!");
        have_weights = true;
      } else if (!r->SkipValue()) {
        return Fail("malformed layer value");
      }
    }
    r->Expect('}');
    if (!have_name) return Fail("Missing name");
    if (!have_index) return Fail("Missing layer index");
    if (!have_weights) return Fail("Layer missing weights (or wrong type).");
    return true;
  }

  Architecture model_;
  LossFunction loss_;
  size_t max_batch_;
  pb_context *ctx_ = nullptr;
  pb_nnet *net_ = nullptr;
  std::vector<bool> host_dirty_, device_newer_;
};

}  // namespace nnet
#endif
