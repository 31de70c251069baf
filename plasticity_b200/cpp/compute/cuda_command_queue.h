// compute::CudaCommandQueue: compute::CommandQueue (compute/command_queue.h:7-11) over the
// CUDA C ABI.  The reference names its generated kernels per layer (nnet/layer.h:91-114,
// nnet/error_layer.h:31-33, nnet/nnet.h:816):
//   evaluate_<i>(I, W, O)                         global = outputs of layer i
//   input_delta_<type>_<i>(I, W, G, dI)           global = inputs  of layer i
//   new_weights_<type>_<i>(I, W, G, Wnew, lr)     global = weights of layer i
//   weight_delta[s]_<type>_<i>(I, W, G, dW, lr)   global = weights of layer i
//   error_value(O, E, e)   error_gradients(O, E, G)   vector_accumulate(a, b)
// There is no program to look the name up in: the name is parsed and the launch goes to the
// layer-type launcher of include/plasticity_b200.h with layer i's dimensions.  The
// work-group size is accepted and ignored (the CUDA kernels choose their own launch shapes);
// a global size that is a multiple of the layer's count runs that many samples in one launch
// (sample-major buffers), which the OpenCL kernels cannot do.
// Errors follow the reference's convention (CL_CHECK): print to std::cerr and exit(1).
#ifndef PLASTICITY_B200_COMPUTE_CUDA_COMMAND_QUEUE_H_
#define PLASTICITY_B200_COMPUTE_CUDA_COMMAND_QUEUE_H_

#include <cstdlib>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "compute/command_queue.h"
#include "plasticity_b200.h"

namespace compute {

class CudaCommandQueue : public CommandQueue {
 public:
  // layers[i] = descriptor of layer i (nnet::Layer::desc()), type_names[i] = its type string
  // ("dense_layer", "convolution_layer", ... nnet/layer_impl.h layer_type()); loss =
  // PB_MEAN_SQUARED / PB_CROSS_ENTROPY for the two error kernels.
  CudaCommandQueue(pb_context *ctx, std::vector<pb_layer_desc> layers,
                   std::vector<std::string> type_names, int loss)
      : ctx_(ctx), layers_(std::move(layers)), types_(std::move(type_names)), loss_(loss) {
    CHECK_NOTNULL(ctx_);
    if (layers_.size() != types_.size()) Die("CudaCommandQueue: one type name per layer");
  }

  void enqueueKernel(const std::string &kernel_name, int global_size, int /*workgroup_size*/,
                     std::unique_ptr<std::vector<compute::ClBuffer>> arguments) override {
    if (!arguments) Die("enqueueKernel: null argument list");
    std::vector<compute::ClBuffer> &a = *arguments;
    if (global_size < 0) Die("enqueueKernel: negative global size");
    if (kernel_name == "vector_accumulate") {
      Need(kernel_name, a, 2);
      PB_CHECK_OK(pb_vector_accumulate(ctx_, Dev(a[0]), Dev(a[1]), (int64_t)global_size));
      return;
    }
    if (kernel_name == "error_value" || kernel_name == "error_gradients") {
      Need(kernel_name, a, 3);
      int64_t n = 0;
      if (layers_.empty() || pb_layer_num_outputs(&layers_.back(), &n) != 0 || n <= 0)
        Die("enqueueKernel: no output layer for " + kernel_name);
      const int64_t batch = Batch(kernel_name, global_size, n);
      if (kernel_name == "error_value")
        PB_CHECK_OK(pb_error_value(ctx_, loss_, n, Dev(a[0]), Dev(a[1]), Dev(a[2]), batch));
      else
        PB_CHECK_OK(pb_error_gradients(ctx_, loss_, n, Dev(a[0]), Dev(a[1]), Dev(a[2]), batch));
      return;
    }
    struct Form { const char *prefix; int kind; };
    // the kernel template is called weight_deltas_* (back_prop.kernel.cl:37-44), the host
    // asks for weight_delta_* (nnet/layer.h:112-114): both spellings are accepted
    static const Form forms[] = {{"evaluate_", 0}, {"input_delta_", 1}, {"new_weights_", 2},
                                 {"weight_delta_", 3}, {"weight_deltas_", 3}};
    for (const Form &f : forms) {
      const std::string prefix(f.prefix);
      if (kernel_name.compare(0, prefix.size(), prefix) != 0) continue;
      const size_t us = kernel_name.rfind('_');
      const std::string digits = kernel_name.substr(us + 1);
      if (digits.empty() || digits.find_first_not_of("0123456789") != std::string::npos) break;
      const size_t i = (size_t)std::strtoull(digits.c_str(), nullptr, 10);
      if (i >= layers_.size()) break;
      if (f.kind != 0) {  // <prefix><type>_<i>: the type must be layer i's
        if (us < prefix.size() || kernel_name.substr(prefix.size(), us - prefix.size()) != types_[i])
          break;
      } else if (us + 1 != prefix.size()) {
        break;
      }
      const pb_layer_desc *l = &layers_[i];
      int64_t nin = 0, nout = 0, nw = 0;
      if (pb_layer_num_inputs(l, &nin) || pb_layer_num_outputs(l, &nout) ||
          pb_layer_num_weights(l, &nw))
        Die(std::string("enqueueKernel: ") + pb_last_error());
      if (f.kind == 0) {
        Need(kernel_name, a, 3);
        const int64_t batch = Batch(kernel_name, global_size, nout);
        PB_CHECK_OK(pb_layer_evaluate(ctx_, l, Dev(a[0]), nw ? Dev(a[1]) : nullptr, Dev(a[2]), batch));
      } else if (f.kind == 1) {
        Need(kernel_name, a, 4);
        const int64_t batch = Batch(kernel_name, global_size, nin);
        PB_CHECK_OK(pb_layer_input_delta(ctx_, l, Dev(a[0]), nw ? Dev(a[1]) : nullptr, Dev(a[2]),
                                         Dev(a[3]), batch));
      } else {
        Need(kernel_name, a, 5);
        if (nw == 0 || (int64_t)global_size != nw)
          Die("enqueueKernel: " + kernel_name + " takes global size = number of weights");
        // One work-item per weight: the global size says nothing about samples.  The
        // reference's callers pass scratch buffers sized max_layer_output_size (nnet.h:86-97),
        // i.e. LARGER than the layer; such a pair is one sample, exactly as the reference's
        // kernel reads it.  A batch is recognised only when BOTH the input and the gradient
        // buffer hold the same whole number of samples.
        const int64_t by_in = nin > 0 ? (int64_t)a[0].size() / nin : 0;
        const int64_t by_out = nout > 0 ? (int64_t)a[2].size() / nout : 0;
        if (by_in < 1 || by_out < 1)
          Die("enqueueKernel: " + kernel_name + ": input / gradient buffer smaller than the layer");
        const bool whole = (int64_t)a[0].size() == by_in * nin && (int64_t)a[2].size() == by_out * nout;
        const int64_t batch = (whole && by_in == by_out) ? by_in : 1;
        PB_CHECK_OK(pb_layer_weight_update(ctx_, l, Dev(a[0]), Dev(a[1]), Dev(a[2]), Dev(a[3]),
                                           Dev(a[4]), f.kind == 2 ? PB_NEW_WEIGHTS : PB_WEIGHT_DELTAS,
                                           batch));
      }
      return;
    }
    Die("enqueueKernel: no kernel named " + kernel_name);  // CL_INVALID_KERNEL_NAME
  }

  void finish() override { PB_CHECK_OK(pb_context_finish(ctx_)); }

 private:
  [[noreturn]] static void Die(const std::string &msg) {
    std::cerr << msg << std::endl;
    std::exit(1);
  }
  static void Need(const std::string &name, const std::vector<compute::ClBuffer> &a, size_t n) {
    if (a.size() != n)
      Die("enqueueKernel: " + name + " takes " + std::to_string(n) + " buffer arguments");
  }
  static int64_t Batch(const std::string &name, int global_size, int64_t per_sample) {
    if (per_sample <= 0 || global_size <= 0 || global_size % per_sample != 0)
      Die("enqueueKernel: global size of " + name + " must be a multiple of " +
          std::to_string(per_sample));
    return (int64_t)global_size / per_sample;
  }
  static float *Dev(compute::ClBuffer &b) { return b.gpu_buffer(); }  // exits unless on the GPU

  pb_context *ctx_;
  std::vector<pb_layer_desc> layers_;
  std::vector<std::string> types_;
  int loss_;
};

}  // namespace compute
#endif
