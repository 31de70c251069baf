// compute::CommandQueue of the B200 host mirror — the reference's abstract launch interface,
// compute/command_queue.h:7-11: a kernel NAME, (global size, work-group size) and positional
// buffer arguments, plus finish().  It is the abstract form of what nnet::Nnet does per launch
// (nnet/nnet.h:243-259, 406-452, 585-607, 816-827).  The reference's own OpenCL implementation
// (compute/cl_command_queue.{h,cc}) is dead code that does not compile (SURVEY.md section 2,
// row 8); compute/cuda_command_queue.h is a working one over a CUDA stream.
#ifndef PLASTICITY_B200_COMPUTE_COMMAND_QUEUE_H_
#define PLASTICITY_B200_COMPUTE_COMMAND_QUEUE_H_

#include <memory>
#include <string>
#include <vector>

#include "compute/cl_buffer.h"

namespace compute {

class CommandQueue {
 public:
  virtual ~CommandQueue() {}
  // Arguments are ClBuffers in the GPU state.  ClBuffer::operator= aliases the device
  // handle (cl_buffer.h:166-179), so a vector filled by assignment refers to the caller's
  // buffers and a kernel's outputs land in them.
  virtual void enqueueKernel(const std::string &kernel_name, int global_size, int workgroup_size,
                             std::unique_ptr<std::vector<compute::ClBuffer>> arguments) = 0;
  virtual void finish() = 0;
};

}  // namespace compute
#endif
