// compute::ClBuffer over the CUDA C ABI — same interface and CPU<->GPU state
// machine as the reference's compute/cl_buffer.h:42-187 / cl_buffer.cc:14-131:
// a buffer is EITHER a host std::vector<double> OR a device allocation (fp32
// here); MoveToGpu / MoveToCpu are blocking; the copy constructor and
// DeepClone copy on the device; operator= aliases the device handle (the
// reference copies the cl::Buffer handle, cl_buffer.h:166-179).  The
// (cl::CommandQueue*, cl::Context*) pair the reference threads through every
// constructor becomes one pb_context*.
#ifndef PLASTICITY_B200_COMPUTE_CL_BUFFER_H_
#define PLASTICITY_B200_COMPUTE_CL_BUFFER_H_

#include <cstdlib>
#include <iostream>
#include <limits>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "plasticity_b200.h"

#if __has_include("geometry/dynamic_matrix.h")
#include "geometry/dynamic_matrix.h"
#define PLASTICITY_B200_HAVE_MATRIX 1
#endif

// The reference's CL_CHECK (compute/cl_buffer.h:19-27): print and exit(1).
#define PB_CHECK_OK(expr)                                                              \
  do {                                                                                 \
    if ((expr) != 0) {                                                                 \
      std::cerr << "plasticity_b200 error: " << pb_last_error() << " (" << #expr << " at " \
                << __FILE__ << ":" << __LINE__ << ")" << std::endl;                    \
      std::exit(1);                                                                    \
    }                                                                                  \
  } while (0)

#ifndef CHECK_NOTNULL
#define CHECK_NOTNULL(x)                                                     \
  if ((x) == nullptr) {                                                      \
    std::cerr << "Error, unexpected nullptr: " << #x << std::endl;           \
    std::exit(1);                                                            \
  }
#endif

namespace compute {

class ClBuffer {
 public:
  enum Location { CPU = 0, GPU };

  ClBuffer() : state_(CPU), cpu_buffer_(0) {}
  explicit ClBuffer(size_t size) : state_(CPU), cpu_buffer_(size) {}
  ClBuffer(const std::vector<double> &values) : state_(CPU), cpu_buffer_(values) {}
  ClBuffer(const std::vector<double> &values, pb_context *ctx)
      : state_(CPU), cpu_buffer_(values), ctx_(ctx) {}
  ClBuffer(size_t size, pb_context *ctx) : state_(CPU), cpu_buffer_(size), ctx_(ctx) {
    CHECK_NOTNULL(ctx_);
  }
  explicit ClBuffer(pb_context *ctx) : state_(CPU), cpu_buffer_(0), ctx_(ctx) { CHECK_NOTNULL(ctx_); }
  // Adopt a device allocation of n floats (used by Nnet for results).
  ClBuffer(pb_context *ctx, float *dptr, size_t n) : state_(GPU), ctx_(ctx), gpu_size_(n) {
    CHECK_NOTNULL(ctx_);
    Adopt(dptr);
  }
  // Non-owning window onto n floats of another device allocation (samples stored back
  // to back): same state machine, nothing is freed when the view dies.
  static std::unique_ptr<ClBuffer> View(pb_context *ctx, float *dptr, size_t n) {
    auto v = std::make_unique<ClBuffer>(ctx);
    v->state_ = GPU;
    v->gpu_size_ = n;
    v->gpu_buffer_ = std::shared_ptr<float>(dptr, [](float *) {});
    return v;
  }
  ClBuffer(const ClBuffer &other)
      : state_(other.state_), cpu_buffer_(other.cpu_buffer_), ctx_(other.ctx_),
        gpu_size_(other.gpu_size_) {
    if (other.state_ == GPU) {  // device-side deep copy, cl_buffer.h:94-111
      CHECK_NOTNULL(ctx_);
      float *p = nullptr;
      PB_CHECK_OK(pb_buffer_alloc(ctx_, gpu_size_, &p));
      PB_CHECK_OK(pb_buffer_copy(ctx_, p, other.gpu_buffer_.get(), gpu_size_));
      Adopt(p);
    }
  }
  ClBuffer(ClBuffer &&other)
      : state_(other.state_), cpu_buffer_(std::move(other.cpu_buffer_)), ctx_(other.ctx_),
        gpu_size_(other.gpu_size_) {
    if (other.state_ == GPU) {
      gpu_buffer_ = std::move(other.gpu_buffer_);
      other.state_ = CPU;
    }
  }
  virtual ~ClBuffer() {}

  ClBuffer DeepClone() {  // cl_buffer.h:52-65
    return ClBuffer(*this);
  }

  static std::unique_ptr<ClBuffer> MakeBufferFromValues(const std::vector<double> &values) {
    return std::make_unique<ClBuffer>(values);
  }
#ifdef PLASTICITY_B200_HAVE_MATRIX
  static std::unique_ptr<ClBuffer> MakeBufferFromColumnVector(Matrix<double> column_vector) {
    std::vector<double> values;  // cl_buffer.cc:5-12
    for (size_t i = 0; i < column_vector.dimensions().rows; ++i) values.push_back(column_vector.at(i, 0));
    return std::make_unique<ClBuffer>(values);
  }
#endif

  void RegisterClBackend(pb_context *ctx) {  // cl_buffer.h:139-144
    MoveToCpu();
    CHECK_NOTNULL(ctx_ = ctx);
    MoveToGpu();
  }

  void MoveToCpu() {  // cl_buffer.cc:75-99
    if (state_ == CPU) return;
    if (!gpu_buffer_) {
      std::cerr << "Error, unexpected nullptr gpu_buffer_" << std::endl;
      std::exit(1);
    }
    cpu_buffer_.resize(gpu_size_);
    if (gpu_size_)
      PB_CHECK_OK(pb_buffer_read_f64(ctx_, gpu_buffer_.get(), cpu_buffer_.data(), gpu_size_));
    gpu_buffer_.reset();
    state_ = CPU;
  }

  void MoveToGpu() {  // cl_buffer.cc:101-127
    if (state_ == GPU) return;
    CHECK_NOTNULL(ctx_);
    float *p = nullptr;
    gpu_size_ = cpu_buffer_.size();
    PB_CHECK_OK(pb_buffer_alloc(ctx_, gpu_size_, &p));
    if (gpu_size_) PB_CHECK_OK(pb_buffer_write_f64(ctx_, p, cpu_buffer_.data(), gpu_size_));
    Adopt(p);
    state_ = GPU;
  }

  Location GetBufferLocation() { return state_; }
  size_t size() const { return state_ == CPU ? cpu_buffer_.size() : gpu_size_; }

  void resize(size_t new_size,
              double default_value = std::numeric_limits<double>::quiet_NaN()) {
    if (new_size == size()) return;  // cl_buffer.cc:14-31
    const Location old_state = state_;
    if (state_ == GPU) MoveToCpu();
    cpu_buffer_.resize(new_size, default_value);
    if (old_state == GPU) MoveToGpu();
  }

  // Only after MoveToCpu (cl_buffer.cc:53-73).
  double &operator[](size_t index) {
    RequireCpu();
    return cpu_buffer_[index];
  }
  const double &operator[](size_t index) const {
    RequireCpu();
    return cpu_buffer_[index];
  }
  double &at(size_t index) { return (*this)[index]; }
  const double &at(size_t index) const { return (*this)[index]; }

  std::string to_string() const {
    RequireCpu();
    std::stringstream out;
    out << "{";
    for (size_t i = 0; i < cpu_buffer_.size(); ++i) out << (i ? ", " : "") << cpu_buffer_[i];
    out << "}";
    return out.str();
  }

  // Only after MoveToGpu.  The reference returns the cl::Buffer; here the fp32
  // device pointer.
  float *gpu_buffer() const {
    if (state_ == CPU) {
      std::cerr << "Requested GPU buffer when in CPU state!" << std::endl;
      std::exit(1);
    }
    return gpu_buffer_.get();
  }

  const ClBuffer &operator=(const ClBuffer &rhs) {  // handle alias, cl_buffer.h:166-179
    state_ = rhs.state_;
    cpu_buffer_ = rhs.cpu_buffer_;
    ctx_ = rhs.ctx_;
    gpu_size_ = rhs.gpu_size_;
    gpu_buffer_ = rhs.state_ == GPU ? rhs.gpu_buffer_ : nullptr;
    return *this;
  }

  pb_context *context() const { return ctx_; }

 private:
  void Adopt(float *p) {
    pb_context *c = ctx_;
    gpu_buffer_ = std::shared_ptr<float>(p, [c](float *q) { pb_buffer_free(c, q); });
  }
  void RequireCpu() const {
    if (state_ == GPU) {
      std::cerr << "Error: [] used while buffer is in GPU." << std::endl;
      std::exit(1);
    }
  }

  Location state_;
  std::vector<double> cpu_buffer_;
  std::shared_ptr<float> gpu_buffer_;
  pb_context *ctx_ = nullptr;
  size_t gpu_size_ = 0;
};

}  // namespace compute
#endif
