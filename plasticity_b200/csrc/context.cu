// Context + buffers: the CUDA implementation behind the reference's
// compute::ClBuffer / command-queue interface (compute/cl_buffer.{h,cc},
// compute/command_queue.h).  See include/plasticity_b200.h for the mapping.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace pb {

static thread_local std::string g_error;

void set_error(const std::string &msg) { g_error = msg; }
int fail(const std::string &msg) {
  g_error = msg;
  return 1;
}

int ensure_stage(pb_context *ctx, size_t elems) {
  if (ctx->stage_elems >= elems) return 0;
  if (ctx->stage) PB_CUDA(cudaFree(ctx->stage));
  ctx->stage = nullptr;
  ctx->stage_elems = 0;
  size_t want = std::max<size_t>(elems, 1 << 16);
  PB_CUDA(cudaMalloc(&ctx->stage, want * sizeof(double)));
  ctx->stage_elems = want;
  return 0;
}

int ensure_workspace(pb_context *ctx, size_t elems) {
  const size_t off = ctx->fork_mode ? ctx->ws_off : 0;
  const size_t need = off + elems;
  if (ctx->ws_cap < need) {
    // The old allocation may still be in use by queued kernels (on either stream of a
    // forked step: only its eager runs grow the workspace, a captured one finds it sized).
    // (this context's streams only: another in-process rank's exchange kernel may be waiting
    // on the device for the very step that is being queued here)
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->fork_stream) PB_CUDA(cudaStreamSynchronize(ctx->fork_stream));
    if (ctx->ws_base) PB_CUDA(cudaFree(ctx->ws_base));
    ctx->ws_base = nullptr;
    ctx->ws_cap = 0;
    size_t want = std::max<size_t>(need, 1 << 20);
    PB_CUDA(cudaMalloc(&ctx->ws_base, want * sizeof(float)));
    ctx->ws_cap = want;
    ctx->alloc_generation++;
  }
  ctx->workspace = ctx->ws_base + off;
  ctx->workspace_elems = ctx->ws_cap - off;
  if (ctx->fork_mode) ctx->ws_off = (need + 63) & ~(size_t)63;   // 256-byte aligned regions
  return 0;
}

int join_forked(pb_context *ctx) {
  for (int i = 0; i < ctx->forks; ++i) PB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->join_ev[i], 0));
  ctx->forks = 0;
  return 0;
}

int ensure_aux(pb_context *ctx, size_t elems) {
  if (ctx->aux_elems >= elems) return 0;
  PB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->aux) PB_CUDA(cudaFree(ctx->aux));
  ctx->aux = nullptr;
  ctx->aux_elems = 0;
  size_t want = std::max<size_t>(elems, 1 << 20);
  PB_CUDA(cudaMalloc(&ctx->aux, want * sizeof(float)));
  ctx->aux_elems = want;
  ctx->alloc_generation++;
  return 0;
}

int check_fault(pb_context *ctx) {
  if (ctx->fault && *(volatile int *)ctx->fault)
    return fail("data-parallel exchange: a peer rank did not arrive within 5 s; the weight "
                "update of that step was skipped on this rank and training must stop");
  return 0;
}

// Mapped pinned host word a kernel can raise without the host having to poll the device.
int ensure_fault_word(pb_context *ctx) {
  if (ctx->fault) return 0;
  PB_CUDA(cudaHostAlloc((void **)&ctx->fault, sizeof(int), cudaHostAllocMapped));
  *ctx->fault = 0;
  return 0;
}

__global__ void f64_to_f32_kernel(const double *__restrict__ src,
                                  float *__restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    dst[i] = (float)src[i];
}

__global__ void f32_to_f64_kernel(const float *__restrict__ src,
                                  double *__restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    dst[i] = (double)src[i];
}

__global__ void fill_kernel(float *__restrict__ dst, float v, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    dst[i] = v;
}

}  // namespace pb

extern "C" {

int pb_abi_version(void) { return PB_ABI_VERSION; }

const char *pb_last_error(void) { return pb::g_error.c_str(); }

int pb_device_count(int *count) {
  PB_CHECK(count, "pb_device_count: null");
  PB_CUDA(cudaGetDeviceCount(count));
  return 0;
}

int pb_context_create(int device, pb_context **out) {
  PB_CHECK(out, "pb_context_create: null out");
  int n = 0;
  PB_CUDA(cudaGetDeviceCount(&n));
  PB_CHECK(n > 0, "pb_context_create: no CUDA device (this library has no CPU path)");
  PB_CHECK(device >= 0 && device < n, "pb_context_create: bad device index");
  PB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PB_CUDA(cudaGetDeviceProperties(&prop, device));
  PB_CHECK(prop.major >= 10,
           "pb_context_create: kernels are built for sm_100a only; device is sm_" +
               std::to_string(prop.major) + std::to_string(prop.minor));
  pb_context *ctx = new pb_context();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete ctx;
    return pb::fail(std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
  }
  *out = ctx;
  return 0;
}

int pb_context_destroy(pb_context *ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->stage) cudaFree(ctx->stage);
  if (ctx->fork_stream) {
    cudaStreamSynchronize(ctx->fork_stream);
    cudaStreamDestroy(ctx->fork_stream);
  }
  for (int i = 0; i < pb_context::kMaxForks; ++i) {
    if (ctx->fork_ev[i]) cudaEventDestroy(ctx->fork_ev[i]);
    if (ctx->join_ev[i]) cudaEventDestroy(ctx->join_ev[i]);
  }
  if (ctx->ws_base) cudaFree(ctx->ws_base);
  if (ctx->aux) cudaFree(ctx->aux);
  if (ctx->fault) cudaFreeHost(ctx->fault);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return 0;
}

int pb_context_finish(pb_context *ctx) {
  PB_CHECK(ctx, "pb_context_finish: null ctx");
  PB_CUDA(cudaStreamSynchronize(ctx->stream));
  for (cudaStream_t st : ctx->side_streams) PB_CUDA(cudaStreamSynchronize(st));
  return pb::check_fault(ctx);
}

int pb_context_device(const pb_context *ctx, int *device) {
  PB_CHECK(ctx && device, "pb_context_device: null");
  *device = ctx->device;
  return 0;
}

int pb_context_stream(const pb_context *ctx, void **stream) {
  PB_CHECK(ctx && stream, "pb_context_stream: null");
  *stream = (void *)ctx->stream;
  return 0;
}

int pb_context_launch_count(const pb_context *ctx, uint64_t *count) {
  PB_CHECK(ctx && count, "pb_context_launch_count: null");
  *count = ctx->launches;
  return 0;
}

int pb_buffer_alloc(pb_context *ctx, size_t n, float **dptr) {
  PB_CHECK(ctx && dptr, "pb_buffer_alloc: null");
  PB_CUDA(cudaSetDevice(ctx->device));
  // Zero-length buffers get a 1-element allocation, like the reference's
  // 1-byte dummy cl::Buffer (compute/cl_buffer.cc:112-120).
  void *p = nullptr;
  PB_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(float)));
  *dptr = (float *)p;
  return 0;
}

int pb_buffer_free(pb_context *ctx, float *dptr) {
  PB_CHECK(ctx, "pb_buffer_free: null ctx");
  if (!dptr) return 0;
  // Kernels queued on the stream may still use the buffer.
  PB_CUDA(cudaStreamSynchronize(ctx->stream));
  PB_CUDA(cudaFree(dptr));
  return 0;
}

int pb_buffer_write_f64(pb_context *ctx, float *dst, const double *host, size_t n) {
  PB_CHECK(ctx && (n == 0 || (dst && host)), "pb_buffer_write_f64: null");
  if (n == 0) return 0;
  if (pb::ensure_stage(ctx, n)) return 1;
  PB_CUDA(cudaMemcpyAsync(ctx->stage, host, n * sizeof(double),
                          cudaMemcpyHostToDevice, ctx->stream));
  pb::f64_to_f32_kernel<<<pb_ew_grid(ctx, n, 256), 256, 0, ctx->stream>>>(
      ctx->stage, dst, n);
  PB_LAUNCHED(ctx);
  PB_CUDA(cudaStreamSynchronize(ctx->stream));  // blocking, like CL_TRUE
  return 0;
}

int pb_buffer_read_f64(pb_context *ctx, const float *src, double *host, size_t n) {
  PB_CHECK(ctx && (n == 0 || (src && host)), "pb_buffer_read_f64: null");
  if (n == 0) return 0;
  if (pb::ensure_stage(ctx, n)) return 1;
  pb::f32_to_f64_kernel<<<pb_ew_grid(ctx, n, 256), 256, 0, ctx->stream>>>(
      src, ctx->stage, n);
  PB_LAUNCHED(ctx);
  PB_CUDA(cudaMemcpyAsync(host, ctx->stage, n * sizeof(double),
                          cudaMemcpyDeviceToHost, ctx->stream));
  PB_CUDA(cudaStreamSynchronize(ctx->stream));
  return pb::check_fault(ctx);
}

int pb_buffer_write_f32(pb_context *ctx, float *dst, const float *host, size_t n) {
  PB_CHECK(ctx && (n == 0 || (dst && host)), "pb_buffer_write_f32: null");
  if (n == 0) return 0;
  PB_CUDA(cudaMemcpyAsync(dst, host, n * sizeof(float), cudaMemcpyHostToDevice,
                          ctx->stream));
  PB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int pb_buffer_read_f32(pb_context *ctx, const float *src, float *host, size_t n) {
  PB_CHECK(ctx && (n == 0 || (src && host)), "pb_buffer_read_f32: null");
  if (n == 0) return 0;
  PB_CUDA(cudaMemcpyAsync(host, src, n * sizeof(float), cudaMemcpyDeviceToHost,
                          ctx->stream));
  PB_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int pb_buffer_write_f32_async(pb_context *ctx, float *dst, const float *host, size_t n) {
  PB_CHECK(ctx && (n == 0 || (dst && host)), "pb_buffer_write_f32_async: null");
  if (n == 0) return 0;
  PB_CUDA(cudaMemcpyAsync(dst, host, n * sizeof(float), cudaMemcpyHostToDevice,
                          ctx->stream));
  return 0;
}

int pb_buffer_read_f32_async(pb_context *ctx, const float *src, float *host, size_t n) {
  PB_CHECK(ctx && (n == 0 || (src && host)), "pb_buffer_read_f32_async: null");
  if (n == 0) return 0;
  PB_CUDA(cudaMemcpyAsync(host, src, n * sizeof(float), cudaMemcpyDeviceToHost,
                          ctx->stream));
  return 0;
}

int pb_buffer_copy(pb_context *ctx, float *dst, const float *src, size_t n) {
  PB_CHECK(ctx && (n == 0 || (dst && src)), "pb_buffer_copy: null");
  if (n == 0) return 0;
  PB_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice,
                          ctx->stream));
  return 0;
}

int pb_buffer_fill(pb_context *ctx, float *dst, float value, size_t n) {
  PB_CHECK(ctx && (n == 0 || dst), "pb_buffer_fill: null");
  if (n == 0) return 0;
  pb::fill_kernel<<<pb_ew_grid(ctx, n, 256), 256, 0, ctx->stream>>>(dst, value, n);
  PB_LAUNCHED(ctx);
  return 0;
}

int pb_host_alloc(size_t bytes, void **ptr) {
  PB_CHECK(ptr, "pb_host_alloc: null");
  PB_CUDA(cudaHostAlloc(ptr, std::max<size_t>(bytes, 1), cudaHostAllocDefault));
  return 0;
}

int pb_host_free(void *ptr) {
  if (ptr) PB_CUDA(cudaFreeHost(ptr));
  return 0;
}

}  // extern "C"
