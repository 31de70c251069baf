// Roofline denominators measured on the device the context owns: FP32 FFMA
// throughput (not in MEASURED_PEAKS.json) and device-to-device copy bandwidth.
#include "common.cuh"
#include "kernels.cuh"

namespace pb {

// 8 independent FFMA chains per thread, 8 warps per CTA, many CTAs per SM:
// issue-bound on the FMA pipes.
__global__ void __launch_bounds__(256) ffma_peak_kernel(float *out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  float x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  const float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456f) out[0] = s;  // never true; keeps the chain alive
}

__global__ void copy_kernel(const float4 *__restrict__ src, float4 *__restrict__ dst, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4;
       i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

}  // namespace pb

extern "C" {

int pb_measure_fp32_tflops(pb_context *ctx, double *tflops) {
  PB_CHECK(ctx && tflops, "pb_measure_fp32_tflops: null");
  float *out = nullptr;
  PB_CUDA(cudaMalloc(&out, sizeof(float)));
  cudaEvent_t e0, e1;
  PB_CUDA(cudaEventCreate(&e0));
  PB_CUDA(cudaEventCreate(&e1));
  const int iters = 2048, grid = ctx->num_sms * 8, T = 256;
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    PB_CUDA(cudaEventRecord(e0, ctx->stream));
    pb::ffma_peak_kernel<<<grid, T, 0, ctx->stream>>>(out, iters, 0.999f, 0.001f);
    PB_LAUNCHED(ctx);
    PB_CUDA(cudaEventRecord(e1, ctx->stream));
    PB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    PB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 8 * 16 * (double)iters * T * grid;
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return 0;
}

int pb_measure_copy_gbs(pb_context *ctx, size_t bytes, double *gbs) {
  PB_CHECK(ctx && gbs, "pb_measure_copy_gbs: null");
  const size_t n4 = bytes / 16;
  PB_CHECK(n4 > 0, "pb_measure_copy_gbs: too small");
  float4 *a = nullptr, *b = nullptr;
  PB_CUDA(cudaMalloc(&a, n4 * 16));
  PB_CUDA(cudaMalloc(&b, n4 * 16));
  PB_CUDA(cudaMemsetAsync(a, 0, n4 * 16, ctx->stream));
  cudaEvent_t e0, e1;
  PB_CUDA(cudaEventCreate(&e0));
  PB_CUDA(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 6; ++rep) {
    PB_CUDA(cudaEventRecord(e0, ctx->stream));
    pb::copy_kernel<<<ctx->num_sms * 16, 512, 0, ctx->stream>>>(a, b, n4);
    PB_LAUNCHED(ctx);
    PB_CUDA(cudaEventRecord(e1, ctx->stream));
    PB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    PB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::max(best, 2.0 * n4 * 16 / (ms * 1e-3) / 1e9);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(a);
  cudaFree(b);
  *gbs = best;
  return 0;
}

}  // extern "C"
