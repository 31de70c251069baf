// Fused classifier head of the batched gradient pass:
//
//   DenseLayer (nnet/dense_layer.cc:12-55) -> ActivationLayer
//   (nnet/activation_layer.cc:6-22) -> SoftmaxLayer (nnet/softmax_layer.cc:9-59)
//   -> ErrorLayer gradient (nnet/error_layer.cc:81-91)
//
// forward AND backward for one sample by one warp, in one launch: the dense
// output has at most 32 values (10 for CIFAR, nnet/cifar_test.cc:343-349), so the
// whole chain lives in one register per lane with warp-shuffle reductions; the
// dense weights sit in shared memory.  Replaces 8 launch-latency-bound kernels
// (dense/act/softmax forward, error_gradients, softmax/act/dense input_delta) —
// and leaves dE/d(dense output) behind for the dense weight gradient, a skinny
// split-K reduction (out <= 32 rows, K = batch).
#include "common.cuh"
#include "kernels.cuh"

namespace pb {
namespace {

__device__ __forceinline__ float wmax(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float wsum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int ACT, int LOSS>
__global__ void __launch_bounds__(256)
head_kernel(const float *__restrict__ X, const float *__restrict__ W, const float *__restrict__ E,
            float *__restrict__ Z, float *__restrict__ A, float *__restrict__ Y,
            float *__restrict__ Gz, float *__restrict__ dX, int in, int out, int64_t batch,
            int early_w) {
  extern __shared__ float w_s[];  // [out][in+1], reference layout
  const int ld = in + 1;
  if (!early_w) chain_wait();  // chained launch, common.cuh
  {
    // L2 loads (see conv_umma.cu), sixteen per thread in flight before the first store
    const int nw = out * ld;
    for (int base = 0; base < nw; base += 16 * 256) {
      float wv[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int i = base + threadIdx.x + k * 256;
        wv[k] = i < nw ? __ldcg(W + i) : 0.0f;
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int i = base + threadIdx.x + k * 256;
        if (i < nw) w_s[i] = wv[k];
      }
    }
  }
  __syncthreads();
  if (early_w) chain_wait();
  chain_release();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  constexpr int KX = 8;  // inputs per lane and pass: a pass covers 32 * KX inputs
  for (int64_t b = warp; b < batch; b += nwarps) {
    const float *x = X + b * in;
    // dense forward.  Every lane keeps a partial sum per output over its own inputs (the
    // loads of a pass are issued together, the outputs' FMA chains are independent), then a
    // halving exchange leaves the total of output o in lane o: 31 shuffles for all outputs
    // instead of five per output.
    float acc[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) acc[o] = 0.0f;
    for (int c0 = 0; c0 < in; c0 += 32 * KX) {
      float xr[KX];
      int ic[KX];
#pragma unroll
      for (int k = 0; k < KX; ++k) {
        const int i = c0 + 32 * k + lane;
        ic[k] = i < in ? i : in - 1;
        xr[k] = i < in ? x[i] : 0.0f;
      }
#pragma unroll
      for (int o = 0; o < 32; ++o) {
        if (o < out) {
          const float *w = w_s + o * ld;
#pragma unroll
          for (int k = 0; k < KX; ++k) acc[o] = fmaf(w[ic[k]], xr[k], acc[o]);
        }
      }
    }
#pragma unroll
    for (int h = 16; h >= 1; h >>= 1) {
      const bool up = (lane & h) != 0;
#pragma unroll
      for (int j = 0; j < h; ++j) {
        const float keep = up ? acc[j + h] : acc[j];
        const float give = up ? acc[j] : acc[j + h];
        acc[j] = keep + __shfl_xor_sync(0xffffffffu, give, h);
      }
    }
    const bool on = lane < out;
    const float z = on ? w_s[lane * ld + in] + acc[0] : 0.0f;
    const float a = act_fwd<ACT>(z);
    // softmax forward (e == 2.71828, max-shifted, ascending-lane sum order irrelevant at 1e-4)
    const float m = wmax(on ? a : -INFINITY);
    const float p = on ? ref_exp(a - m) : 0.0f;
    const float s = wsum(p);
    const float y = p / s;
    // error gradient dE/dy
    const float e = on ? E[b * out + lane] : 0.0f;
    float g;
    if (LOSS == PB_MEAN_SQUARED) g = y - e;
    else g = -1.0f * (e * (1.0f / (y + kRefEps)));
    if (!on) g = 0.0f;
    // softmax backward: dI[k] = sum_i (O_i * (delta_ik - O_k)) * G_i, ascending i
    float gs = 0.0f;
    for (int i = 0; i < out; ++i) {
      const float yi = __shfl_sync(0xffffffffu, y, i);
      const float gi = __shfl_sync(0xffffffffu, g, i);
      gs += (yi * (((i == lane) ? 1.0f : 0.0f) - y)) * gi;
    }
    const float gz = on ? act_bwd<ACT>(z, gs) : 0.0f;
    if (on) {
      Z[b * out + lane] = z;
      A[b * out + lane] = a;
      Y[b * out + lane] = y;
      Gz[b * out + lane] = gz;
    }
    // dense input gradient: dX[i] = sum_o gz[o] * W[o][i], ascending o; gz is broadcast to
    // registers once (acc is free again)
#pragma unroll
    for (int o = 0; o < 32; ++o) acc[o] = __shfl_sync(0xffffffffu, gz, o);
    for (int c0 = 0; c0 < in; c0 += 32 * KX) {
      float d[KX];
      int ic[KX];
#pragma unroll
      for (int k = 0; k < KX; ++k) {
        const int i = c0 + 32 * k + lane;
        ic[k] = i < in ? i : in - 1;
        d[k] = 0.0f;
      }
#pragma unroll
      for (int o = 0; o < 32; ++o) {
        if (o < out) {
          const float *w = w_s + o * ld;
#pragma unroll
          for (int k = 0; k < KX; ++k) d[k] = fmaf(acc[o], w[ic[k]], d[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < KX; ++k) {
        const int i = c0 + 32 * k + lane;
        if (i < in) dX[b * in + i] = d[k];
      }
    }
  }
}

// partial[split][o*(in+1) + i] = sum_{b in split} G[b][o] * (i < in ? X[b][i] : 1)
// CTA: 32 columns i (lanes) x 8 warps each walking 1/8 of the split's samples.
template <int OUT>
__global__ void __launch_bounds__(256)
dense_wgrad_skinny_kernel(const float *__restrict__ X, const float *__restrict__ G,
                          float *__restrict__ partial, int in, int out, int64_t batch,
                          int64_t per_split) {
  __shared__ float red[8][OUT][33];
  chain_wait();
  chain_release();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  const int ld = in + 1;
  const int64_t b0 = blockIdx.y * per_split;
  const int64_t b1 = min(batch, b0 + per_split);
  float acc[OUT];
#pragma unroll
  for (int o = 0; o < OUT; ++o) acc[o] = 0.0f;
  for (int64_t b = b0 + w; b < b1; b += 8) {
    const float x = i < in ? X[b * in + i] : 1.0f;
    const float gl = lane < out ? G[b * out + lane] : 0.0f;
#pragma unroll
    for (int o = 0; o < OUT; ++o) acc[o] = fmaf(__shfl_sync(0xffffffffu, gl, o), x, acc[o]);
  }
#pragma unroll
  for (int o = 0; o < OUT; ++o) red[w][o][lane] = acc[o];
  __syncthreads();
  for (int t = threadIdx.x; t < out * 32; t += blockDim.x) {
    const int o = t >> 5, l = t & 31;
    const int col = blockIdx.x * 32 + l;
    if (col >= ld) continue;
    float sum = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += red[k][o][l];
    partial[(int64_t)blockIdx.y * out * ld + (int64_t)o * ld + col] = sum;
  }
}

template <int ACT>
int launch_head(pb_context *ctx, int loss, const float *X, const float *W, const float *E, float *Z,
                float *A, float *Y, float *Gz, float *dX, int in, int out, int64_t batch) {
  const size_t smem = sizeof(float) * (size_t)out * (in + 1);
  const int grid = (int)std::min<int64_t>((batch + 7) / 8, (int64_t)ctx->num_sms * 2);
  const int early_w = ctx->chain == 2;
  if (loss == PB_MEAN_SQUARED) {
    auto k = head_kernel<ACT, PB_MEAN_SQUARED>;
    if (smem > 48 * 1024) PB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PB_CUDA(launch_chained(ctx, k, dim3(grid), dim3(256), smem, X, W, E, Z, A, Y, Gz, dX, in, out,
                           batch, early_w));
  } else {
    auto k = head_kernel<ACT, PB_CROSS_ENTROPY>;
    if (smem > 48 * 1024) PB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PB_CUDA(launch_chained(ctx, k, dim3(grid), dim3(256), smem, X, W, E, Z, A, Y, Gz, dX, in, out,
                           batch, early_w));
  }
  PB_LAUNCHED(ctx);
  return 0;
}

}  // namespace

// Returns -1 outside the envelope (out <= 32 == softmax width, weights fit shared memory).
int head_forward_backward(pb_context *ctx, const pb_layer_desc *dense, int act, int64_t softmax_n,
                          int loss, const float *X, const float *W, const float *E, float *Z,
                          float *A, float *Y, float *Gz, float *dX, int64_t batch) {
  const int64_t in = dense->in, out = dense->out;
  if (out < 1 || out > 32 || softmax_n != out || in < 1) return -1;
  if ((size_t)out * (in + 1) * sizeof(float) > 160 * 1024 || batch < 1) return -1;
  switch (act) {
    case PB_IDENTITY: return launch_head<PB_IDENTITY>(ctx, loss, X, W, E, Z, A, Y, Gz, dX, (int)in, (int)out, batch);
    case PB_RELU: return launch_head<PB_RELU>(ctx, loss, X, W, E, Z, A, Y, Gz, dX, (int)in, (int)out, batch);
    case PB_LEAKY_RELU: return launch_head<PB_LEAKY_RELU>(ctx, loss, X, W, E, Z, A, Y, Gz, dX, (int)in, (int)out, batch);
    case PB_SIGMOID: return launch_head<PB_SIGMOID>(ctx, loss, X, W, E, Z, A, Y, Gz, dX, (int)in, (int)out, batch);
  }
  return -1;
}

// Dense weight gradient for a layer with few outputs and a long batch: split-K
// partial sums + the deterministic reduce/update kernel.  Returns -1 outside the envelope.
int dense_weight_update_skinny(pb_context *ctx, const pb_layer_desc *l, const float *X,
                               const float *W, const float *G, float *out, const float *lr,
                               int mode, int64_t batch) {
  const int64_t in = l->in, nout = l->out;
  if (nout < 1 || nout > 16 || batch < 64) return -1;
  const int64_t ld = in + 1, nw = ld * nout;
  const int col_tiles = pb_div_up(ld, 32);
  int splits = (int)std::max<int64_t>(1, std::min<int64_t>(batch / 32, (2 * ctx->num_sms) / col_tiles));
  const int64_t per_split = (batch + splits - 1) / splits;
  splits = (int)((batch + per_split - 1) / per_split);
  if (ensure_workspace(ctx, (size_t)splits * nw)) return 1;
  dim3 grid(col_tiles, splits);
  if (ctx->fork_operands_stable)
    PB_CUDA(launch_forked(ctx, dense_wgrad_skinny_kernel<16>, grid, dim3(256), 0, X, G,
                          ctx->workspace, (int)in, (int)nout, batch, per_split));
  else
    PB_CUDA(launch_chained(ctx, dense_wgrad_skinny_kernel<16>, grid, dim3(256), 0, X, G,
                           ctx->workspace, (int)in, (int)nout, batch, per_split));
  PB_LAUNCHED(ctx);
  return reduce_partials_update(ctx, ctx->workspace, splits, (int)nw, W, out, lr, mode);
}

}  // namespace pb
