// HBM-bound layers: activation, max-pool, softmax, error, vector_accumulate.
// One fp32 value in, one out; grid-stride loops, float4 where the layout
// allows; softmax and the error sum use warp-shuffle reductions.
//
// Semantics follow the reference's generated kernels:
//   activation  nnet/activation_layer.cc:6-22 + symbolic/symbolic_util.cc:10-28
//   max pool    nnet/max_pool_layer.cc:40-141
//   softmax     nnet/softmax_layer.cc:9-59
//   error       nnet/error_layer.cc:81-91, nnet/kernels/error.kernel.cl
//   accumulate  nnet/kernels/combine.kernel.cl:2-5
#include "common.cuh"
#include "kernels.cuh"

namespace pb {

// ---- activation -----------------------------------------------------------

template <int ACT>
__global__ void act_fwd_kernel(const float *__restrict__ I, float *__restrict__ O,
                               int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t n4 = n >> 2;
  const float4 *I4 = reinterpret_cast<const float4 *>(I);
  float4 *O4 = reinterpret_cast<float4 *>(O);
  for (int64_t i = tid; i < n4; i += stride) {
    float4 v = I4[i];
    v.x = act_fwd<ACT>(v.x);
    v.y = act_fwd<ACT>(v.y);
    v.z = act_fwd<ACT>(v.z);
    v.w = act_fwd<ACT>(v.w);
    O4[i] = v;
  }
  for (int64_t i = (n4 << 2) + tid; i < n; i += stride) O[i] = act_fwd<ACT>(I[i]);
}

template <int ACT>
__global__ void act_bwd_kernel(const float *__restrict__ I, const float *__restrict__ G,
                               float *__restrict__ dI, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t n4 = n >> 2;
  const float4 *I4 = reinterpret_cast<const float4 *>(I);
  const float4 *G4 = reinterpret_cast<const float4 *>(G);
  float4 *D4 = reinterpret_cast<float4 *>(dI);
  for (int64_t i = tid; i < n4; i += stride) {
    const float4 x = I4[i], g = G4[i];
    float4 d;
    d.x = act_bwd<ACT>(x.x, g.x);
    d.y = act_bwd<ACT>(x.y, g.y);
    d.z = act_bwd<ACT>(x.z, g.z);
    d.w = act_bwd<ACT>(x.w, g.w);
    D4[i] = d;
  }
  for (int64_t i = (n4 << 2) + tid; i < n; i += stride) dI[i] = act_bwd<ACT>(I[i], G[i]);
}

static inline bool aligned16(const void *p) { return (((uintptr_t)p) & 15) == 0; }

// Scalar-only variants for unaligned sub-buffers.
template <int ACT>
__global__ void act_fwd_scalar_kernel(const float *__restrict__ I, float *__restrict__ O,
                                      int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    O[i] = act_fwd<ACT>(I[i]);
}
template <int ACT>
__global__ void act_bwd_scalar_kernel(const float *__restrict__ I,
                                      const float *__restrict__ G,
                                      float *__restrict__ dI, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    dI[i] = act_bwd<ACT>(I[i], G[i]);
}

int activation_evaluate(pb_context *ctx, int act, const float *I, float *O, int64_t n) {
  if (n == 0) return 0;
  const int T = 256;
  const bool vec = aligned16(I) && aligned16(O);
  const int grid = pb_ew_grid(ctx, vec ? (n + 3) / 4 : n, T);
#define PB_ACT_FWD(A)                                                        \
  case A:                                                                    \
    if (vec) act_fwd_kernel<A><<<grid, T, 0, ctx->stream>>>(I, O, n);        \
    else act_fwd_scalar_kernel<A><<<grid, T, 0, ctx->stream>>>(I, O, n);     \
    break;
  switch (act) {
    PB_ACT_FWD(PB_IDENTITY)
    PB_ACT_FWD(PB_RELU)
    PB_ACT_FWD(PB_LEAKY_RELU)
    PB_ACT_FWD(PB_SIGMOID)
    default:
      return fail("activation: only Identity/Relu/LeakyRelu/Sigmoid exist without the "
                  "reference's code generator");
  }
#undef PB_ACT_FWD
  PB_LAUNCHED(ctx);
  return 0;
}

int activation_input_delta(pb_context *ctx, int act, const float *I, const float *G,
                           float *dI, int64_t n) {
  if (n == 0) return 0;
  const int T = 256;
  const bool vec = aligned16(I) && aligned16(G) && aligned16(dI);
  const int grid = pb_ew_grid(ctx, vec ? (n + 3) / 4 : n, T);
#define PB_ACT_BWD(A)                                                          \
  case A:                                                                      \
    if (vec) act_bwd_kernel<A><<<grid, T, 0, ctx->stream>>>(I, G, dI, n);      \
    else act_bwd_scalar_kernel<A><<<grid, T, 0, ctx->stream>>>(I, G, dI, n);   \
    break;
  switch (act) {
    PB_ACT_BWD(PB_IDENTITY)
    PB_ACT_BWD(PB_RELU)
    PB_ACT_BWD(PB_LEAKY_RELU)
    PB_ACT_BWD(PB_SIGMOID)
    default:
      return fail("activation: unknown activation function");
  }
#undef PB_ACT_BWD
  PB_LAUNCHED(ctx);
  return 0;
}

// ---- max pool ---------------------------------------------------------------
// Forward: one thread per output; strict '>' scan rows-then-cols from -inf, so
// the first maximum wins and only the VALUE is kept (no index), as in the
// reference.  Backward: one thread per input; gradient flows to every input
// that no window element strictly exceeds (all tied maxima), exact 0 elsewhere.

__global__ void pool_fwd_kernel(const float *__restrict__ I, float *__restrict__ O,
                                int W, int H, int D, int ow, int oh, int64_t total) {
  const int gw = W / ow, gh = H / oh;
  const int oplane = ow * oh;
  const int64_t per_out = (int64_t)oplane * D, per_in = (int64_t)W * H * D;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / per_out;
    const int idx = (int)(t - b * per_out);
    const int z = idx / oplane, rem = idx - z * oplane;
    const int orow = rem / ow, ocol = rem - orow * ow;
    const float *src = I + b * per_in + (int64_t)z * W * H + (orow * gh) * W + ocol * gw;
    float m = -INFINITY;
    for (int r = 0; r < gh; ++r)
      for (int c = 0; c < gw; ++c) {
        const float v = src[r * W + c];
        if (v > m) m = v;
      }
    O[t] = m;
  }
}

// Exact 2x2 windows, W % 4 == 0: one thread reads a float4 from each of the two window
// rows and writes two outputs — every access is a full-width vector, the kernel runs at
// HBM speed.  Same strict '>' scan order as above (row 0 then row 1, left to right).
__global__ void pool2x2_fwd_kernel(const float *__restrict__ I, float *__restrict__ O, int W,
                                   int H, int64_t total /* batch*D*(H/2)*(W/4) */) {
  const int qw = W >> 2, oh = H >> 1;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t plane = t / (qw * oh);
    const int rem = (int)(t - plane * (qw * oh));
    const int orow = rem / qw, q = rem - orow * qw;
    const float *src = I + plane * (int64_t)W * H + (int64_t)(2 * orow) * W + 4 * q;
    const float4 a = __ldg(reinterpret_cast<const float4 *>(src));
    const float4 b = __ldg(reinterpret_cast<const float4 *>(src + W));
    float m0 = -INFINITY, m1 = -INFINITY;
    if (a.x > m0) m0 = a.x;
    if (a.y > m0) m0 = a.y;
    if (b.x > m0) m0 = b.x;
    if (b.y > m0) m0 = b.y;
    if (a.z > m1) m1 = a.z;
    if (a.w > m1) m1 = a.w;
    if (b.z > m1) m1 = b.z;
    if (b.w > m1) m1 = b.w;
    *reinterpret_cast<float2 *>(O + plane * (int64_t)(W >> 1) * oh + (int64_t)orow * (W >> 1) +
                                2 * q) = make_float2(m0, m1);
  }
}

__global__ void pool_bwd_kernel(const float *__restrict__ I, const float *__restrict__ G,
                                float *__restrict__ dI, int W, int H, int D, int ow,
                                int oh, int64_t total) {
  const int gw = W / ow, gh = H / oh;
  const int iplane = W * H;
  const int64_t per_in = (int64_t)iplane * D, per_out = (int64_t)ow * oh * D;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / per_in;
    const int idx = (int)(t - b * per_in);
    const int z = idx / iplane, rem = idx - z * iplane;
    const int r = rem / W, c = rem - r * W;
    const int orow = r / gh, ocol = c / gw;
    float out = 0.0f;
    if (orow < oh && ocol < ow) {  // past the last full window: feeds no output
      const float self = I[t];
      const float *win = I + b * per_in + (int64_t)z * iplane + (orow * gh) * W + ocol * gw;
      bool beaten = false;
      for (int gr = 0; gr < gh; ++gr)
        for (int gc = 0; gc < gw; ++gc) beaten |= (win[gr * W + gc] > self);
      if (!beaten) out = G[b * per_out + (int64_t)z * ow * oh + orow * ow + ocol];
    }
    dI[t] = out;
  }
}

int pool_evaluate(pb_context *ctx, const pb_layer_desc *l, const float *I, float *O,
                  int64_t batch) {
  const int64_t total = batch * l->out_width * l->out_height * l->depth;
  if (total == 0) return 0;
  if (l->out_width * 2 == l->width && l->out_height * 2 == l->height && (l->width & 3) == 0 &&
      (reinterpret_cast<uintptr_t>(I) & 15) == 0 && (reinterpret_cast<uintptr_t>(O) & 7) == 0) {
    const int64_t quads = total / 2;
    pool2x2_fwd_kernel<<<pb_ew_grid(ctx, quads, 256), 256, 0, ctx->stream>>>(
        I, O, (int)l->width, (int)l->height, quads);
    PB_LAUNCHED(ctx);
    return 0;
  }
  pool_fwd_kernel<<<pb_ew_grid(ctx, total, 256), 256, 0, ctx->stream>>>(
      I, O, (int)l->width, (int)l->height, (int)l->depth, (int)l->out_width,
      (int)l->out_height, total);
  PB_LAUNCHED(ctx);
  return 0;
}

int pool_input_delta(pb_context *ctx, const pb_layer_desc *l, const float *I,
                     const float *G, float *dI, int64_t batch) {
  const int64_t total = batch * l->width * l->height * l->depth;
  if (total == 0) return 0;
  pool_bwd_kernel<<<pb_ew_grid(ctx, total, 256), 256, 0, ctx->stream>>>(
      I, G, dI, (int)l->width, (int)l->height, (int)l->depth, (int)l->out_width,
      (int)l->out_height, total);
  PB_LAUNCHED(ctx);
  return 0;
}

// ---- fused activation + 2x2 max-pool backward -------------------------------
// dO[x] = act'(O[x]) * (act(O[x]) is a maximum of its 2x2 window ? Gp[window] : 0)
// == activation_input_delta(pool_input_delta(...)) of the unfused path, value
// for value, without materialising the activation output or the pool's input
// gradient.  A thread owns two windows (2 rows x 4 columns): float4 loads.
template <int ACT>
__global__ void act_pool_bwd_kernel(const float *__restrict__ O, const float *__restrict__ Gp,
                                    float *__restrict__ dO, int W, int H, int64_t total) {
  chain_wait();  // chained launch, common.cuh
  chain_release();
  const int wq = W >> 2, oh = H >> 1, ow = W >> 1;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int xq = (int)(t % wq);
    const int64_t u = t / wq;
    const int orow = (int)(u % oh);
    const int64_t bz = u / oh;
    const int64_t base = bz * H * W + (int64_t)(2 * orow) * W + 4 * xq;
    const float4 o0 = *reinterpret_cast<const float4 *>(O + base);
    const float4 o1 = *reinterpret_cast<const float4 *>(O + base + W);
    const float2 g = *reinterpret_cast<const float2 *>(Gp + bz * oh * ow + (int64_t)orow * ow + 2 * xq);
    const float a00 = act_fwd<ACT>(o0.x), a01 = act_fwd<ACT>(o0.y), a02 = act_fwd<ACT>(o0.z),
                a03 = act_fwd<ACT>(o0.w);
    const float a10 = act_fwd<ACT>(o1.x), a11 = act_fwd<ACT>(o1.y), a12 = act_fwd<ACT>(o1.z),
                a13 = act_fwd<ACT>(o1.w);
    const float m0 = fmaxf(fmaxf(a00, a01), fmaxf(a10, a11));
    const float m1 = fmaxf(fmaxf(a02, a03), fmaxf(a12, a13));
    float4 d0, d1;
    d0.x = act_bwd<ACT>(o0.x, a00 < m0 ? 0.0f : g.x);
    d0.y = act_bwd<ACT>(o0.y, a01 < m0 ? 0.0f : g.x);
    d0.z = act_bwd<ACT>(o0.z, a02 < m1 ? 0.0f : g.y);
    d0.w = act_bwd<ACT>(o0.w, a03 < m1 ? 0.0f : g.y);
    d1.x = act_bwd<ACT>(o1.x, a10 < m0 ? 0.0f : g.x);
    d1.y = act_bwd<ACT>(o1.y, a11 < m0 ? 0.0f : g.x);
    d1.z = act_bwd<ACT>(o1.z, a12 < m1 ? 0.0f : g.y);
    d1.w = act_bwd<ACT>(o1.w, a13 < m1 ? 0.0f : g.y);
    *reinterpret_cast<float4 *>(dO + base) = d0;
    *reinterpret_cast<float4 *>(dO + base + W) = d1;
  }
}

// Returns -1 when the block is not a (any activation, exact 2x2 pool) pair the
// fused kernel covers.
int act_pool_input_delta(pb_context *ctx, int act, const pb_layer_desc *pool, const float *O,
                         const float *Gp, float *dO, int64_t batch) {
  const int W = (int)pool->width, H = (int)pool->height;
  if (pool->out_width * 2 != W || pool->out_height * 2 != H || (W & 3)) return -1;
  if (!aligned16(O) || !aligned16(dO) || (((uintptr_t)Gp) & 7)) return -1;
  const int64_t total = batch * pool->depth * (H / 2) * (W / 4);
  if (total == 0) return 0;
  const int grid = pb_ew_grid(ctx, total, 256);
  cudaError_t err = cudaSuccess;
  switch (act) {
    case PB_IDENTITY: err = launch_chained(ctx, act_pool_bwd_kernel<PB_IDENTITY>, dim3(grid), dim3(256), 0, O, Gp, dO, W, H, total); break;
    case PB_RELU: err = launch_chained(ctx, act_pool_bwd_kernel<PB_RELU>, dim3(grid), dim3(256), 0, O, Gp, dO, W, H, total); break;
    case PB_LEAKY_RELU: err = launch_chained(ctx, act_pool_bwd_kernel<PB_LEAKY_RELU>, dim3(grid), dim3(256), 0, O, Gp, dO, W, H, total); break;
    case PB_SIGMOID: err = launch_chained(ctx, act_pool_bwd_kernel<PB_SIGMOID>, dim3(grid), dim3(256), 0, O, Gp, dO, W, H, total); break;
    default: return -1;
  }
  PB_CUDA(err);
  PB_LAUNCHED(ctx);
  return 0;
}

// ---- softmax ----------------------------------------------------------------
// One warp per sample.  max and sum are warp-shuffle reductions.
//   O[k]  = e^(I[k]-max) / sum_i e^(I[i]-max)           (e == 2.71828)
//   dI[k] = sum_i (O[i] * (delta_ik - O[k])) * G[i]     term order of the reference

__device__ __forceinline__ float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void softmax_fwd_kernel(const float *__restrict__ I, float *__restrict__ O,
                                   int n, int64_t batch) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp; b < batch; b += nwarps) {
    const float *x = I + b * n;
    float m = -INFINITY;
    for (int i = lane; i < n; i += 32) m = fmaxf(m, x[i]);
    m = warp_max(m);
    float s = 0.0f;
    for (int i = lane; i < n; i += 32) s += ref_exp(x[i] - m);
    s = warp_sum(s);
    for (int i = lane; i < n; i += 32) O[b * n + i] = ref_exp(x[i] - m) / s;
  }
}

// n <= 32: O and G live in registers, one value per lane.
__global__ void softmax_bwd_small_kernel(const float *__restrict__ I,
                                         const float *__restrict__ G,
                                         float *__restrict__ dI, int n, int64_t batch) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp; b < batch; b += nwarps) {
    const bool on = lane < n;
    const float x = on ? I[b * n + lane] : -INFINITY;
    const float g = on ? G[b * n + lane] : 0.0f;
    const float m = warp_max(x);
    const float p = on ? ref_exp(x - m) : 0.0f;
    const float s = warp_sum(p);
    const float o = p / s;
    float acc = 0.0f;
    for (int i = 0; i < n; ++i) {  // ascending i, like the reference's sum
      const float oi = __shfl_sync(0xffffffffu, o, i);
      const float gi = __shfl_sync(0xffffffffu, g, i);
      acc += (oi * (((i == lane) ? 1.0f : 0.0f) - o)) * gi;
    }
    if (on) dI[b * n + lane] = acc;
  }
}

// general n: outputs staged in shared memory, one warp per sample.
__global__ void softmax_bwd_kernel(const float *__restrict__ I, const float *__restrict__ G,
                                   float *__restrict__ dI, int n, int64_t batch) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float *o = sm + (size_t)wib * 2 * n, *g = o + n;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp; b < batch; b += nwarps) {
    const float *x = I + b * n;
    float m = -INFINITY;
    for (int i = lane; i < n; i += 32) m = fmaxf(m, x[i]);
    m = warp_max(m);
    float s = 0.0f;
    for (int i = lane; i < n; i += 32) {
      const float p = ref_exp(x[i] - m);
      o[i] = p;
      g[i] = G[b * n + i];
      s += p;
    }
    s = warp_sum(s);
    __syncwarp();
    for (int i = lane; i < n; i += 32) o[i] = o[i] / s;
    __syncwarp();
    for (int k = lane; k < n; k += 32) {
      const float ok = o[k];
      float acc = 0.0f;
      for (int i = 0; i < n; ++i) acc += (o[i] * (((i == k) ? 1.0f : 0.0f) - ok)) * g[i];
      dI[b * n + k] = acc;
    }
    __syncwarp();
  }
}

// any n: the same arithmetic with the outputs staged in a global workspace ([sample][n]) instead
// of shared memory; one CTA per sample.  For layers wider than the shared-memory path holds.
__global__ void softmax_bwd_wide_kernel(const float *__restrict__ I, const float *__restrict__ G,
                                        float *__restrict__ dI, float *__restrict__ O, int n,
                                        int64_t batch) {
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int64_t b = blockIdx.x; b < batch; b += gridDim.x) {
    const float *x = I + b * n, *g = G + b * n;
    float *o = O + b * n;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, x[i]);
    m = warp_max(m);
    if (lane == 0) red[wib] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < nw; ++w) m = fmaxf(m, red[w]);
    __syncthreads();
    float s = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float p = ref_exp(x[i] - m);
      o[i] = p;
      s += p;
    }
    s = warp_sum(s);
    if (lane == 0) red[wib] = s;
    __syncthreads();
    s = 0.0f;
    for (int w = 0; w < nw; ++w) s += red[w];
    for (int i = threadIdx.x; i < n; i += blockDim.x) o[i] = o[i] / s;
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
      const float ok = o[k];
      float acc = 0.0f;
      for (int i = 0; i < n; ++i) acc += (o[i] * (((i == k) ? 1.0f : 0.0f) - ok)) * g[i];
      dI[b * n + k] = acc;
    }
    __syncthreads();
  }
}

int softmax_evaluate(pb_context *ctx, int64_t n, const float *I, float *O, int64_t batch) {
  if (n == 0 || batch == 0) return 0;
  const int T = 128;
  const int grid = (int)std::min<int64_t>((batch * 32 + T - 1) / T, (int64_t)ctx->num_sms * 16);
  softmax_fwd_kernel<<<grid, T, 0, ctx->stream>>>(I, O, (int)n, batch);
  PB_LAUNCHED(ctx);
  return 0;
}

int softmax_input_delta(pb_context *ctx, int64_t n, const float *I, const float *G,
                        float *dI, int64_t batch) {
  if (n == 0 || batch == 0) return 0;
  const int T = 128;
  const int grid = (int)std::min<int64_t>((batch * 32 + T - 1) / T, (int64_t)ctx->num_sms * 16);
  if (n <= 32) {
    softmax_bwd_small_kernel<<<grid, T, 0, ctx->stream>>>(I, G, dI, (int)n, batch);
  } else {
    const size_t smem = (size_t)(T / 32) * 2 * n * sizeof(float);
    if (smem > 200 * 1024) {
      if (ensure_workspace(ctx, (size_t)batch * n)) return 1;
      softmax_bwd_wide_kernel<<<(int)std::min<int64_t>(batch, (int64_t)ctx->num_sms * 4), 256, 0,
                                ctx->stream>>>(I, G, dI, ctx->workspace, (int)n, batch);
      PB_LAUNCHED(ctx);
      return 0;
    }
    if (smem > 48 * 1024)
      PB_CUDA(cudaFuncSetAttribute(softmax_bwd_kernel,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    softmax_bwd_kernel<<<grid, T, smem, ctx->stream>>>(I, G, dI, (int)n, batch);
  }
  PB_LAUNCHED(ctx);
  return 0;
}

// ---- error layer --------------------------------------------------------------
//   MSE: e = (E-O)^2/2, dE/dO = O-E
//   CE : e = -E*log(O+2.22045e-16)/log(2.71828), dE/dO = -E*(1/(O+2.22045e-16))

template <int LOSS>
__global__ void error_value_kernel(const float *__restrict__ O, const float *__restrict__ E,
                                   float *__restrict__ comp, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (LOSS == PB_MEAN_SQUARED) {
      const float d = E[i] - O[i];
      comp[i] = (d * d) / 2.0f;
    } else {
      // log_b(x) = ln(x) / ln(2.71828); the divisor is 0.999999327, not 1.
      comp[i] = -1.0f * (E[i] * (logf(O[i] + kRefEps) / 0.999999327347282f));
    }
  }
}

template <int LOSS>
__global__ void error_grad_kernel(const float *__restrict__ O, const float *__restrict__ E,
                                  float *__restrict__ g, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (LOSS == PB_MEAN_SQUARED) {
      g[i] = O[i] - E[i];
    } else {
      g[i] = -1.0f * (E[i] * (1.0f / (O[i] + kRefEps)));
    }
  }
}

__global__ void accumulate_kernel(float *__restrict__ a, const float *__restrict__ b,
                                  int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    a[i] += b[i];
}

}  // namespace pb

extern "C" {

int pb_error_value(pb_context *ctx, int loss, int64_t n, const float *O, const float *E,
                   float *components, int64_t batch) {
  PB_CHECK(ctx && O && E && components, "pb_error_value: null");
  const int64_t total = n * batch;
  if (total == 0) return 0;
  const int grid = pb_ew_grid(ctx, total, 256);
  if (loss == PB_MEAN_SQUARED)
    pb::error_value_kernel<PB_MEAN_SQUARED><<<grid, 256, 0, ctx->stream>>>(O, E, components, total);
  else if (loss == PB_CROSS_ENTROPY)
    pb::error_value_kernel<PB_CROSS_ENTROPY><<<grid, 256, 0, ctx->stream>>>(O, E, components, total);
  else
    return pb::fail("pb_error_value: unknown loss function");
  PB_LAUNCHED(ctx);
  return 0;
}

int pb_error_gradients(pb_context *ctx, int loss, int64_t n, const float *O, const float *E,
                       float *gradients, int64_t batch) {
  PB_CHECK(ctx && O && E && gradients, "pb_error_gradients: null");
  const int64_t total = n * batch;
  if (total == 0) return 0;
  const int grid = pb_ew_grid(ctx, total, 256);
  if (loss == PB_MEAN_SQUARED)
    pb::error_grad_kernel<PB_MEAN_SQUARED><<<grid, 256, 0, ctx->stream>>>(O, E, gradients, total);
  else if (loss == PB_CROSS_ENTROPY)
    pb::error_grad_kernel<PB_CROSS_ENTROPY><<<grid, 256, 0, ctx->stream>>>(O, E, gradients, total);
  else
    return pb::fail("pb_error_gradients: unknown loss function");
  PB_LAUNCHED(ctx);
  return 0;
}

int pb_vector_accumulate(pb_context *ctx, float *a, const float *b, int64_t n) {
  PB_CHECK(ctx && (n == 0 || (a && b)), "pb_vector_accumulate: null");
  if (n == 0) return 0;
  pb::accumulate_kernel<<<pb_ew_grid(ctx, n, 256), 256, 0, ctx->stream>>>(a, b, n);
  PB_LAUNCHED(ctx);
  return 0;
}

}  // extern "C"
