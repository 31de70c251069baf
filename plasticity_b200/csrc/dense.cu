// DenseLayer kernels (nnet/dense_layer.cc:12-55).
//
//   evaluate      O[b][o]  = W[o][in] + sum_i W[o][i] * I[b][i]
//   input_delta   dI[b][i] = sum_o G[b][o] * W[o][i]
//   weight update g[o][i]  = sum_b G[b][o] * (i < in ? I[b][i] : 1)
//
// Weights stay in the reference layout: row-major [out][in+1], bias last
// (nnet/symbol_generator.h:58-86).  All three are one FP32 FFMA GEMM with
// different operand majors; small batches (per-sample SGD) take GEMV-shaped
// kernels that stream W once at HBM speed.
#include "common.cuh"
#include "kernels.cuh"

namespace pb {

// ---------------------------------------------------------------------------
// Tiled FP32 GEMM: C(m,n) = sum_k A(m,k) * B(k,n).  64x64x16 tile, 256
// threads, 4x4 register tile per thread, operands staged through shared
// memory as [k][m] / [k][n] so the inner loop reads two float4 per 16 FFMA.
// Operand access goes through functors so one kernel serves the three dense
// products; each functor says which index is contiguous in memory so the
// global loads coalesce either way.
// ---------------------------------------------------------------------------
constexpr int GBM = 64, GBN = 64, GBK = 16, GPAD = 4;

template <class OpA, class OpB, class Epi>
__global__ void __launch_bounds__(256)
gemm_kernel(int M, int N, int K, OpA opa, OpB opb, Epi epi) {
  __shared__ __align__(16) float As[GBK][GBM + GPAD];
  __shared__ __align__(16) float Bs[GBK][GBN + GPAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int k0 = 0; k0 < K; k0 += GBK) {
#pragma unroll
    for (int it = 0; it < (GBM * GBK) / 256; ++it) {
      const int e = tid + it * 256;
      int m, k;
      if (OpA::kContiguousK) { m = e / GBK; k = e % GBK; }
      else { m = e % GBM; k = e / GBM; }
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < M && gk < K) ? opa(gm, gk) : 0.0f;
    }
#pragma unroll
    for (int it = 0; it < (GBN * GBK) / 256; ++it) {
      const int e = tid + it * 256;
      int n, k;
      if (OpB::kContiguousK) { n = e / GBK; k = e % GBK; }
      else { n = e % GBN; k = e / GBN; }
      const int gn = n0 + n, gk = k0 + k;
      Bs[k][n] = (gn < N && gk < K) ? opb(gk, gn) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      const float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < N) epi(gm, gn, acc[i][j]);
    }
  }
}

// ---- operand / epilogue functors ---------------------------------------------

struct FwdA {  // A(m=b, k=i) = I[b][i]
  static constexpr bool kContiguousK = true;
  const float *I; int64_t in;
  __device__ float operator()(int b, int i) const { return I[(int64_t)b * in + i]; }
};
struct FwdB {  // B(k=i, n=o) = W[o][i]
  static constexpr bool kContiguousK = true;
  const float *W; int64_t ld;
  __device__ float operator()(int i, int o) const { return W[(int64_t)o * ld + i]; }
};
struct FwdEpi {
  const float *W; float *O; int64_t ld, in, out;
  __device__ void operator()(int b, int o, float acc) const {
    O[(int64_t)b * out + o] = W[(int64_t)o * ld + in] + acc;
  }
};

struct BwdA {  // A(m=b, k=o) = G[b][o]
  static constexpr bool kContiguousK = true;
  const float *G; int64_t out;
  __device__ float operator()(int b, int o) const { return G[(int64_t)b * out + o]; }
};
struct BwdB {  // B(k=o, n=i) = W[o][i]
  static constexpr bool kContiguousK = false;
  const float *W; int64_t ld;
  __device__ float operator()(int o, int i) const { return W[(int64_t)o * ld + i]; }
};
struct BwdEpi {
  float *dI; int64_t in;
  __device__ void operator()(int b, int i, float acc) const { dI[(int64_t)b * in + i] = acc; }
};

struct WgA {  // A(m=o, k=b) = G[b][o]
  static constexpr bool kContiguousK = false;
  const float *G; int64_t out;
  __device__ float operator()(int o, int b) const { return G[(int64_t)b * out + o]; }
};
struct WgB {  // B(k=b, n=i) = i < in ? I[b][i] : 1   (the bias column)
  static constexpr bool kContiguousK = false;
  const float *I; int64_t in;
  __device__ float operator()(int b, int i) const {
    return i < in ? I[(int64_t)b * in + i] : 1.0f;
  }
};
struct WgEpi {
  const float *W; float *out; const float *lr; int64_t ld; int mode;
  __device__ void operator()(int o, int i, float acc) const {
    apply_update(mode, out, W, (int64_t)o * ld + i, lr[0], acc);
  }
};

// ---- small-batch (GEMV) kernels -------------------------------------------------

// One warp per (b, o): lanes stride over the row of W (coalesced), shuffle sum.
__global__ void dense_fwd_gemv_kernel(const float *__restrict__ I, const float *__restrict__ W,
                                      float *__restrict__ O, int in, int out, int64_t batch) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t ld = (int64_t)in + 1;
  for (int64_t t = warp; t < batch * out; t += nwarps) {
    const int64_t b = t / out;
    const int o = (int)(t - b * out);
    const float *w = W + o * ld, *x = I + b * in;
    float acc = 0.0f;
    for (int i = lane; i < in; i += 32) acc = fmaf(w[i], x[i], acc);
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) O[t] = w[in] + acc;
  }
}

// One thread per (b, i): walks the outputs; lanes read consecutive i of a row.
__global__ void dense_bwd_gemv_kernel(const float *__restrict__ W, const float *__restrict__ G,
                                      float *__restrict__ dI, int in, int out, int64_t batch) {
  const int64_t ld = (int64_t)in + 1;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < batch * in;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / in;
    const int i = (int)(t - b * in);
    const float *g = G + b * out;
    float acc = 0.0f;
    for (int o = 0; o < out; ++o) acc = fmaf(g[o], W[o * ld + i], acc);
    dI[t] = acc;
  }
}

// One thread per weight; sums over the (small) batch.
__global__ void dense_wupd_small_kernel(const float *__restrict__ I, const float *W,
                                        const float *__restrict__ G, float *out,
                                        const float *__restrict__ lr, int in, int nout,
                                        int64_t batch, int mode) {
  const int64_t ld = (int64_t)in + 1, nw = ld * nout;
  const float rate = lr[0];
  for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < nw;
       w += (int64_t)gridDim.x * blockDim.x) {
    const int64_t node = w / ld;
    const int edge = (int)(w - node * ld);
    float g = 0.0f;
    for (int64_t b = 0; b < batch; ++b)
      g += (edge < in) ? G[b * nout + node] * I[b * in + edge] : G[b * nout + node];
    apply_update(mode, out, W, w, rate, g);
  }
}

constexpr int64_t kGemvMaxBatch = 8;

int dense_evaluate(pb_context *ctx, const pb_layer_desc *l, const float *I, const float *W,
                   float *O, int64_t batch) {
  const int64_t in = l->in, out = l->out;
  if (batch * out == 0) return 0;
  if (batch <= kGemvMaxBatch) {
    const int T = 256;
    const int grid = (int)std::min<int64_t>((batch * out * 32 + T - 1) / T,
                                            (int64_t)ctx->num_sms * 8);
    dense_fwd_gemv_kernel<<<grid, T, 0, ctx->stream>>>(I, W, O, (int)in, (int)out, batch);
  } else {
    const int rc = dense_evaluate_umma(ctx, l, I, W, O, batch);
    if (rc >= 0) return rc;
    dim3 grid(pb_div_up(out, GBN), pb_div_up(batch, GBM));
    gemm_kernel<<<grid, 256, 0, ctx->stream>>>((int)batch, (int)out, (int)in, FwdA{I, in},
                                               FwdB{W, in + 1},
                                               FwdEpi{W, O, in + 1, in, out});
  }
  PB_LAUNCHED(ctx);
  return 0;
}

int dense_input_delta(pb_context *ctx, const pb_layer_desc *l, const float *W,
                      const float *G, float *dI, int64_t batch) {
  const int64_t in = l->in, out = l->out;
  if (batch * in == 0) return 0;
  if (batch <= kGemvMaxBatch) {
    dense_bwd_gemv_kernel<<<pb_ew_grid(ctx, batch * in, 128), 128, 0, ctx->stream>>>(
        W, G, dI, (int)in, (int)out, batch);
  } else {
    const int rc = dense_input_delta_umma(ctx, l, W, G, dI, batch);
    if (rc >= 0) return rc;
    dim3 grid(pb_div_up(in, GBN), pb_div_up(batch, GBM));
    gemm_kernel<<<grid, 256, 0, ctx->stream>>>((int)batch, (int)in, (int)out, BwdA{G, out},
                                               BwdB{W, in + 1}, BwdEpi{dI, in});
  }
  PB_LAUNCHED(ctx);
  return 0;
}

int dense_weight_update(pb_context *ctx, const pb_layer_desc *l, const float *I,
                        const float *W, const float *G, float *out, const float *lr,
                        int mode, int64_t batch) {
  const int64_t in = l->in, nout = l->out;
  const int64_t nw = (in + 1) * nout;
  if (nw == 0) return 0;
  {
    const int rc = dense_weight_update_skinny(ctx, l, I, W, G, out, lr, mode, batch);
    if (rc >= 0) return rc;
  }
  if (batch <= kGemvMaxBatch) {
    dense_wupd_small_kernel<<<pb_ew_grid(ctx, nw, 256), 256, 0, ctx->stream>>>(
        I, W, G, out, lr, (int)in, (int)nout, batch, mode);
  } else {
    const int rc = dense_weight_update_umma(ctx, l, I, W, G, out, lr, mode, batch);
    if (rc >= 0) return rc;
    dim3 grid(pb_div_up(in + 1, GBN), pb_div_up(nout, GBM));
    gemm_kernel<<<grid, 256, 0, ctx->stream>>>((int)nout, (int)(in + 1), (int)batch,
                                               WgA{G, nout}, WgB{I, in},
                                               WgEpi{W, out, lr, in + 1, mode});
  }
  PB_LAUNCHED(ctx);
  return 0;
}

}  // namespace pb
