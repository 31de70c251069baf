// 3xTF32 tcgen05 GEMM for the batched DenseLayer products
// (nnet/dense_layer.cc:12-55) — the "real contraction" shapes of the wide MLP
// (4096-4096-4096-10, batch 1024) and of the YOLOv1 classifier.
//
//   evaluate      O[b][o]  = W[o][in] + sum_i I[b][i] * W[o][i]      M=b N=o K=i
//   input_delta   dI[b][i] = sum_o G[b][o] * W[o][i]                 M=b N=i K=o
//   weight update g[o][i]  = sum_b G[b][o] * I[b][i]                 M=o N=i K=b
//
// All three run on the kernel skeleton of umma_gemm.cuh; what differs is how each
// operand lies in HBM.  The reference layouts are kept verbatim — weights are
// row-major [out][in+1], an ODD pitch for even widths, so TMA could not fetch them;
// the producer warps' register path (float4, scalar or transposing loads) can.
#include "umma_gemm.cuh"

namespace pb {
namespace {

using namespace ugemm;

enum { EPI_PLAIN = 0, EPI_BIAS = 1, EPI_UPDATE = 2 };

struct DenseArgs : GemmCore {
  const float *A, *B;
  int64_t lda, ldb;
  int K;
  int epi;
  // EPI_PLAIN / EPI_BIAS: C[m*ldc + n] = acc (+ bias[n*ldbias])
  // EPI_UPDATE: apply_update(mode, C, Wold, m*ldc + n, lr, acc)
  float *C;
  int64_t ldc;
  const float *bias;
  int64_t ldbias;
  const float *Wold, *lr;
  int mode;
};

// TRANSPOSE: the epilogue writes C through the warp transpose tile (whole-warp accesses
// of one row: the read-modify-write of the SGD update); otherwise one row per thread.
template <int AMODE, int BMODE, bool TRANSPOSE = false>
struct DensePolicy {
  using Args = DenseArgs;
  static constexpr bool kEpiTranspose = TRANSPOSE;
  static constexpr int kEpiCols = 32;
  struct A : MatrixOperand<AMODE, BM, LA> {
    __device__ __forceinline__ void init(const Args &g) { this->setup(g.A, g.lda, g.M, g.K, BM); }
    __device__ __forceinline__ void tile(const Args &, int t, int) { this->set_tile(t); }
    __device__ __forceinline__ void load(const Args &, int ks, int p) { this->load_stage(ks, p); }
  };
  // the transposing (MN) loader maps one row per thread of the 128-thread group
  static constexpr int kMinBN = BMODE == OP_MN ? 128 : 64;
  template <int BN_T>
  struct B : MatrixOperand<BMODE, BN_T, BN_T + 2> {
    __device__ __forceinline__ void init(const Args &g) { this->setup(g.B, g.ldb, g.N, g.K, g.mma_n); }
    __device__ __forceinline__ void tile(const Args &, int t, int) { this->set_tile(t); }
    __device__ __forceinline__ void load(const Args &, int ks, int p) { this->load_stage(ks, p); }
  };

  static __device__ __forceinline__ void store1(const Args &g, int64_t m, int n, float acc) {
    if (g.epi == EPI_BIAS) g.C[m * g.ldc + n] = __ldg(g.bias + (int64_t)n * g.ldbias) + acc;
    else if (g.epi == EPI_UPDATE) apply_update(g.mode, g.C, g.Wold, m * g.ldc + n, g.lr[0], acc);
    else g.C[m * g.ldc + n] = acc;
  }

  // The accumulator arrives one ROW per lane; C is row-major, so the 32 x 32 block goes
  // through the warp's transpose tile and every global access of the warp is one
  // contiguous 128-byte run of one row (also the read-modify-write of the SGD update).
  static __device__ __forceinline__ void store_cols(const Args &g, int row0, int lane, int n0,
                                                 const float (&v)[32], int nv, float *stage) {
    if constexpr (!TRANSPOSE) {
      const int m = row0 + lane;
      if (m >= g.M) return;
      float *crow = g.C + (int64_t)m * g.ldc + n0;
      if (g.epi != EPI_UPDATE && nv == 32 && (g.ldc & 3) == 0 &&
          (reinterpret_cast<uintptr_t>(g.C) & 15) == 0) {
        const float *bp = g.bias + (int64_t)n0 * g.ldbias;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float b[4] = {0.f, 0.f, 0.f, 0.f};
          if (g.epi == EPI_BIAS) {
#pragma unroll
            for (int i = 0; i < 4; ++i, bp += g.ldbias) b[i] = __ldg(bp);
          }
          *reinterpret_cast<float4 *>(crow + 4 * q) = make_float4(
              v[4 * q] + b[0], v[4 * q + 1] + b[1], v[4 * q + 2] + b[2], v[4 * q + 3] + b[3]);
        }
        return;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nv) store1(g, m, n0 + j, v[j]);
      return;
    } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) stage[lane * 33 + j] = v[j];
    __syncwarp();
    const int rows = min(32, g.M - row0);
    const bool col_ok = lane < nv;
    float *c = g.C + (int64_t)row0 * g.ldc + n0 + lane;
    if (g.epi == EPI_UPDATE) {
      const float rate = g.lr[0];
      if (g.mode == PB_WEIGHT_DELTAS) {
#pragma unroll 8
        for (int r = 0; r < rows; ++r, c += g.ldc)
          if (col_ok) *c = -rate * stage[r * 33 + lane];
      } else {
        // out = W - lr*g (new_weights_*) or out += -lr*g (weight_deltas_* + vector_accumulate)
        const float *w = g.mode == PB_NEW_WEIGHTS ? g.Wold + (int64_t)row0 * g.ldc + n0 + lane : c;
        for (int r8 = 0; r8 < rows; r8 += 8) {
          float old[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            old[i] = (col_ok && r8 + i < rows) ? w[(int64_t)(r8 + i) * g.ldc] : 0.0f;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (col_ok && r8 + i < rows)
              c[(int64_t)(r8 + i) * g.ldc] = old[i] - rate * stage[(r8 + i) * 33 + lane];
        }
      }
    } else {
      const float b = (g.epi == EPI_BIAS && col_ok) ? __ldg(g.bias + (int64_t)(n0 + lane) * g.ldbias) : 0.0f;
#pragma unroll 8
      for (int r = 0; r < rows; ++r, c += g.ldc)
        if (col_ok) *c = stage[r * 33 + lane] + b;
    }
    __syncwarp();
    }
  }
};

// bias column of the dense weight gradient: g[o][in] = sum_b G[b][o]
// (nnet/dense_layer.cc:44-55, the "edge == in" branch).  One thread per output,
// lanes along o: every load is coalesced.
__global__ void dense_bias_update_kernel(const float *__restrict__ G, const float *W, float *out,
                                         const float *__restrict__ lr, int in, int nout,
                                         int64_t batch, int mode) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= nout) return;
  float acc = 0.0f;
  for (int64_t b = 0; b < batch; ++b) acc += G[b * nout + o];
  apply_update(mode, out, W, (int64_t)o * (in + 1) + in, lr[0], acc);
}

bool aligned16(const float *p, int64_t ld) {
  return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % 4 == 0;
}

// Below this much work the FFMA kernels (or the skinny / head kernels) win.
bool worth_it(int64_t M, int64_t N, int64_t K) {
  return M * N * K >= (int64_t)1 << 24 && K >= 8 && M < (1 << 30) && N < (1 << 30) &&
         K < (1 << 30);
}

// ---- convolution weight gradient as a GEMM with gathered operands ------------------------
// (nnet/convolution_layer.cc:226-270, any filter / stride / padding / channel count)
//   g[oc][c][fy][fx] = sum_{b,r,q} G[b,oc,r,q] * In0[b,c,r*s-p+fy,q*s-p+fx]
// M = filter, N = (channel, tap) in the reference's own weight order, K = (sample, output
// pixel): C is the filter-block matrix itself (row pitch = taps*C + 1), so the epilogue is the
// dense SGD update.  A = the gradient, one filter row per thread, k runs along the pixels of a
// plane and on into the next sample; B = the im2col matrix, never materialised: one (channel,
// tap) row per thread whose k walk is a walk along the image row (zero outside the image).
struct ConvWgradArgs : DenseArgs {
  const float *G, *In;
  int OCn, Cn, H, W, OH, OW, FW, taps, stride, pad, P;
};

struct ConvWgradPolicy : DensePolicy<OP_MN, OP_MN, true> {
  using Args = ConvWgradArgs;
  static constexpr int kMinBN = 128;
  struct A : MatrixOperand<OP_MN, BM, LA> {
    int oc;
    bool row_ok;
    __device__ __forceinline__ void init(const Args &g) { this->setup(nullptr, 0, g.M, g.K, BM); }
    __device__ __forceinline__ void tile(const Args &g, int t, int p) {
      oc = t * BM + p;
      row_ok = oc < g.M;
    }
    __device__ __forceinline__ void load(const Args &g, int ks, int) {
      const int k0 = ks * BK;
      int b = k0 / g.P, pix = k0 - b * g.P;
      const float *q = g.G + ((int64_t)b * g.OCn + (row_ok ? oc : 0)) * g.P + pix;
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        this->v[kk] = (row_ok && k0 + kk < g.K) ? __ldg(q) : 0.0f;
        ++q;
        if (++pix == g.P) {   // next sample: skip the other filters' planes
          pix = 0;
          q += (int64_t)(g.OCn - 1) * g.P;
        }
      }
    }
  };
  template <int BN_T>
  struct B : MatrixOperand<OP_MN, BN_T, BN_T + 2> {
    static constexpr int H = BN_T / kGroupThreads;
    int plane[H], fy[H], fx[H];
    bool n_ok[H];
    __device__ __forceinline__ void init(const Args &g) { this->setup(nullptr, 0, g.N, g.K, g.mma_n); }
    __device__ __forceinline__ void tile(const Args &g, int t, int p) {
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const int n = t * BN_T + p + h * kGroupThreads;
        n_ok[h] = n < g.N && h * kGroupThreads < this->rows_active;
        const int nn = n_ok[h] ? n : 0;
        const int c = nn / g.taps, tap = nn - c * g.taps;
        fy[h] = tap / g.FW - g.pad;
        fx[h] = tap % g.FW - g.pad;
        plane[h] = c * g.H * g.W;
      }
    }
    __device__ __forceinline__ void load(const Args &g, int ks, int) {
      const int k0 = ks * BK;
      int b = k0 / g.P;
      const int pix = k0 - b * g.P;
      int r = pix / g.OW, q = pix - r * g.OW;
      int64_t sample = (int64_t)b * g.Cn * g.H * g.W;
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const bool k_ok = k0 + kk < g.K;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const int iy = r * g.stride + fy[h], ix = q * g.stride + fx[h];
          const bool ok = k_ok && n_ok[h] && (unsigned)iy < (unsigned)g.H && (unsigned)ix < (unsigned)g.W;
          this->v[4 * (H * (kk >> 2) + h) + (kk & 3)] =
              ok ? __ldg(g.In + sample + plane[h] + iy * g.W + ix) : 0.0f;
        }
        if (++q == g.OW) {
          q = 0;
          if (++r == g.OH) {
            r = 0;
            sample += (int64_t)g.Cn * g.H * g.W;
          }
        }
      }
    }
  };
};

// bias of filter oc: sum over samples and pixels of G[b][oc][.], one block per filter,
// fixed summation order (thread-strided partial sums, then a shared-memory tree)
__global__ void conv_bias_update_kernel(const float *__restrict__ G, const float *W, float *out,
                                        const float *__restrict__ lr, int OCn, int P, int64_t batch,
                                        int64_t K1, int mode) {
  __shared__ float red[256];
  const int oc = blockIdx.x;
  float acc = 0.0f;
  for (int64_t b = 0; b < batch; ++b) {
    const float *g = G + (b * OCn + oc) * P;
    for (int i = threadIdx.x; i < P; i += blockDim.x) acc += g[i];
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) apply_update(mode, out, W, (int64_t)oc * K1 + K1 - 1, lr[0], red[0]);
}

}  // namespace

int dense_evaluate_umma(pb_context *ctx, const pb_layer_desc *l, const float *I, const float *W,
                        float *O, int64_t batch) {
  const int64_t in = l->in, out = l->out;
  if (!umma_gemm_enabled() || !worth_it(batch, out, in)) return -1;
  DenseArgs g{};
  g.A = I; g.lda = in; g.B = W; g.ldb = in + 1; g.K = (int)in;
  g.epi = EPI_BIAS; g.C = O; g.ldc = out; g.bias = W + in; g.ldbias = in + 1;
  if (aligned16(I, in)) return launch_gemm<DensePolicy<OP_KV, OP_KS>>(ctx, g, (int)batch, (int)out, (int)in);
  if (aligned16(W, in + 1)) return launch_gemm<DensePolicy<OP_KS, OP_KV>>(ctx, g, (int)batch, (int)out, (int)in);
  return launch_gemm<DensePolicy<OP_KS, OP_KS>>(ctx, g, (int)batch, (int)out, (int)in);
}

int dense_input_delta_umma(pb_context *ctx, const pb_layer_desc *l, const float *W, const float *G,
                           float *dI, int64_t batch) {
  const int64_t in = l->in, out = l->out;
  if (!umma_gemm_enabled() || !worth_it(batch, in, out)) return -1;
  DenseArgs g{};
  g.A = G; g.lda = out; g.B = W; g.ldb = in + 1; g.K = (int)out;
  g.epi = EPI_PLAIN; g.C = dI; g.ldc = in;
  if (aligned16(G, out)) return launch_gemm<DensePolicy<OP_KV, OP_MN>>(ctx, g, (int)batch, (int)in, (int)out);
  return launch_gemm<DensePolicy<OP_KS, OP_MN>>(ctx, g, (int)batch, (int)in, (int)out);
}

int dense_weight_update_umma(pb_context *ctx, const pb_layer_desc *l, const float *I,
                             const float *W, const float *G, float *out, const float *lr,
                             int mode, int64_t batch) {
  const int64_t in = l->in, nout = l->out;
  if (!umma_gemm_enabled() || !worth_it(nout, in, batch)) return -1;
  DenseArgs g{};
  g.A = G; g.lda = nout; g.B = I; g.ldb = in; g.K = (int)batch;
  g.epi = EPI_UPDATE; g.C = out; g.ldc = in + 1; g.Wold = W; g.lr = lr; g.mode = mode;
  if (launch_gemm<DensePolicy<OP_MN, OP_MN, true>>(ctx, g, (int)nout, (int)in, (int)batch)) return 1;
  dense_bias_update_kernel<<<pb_div_up(nout, 128), 128, 0, ctx->stream>>>(G, W, out, lr, (int)in,
                                                                         (int)nout, batch, mode);
  PB_LAUNCHED(ctx);
  return 0;
}

// Convolution weight gradient + SGD update for any shape on the tensor cores; -1 when the
// product is too small to be worth it (or PB_DISABLE_UMMA=1).
int conv_weight_update_gemm_umma(pb_context *ctx, const pb_layer_desc *l, const float *I,
                                 const float *Wt, const float *G, float *out, const float *lr,
                                 int mode, int64_t batch) {
  const int64_t C = l->depth, OC = l->num_filters;
  const int64_t OW = conv_out_w(l), OH = conv_out_h(l), P = OW * OH;
  const int64_t taps = l->filter_width * l->filter_height;
  const int64_t N = C * taps, K = batch * P, K1 = N + 1;
  if (!umma_gemm_enabled() || !worth_it(OC, N, K) || K >= ((int64_t)1 << 31) - BK ||
      C * l->height * l->width >= ((int64_t)1 << 31))
    return -1;
  ConvWgradArgs g{};
  g.K = (int)K;
  g.epi = EPI_UPDATE; g.C = out; g.ldc = K1; g.Wold = Wt; g.lr = lr; g.mode = mode;
  g.G = G; g.In = I;
  g.OCn = (int)OC; g.Cn = (int)C; g.H = (int)l->height; g.W = (int)l->width;
  g.OH = (int)OH; g.OW = (int)OW; g.FW = (int)l->filter_width; g.taps = (int)taps;
  g.stride = (int)l->stride; g.pad = (int)l->padding; g.P = (int)P;
  if (launch_gemm<ConvWgradPolicy>(ctx, g, (int)OC, (int)N, (int)K)) return 1;
  conv_bias_update_kernel<<<(unsigned)OC, 256, 0, ctx->stream>>>(G, Wt, out, lr, (int)OC, (int)P, batch,
                                                                 K1, mode);
  PB_LAUNCHED(ctx);
  return 0;
}

}  // namespace pb
