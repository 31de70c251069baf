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

}  // namespace pb
