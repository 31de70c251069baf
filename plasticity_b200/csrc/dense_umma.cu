// 3xTF32 tcgen05 GEMM for the batched DenseLayer products
// (nnet/dense_layer.cc:12-55) — the "real contraction" shapes of the wide MLP
// (4096-4096-4096-10, batch 1024) and of the YOLOv1 classifier.
//
//   evaluate      O[b][o]  = W[o][in] + sum_i I[b][i] * W[o][i]      M=b N=o K=i
//   input_delta   dI[b][i] = sum_o G[b][o] * W[o][i]                 M=b N=i K=o
//   weight update g[o][i]  = sum_b G[b][o] * I[b][i]                 M=o N=i K=b
//
// One kernel computes C(m,n) = sum_k A(m,k) * B(n,k) for all three; what differs
// is how each operand lies in HBM (the reference layouts are kept verbatim —
// weights are row-major [out][in+1], an ODD pitch for even widths, so TMA cannot
// fetch them):
//   KV  k contiguous, 16-byte aligned rows  -> float4 loads
//   KS  k contiguous, any alignment         -> scalar loads (lanes along k)
//   MN  m (or n) contiguous                 -> scalar loads (lanes along m)
// Producer warps load fp32 straight from global memory, split every value into
// hi = tf32(x) and lo = x - hi in registers, and store both into shared memory
// in tcgen05's K-major no-swizzle canonical layout (chunk planes of [row][4
// floats], LBO = plane pitch, SBO = 128 B): the transpose of an MN operand is
// free — it is just which lane loads what.  No pre-pass, no extra HBM traffic.
//
// Per K=8 step the single MMA-issuing thread emits three tcgen05.mma.kind::tf32
// into one fp32 TMEM accumulator:  A_hi*B_hi + A_lo*B_hi + A_hi*B_lo  (3xTF32;
// the dropped lo*lo term is ~2^-22 relative), inside the |rel| <= 1e-4 bar.
//
// CTA tile 128 x 256 (one M=128, N=256 MMA), K stage 16, 4 shared-memory
// stages, 2 TMEM accumulator stages (2 x 256 columns = all of TMEM): persistent
// CTAs, one per SM.  Warp roles (544 threads):
//   warps 0-3   epilogue  (tcgen05.ld -> bias / SGD update -> global)
//   warp  4     MMA issuer + TMEM allocator
//   warps 5-16  producers, three groups of four warps; group g fills K stages
//               g, g+3, g+6, ... so that one group's global loads are in flight
//               while another converts and stores.
// Small tile counts are split along K into a workspace and reduced in a fixed
// order (deterministic).
#include "common.cuh"
#include "kernels.cuh"
#include "umma.cuh"

namespace pb {
namespace {

using namespace umma;

constexpr int BM = 128, BN = 256, BK = 16, STAGES = 4;
constexpr int kEpiWarps = 4, kMmaWarp = 4, kGroups = 3, kGroupThreads = 128;
constexpr int kThreads = (kEpiWarps + 1) * 32 + kGroups * kGroupThreads;  // 544
// Plane pitches in 16-byte units.  = 2 (mod 8): the float4 stores of a quarter
// warp (2 rows x 4 chunks) and the scalar stores of a warp (2 rows x 16 k) hit
// distinct banks.
constexpr int LA = BM + 2, LB = BN + 2;
constexpr int CHUNKS = BK / 4;
constexpr int STAGE_UNITS = 2 * CHUNKS * (LA + LB);  // float4 per stage
constexpr size_t kSmem = 16 * (size_t)STAGES * STAGE_UNITS + 8 * (2 * STAGES + 4) + 16;

enum { OP_KV = 0, OP_KS = 1, OP_MN = 2 };
enum { EPI_PLAIN = 0, EPI_BIAS = 1, EPI_UPDATE = 2, EPI_PARTIAL = 3 };

struct GemmArgs {
  const float *A, *B;
  int64_t lda, ldb;
  int M, N, K;
  int splits, k_per_split;  // k_per_split is a multiple of BK
  int epi;
  // EPI_PLAIN / EPI_BIAS: C[m*ldc + n] = acc (+ bias[n*ldbias])
  // EPI_UPDATE: apply_update(mode, C, Wold, m*ldc + n, lr, acc)
  // EPI_PARTIAL: C[(split*M + m)*N + n] = acc (ldc == N)
  float *C;
  int64_t ldc;
  const float *bias;
  int64_t ldbias;
  const float *Wold, *lr;
  int mode;
};

// One operand tile (ROWS x BK) of one K stage: global -> registers.
template <int MODE, int ROWS>
__device__ __forceinline__ void load_tile(float (&v)[ROWS * BK / kGroupThreads], const float *base,
                                          int64_t ld, int row0, int rows_total, int k0, int k_end,
                                          int p) {
  constexpr int NV = ROWS * BK / kGroupThreads;
  if (MODE == OP_KV) {
#pragma unroll
    for (int j = 0; j < NV / 4; ++j) {
      const int e = p + j * kGroupThreads;
      const int row = row0 + (e >> 2), k = k0 + 4 * (e & 3);
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < rows_total && k < k_end)
        x = __ldg(reinterpret_cast<const float4 *>(base + (int64_t)row * ld + k));
      v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
    }
  } else if (MODE == OP_MN) {
#pragma unroll
    for (int j = 0; j < NV / 4; ++j) {
      const int e = p + j * kGroupThreads;
      const int row = row0 + (e % ROWS), k = k0 + 4 * (e / ROWS);
      const bool ok = row < rows_total;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        v[4 * j + i] = (ok && k + i < k_end) ? __ldg(base + (int64_t)(k + i) * ld + row) : 0.0f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int e = p + j * kGroupThreads;
      const int row = row0 + (e >> 4), k = k0 + (e & 15);
      v[j] = (row < rows_total && k < k_end) ? __ldg(base + (int64_t)row * ld + k) : 0.0f;
    }
  }
}

// registers -> hi / lo planes of the stage (L = plane pitch in float4 units).
template <int MODE, int ROWS, int L>
__device__ __forceinline__ void store_tile(const float (&v)[ROWS * BK / kGroupThreads], float4 *hi,
                                           float4 *lo, int p) {
  constexpr int NV = ROWS * BK / kGroupThreads;
  if (MODE == OP_KS) {
    float *h = reinterpret_cast<float *>(hi), *l = reinterpret_cast<float *>(lo);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int e = p + j * kGroupThreads;
      const int row = e >> 4, k = e & 15;
      const int idx = ((k >> 2) * L + row) * 4 + (k & 3);
      const float x = v[j], xh = tf32_hi(x);
      h[idx] = xh;
      l[idx] = x - xh;
    }
  } else {
#pragma unroll
    for (int j = 0; j < NV / 4; ++j) {
      const int e = p + j * kGroupThreads;
      const int row = MODE == OP_KV ? (e >> 2) : (e % ROWS);
      const int c = MODE == OP_KV ? (e & 3) : (e / ROWS);
      const float h0 = tf32_hi(v[4 * j]), h1 = tf32_hi(v[4 * j + 1]), h2 = tf32_hi(v[4 * j + 2]),
                  h3 = tf32_hi(v[4 * j + 3]);
      hi[c * L + row] = make_float4(h0, h1, h2, h3);
      lo[c * L + row] =
          make_float4(v[4 * j] - h0, v[4 * j + 1] - h1, v[4 * j + 2] - h2, v[4 * j + 3] - h3);
    }
  }
}

template <int AMODE, int BMODE>
__global__ void __launch_bounds__(kThreads, 1) dense_umma_kernel(const GemmArgs g) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float4 *stage_s = reinterpret_cast<float4 *>(smem_raw);  // [STAGES][STAGE_UNITS]
  uint64_t *bars = reinterpret_cast<uint64_t *>(stage_s + STAGES * STAGE_UNITS);
  uint64_t *full = bars, *empty = bars + STAGES, *d_full = bars + 2 * STAGES,
           *d_empty = bars + 2 * STAGES + 2;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], kGroupThreads); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], kEpiWarps * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_ptr)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int tiles_m = (g.M + BM - 1) / BM, tiles_n = (g.N + BN - 1) / BN;
  const int n_work = tiles_m * tiles_n * g.splits;
  // work item -> (split, tile_n, tile_m), tile_m fastest: CTAs running at the same
  // time share B tiles (the weights) through L2.
  auto decode = [&](int w, int &tm, int &tn, int &k0, int &k1) {
    tm = w % tiles_m;
    const int r = w / tiles_m;
    tn = r % tiles_n;
    const int sp = r / tiles_n;
    k0 = sp * g.k_per_split;
    k1 = min(g.K, k0 + g.k_per_split);
  };

  if (warp > kMmaWarp) {
    // ---- producers ------------------------------------------------------------
    const int pt = tid - (kEpiWarps + 1) * 32;
    const int grp = pt / kGroupThreads, p = pt % kGroupThreads;
    int it = 0;  // K-stage counter across all work items of this CTA
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      int tm, tn, k0, k1;
      decode(w, tm, tn, k0, k1);
      for (int k = k0; k < k1; k += BK, ++it) {
        if (it % kGroups != grp) continue;
        float va[BM * BK / kGroupThreads], vb[BN * BK / kGroupThreads];
        load_tile<AMODE, BM>(va, g.A, g.lda, tm * BM, g.M, k, k1, p);
        load_tile<BMODE, BN>(vb, g.B, g.ldb, tn * BN, g.N, k, k1, p);
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        float4 *a_hi = stage_s + s * STAGE_UNITS;
        float4 *a_lo = a_hi + CHUNKS * LA;
        float4 *b_hi = a_lo + CHUNKS * LA;
        float4 *b_lo = b_hi + CHUNKS * LB;
        store_tile<AMODE, BM, LA>(va, a_hi, a_lo, p);
        store_tile<BMODE, BN, LB>(vb, b_hi, b_lo, p);
        fence_async_proxy();
        mbar_arrive(&full[s]);
      }
    }
  } else if (warp == kMmaWarp) {
    // ---- MMA issuer -------------------------------------------------------------
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc(BM, BN);
    const uint32_t s_base = smem_u32(stage_s) >> 4;
    int it = 0, tc = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++tc) {
      int tm, tn, k0, k1;
      decode(w, tm, tn, k0, k1);
      const int acc = tc & 1;
      mbar_wait(&d_empty[acc], ((tc >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(acc * BN);
      for (int k = k0; k < k1; k += BK, ++it) {
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_hi = s_base + (uint32_t)(s * STAGE_UNITS);
        const uint32_t a_lo = a_hi + CHUNKS * LA;
        const uint32_t b_hi = a_lo + CHUNKS * LA;
        const uint32_t b_lo = b_hi + CHUNKS * LB;
#pragma unroll
        for (int kk = 0; kk < BK / 8; ++kk) {
          const uint64_t dah = smem_desc(a_hi + kk * 2 * LA, LA, 8);
          const uint64_t dal = smem_desc(a_lo + kk * 2 * LA, LA, 8);
          const uint64_t dbh = smem_desc(b_hi + kk * 2 * LB, LB, 8);
          const uint64_t dbl = smem_desc(b_lo + kk * 2 * LB, LB, 8);
          if (leader) {
            umma_tf32(d, dal, dbh, idesc, (k > k0 || kk > 0) ? 1u : 0u);
            umma_tf32(d, dah, dbl, idesc, 1u);
            umma_tf32(d, dah, dbh, idesc, 1u);
          }
        }
        if (leader) umma_commit(&empty[s]);
        __syncwarp();
      }
      if (leader) umma_commit(&d_full[acc]);
      __syncwarp();
    }
  } else {
    // ---- epilogue ---------------------------------------------------------------
    int tc = 0;
    const float rate = g.epi == EPI_UPDATE ? g.lr[0] : 0.0f;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++tc) {
      int tm, tn, k0, k1;
      decode(w, tm, tn, k0, k1);
      const int acc = tc & 1;
      mbar_wait(&d_full[acc], (tc >> 1) & 1);
      tc_fence_after();
      const int m = tm * BM + warp * 32 + lane;
      const bool m_ok = m < g.M;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * BN);
      const int n_lim = min(BN, g.N - tn * BN);
      const int sp = k0 / g.k_per_split;
      float *crow = g.epi == EPI_PARTIAL ? g.C + ((int64_t)sp * g.M + m) * g.N
                                         : g.C + (int64_t)m * g.ldc;
      const bool vec = (g.epi == EPI_PLAIN || g.epi == EPI_PARTIAL) &&
                       ((g.epi == EPI_PARTIAL ? g.N : g.ldc) % 4 == 0) &&
                       ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
#pragma unroll 1
      for (int c0 = 0; c0 < n_lim; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
        if (!m_ok) continue;
        const int n0 = tn * BN + c0;
        if (vec && c0 + 16 <= n_lim) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4 *>(crow + n0 + 4 * q) =
                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = n0 + j;
            if (c0 + j < n_lim) {
              if (g.epi == EPI_BIAS) crow[n] = __ldg(g.bias + (int64_t)n * g.ldbias) + v[j];
              else if (g.epi == EPI_UPDATE)
                apply_update(g.mode, g.C, g.Wold, (int64_t)m * g.ldc + n, rate, v[j]);
              else crow[n] = v[j];
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&d_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u)
                 : "memory");
  }
}

// Fixed-order sum of the split-K partials + the real epilogue.
__global__ void splitk_reduce_kernel(const float *__restrict__ part, int splits, const GemmArgs g) {
  const int64_t total = (int64_t)g.M * g.N;
  const float rate = g.epi == EPI_UPDATE ? g.lr[0] : 0.0f;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.0f;
    for (int s = 0; s < splits; ++s) acc += part[(int64_t)s * total + t];
    const int64_t m = t / g.N;
    const int n = (int)(t - m * g.N);
    if (g.epi == EPI_BIAS) g.C[m * g.ldc + n] = __ldg(g.bias + (int64_t)n * g.ldbias) + acc;
    else if (g.epi == EPI_UPDATE) apply_update(g.mode, g.C, g.Wold, m * g.ldc + n, rate, acc);
    else g.C[m * g.ldc + n] = acc;
  }
}

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

template <int AMODE, int BMODE>
int launch(pb_context *ctx, GemmArgs g) {
  auto kern = dense_umma_kernel<AMODE, BMODE>;
  static bool configured = false;
  if (!configured) {
    PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    configured = true;
  }
  const int tiles = pb_div_up(g.M, BM) * pb_div_up(g.N, BN);
  const int k_stages = pb_div_up(g.K, BK);
  // Split K when the tiles alone leave most SMs idle; every split keeps at least
  // 16 K stages so the pipeline fill stays amortised.
  int splits = 1;
  if (tiles * 2 <= ctx->num_sms) splits = std::max(1, std::min(ctx->num_sms / tiles, k_stages / 16));
  const int per = pb_div_up(k_stages, splits) * BK;
  splits = pb_div_up(g.K, per);
  g.splits = splits;
  g.k_per_split = per;
  GemmArgs final_args = g;
  if (splits > 1) {
    if (ensure_workspace(ctx, (size_t)splits * g.M * g.N)) return 1;
    g.epi = EPI_PARTIAL;
    g.C = ctx->workspace;
    g.ldc = g.N;
  }
  const int grid = std::min(tiles * splits, ctx->num_sms);
  kern<<<grid, kThreads, kSmem, ctx->stream>>>(g);
  PB_LAUNCHED(ctx);
  if (splits > 1) {
    splitk_reduce_kernel<<<pb_ew_grid(ctx, (int64_t)g.M * g.N, 256), 256, 0, ctx->stream>>>(
        ctx->workspace, splits, final_args);
    PB_LAUNCHED(ctx);
  }
  return 0;
}

bool dense_umma_enabled() {
  static const bool on = [] {
    const char *e = getenv("PB_DISABLE_UMMA");
    return !(e && e[0] == '1');
  }();
  return on;
}

// Below this much work the FFMA kernels (or the skinny / head kernels) win.
bool worth_it(int64_t M, int64_t N, int64_t K) { return M * N * K >= (int64_t)1 << 24 && K >= 8; }

}  // namespace

int dense_evaluate_umma(pb_context *ctx, const pb_layer_desc *l, const float *I, const float *W,
                        float *O, int64_t batch) {
  const int64_t in = l->in, out = l->out;
  if (!dense_umma_enabled() || !worth_it(batch, out, in) || batch > (1 << 30) / 2) return -1;
  GemmArgs g{};
  g.A = I; g.lda = in; g.B = W; g.ldb = in + 1;
  g.M = (int)batch; g.N = (int)out; g.K = (int)in;
  g.epi = EPI_BIAS; g.C = O; g.ldc = out; g.bias = W + in; g.ldbias = in + 1;
  const bool av = aligned16(I, in), bv = aligned16(W, in + 1);
  if (av && bv) return launch<OP_KV, OP_KV>(ctx, g);
  if (av) return launch<OP_KV, OP_KS>(ctx, g);
  return launch<OP_KS, OP_KS>(ctx, g);
}

int dense_input_delta_umma(pb_context *ctx, const pb_layer_desc *l, const float *W, const float *G,
                           float *dI, int64_t batch) {
  const int64_t in = l->in, out = l->out;
  if (!dense_umma_enabled() || !worth_it(batch, in, out) || batch > (1 << 30) / 2) return -1;
  GemmArgs g{};
  g.A = G; g.lda = out; g.B = W; g.ldb = in + 1;
  g.M = (int)batch; g.N = (int)in; g.K = (int)out;
  g.epi = EPI_PLAIN; g.C = dI; g.ldc = in;
  if (aligned16(G, out)) return launch<OP_KV, OP_MN>(ctx, g);
  return launch<OP_KS, OP_MN>(ctx, g);
}

int dense_weight_update_umma(pb_context *ctx, const pb_layer_desc *l, const float *I,
                             const float *W, const float *G, float *out, const float *lr,
                             int mode, int64_t batch) {
  const int64_t in = l->in, nout = l->out;
  if (!dense_umma_enabled() || !worth_it(nout, in, batch) || batch > (1 << 30) / 2) return -1;
  GemmArgs g{};
  g.A = G; g.lda = nout; g.B = I; g.ldb = in;
  g.M = (int)nout; g.N = (int)in; g.K = (int)batch;
  g.epi = EPI_UPDATE; g.C = out; g.ldc = in + 1; g.Wold = W; g.lr = lr; g.mode = mode;
  if (launch<OP_MN, OP_MN>(ctx, g)) return 1;
  dense_bias_update_kernel<<<pb_div_up(nout, 128), 128, 0, ctx->stream>>>(G, W, out, lr, (int)in,
                                                                         (int)nout, batch, mode);
  PB_LAUNCHED(ctx);
  return 0;
}

}  // namespace pb
