// Persistent warp-specialised 3xTF32 tcgen05 GEMM skeleton shared by the dense
// products (dense_umma.cu) and the general implicit-GEMM convolution
// (conv_gemm_umma.cu):  C(m,n) = sum_k A(m,k) * B(n,k)  with fp32 operands that
// lie in HBM in the reference's own layouts.
//
// Producer warps load fp32 straight from global memory, split every value into
// hi = x with the 13 low mantissa bits cleared (a TF32 value) and lo = x - hi in
// registers, and store both into shared memory in tcgen05's K-major no-swizzle
// canonical layout (chunk planes of [row][4 floats], LBO = plane pitch, SBO = 128
// B): transposed operands and im2col gathers are free — they are just which lane
// loads what.  No pre-pass over the activations, no extra HBM traffic.
//
// Per K=8 step the single MMA-issuing thread emits three tcgen05.mma.kind::tf32
// into one fp32 TMEM accumulator:  A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  (3xTF32; the
// dropped lo*lo term is ~2^-22 relative), inside the |rel| <= 1e-4 bar.
//
// CTA tile 128 x (up to) 256 — one M=128, N=mma_n MMA — K stage 16, 4 shared-memory
// stages, 2 TMEM accumulator stages (2 x 256 columns = all of TMEM): persistent
// CTAs, one per SM.  Warp roles (544 threads):
//   warps 0-3   epilogue  (tcgen05.ld -> policy store -> global)
//   warp  4     MMA issuer + TMEM allocator
//   warps 5-16  producers, three groups of four warps; group g fills K stages
//               g, g+3, g+6, ... so that one group's global loads are in flight
//               while another converts and stores.
// Small tile counts are split along K into a workspace and reduced in a fixed
// order (deterministic).
//
// A policy P provides:
//   struct Args : GemmCore {...};
//   struct A { void init(const Args&); void tile(const Args&, int tile_m, int p);
//              void load(const Args&, int k_stage, int p); void store(float4 *hi, int p) const; };
//   template <int BN_T> struct B { same, with tile_n; BN_T rows per tile };
//                                              (p = thread index in the 128-thread group)
//   static constexpr int kEpiCols (16 or 32); static constexpr bool kEpiTranspose;
//   static void store_cols(const Args&, int row0, int lane, int n0, const float (&v)[kEpiCols],
//                          int n_valid, float *stage);   (v = columns n0.. of row row0+lane;
//                          stage = this warp's 32x33 transpose tile if kEpiTranspose)
//   static void store1(const Args&, int64_t m, int n, float acc);     (split-K reduce)
#pragma once
#include "common.cuh"
#include "kernels.cuh"
#include "umma.cuh"

namespace pb {
namespace ugemm {

using namespace umma;

constexpr int BM = 128, BN = 256, BK = 16, STAGES = 4;
#ifndef PB_GEMM_GROUPS
#define PB_GEMM_GROUPS 3
#endif
constexpr int kEpiWarps = 4, kMmaWarp = 4, kGroups = PB_GEMM_GROUPS, kGroupThreads = 128;
constexpr int kThreads = (kEpiWarps + 1) * 32 + kGroups * kGroupThreads;  // 544
// Plane pitches in 16-byte units.  = 2 (mod 8): the float4 stores of a quarter
// warp (2 rows x 4 chunks) and the scalar stores of a warp (2 rows x 16 k) hit
// distinct banks.
constexpr int LA = BM + 2;
constexpr int CHUNKS = BK / 4;
// The N extent of the CTA tile is a template parameter (256 / 128 / 64): problems with few
// columns (a convolution with 64 filters) stage fewer B rows per K step and get more
// shared-memory stages for the same budget, i.e. more load latency hidden per MMA cycle.
template <int BN_T>
struct Geo {
  static constexpr int LB = BN_T + 2;
  static constexpr int STAGE_UNITS = 2 * CHUNKS * (LA + LB);  // float4 per stage
  static constexpr int STAGES = BN_T == 256 ? 4 : (BN_T == 128 ? 5 : 7);
  // + 1 KiB the system reserves per CTA <= 196 KiB: stays under that shared-memory
  // carve-out step, which leaves 60 KiB of L1 for the operand gathers (the next step,
  // 228 KiB, leaves 28 and costs 10-40 % on the gather-heavy layers).
  static constexpr size_t kSmem = 16 * (size_t)STAGES * STAGE_UNITS + 8 * (2 * STAGES + 4) + 16;
  static constexpr uint32_t TMEM_COLS = 2 * BN_T;  // two accumulator stages
  static_assert(kSmem + 1024 <= 196 * 1024, "shared-memory budget");
};
constexpr int LB = Geo<BN>::LB;  // plane pitch of the full-width tile (policies' default)
// Optional (policy P::kEpiTranspose): per epilogue warp a 32 x 32 transpose tile, pitch 33
// (conflict-free both ways), for policies whose C rows must be written by whole warps.
constexpr int kEpiStage = 32 * 33;
template <class P, int BN_T>
constexpr size_t smem_bytes() {
  return Geo<BN_T>::kSmem + (P::kEpiTranspose ? 4 * (size_t)kEpiWarps * kEpiStage : 0);
}

struct GemmCore {
  int M, N;
  int k_stages;          // ceil(K / BK)
  int splits, stages_per_split;
  int mma_n;             // MMA N: multiple of 16, 16..256 (>= min(N, 256))
  float *partial;        // split-K workspace [splits][M][N]
};

// hi = x with the 13 low mantissa bits cleared (a valid TF32 value: what the tensor
// core reads from x anyway), lo = x - hi (exact in fp32).
__device__ __forceinline__ float split_hi(float x) {
  return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
}

// 16 values of one row (k = 0..15 of the stage, four float4 groups) -> hi / lo planes.
template <int L>
__device__ __forceinline__ void store_row16(const float (&v)[16], float4 *hi, int row) {
  float4 *lo = hi + CHUNKS * L;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float h0 = split_hi(v[4 * c]), h1 = split_hi(v[4 * c + 1]), h2 = split_hi(v[4 * c + 2]),
                h3 = split_hi(v[4 * c + 3]);
    hi[c * L + row] = make_float4(h0, h1, h2, h3);
    lo[c * L + row] =
        make_float4(v[4 * c] - h0, v[4 * c + 1] - h1, v[4 * c + 2] - h2, v[4 * c + 3] - h3);
  }
}

enum { OP_KV = 0, OP_KS = 1, OP_MN = 2 };

// One plain matrix operand tile (ROWS x BK) of one K stage: global -> registers ->
// hi / lo planes.  How the operand lies in HBM:
//   KV  k contiguous, 16-byte aligned rows  -> float4 loads
//   KS  k contiguous, any alignment         -> scalar loads (lanes along k)
//   MN  m (or n) contiguous                 -> scalar loads (lanes along m)
// Thread p of a 128-thread producer group owns ROWS*BK/128 values:
//   KV  row (p>>2) + 32j, chunk p&3            (one float4 load each)
//   MN  row p (+128), k = 0..15                (pointer walks k)
//   KS  row (p>>4) + 8j, k = p&15              (pointer walks the rows)
// The interior of the matrix takes an unpredicated pointer-walking path; tiles that
// touch an edge (row or K tail) take the bounds-checked one.
template <int MODE, int ROWS, int L>
struct MatrixOperand {
  static constexpr int NV = ROWS * BK / kGroupThreads;
  const float *base;
  int64_t ld;
  int rows_total, k_total, row0;
  int rows_active;  // rows of the tile the MMA reads (mma_n for B): the rest is not touched
  float v[NV];

  __device__ __forceinline__ void setup(const float *b, int64_t l, int rows, int k, int active) {
    base = b; ld = l; rows_total = rows; k_total = k; rows_active = active;
  }
  __device__ __forceinline__ void set_tile(int tile) { row0 = tile * ROWS; }

  __device__ __forceinline__ void load_stage(int ks, int p) {
    const int k0 = ks * BK, k_end = k_total;
    const bool interior =
        row0 + ((rows_active + 31) & ~31) <= rows_total && k0 + BK <= k_end;
    if (MODE == OP_KV) {
      const int r = row0 + (p >> 2), k = k0 + 4 * (p & 3);
      const float *q = base + (int64_t)r * ld + k;
      const int64_t step = 32 * ld;
      if (interior) {
#pragma unroll
        for (int j = 0; j < NV / 4; ++j, q += step) {
          if (32 * j < rows_active) {
            const float4 x = __ldg(reinterpret_cast<const float4 *>(q));
            v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < NV / 4; ++j, q += step) {
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
          if (32 * j < rows_active && r + 32 * j < rows_total && k < k_end)
            x = __ldg(reinterpret_cast<const float4 *>(q));
          v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
        }
      }
    } else if (MODE == OP_MN) {
      constexpr int H = ROWS >= kGroupThreads ? ROWS / kGroupThreads : 1;  // row halves per thread
      const int r = row0 + p;
      const float *q = base + (int64_t)k0 * ld + r;
      if (interior) {
#pragma unroll
        for (int kk = 0; kk < BK; ++kk, q += ld)
#pragma unroll
          for (int h = 0; h < H; ++h)
            if (h * kGroupThreads < rows_active)
              v[4 * (H * (kk >> 2) + h) + (kk & 3)] = __ldg(q + h * kGroupThreads);
      } else {
#pragma unroll
        for (int kk = 0; kk < BK; ++kk, q += ld)
#pragma unroll
          for (int h = 0; h < H; ++h)
            v[4 * (H * (kk >> 2) + h) + (kk & 3)] =
                (h * kGroupThreads < rows_active && r + h * kGroupThreads < rows_total &&
                 k0 + kk < k_end)
                    ? __ldg(q + h * kGroupThreads) : 0.0f;
      }
    } else {
      const int r = row0 + (p >> 4), k = k0 + (p & 15);
      const float *q = base + (int64_t)r * ld + k;
      const int64_t step = 8 * ld;
      if (interior) {
#pragma unroll
        for (int j = 0; j < NV; ++j, q += step)
          if (8 * j < rows_active) v[j] = __ldg(q);
      } else {
#pragma unroll
        for (int j = 0; j < NV; ++j, q += step)
          v[j] = (8 * j < rows_active && r + 8 * j < rows_total && k < k_end) ? __ldg(q) : 0.0f;
      }
    }
  }

  __device__ __forceinline__ void store(float4 *hi, int p) const {
    float4 *lo = hi + CHUNKS * L;
    if (MODE == OP_KS) {
      float *h = reinterpret_cast<float *>(hi) + (((p & 15) >> 2) * L + (p >> 4)) * 4 + (p & 3);
      float *l = h + CHUNKS * L * 4;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if (8 * j < rows_active) {
          const float xh = split_hi(v[j]);
          h[32 * j] = xh;
          l[32 * j] = v[j] - xh;
        }
      }
    } else {
      // KV: value group j = (row (p>>2) + 32j, chunk p&3); MN: group j = (chunk j/H, row p + 128(j%H))
      constexpr int H = ROWS >= kGroupThreads ? ROWS / kGroupThreads : 1;
      hi += MODE == OP_KV ? (p & 3) * L + (p >> 2) : p;
      lo += MODE == OP_KV ? (p & 3) * L + (p >> 2) : p;
#pragma unroll
      for (int j = 0; j < NV / 4; ++j) {
        const int off = MODE == OP_KV ? 32 * j : (j / H) * L + (j % H) * kGroupThreads;
        if ((MODE == OP_KV ? 32 * j : (j % H) * kGroupThreads) >= rows_active) continue;
        const float h0 = split_hi(v[4 * j]), h1 = split_hi(v[4 * j + 1]),
                    h2 = split_hi(v[4 * j + 2]), h3 = split_hi(v[4 * j + 3]);
        hi[off] = make_float4(h0, h1, h2, h3);
        lo[off] = make_float4(v[4 * j] - h0, v[4 * j + 1] - h1, v[4 * j + 2] - h2, v[4 * j + 3] - h3);
      }
    }
  }
};

#ifdef PB_GEMM_MAXNREG
#define PB_GEMM_BOUNDS __maxnreg__(PB_GEMM_MAXNREG)
#else
#define PB_GEMM_BOUNDS __launch_bounds__(kThreads, 1)
#endif
template <class P, int BN_T>
__global__ void PB_GEMM_BOUNDS umma_gemm_kernel(const __grid_constant__ typename P::Args g) {
  using G_ = Geo<BN_T>;
  constexpr int STAGES = G_::STAGES, STAGE_UNITS = G_::STAGE_UNITS, LB = G_::LB, BN = BN_T;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float4 *stage_s = reinterpret_cast<float4 *>(smem_raw);  // [STAGES][STAGE_UNITS]
  uint64_t *bars = reinterpret_cast<uint64_t *>(stage_s + STAGES * STAGE_UNITS);
  uint64_t *full = bars, *empty = bars + STAGES, *d_full = bars + 2 * STAGES,
           *d_empty = bars + 2 * STAGES + 2;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 4);
  float *epi_stage = reinterpret_cast<float *>(tmem_ptr + 4);  // [kEpiWarps][kEpiStage]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], kGroupThreads / 32); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], kEpiWarps * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_ptr)),
                 "r"(G_::TMEM_COLS)
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
  auto decode = [&](int w, int &tm, int &tn, int &sp, int &ks0, int &ks1) {
    tm = w % tiles_m;
    const int r = w / tiles_m;
    tn = r % tiles_n;
    sp = r / tiles_n;
    ks0 = sp * g.stages_per_split;
    ks1 = min(g.k_stages, ks0 + g.stages_per_split);
  };

  if (warp > kMmaWarp) {
    // ---- producers ------------------------------------------------------------
    const int pt = tid - (kEpiWarps + 1) * 32;
    const int grp = pt / kGroupThreads, p = pt % kGroupThreads;
    typename P::A opa;
    typename P::template B<BN_T> opb;
    opa.init(g);
    opb.init(g);
    int it = 0;  // K-stage counter across all work items of this CTA
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      int tm, tn, sp, ks0, ks1;
      decode(w, tm, tn, sp, ks0, ks1);
      opa.tile(g, tm, p);
      opb.tile(g, tn, p);
      for (int ks = ks0; ks < ks1; ++ks, ++it) {
        if (it % kGroups != grp) continue;
        opa.load(g, ks, p);
        opb.load(g, ks, p);
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        float4 *a_hi = stage_s + s * STAGE_UNITS;
        opa.store(a_hi, p);
        opb.store(a_hi + 2 * CHUNKS * LA, p);
        // every thread publishes its own stores to the async proxy; one lane per warp
        // arrives (128 same-address arrivals per stage would serialise in shared memory)
        fence_async_proxy();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
      }
    }
  } else if (warp == kMmaWarp) {
    // ---- MMA issuer -------------------------------------------------------------
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc(BM, g.mma_n);
    const uint32_t s_base = smem_u32(stage_s) >> 4;
    int it = 0, tc = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++tc) {
      int tm, tn, sp, ks0, ks1;
      decode(w, tm, tn, sp, ks0, ks1);
      const int acc = tc & 1;
      mbar_wait(&d_empty[acc], ((tc >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(acc * BN);
      for (int ks = ks0; ks < ks1; ++ks, ++it) {
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
            umma_tf32(d, dal, dbh, idesc, (ks > ks0 || kk > 0) ? 1u : 0u);
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
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++tc) {
      int tm, tn, sp, ks0, ks1;
      decode(w, tm, tn, sp, ks0, ks1);
      const int acc = tc & 1;
      mbar_wait(&d_full[acc], (tc >> 1) & 1);
      tc_fence_after();
      const int row0 = tm * BM + warp * 32;
      const int m = row0 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * BN);
      const int n_lim = min(BN, g.N - tn * BN);
      constexpr int EC = P::kEpiCols;  // accumulator columns per epilogue step: 16 or 32
#pragma unroll 1
      for (int c0 = 0; c0 < n_lim; c0 += EC) {
        float v[EC];
#pragma unroll
        for (int q = 0; q < EC / 16; ++q) tmem_ld16(taddr + c0 + 16 * q, v + 16 * q);
        tmem_ld_wait();
        const int n0 = tn * BN + c0, nv = min(EC, n_lim - c0);
        if (g.splits > 1) {
          if (m < g.M) {
            float *dst = g.partial + ((int64_t)sp * g.M + m) * g.N + n0;
#pragma unroll
            for (int j = 0; j < EC; ++j)
              if (j < nv) dst[j] = v[j];
          }
        } else {
          P::store_cols(g, row0, lane, n0, v, nv, epi_stage + warp * kEpiStage);
        }
      }
      tc_fence_before();
      mbar_arrive(&d_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(G_::TMEM_COLS)
                 : "memory");
  }
}

// Fixed-order sum of the split-K partials + the policy's store.
template <class P>
__global__ void splitk_reduce_kernel(const typename P::Args g) {
  const int64_t total = (int64_t)g.M * g.N;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.0f;
    for (int s = 0; s < g.splits; ++s) acc += g.partial[(int64_t)s * total + t];
    const int64_t m = t / g.N;
    P::store1(g, m, (int)(t - m * g.N), acc);
  }
}

// Fills the GemmCore fields (tiling, split-K) and launches.  K = number of k values.
template <class P, int BN_T>
int launch_gemm_bn(pb_context *ctx, typename P::Args g, int M, int N, int K) {
  auto kern = umma_gemm_kernel<P, BN_T>;
  static bool configured = false;
  if (!configured) {
    PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem_bytes<P, BN_T>()));
    configured = true;
  }
  g.M = M;
  g.N = N;
  g.k_stages = pb_div_up(K, BK);
  g.mma_n = std::min(BN_T, pb_div_up(N, 16) * 16);
  const int tiles = pb_div_up(M, BM) * pb_div_up(N, BN_T);
  // Split K when the tiles alone leave most SMs idle; every split keeps at least
  // 16 K stages so the pipeline fill stays amortised.
  int splits = 1;
  if (tiles * 2 <= ctx->num_sms)
    splits = std::max(1, std::min(ctx->num_sms / tiles, g.k_stages / 16));
  g.stages_per_split = pb_div_up(g.k_stages, splits);
  g.splits = pb_div_up(g.k_stages, g.stages_per_split);
  g.partial = nullptr;
  if (g.splits > 1) {
    if (ensure_workspace(ctx, (size_t)g.splits * M * N)) return 1;
    g.partial = ctx->workspace;
  }
  const int grid = std::min(tiles * g.splits, ctx->num_sms);
  kern<<<grid, kThreads, smem_bytes<P, BN_T>(), ctx->stream>>>(g);
  PB_LAUNCHED(ctx);
  if (g.splits > 1) {
    splitk_reduce_kernel<P><<<pb_ew_grid(ctx, (int64_t)M * N, 256), 256, 0, ctx->stream>>>(g);
    PB_LAUNCHED(ctx);
  }
  return 0;
}

// Tile width by problem width.  P::kMinBN: the narrowest tile the policy's B loader
// supports (its row mapping needs at least that many rows).
template <class P>
int launch_gemm(pb_context *ctx, const typename P::Args &g, int M, int N, int K) {
  if (N <= 64 && P::kMinBN <= 64) return launch_gemm_bn<P, 64>(ctx, g, M, N, K);
  if (N <= 128 && P::kMinBN <= 128) return launch_gemm_bn<P, 128>(ctx, g, M, N, K);
  if (P::kMinBN <= 128) {
    // Tile quantisation: a problem whose 128 x 256 tiles fill the last wave badly (sharded
    // inference at a small per-GPU batch: 196 tiles on 148 SMs) runs on 128 x 128 tiles when
    // that fills the SMs better by more than what the narrower tile costs (it re-stages A
    // twice as often; ~0.88 of the wide tile's rate).
    auto fill = [&](int bn) {
      const int64_t tiles = (int64_t)pb_div_up(M, BM) * pb_div_up(N, bn);
      if (tiles * 2 <= ctx->num_sms) return 0.0;  // split-K territory: leave it to the wide tile
      const int64_t waves = (tiles + ctx->num_sms - 1) / ctx->num_sms;
      return (double)tiles / (double)(waves * ctx->num_sms);
    };
    const double wide = fill(256), narrow = fill(128) * 0.88;
    if (wide > 0.0 && narrow > wide * 1.05) return launch_gemm_bn<P, 128>(ctx, g, M, N, K);
  }
  return launch_gemm_bn<P, 256>(ctx, g, M, N, K);
}

inline bool umma_gemm_enabled() {
  static const bool on = [] {
    const char *e = getenv("PB_DISABLE_UMMA");
    return !(e && e[0] == '1');
  }();
  return on;
}

}  // namespace ugemm
}  // namespace pb
