// 3xTF32 tensor-core weight gradient for the batched stride-1 "same" 5x5 convolutions
// of the CIFAR network whose channel counts fill an MMA tile (layers 4 and 7 of
// nnet/cifar_test.cc:293-350; semantics nnet/convolution_layer.cc:226-270):
//
//   g[oc][ic][fy][fx] = sum_{b,y,x} G[b,oc,y,x] * In0[b,ic,y+fy-2,x+fx-2],   bias: sum G
//
// The contraction runs over (sample, pixel) and the free indices are tiny — 16..20
// input channels x 20 filters per tap — far below tcgen05's 64/128-row tiles (DESIGN.md
// section 3.2), but exactly the granularity of the warp-level mma.sync.m16n8k8: per tap,
//   D[ic 16][oc 8] += A[ic 16][pixel 8] * B[pixel 8][oc 8]
// with A read from the zero-haloed input tile at the tap's shift and B from the gradient
// tile.  Tiles are raw fp32 in shared memory (cp.async, double-buffered per sample); the
// hi/lo split (hi = x & 0xFFFFE000, lo = x - hi) happens in registers at fragment-load
// time and every product is three MMAs (lo*hi + hi*lo + hi*hi), fp32 accumulate.
//
// A CTA (20 warps, five per scheduler) walks its share of the batch.  Warp = (filter row fy, 16-channel M
// tile) x K slice (a band of output rows): it keeps the accumulators of its 5 fx taps x
// all filter tiles in registers for the whole walk, loads the B fragments of a K step
// once and reuses them for the 5 taps.  Plane pitches are = 4 (mod 8) words, which makes
// every fragment load bank-conflict-free.  Each (CTA, K slice) writes one partial in the
// reference weight layout; the fixed-order reduce + SGD update kernel of conv_tiled.cu
// finishes (deterministic).
#include <cuda_pipeline.h>

#include "common.cuh"
#include "kernels.cuh"
#include "umma.cuh"

namespace pb {
namespace {

constexpr int kWarps = 20, kThreads = kWarps * 32;

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// arrives on `bar` once all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(umma::smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void split(float x, uint32_t &hi, uint32_t &lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

template <int IC_, int OC_, int S_>
struct WCfg {
  static constexpr int IC = IC_, OC = OC_, S = S_, F = 5, PADW = 2;
  static constexpr int MT = (IC + 15) / 16;        // 16-channel M tiles
  static constexpr int NT = (OC + 7) / 8;          // 8-filter N tiles
  static constexpr int UNITS = F * MT;             // (fy, m tile)
  static constexpr int SLICES = kWarps / UNITS;    // K slices (bands of output rows)
  static constexpr int RPS = S / SLICES;           // rows per slice
  // input tile: rows y-2..y+S+1, row pitch S+8 with the image at column 4 (16-byte aligned
  // rows for cp.async; the left halo is columns 2,3, the right halo S+4, S+5)
  static constexpr int LEFT = 4;
  static constexpr int PITCH = S + 8;
  static constexpr int ROWS = S + 2 * PADW;
  static constexpr int PLANE = ((ROWS * PITCH + 7) / 8) * 8 + 4;   // = 4 (mod 8)
  static constexpr int GPLANE = S * S + 4;                          // = 4 (mod 8)
  static constexpr int IN_ELEMS = MT * 16 * PLANE, G_ELEMS = NT * 8 * GPLANE;
  static constexpr int STAGE = IN_ELEMS + G_ELEMS;
  static constexpr int MAIN_ELEMS = 2 * STAGE;
  static constexpr int K1 = F * F * IC + 1;        // reference filter-block stride
  // one partial per CTA, in FRAGMENT order so that every store of the kernel is
  // one contiguous 128-byte line: [unit][fx][nt][i][lane], then the bias sums [nt][g]
  static constexpr int FRAG_ELEMS = UNITS * F * NT * 4 * 32;
  static constexpr int ROW = FRAG_ELEMS + NT * 8;
  // the K slices of a CTA are summed through shared memory (fixed tree order) first
  static constexpr int RED_ELEMS = (SLICES / 2) * ROW;
  static constexpr size_t SMEM =
      sizeof(float) * (MAIN_ELEMS > RED_ELEMS ? MAIN_ELEMS : RED_ELEMS);
  static_assert(UNITS * SLICES == kWarps && RPS * SLICES == S, "warp layout");
  static_assert(S % 8 == 0 && (S * S) % 8 == 0, "K steps of 8 pixels");
  static_assert(SMEM <= 220 * 1024, "shared memory");
};

template <class C>
__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_mma_kernel(const float *__restrict__ In, const float *__restrict__ G,
                      float *__restrict__ partial, int batch, int samples_per_cta) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int unit = warp % C::UNITS, slice = warp / C::UNITS;
  const int fy = unit / C::MT, mt = unit % C::MT;

  // zero once: halos, the channel planes past IC and the filter planes past OC stay zero
  for (int i = tid; i < C::MAIN_ELEMS; i += kThreads) smem[i] = 0.0f;
  __syncthreads();
  chain_wait();  // chained launch (common.cuh): the zero fill above ran ahead of it
  chain_release();

  const int b0 = blockIdx.x * samples_per_cta;
  const int b1 = min(batch, b0 + samples_per_cta);

  auto load = [&](int st, int b) {
    float *in_s = smem + st * C::STAGE, *g_s = in_s + C::IN_ELEMS;
    const float *src = In + (int64_t)b * C::IC * C::S * C::S;
    // 16-byte copies only: 4-byte cp.async (one per pixel) saturates the LSU queue and
    // starves the warps that should be issuing MMAs
    for (int e = tid; e < C::IC * C::S * C::S / 4; e += kThreads) {
      const int x4 = e % (C::S / 4), r = e / (C::S / 4);
      const int y = r % C::S, ic = r / C::S;
      __pipeline_memcpy_async(in_s + ic * C::PLANE + (y + C::PADW) * C::PITCH + C::LEFT + 4 * x4,
                              src + 4 * e, 16);
    }
    const float *gs = G + (int64_t)b * C::OC * C::S * C::S;
    for (int e = tid; e < C::OC * C::S * C::S / 4; e += kThreads) {
      const int q = e % (C::S * C::S / 4), oc = e / (C::S * C::S / 4);
      __pipeline_memcpy_async(g_s + oc * C::GPLANE + 4 * q, gs + 4 * e, 16);
    }
  };

  float acc[C::F][C::NT][4];
#pragma unroll
  for (int fx = 0; fx < C::F; ++fx)
#pragma unroll
    for (int nt = 0; nt < C::NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[fx][nt][i] = 0.0f;
  float bias_acc[C::NT];
#pragma unroll
  for (int nt = 0; nt < C::NT; ++nt) bias_acc[nt] = 0.0f;

  if (b0 < b1) load(0, b0);
  __pipeline_commit();
  for (int b = b0; b < b1; ++b) {
    const int st = (b - b0) & 1;
    if (b + 1 < b1) load(st ^ 1, b + 1);
    __pipeline_commit();
    __pipeline_wait_prior(1);
    __syncthreads();
    const float *in_s = smem + st * C::STAGE, *g_s = in_s + C::IN_ELEMS;
    const float *a_base = in_s + (mt * 16 + g) * C::PLANE + t;
    const float *b_base = g_s + g * C::GPLANE + t;
#pragma unroll 1
    for (int yy = 0; yy < C::RPS; ++yy) {
      const int y = slice * C::RPS + yy;
#pragma unroll 1
      for (int x0 = 0; x0 < C::S; x0 += 8) {
        uint32_t bh[C::NT][2], bl[C::NT][2];
#pragma unroll
        for (int nt = 0; nt < C::NT; ++nt) {
          const float *p = b_base + nt * 8 * C::GPLANE + y * C::S + x0;
          const float v0 = p[0], v1 = p[4];
          split(v0, bh[nt][0], bl[nt][0]);
          split(v1, bh[nt][1], bl[nt][1]);
          bias_acc[nt] += v0 + v1;
        }
        const float *ap = a_base + (y + fy) * C::PITCH + x0 + (C::LEFT - C::PADW);
#pragma unroll
        for (int fx = 0; fx < C::F; ++fx) {
          uint32_t ah[4], al[4];
          split(ap[fx], ah[0], al[0]);
          split(ap[fx + 8 * C::PLANE], ah[1], al[1]);
          split(ap[fx + 4], ah[2], al[2]);
          split(ap[fx + 8 * C::PLANE + 4], ah[3], al[3]);
          // pass-major: consecutive MMAs on one accumulator are NT instructions apart
#pragma unroll
          for (int nt = 0; nt < C::NT; ++nt) mma_tf32(acc[fx][nt], al, bh[nt]);
#pragma unroll
          for (int nt = 0; nt < C::NT; ++nt) mma_tf32(acc[fx][nt], ah, bl[nt]);
#pragma unroll
          for (int nt = 0; nt < C::NT; ++nt) mma_tf32(acc[fx][nt], ah, bh[nt]);
        }
      }
    }
    __syncthreads();  // everyone is done with stage st before it is refilled
  }
  __pipeline_wait_prior(0);

  // bias: lanes t = 0..3 of a group g hold disjoint pixel subsets of filter 8nt+g, and a
  // unit-0 warp's B fragments covered every (filter, pixel) of its row band exactly once
#pragma unroll
  for (int nt = 0; nt < C::NT; ++nt) {
    bias_acc[nt] += __shfl_xor_sync(0xffffffffu, bias_acc[nt], 1);
    bias_acc[nt] += __shfl_xor_sync(0xffffffffu, bias_acc[nt], 2);
  }
  // sum the K slices of the CTA in a fixed tree order through shared memory
  // (slice s += slice s + half), in fragment order: conflict-free and coalesced
  for (int half = C::SLICES / 2; half >= 1; half >>= 1) {
    if (slice >= half && slice < 2 * half) {
      float *row = smem + (slice - half) * C::ROW;
#pragma unroll
      for (int fx = 0; fx < C::F; ++fx)
#pragma unroll
        for (int nt = 0; nt < C::NT; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i)
            row[(((unit * C::F + fx) * C::NT + nt) * 4 + i) * 32 + lane] = acc[fx][nt][i];
      if (unit == 0 && t == 0)
#pragma unroll
        for (int nt = 0; nt < C::NT; ++nt) row[C::FRAG_ELEMS + nt * 8 + g] = bias_acc[nt];
    }
    __syncthreads();
    if (slice < half) {
      const float *row = smem + slice * C::ROW;
#pragma unroll
      for (int fx = 0; fx < C::F; ++fx)
#pragma unroll
        for (int nt = 0; nt < C::NT; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i)
            acc[fx][nt][i] += row[(((unit * C::F + fx) * C::NT + nt) * 4 + i) * 32 + lane];
      if (unit == 0)
#pragma unroll
        for (int nt = 0; nt < C::NT; ++nt) bias_acc[nt] += row[C::FRAG_ELEMS + nt * 8 + g];
    }
    __syncthreads();
  }
  if (slice != 0) return;
  // Accumulators leave in fragment order (coalesced); the reduce kernel below maps
  // fragment slots to weights.  (Scattering 60 four-byte stores per thread into the
  // reference layout here costs 32 LSU cycles per store: 20 us per CTA.)
  float *out = partial + (int64_t)blockIdx.x * C::ROW;
#pragma unroll
  for (int fx = 0; fx < C::F; ++fx)
#pragma unroll
    for (int nt = 0; nt < C::NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i)
        out[(((unit * C::F + fx) * C::NT + nt) * 4 + i) * 32 + lane] = acc[fx][nt][i];
  if (unit == 0 && t == 0)
#pragma unroll
    for (int nt = 0; nt < C::NT; ++nt) out[C::FRAG_ELEMS + nt * 8 + g] = bias_acc[nt];
}

// partial[p][slot] summed over p = grp, grp + 8, ... in ascending order.  One partial row
// per CTA of the gradient kernel (at most one per SM): the first 20 terms — all of them up
// to 160 rows — are loaded before the first addition, so the thread waits for one L2 round
// trip instead of one per unrolled group.
__device__ __forceinline__ float sum_partial_rows(const float *__restrict__ partial, int n_partial,
                                                  int row_len, int slot, int grp) {
  constexpr int kAhead = 20;
  float v[kAhead];
#pragma unroll
  for (int k = 0; k < kAhead; ++k) {
    const int p = grp + 8 * k;
    v[k] = p < n_partial ? partial[(int64_t)p * row_len + slot] : 0.0f;
  }
  float part = 0.0f;
#pragma unroll
  for (int k = 0; k < kAhead; ++k) part += v[k];
  for (int p = grp + 8 * kAhead; p < n_partial; p += 8) part += partial[(int64_t)p * row_len + slot];
  return part;
}

// Fixed-order sum of the per-CTA partials + the SGD update, one thread per FRAGMENT
// slot (coalesced reads of every partial row; one scattered write per weight at the end).
// D fragment: slot i of lane (g,t) is row g + 8*(i>>1), column 2t + (i&1); row = input
// channel within the M tile, column = filter within the N tile.
template <class C>
__global__ void wgrad_mma_reduce_kernel(const float *__restrict__ partial, int n_partial,
                                        const float *Wt, float *out, const float *__restrict__ lr,
                                        int mode) {
  // block = 32 slots x 8 row groups: warp j sums partial rows j, j+8, ... of its 32 slots,
  // then the eight sums are added in order (deterministic)
  __shared__ float red[8][32];
  chain_wait();
  chain_release();
  const int lane_ = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int slot = blockIdx.x * 32 + lane_;
  const float part = slot < C::ROW ? sum_partial_rows(partial, n_partial, C::ROW, slot, grp) : 0.0f;
  red[grp][lane_] = part;
  __syncthreads();
  if (grp != 0 || slot >= C::ROW) return;
  int w;
  if (slot >= C::FRAG_ELEMS) {
    const int oc = slot - C::FRAG_ELEMS;  // [nt][g] with oc = 8nt + g
    if (oc >= C::OC) return;
    w = oc * C::K1 + C::K1 - 1;
  } else {
    const int lane = slot & 31, i = (slot >> 5) & 3;
    int r = slot >> 7;
    const int nt = r % C::NT; r /= C::NT;
    const int fx = r % C::F; r /= C::F;
    const int mt = r % C::MT, fy = r / C::MT;
    const int ic = mt * 16 + (lane >> 2) + ((i & 2) ? 8 : 0);
    const int oc = nt * 8 + 2 * (lane & 3) + (i & 1);
    if (ic >= C::IC || oc >= C::OC) return;
    w = oc * C::K1 + ic * C::F * C::F + fy * C::F + fx;
  }
  float acc = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) acc += red[j][lane_];
  apply_update(mode, out, Wt, w, lr[0], acc);
}

template <class C>
int launch(pb_context *ctx, const float *I, const float *Wt, const float *G, float *out,
           const float *lr, int mode, int batch) {
  auto kern = conv_wgrad_mma_kernel<C>;
  static bool configured = false;
  if (!configured) {
    PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    configured = true;
  }
  const int n_cta = std::min(batch, ctx->num_sms);
  const int spc = (batch + n_cta - 1) / n_cta;
  const int grid = (batch + spc - 1) / spc;
  if (ensure_workspace(ctx, (size_t)grid * C::ROW)) return 1;
  PB_CUDA(launch_chained(ctx, kern, dim3(grid), dim3(kThreads), C::SMEM, I, G, ctx->workspace,
                         batch, spc));
  PB_LAUNCHED(ctx);
  PB_CUDA(launch_forked(ctx, wgrad_mma_reduce_kernel<C>, dim3(pb_div_up(C::ROW, 32)), dim3(256), 0,
                         ctx->workspace, grid, Wt, out, lr, mode));
  PB_LAUNCHED(ctx);
  return 0;
}

// ---------------------------------------------------------------------------------
// Few input channels (the RGB layer: 3 channels would fill 19 % of a 16-row M tile).
// Here the filters are the M tile and (channel, fx) is the N dimension:
//   D[oc 16][(ic,fx) 8] += A[oc 16][pixel 8] * B[pixel 8][(ic,fx) 8]   per filter row fy
// 3 channels x 5 taps = 15 of 16 columns used.  A (the gradient tile) does not depend on
// the tap: its fragments are loaded once per K step and reused for the 5 filter rows; B is
// the input tile read at a per-lane (channel plane, fx) offset.  16 warps = 16 bands of
// output rows, each warp holds all 5 x 2 accumulator tiles.
// ---------------------------------------------------------------------------------
template <int IC_, int OC_, int S_>
struct NCfg {
  static constexpr int IC = IC_, OC = OC_, S = S_, F = 5, PADW = 2;
  static constexpr int WARPS = 16, THREADS = WARPS * 32;
  static constexpr int NCOL = IC * F;               // (ic, fx) columns
  static constexpr int NT = (NCOL + 7) / 8;
  static constexpr int SLICES = WARPS, RPS = S / SLICES;
  static constexpr int LEFT = 4, PITCH = S + 8, ROWS = S + 2 * PADW;
  static constexpr int PLANE = ((ROWS * PITCH + 31) / 32) * 32 + 8;  // = 8 (mod 32)
  static constexpr int GPLANE = S * S + 4;                            // = 4 (mod 8)
  static constexpr int IN_ELEMS = IC * PLANE, G_ELEMS = 16 * GPLANE;
  static constexpr int STAGE = IN_ELEMS + G_ELEMS;
  static constexpr int MAIN_ELEMS = 2 * STAGE;
  static constexpr int K1 = F * F * IC + 1;
  static constexpr int FRAG_ELEMS = F * NT * 4 * 32;  // [fy][nt][i][lane]
  static constexpr int ROW = FRAG_ELEMS + 16;          // + bias sums per filter
  static constexpr int RED_ELEMS = (SLICES / 2) * ROW;
  static constexpr size_t SMEM =
      sizeof(float) * (MAIN_ELEMS > RED_ELEMS ? MAIN_ELEMS : RED_ELEMS);
  static_assert(OC <= 16 && RPS * SLICES == S && S % 8 == 0, "envelope");
  static_assert(SMEM <= 220 * 1024, "shared memory");
};

// ACT >= 0: the gradient tile is FORMED by the (then four) producer warps from the stored
// convolution output O (passed as `G`) and the gradient Gp of the 2x2 max-pool's output behind
// the activation - act_pool_bwd_kernel (elementwise.cu) value for value - so dE/d(conv output)
// of the first convolution never travels through HBM.  nnet.cu uses it when nothing reads the
// layer's input gradient (the network's first convolution: BatchTrain has no input_gradients
// parameter, nnet/nnet.h:617-622).
constexpr int kFormWarps = 4;
template <class C, int ACT>
__global__ void __launch_bounds__(C::THREADS + (ACT >= 0 ? 32 * kFormWarps : 32), 1)
conv_wgrad_mma_rgb_kernel(const float *__restrict__ In, const float *__restrict__ G,
                          const float *__restrict__ Gp, float *__restrict__ partial, int batch,
                          int samples_per_cta) {
  extern __shared__ __align__(16) float smem[];
  constexpr int kProd = ACT >= 0 ? 32 * kFormWarps : 32, kAll = C::THREADS + kProd;
  const int tid = threadIdx.x, slice = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  __shared__ uint64_t full_bar[2], empty_bar[2];
  for (int i = tid; i < C::MAIN_ELEMS; i += kAll) smem[i] = 0.0f;
  if (tid == 0) {
    for (int st = 0; st < 2; ++st) {
      umma::mbar_init(&full_bar[st], ACT >= 0 ? 2 * kProd : kProd);  // cp.async + plain arrivals
      umma::mbar_init(&empty_bar[st], C::WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  chain_wait();  // chained launch (common.cuh): zero fill and barrier setup ran ahead of it
  chain_release();
  const int b0 = blockIdx.x * samples_per_cta;
  const int b1 = min(batch, b0 + samples_per_cta);

  if (slice >= C::WARPS) {  // producers
    const int pt = tid - C::THREADS;
    for (int b = b0; b < b1; ++b) {
      const int i = b - b0, st = i & 1;
      umma::mbar_wait(&empty_bar[st], ((i >> 1) & 1) ^ 1);
      float *in_s = smem + st * C::STAGE, *g_s = in_s + C::IN_ELEMS;
      const float *src = In + (int64_t)b * C::IC * C::S * C::S;
      for (int e = pt; e < C::IC * C::S * C::S / 4; e += kProd) {
        const int x4 = e % (C::S / 4), r = e / (C::S / 4);
        const int y = r % C::S, ic = r / C::S;
        __pipeline_memcpy_async(in_s + ic * C::PLANE + (y + C::PADW) * C::PITCH + C::LEFT + 4 * x4,
                                src + 4 * e, 16);
      }
      if constexpr (ACT < 0) {
        const float *gs = G + (int64_t)b * C::OC * C::S * C::S;
        for (int e = pt; e < C::OC * C::S * C::S / 4; e += kProd) {
          const int q = e % (C::S * C::S / 4), oc = e / (C::S * C::S / 4);
          __pipeline_memcpy_async(g_s + oc * C::GPLANE + 4 * q, gs + 4 * e, 16);
        }
        cp_async_arrive(&full_bar[st]);
      } else {
        cp_async_arrive(&full_bar[st]);
        // a thread owns two pool windows (2 rows x 4 columns) of every filter plane, four
        // planes in flight (constant strides: one base address per thread)
        constexpr int HS = C::S / 2, WQ = C::S / 4;
        static_assert(HS * WQ == kProd && C::OC % 4 == 0, "one item per thread and plane");
        const int xq = pt % WQ, orow = pt / WQ;
        const float *o = G + (int64_t)b * C::OC * C::S * C::S + 2 * orow * C::S + 4 * xq;
        const float *gp = Gp + (int64_t)b * C::OC * HS * HS + orow * HS + 2 * xq;
        float *gd = g_s + 2 * orow * C::S + 4 * xq;
#pragma unroll 1
        for (int oc0 = 0; oc0 < C::OC; oc0 += 4) {
          float4 o0[4], o1[4];
          float2 gg[4];
          int dst[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            o0[k] = *reinterpret_cast<const float4 *>(o + (oc0 + k) * C::S * C::S);
            o1[k] = *reinterpret_cast<const float4 *>(o + (oc0 + k) * C::S * C::S + C::S);
            gg[k] = *reinterpret_cast<const float2 *>(gp + (oc0 + k) * HS * HS);
            dst[k] = (oc0 + k) * C::GPLANE;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 p0 = o0[k], p1 = o1[k];
            const float a00 = act_fwd<ACT>(p0.x), a01 = act_fwd<ACT>(p0.y), a02 = act_fwd<ACT>(p0.z),
                        a03 = act_fwd<ACT>(p0.w);
            const float a10 = act_fwd<ACT>(p1.x), a11 = act_fwd<ACT>(p1.y), a12 = act_fwd<ACT>(p1.z),
                        a13 = act_fwd<ACT>(p1.w);
            const float m0 = fmaxf(fmaxf(a00, a01), fmaxf(a10, a11));
            const float m1 = fmaxf(fmaxf(a02, a03), fmaxf(a12, a13));
            float4 d0, d1;
            d0.x = act_bwd<ACT>(p0.x, a00 < m0 ? 0.0f : gg[k].x);
            d0.y = act_bwd<ACT>(p0.y, a01 < m0 ? 0.0f : gg[k].x);
            d0.z = act_bwd<ACT>(p0.z, a02 < m1 ? 0.0f : gg[k].y);
            d0.w = act_bwd<ACT>(p0.w, a03 < m1 ? 0.0f : gg[k].y);
            d1.x = act_bwd<ACT>(p1.x, a10 < m0 ? 0.0f : gg[k].x);
            d1.y = act_bwd<ACT>(p1.y, a11 < m0 ? 0.0f : gg[k].x);
            d1.z = act_bwd<ACT>(p1.z, a12 < m1 ? 0.0f : gg[k].y);
            d1.w = act_bwd<ACT>(p1.w, a13 < m1 ? 0.0f : gg[k].y);
            *reinterpret_cast<float4 *>(gd + dst[k]) = d0;
            *reinterpret_cast<float4 *>(gd + dst[k] + C::S) = d1;
          }
        }
        umma::mbar_arrive(&full_bar[st]);
      }
    }
    __pipeline_commit();
    __pipeline_wait_prior(0);
  }

  // this lane's column n = 8nt + g of the B fragments = (channel, fx); columns past
  // IC*F read channel IC-1 (their products land in D columns that are never stored)
  int boff[C::NT];
#pragma unroll
  for (int nt = 0; nt < C::NT; ++nt) {
    const int n = min(nt * 8 + g, C::NCOL - 1);
    boff[nt] = (n / C::F) * C::PLANE + (n % C::F) + (C::LEFT - C::PADW) + t;
  }

  float acc[C::F][C::NT][4];
#pragma unroll
  for (int fy = 0; fy < C::F; ++fy)
#pragma unroll
    for (int nt = 0; nt < C::NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[fy][nt][i] = 0.0f;
  float bias_lo = 0.0f, bias_hi = 0.0f;  // filters g and g + 8

  for (int b = b0; b < b1 && slice < C::WARPS; ++b) {
    const int st = (b - b0) & 1;
    umma::mbar_wait(&full_bar[st], ((b - b0) >> 1) & 1);
    const float *in_s = smem + st * C::STAGE, *g_s = in_s + C::IN_ELEMS;
#pragma unroll 1
    for (int yy = 0; yy < C::RPS; ++yy) {
      const int y = slice * C::RPS + yy;
#pragma unroll 1
      for (int x0 = 0; x0 < C::S; x0 += 8) {
        uint32_t ah[4], al[4];
        {
          const float *p = g_s + g * C::GPLANE + y * C::S + x0 + t;
          const float v0 = p[0], v1 = p[8 * C::GPLANE], v2 = p[4], v3 = p[8 * C::GPLANE + 4];
          split(v0, ah[0], al[0]);
          split(v1, ah[1], al[1]);
          split(v2, ah[2], al[2]);
          split(v3, ah[3], al[3]);
          bias_lo += v0 + v2;
          bias_hi += v1 + v3;
        }
        uint32_t bh[C::F][C::NT][2], bl[C::F][C::NT][2];
#pragma unroll
        for (int fy = 0; fy < C::F; ++fy)
#pragma unroll
          for (int nt = 0; nt < C::NT; ++nt) {
            const float *p = in_s + boff[nt] + (y + fy) * C::PITCH + x0;
            split(p[0], bh[fy][nt][0], bl[fy][nt][0]);
            split(p[4], bh[fy][nt][1], bl[fy][nt][1]);
          }
        // pass-major: consecutive MMAs on one accumulator are F*NT instructions apart
#pragma unroll
        for (int fy = 0; fy < C::F; ++fy)
#pragma unroll
          for (int nt = 0; nt < C::NT; ++nt) mma_tf32(acc[fy][nt], al, bh[fy][nt]);
#pragma unroll
        for (int fy = 0; fy < C::F; ++fy)
#pragma unroll
          for (int nt = 0; nt < C::NT; ++nt) mma_tf32(acc[fy][nt], ah, bl[fy][nt]);
#pragma unroll
        for (int fy = 0; fy < C::F; ++fy)
#pragma unroll
          for (int nt = 0; nt < C::NT; ++nt) mma_tf32(acc[fy][nt], ah, bh[fy][nt]);
      }
    }
    __syncwarp();
    if (lane == 0) umma::mbar_arrive(&empty_bar[st]);
  }
  __syncthreads();

  // lanes t = 0..3 hold disjoint pixel subsets of filters g / g + 8
  bias_lo += __shfl_xor_sync(0xffffffffu, bias_lo, 1);
  bias_lo += __shfl_xor_sync(0xffffffffu, bias_lo, 2);
  bias_hi += __shfl_xor_sync(0xffffffffu, bias_hi, 1);
  bias_hi += __shfl_xor_sync(0xffffffffu, bias_hi, 2);
  // row bands summed in a fixed tree order through shared memory (fragment order)
  for (int half = C::SLICES / 2; half >= 1; half >>= 1) {
    if (slice >= half && slice < 2 * half) {
      float *row = smem + (slice - half) * C::ROW;
#pragma unroll
      for (int fy = 0; fy < C::F; ++fy)
#pragma unroll
        for (int nt = 0; nt < C::NT; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) row[((fy * C::NT + nt) * 4 + i) * 32 + lane] = acc[fy][nt][i];
      if (t == 0) { row[C::FRAG_ELEMS + g] = bias_lo; row[C::FRAG_ELEMS + g + 8] = bias_hi; }
    }
    __syncthreads();
    if (slice < half) {
      const float *row = smem + slice * C::ROW;
#pragma unroll
      for (int fy = 0; fy < C::F; ++fy)
#pragma unroll
        for (int nt = 0; nt < C::NT; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[fy][nt][i] += row[((fy * C::NT + nt) * 4 + i) * 32 + lane];
      bias_lo += row[C::FRAG_ELEMS + g];
      bias_hi += row[C::FRAG_ELEMS + g + 8];
    }
    __syncthreads();
  }
  if (slice != 0) return;
  float *out = partial + (int64_t)blockIdx.x * C::ROW;
#pragma unroll
  for (int fy = 0; fy < C::F; ++fy)
#pragma unroll
    for (int nt = 0; nt < C::NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) out[((fy * C::NT + nt) * 4 + i) * 32 + lane] = acc[fy][nt][i];
  if (t == 0) { out[C::FRAG_ELEMS + g] = bias_lo; out[C::FRAG_ELEMS + g + 8] = bias_hi; }
}

// slot i of lane (g,t): row g + 8*(i>>1) = filter, column 8nt + 2t + (i&1) = (channel, fx)
template <class C>
__global__ void wgrad_mma_rgb_reduce_kernel(const float *__restrict__ partial, int n_partial,
                                            const float *Wt, float *out,
                                            const float *__restrict__ lr, int mode) {
  __shared__ float red[8][32];
  chain_wait();
  chain_release();
  const int lane_ = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int slot = blockIdx.x * 32 + lane_;
  const float part = slot < C::ROW ? sum_partial_rows(partial, n_partial, C::ROW, slot, grp) : 0.0f;
  red[grp][lane_] = part;
  __syncthreads();
  if (grp != 0 || slot >= C::ROW) return;
  int w;
  if (slot >= C::FRAG_ELEMS) {
    const int oc = slot - C::FRAG_ELEMS;
    if (oc >= C::OC) return;
    w = oc * C::K1 + C::K1 - 1;
  } else {
    const int lane = slot & 31, i = (slot >> 5) & 3;
    const int r = slot >> 7, nt = r % C::NT, fy = r / C::NT;
    const int oc = (lane >> 2) + ((i & 2) ? 8 : 0);
    const int n = nt * 8 + 2 * (lane & 3) + (i & 1);
    if (oc >= C::OC || n >= C::NCOL) return;
    w = oc * C::K1 + (n / C::F) * C::F * C::F + fy * C::F + (n % C::F);
  }
  float acc = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) acc += red[j][lane_];
  apply_update(mode, out, Wt, w, lr[0], acc);
}

template <class C, int ACT>
int launch_rgb(pb_context *ctx, const float *I, const float *Wt, const float *G, const float *Gp,
               float *out, const float *lr, int mode, int batch) {
  auto kern = conv_wgrad_mma_rgb_kernel<C, ACT>;
  static bool configured = false;
  if (!configured) {
    PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    configured = true;
  }
  const int n_cta = std::min(batch, ctx->num_sms);
  const int spc = (batch + n_cta - 1) / n_cta;
  const int grid = (batch + spc - 1) / spc;
  if (ensure_workspace(ctx, (size_t)grid * C::ROW)) return 1;
  PB_CUDA(launch_chained(ctx, kern, dim3(grid),
                         dim3(C::THREADS + (ACT >= 0 ? 32 * kFormWarps : 32)), C::SMEM, I, G, Gp,
                         ctx->workspace, batch, spc));
  PB_LAUNCHED(ctx);
  PB_CUDA(launch_forked(ctx, wgrad_mma_rgb_reduce_kernel<C>, dim3(pb_div_up(C::ROW, 32)), dim3(256),
                         0, ctx->workspace, grid, Wt, out, lr, mode));
  PB_LAUNCHED(ctx);
  return 0;
}

}  // namespace

// Returns -1 outside the envelope (the FFMA kernels of conv_tiled.cu then run).
int conv_weight_update_mma(pb_context *ctx, const pb_layer_desc *l, const float *I, const float *Wt,
                           const float *G, float *out, const float *lr, int mode, int64_t batch) {
  static const bool on = [] {
    const char *e = getenv("PB_DISABLE_UMMA"), *f = getenv("PB_WGRAD_MMA");
    return !(e && e[0] == '1') && !(f && f[0] == '0');
  }();
  if (!on || l->stride != 1 || l->filter_width != 5 || l->filter_height != 5 || l->padding != 2 ||
      l->width != l->height || batch < 64 || batch > (1 << 24))
    return -1;
  const int w = (int)l->width, d = (int)l->depth, f = (int)l->num_filters;
  if (w == 16 && d == 16 && f == 20) return launch<WCfg<16, 20, 16>>(ctx, I, Wt, G, out, lr, mode, (int)batch);
  if (w == 8 && d == 20 && f == 20) return launch<WCfg<20, 20, 8>>(ctx, I, Wt, G, out, lr, mode, (int)batch);
  if (w == 32 && d == 3 && f == 16)
    return launch_rgb<NCfg<3, 16, 32>, -1>(ctx, I, Wt, G, nullptr, out, lr, mode, (int)batch);
  return -1;
}

// The weight gradient of a (convolution, activation, 2x2 max-pool) block straight from the
// convolution's stored output O and the gradient Gp of the pool's output; dE/d(conv output) is
// formed inside the kernel and the convolution's input gradient is NOT computed (the caller
// has no reader for it).  Returns -1 outside the envelope.
int conv_act_pool_weight_update_mma(pb_context *ctx, const pb_layer_desc *l, int act,
                                    const pb_layer_desc *pool, const float *I, const float *Wt,
                                    const float *O, const float *Gp, float *out, const float *lr,
                                    int mode, int64_t batch) {
  static const bool on = [] {
    const char *e = getenv("PB_DISABLE_UMMA"), *f = getenv("PB_WGRAD_MMA"), *h = getenv("PB_WGRAD_FORM");
    return !(e && e[0] == '1') && !(f && f[0] == '0') && !(h && h[0] == '0');
  }();
  if (!on || l->stride != 1 || l->filter_width != 5 || l->filter_height != 5 || l->padding != 2 ||
      l->width != l->height || batch < 64 || batch > (1 << 24))
    return -1;
  const int w = (int)l->width, d = (int)l->depth, f = (int)l->num_filters;
  if (pool->width != w || pool->height != w || pool->depth != f || pool->out_width * 2 != w ||
      pool->out_height * 2 != w)
    return -1;
  if ((((uintptr_t)O) & 15) || (((uintptr_t)Gp) & 7)) return -1;
  if (!(w == 32 && d == 3 && f == 16)) return -1;
  using C = NCfg<3, 16, 32>;
  switch (act) {
    case PB_IDENTITY: return launch_rgb<C, PB_IDENTITY>(ctx, I, Wt, O, Gp, out, lr, mode, (int)batch);
    case PB_RELU: return launch_rgb<C, PB_RELU>(ctx, I, Wt, O, Gp, out, lr, mode, (int)batch);
    case PB_LEAKY_RELU: return launch_rgb<C, PB_LEAKY_RELU>(ctx, I, Wt, O, Gp, out, lr, mode, (int)batch);
    case PB_SIGMOID: return launch_rgb<C, PB_SIGMOID>(ctx, I, Wt, O, Gp, out, lr, mode, (int)batch);
    default: return -1;
  }
}

}  // namespace pb
