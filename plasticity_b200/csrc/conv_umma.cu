// 3xTF32 tcgen05 implicit-GEMM convolution (forward and input-gradient) for the
// batched stride-1 "same" convolutions of the CIFAR network
// (nnet/cifar_test.cc:293-350; semantics nnet/convolution_layer.cc:50-224).
//
//   Out[b,oc,r,c] = bias_oc + sum_{ic,fy,fx} Wk[oc,ic,fy,fx] * In0[b,ic,r-p+fy,c-p+fx]
//
// (input-gradient: In = G, Wk = flipped / channel-transposed filters, no bias;
// with pad == filter/2 the reference's centre guard, convolution_layer.cc:209,
// is the identity.)
//
// GEMM view: M = output pixels, N = output channels, K = (tap, input channel).
// The image of one work item (a sample, or a 16-row block of it) is stored
// ONCE in shared memory as [4-channel chunk][padded pixel][4 floats], split
// into a TF32 "hi" plane set and an fp32-residual "lo" plane set.  That layout
// is exactly tcgen05's K-major no-swizzle canonical operand layout
//   byte(row m, chunk) = chunk*LBO + (m/8)*SBO + (m%8)*16
// with SBO = padded image-row pitch: an M-tile is an 8-pixel-wide strip of 16
// (or 8) image rows, and moving the descriptor's start address by
// (fy*pitch + fx)*16 bytes selects filter tap (fy,fx) — no im2col copy exists
// anywhere.  One K=8 MMA covers two 4-channel chunks (LBO = their distance).
// (Layout rules validated on B200 by tools/umma_probe.cu.)
//
// 3xTF32: x = hi + lo, hi = tf32(x).  Per K block two MMAs:
//   D[:, 0:N1]  += A_hi * [B_hi | B_lo]      (hi*hi and hi*lo side by side in N)
//   D[:, 0:N2]  += A_lo * B_hi               (accumulated onto the hi*hi columns)
// and the epilogue adds the two column groups: fp32-level accuracy (dropped
// term lo*lo ~ 2^-22), inside the |rel| <= 1e-4 bar.
//
// Warp roles (544 threads, one persistent CTA per SM):
//   warps 0-7  epilogue: tcgen05.ld accumulators -> registers -> bias -> global
//              (warp w reads TMEM lane quadrant w&3; warps 0-3 / 4-7 split the channels)
//   warp  8    MMA issuer (one elected lane), TMEM allocator
//   warps 9-16 producers: global CHW fp32 -> hi/lo split -> shared memory
// Two shared-memory stages and two TMEM accumulator stages, mbarrier
// full/empty pairs (tcgen05.commit arrives on the "empty"/"full" barriers).
// Measured operand-fetch bound of this instruction shape: (M+N)/4 cycles per
// MMA (128 B/clk of shared-memory operand bandwidth).
#include "common.cuh"
#include "kernels.cuh"
#include "umma.cuh"

namespace pb {
namespace {

using namespace umma;

// warps 0-7 epilogue, warp 8 MMA issuer, warps 9-16 producers
constexpr int kEpilogue = 256;
constexpr int kMmaWarp = kEpilogue / 32;
constexpr int kProducers = 256;
constexpr int kThreads = kEpilogue + 32 + kProducers;

struct KBlock {
  int first, second;  // chunk ids (tap*CH + c); second == -1: zero-weight dummy
  int off;            // A start offset of `first`, 16-byte units (= pixels)
  int lbo;            // A distance first -> second, 16-byte units
};

// NS ("N-stacked", for very few output channels): the fx taps are moved from K
// into N.  The MMA computes, per pixel (r,x'), Q[(fx,oc)] = sum_{fy,ic} W[oc,ic,fy,fx] *
// In0[ic, r-p+fy, x'] (no horizontal shift), K shrinks F-fold, and the epilogue
// gathers Out[oc,r,c] = sum_fx Q[(r, c-p+fx)][(fx,oc)] through shared memory.
template <int IC_, int OC_, int W_, int H_, int F_, bool BWD_, bool NS_ = false>
struct UCfg {
  static constexpr int IC = IC_, OC = OC_, W = W_, H = H_, F = F_;
  static constexpr bool BWD = BWD_, NS = NS_;
  static constexpr int NOUT = NS ? F * OC : OC;    // GEMM columns (before hi/lo doubling)
  static constexpr int PAD = F / 2;
  static constexpr int P = W + F - 1;              // shared-memory row pitch, pixels
  static constexpr int RI = H >= 16 ? 16 : 8;      // output rows per work item
  static constexpr int SR = RI + F - 1;            // shared-memory rows per item
  static constexpr int PL = ((SR * P + 7) / 8) * 8;  // pixels per chunk plane
  static constexpr int CH = (IC + 3) / 4;          // 4-channel chunks
  static constexpr int NCHUNK = (NS ? F : F * F) * CH;
  static constexpr int NKB = (NCHUNK + 1) / 2;     // K=8 blocks
  static constexpr int MT = RI * 8;                // MMA M: 8-pixel strip x RI rows
  static constexpr int NT = W / 8;                 // strips (tiles) per item
  static constexpr int NP = ((NOUT + 7) / 8) * 8;  // rows of the hi (and lo) weight block
  static constexpr int N1 = ((2 * NP + 15) / 16) * 16;
  static constexpr int N2 = MT == 128 ? ((NP + 15) / 16) * 16 : NP;
  static constexpr int A_STAGE = 2 * CH * PL;      // float4 per stage: hi planes, lo planes
  static constexpr int B_UNITS = NKB * 2 * N1;     // float4
  static constexpr int IPS = H / RI;               // items per sample
  static constexpr int TMEM_STAGE = 128;           // accumulator columns per stage
  static constexpr int STG_REC = NOUT | 1;         // odd record pitch: conflict-free staging
  static constexpr int STG_FLOATS = NS ? RI * W * STG_REC : 0;
  static constexpr size_t SMEM =
      16 * (size_t)(2 * A_STAGE + B_UNITS) + 32 * 4 + 8 * 8 + 16 + 4 * (size_t)STG_FLOATS;
  static_assert(!NS || (BWD && MT == 128 && NOUT <= 32), "N-stacked variant envelope");
  static_assert(W % 8 == 0 && H % RI == 0, "strip tiling");
  static_assert(NT * N1 <= TMEM_STAGE, "accumulators of one item must fit a TMEM stage");
  static_assert(N2 <= N1 && N1 % 16 == 0 && N2 % 8 == 0, "MMA N");
  static_assert(SMEM <= 227 * 1024, "shared memory");

  static constexpr int chunk_off(int j) {
    return NS ? (j / CH) * P + PAD + (j % CH) * PL
              : ((j / CH) / F) * P + ((j / CH) % F) + (j % CH) * PL;
  }
  static constexpr KBlock kblock(int kb) {
    const int j0 = 2 * kb, j1 = 2 * kb + 1;
    if (j1 >= NCHUNK) return KBlock{j0, -1, chunk_off(j0), 1};
    const int o0 = chunk_off(j0), o1 = chunk_off(j1);
    return o1 > o0 ? KBlock{j0, j1, o0, o1 - o0} : KBlock{j1, j0, o1, o0 - o1};
  }
};


// Epilogue of one accumulator tile for one half of the output channels:
// out = (hi*hi + hi*lo columns) + bias -> global; optionally the 2x2 max-pool of
// act(out) (act is monotone, so pool(act(x)) == act(pool(x)) exactly: one
// activation per window instead of four).
template <class C, int ACT, int HALF>
__device__ __forceinline__ void epilogue_tile(uint32_t taddr, const float *bias_s, float *dst,
                                              float *dst2, bool row_ok, bool writer) {
  constexpr int FHALF = (C::OC + 1) / 2;
  constexpr int F0 = HALF * FHALF;
  constexpr int NF = (C::OC - F0) < FHALF ? (C::OC - F0) : FHALF;
  constexpr int C0 = F0 & ~7;                       // 8-aligned first column loaded
  constexpr int LW = (F0 - C0 + NF <= 8) ? 8 : 16;  // columns per tcgen05.ld
  static_assert(F0 - C0 + NF <= 16, "channel half must fit one 16-column load");
  static_assert(C::NP + C0 + LW <= C::N1, "loads stay inside the tile's accumulator columns");
  if (NF <= 0) return;
  float vh[LW], vl[LW];
  if (LW == 8) { tmem_ld8(taddr + C0, vh); tmem_ld8(taddr + C::NP + C0, vl); }
  else { tmem_ld16(taddr + C0, vh); tmem_ld16(taddr + C::NP + C0, vl); }
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < NF; ++j) {
    const int f = F0 + j;
    const float v = (vh[f - C0] + vl[f - C0]) + bias_s[f];
    if (row_ok) dst[(int64_t)f * C::H * C::W] = v;
    if (ACT >= 0) {
      // 2x2 window partners: x^1 is lane^1, row^1 is lane^8 (row groups of 8 lanes)
      float m = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
      if (writer) dst2[(int64_t)f * (C::H / 2) * (C::W / 2)] = act_fwd<(ACT >= 0 ? ACT : 0)>(m);
    }
  }
}

// ACT >= 0, forward: fused epilogue - additionally writes Out2 = 2x2 max-pool of
// act(conv output) (MaxPoolLayer forward of ActivationLayer forward).
// ACT >= 0, input gradient (BWD): fused PRODUCER - `In` is the gradient of the 2x2 max-pool's
// OUTPUT (H/2 x W/2 per channel) and Aux the stored convolution output of the layer whose
// gradient this is; the producers form dE/d(conv output) = act'(O) * (act(O) is a maximum of
// its window ? Gp : 0) on the fly (max_pool_layer.cc:88-141 + activation_layer.cc:13-22,
// value for value what act_pool_bwd_kernel computes), feed it to the MMAs and also write it
// to Out2 for the weight-gradient kernel: no stand-alone pool / activation backward pass.
template <class C, int ACT>
__global__ void __launch_bounds__(kThreads, 1)
conv_umma_kernel(const float *__restrict__ In, const float *__restrict__ Wt,
                 float *__restrict__ Out, float *__restrict__ Out2, int batch, int early_w,
                 const float *__restrict__ Aux) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float4 *a_s = reinterpret_cast<float4 *>(smem_raw);         // [2][A_STAGE]
  float4 *b_s = a_s + 2 * C::A_STAGE;                         // [B_UNITS]
  float *bias_s = reinterpret_cast<float *>(b_s + C::B_UNITS);  // [32]
  uint64_t *bars = reinterpret_cast<uint64_t *>(bias_s + 32);
  uint64_t *a_full = bars, *a_empty = bars + 2, *d_full = bars + 4, *d_empty = bars + 6;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 8);
  float *stg = reinterpret_cast<float *>(tmem_ptr + 4);      // [RI*W][STG_REC] (NS only)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int TAPS = C::F * C::F;
  constexpr int K1 = C::BWD ? TAPS * C::OC + 1 : TAPS * C::IC + 1;  // reference filter-block stride

  // ---- one-time setup -----------------------------------------------------
  // Raw filter block (reference layout) -> shared-memory scratch (the A stages,
  // not yet in use) with coalesced loads, then the hi/lo, chunk-ordered B image
  // is built from shared memory.
  constexpr int NWT = (C::BWD ? C::IC : C::OC) * K1;
  static_assert(NWT <= 8 * C::A_STAGE, "weight scratch must fit the A stages");
  // Chained launch (common.cuh): when the launch before this one does not write this
  // layer's filters, the whole setup runs while that kernel is still finishing.
  if (!early_w) chain_wait();
  float *w_raw = reinterpret_cast<float *>(a_s);
  // L2 loads (ld.global.cg): ahead of the dependency wait nothing may rely on this SM's L1
  // having been invalidated since the filters were last updated
  {
    // every load of the thread in flight before the first store (one L2 round trip)
    constexpr int WL = (NWT + kThreads - 1) / kThreads;
    float wv[WL];
#pragma unroll
    for (int k = 0; k < WL; ++k) {
      const int i = tid + k * kThreads;
      wv[k] = i < NWT ? __ldcg(Wt + i) : 0.0f;
    }
#pragma unroll
    for (int k = 0; k < WL; ++k) {
      const int i = tid + k * kThreads;
      if (i < NWT) w_raw[i] = wv[k];
    }
  }
  __syncthreads();
  for (int u = tid; u < C::B_UNITS; u += blockDim.x) {
    const int slot = u / C::N1, n = u - slot * C::N1;
    const KBlock kb = C::kblock(slot >> 1);
    const int j = (slot & 1) ? kb.second : kb.first;
    const bool is_lo = n >= C::NP;
    const int col = is_lo ? n - C::NP : n;          // GEMM column
    float w[4] = {0.f, 0.f, 0.f, 0.f};
    if (j >= 0 && col < C::NOUT && n < 2 * C::NP) {
      const int c = j % C::CH;
      // NS: column = (fx, oc), chunk = (fy, c).  Otherwise column = oc, chunk = (tap, c).
      const int oc = C::NS ? col % C::OC : col;
      const int tap = C::NS ? (j / C::CH) * C::F + col / C::OC : j / C::CH;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ic = 4 * c + i;
        if (ic < C::IC) {
          const float x = C::BWD ? w_raw[ic * K1 + oc * TAPS + (TAPS - 1 - tap)]
                                 : w_raw[oc * K1 + ic * TAPS + tap];
          const float h = tf32_hi(x);
          w[i] = is_lo ? x - h : h;
        }
      }
    }
    b_s[u] = make_float4(w[0], w[1], w[2], w[3]);
  }
  if (tid < 32) bias_s[tid] = (!C::BWD && tid < C::OC) ? w_raw[tid * K1 + K1 - 1] : 0.0f;
  __syncthreads();
  for (int i = tid; i < 2 * C::A_STAGE; i += blockDim.x) a_s[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) {
    mbar_init(&a_full[0], kProducers); mbar_init(&a_full[1], kProducers);
    mbar_init(&a_empty[0], 1);  mbar_init(&a_empty[1], 1);
    mbar_init(&d_full[0], 1);   mbar_init(&d_full[1], 1);
    mbar_init(&d_empty[0], kEpilogue); mbar_init(&d_empty[1], kEpilogue);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_ptr)),
                 "r"(2u * C::TMEM_STAGE)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_proxy();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int n_items = batch * C::IPS;
  if (early_w) chain_wait();
  chain_release();

  if (warp > kMmaWarp) {
    // ---- producers: global CHW -> [chunk][pixel][4] hi / lo ----------------
    // All global loads of an item are issued before the first conversion so the
    // memory latency is paid once per item, not once per element.
    const int ptid = tid - (kEpilogue + 32);
    constexpr int TOTAL = C::CH * C::SR * C::W;
    constexpr int ITERS = (TOTAL + kProducers - 1) / kProducers;
    int it = 0;
    if constexpr (C::BWD && ACT >= 0) {
      // one thread = one 2x2 pool window x one 4-channel chunk: 16 gradient values from
      // 8 float2 loads of the convolution output and 4 pooled gradients
      constexpr int ACTV = ACT >= 0 ? ACT : 0;
      constexpr int WW = C::W / 2, WR = C::SR / 2;       // windows per row, window rows per item
      constexpr int WTOTAL = C::CH * WR * WW;
      constexpr int WITERS = (WTOTAL + kProducers - 1) / kProducers;
      static_assert(C::SR % 2 == 0 && C::PAD % 2 == 0 && C::RI % 2 == 0, "window-aligned rows");
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int s = it & 1, ph = (it >> 1) & 1;
        const int b = item / C::IPS, rb = item - b * C::IPS;
        const int y0 = rb * C::RI - C::PAD;               // even: shared-memory rows pair up as windows
        float4 *hi = a_s + s * C::A_STAGE;
        float4 *lo = hi + C::CH * C::PL;
        const float *o_src = Aux + (int64_t)b * C::IC * C::H * C::W;
        const float *g_src = In + (int64_t)b * C::IC * (C::H / 2) * (C::W / 2);
        float *g_dst = Out2 + (int64_t)b * C::IC * C::H * C::W;
        float2 o0[WITERS][4], o1[WITERS][4];
        float gp[WITERS][4];
#pragma unroll
        for (int k = 0; k < WITERS; ++k) {
          const int e = ptid + k * kProducers;
          const int xw = e % WW, t = e / WW;
          const int wr = t % WR, c = t / WR;
          const int y = y0 + 2 * wr;
          const bool ok = e < WTOTAL && y >= 0 && y < C::H;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int ic = 4 * c + i;
            const bool live = ok && ic < C::IC;
            const float *po = o_src + ((int64_t)ic * C::H + y) * C::W + 2 * xw;
            o0[k][i] = live ? __ldg(reinterpret_cast<const float2 *>(po)) : make_float2(0.f, 0.f);
            o1[k][i] = live ? __ldg(reinterpret_cast<const float2 *>(po + C::W)) : make_float2(0.f, 0.f);
            gp[k][i] = live ? __ldg(g_src + ((int64_t)ic * (C::H / 2) + (y >> 1)) * (C::W / 2) + xw) : 0.0f;
          }
        }
        mbar_wait(&a_empty[s], ph ^ 1);
#pragma unroll
        for (int k = 0; k < WITERS; ++k) {
          const int e = ptid + k * kProducers;
          if (e < WTOTAL) {
            const int xw = e % WW, t = e / WW;
            const int wr = t % WR, c = t / WR;
            const int y = y0 + 2 * wr;
            const bool inside = y >= 0 && y < C::H;
            // rows of this item's own band (not the halo shared with its neighbours)
            const bool own = inside && 2 * wr >= C::PAD && 2 * wr < C::PAD + C::RI;
            float d[4][4];  // [pixel 00 01 10 11][channel]
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float a00 = act_fwd<ACTV>(o0[k][i].x), a01 = act_fwd<ACTV>(o0[k][i].y);
              const float a10 = act_fwd<ACTV>(o1[k][i].x), a11 = act_fwd<ACTV>(o1[k][i].y);
              const float m = fmaxf(fmaxf(a00, a01), fmaxf(a10, a11));
              const float g = gp[k][i];
              d[0][i] = act_bwd<ACTV>(o0[k][i].x, a00 < m ? 0.0f : g);
              d[1][i] = act_bwd<ACTV>(o0[k][i].y, a01 < m ? 0.0f : g);
              d[2][i] = act_bwd<ACTV>(o1[k][i].x, a10 < m ? 0.0f : g);
              d[3][i] = act_bwd<ACTV>(o1[k][i].y, a11 < m ? 0.0f : g);
              const int ic = 4 * c + i;
              if (own && ic < C::IC) {
                float *pg = g_dst + ((int64_t)ic * C::H + y) * C::W + 2 * xw;
                *reinterpret_cast<float2 *>(pg) = make_float2(d[0][i], d[1][i]);
                *reinterpret_cast<float2 *>(pg + C::W) = make_float2(d[2][i], d[3][i]);
              }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float v0 = inside ? d[q][0] : 0.0f, v1 = inside ? d[q][1] : 0.0f,
                          v2 = inside ? d[q][2] : 0.0f, v3 = inside ? d[q][3] : 0.0f;
              const float h0 = tf32_hi(v0), h1 = tf32_hi(v1), h2 = tf32_hi(v2), h3 = tf32_hi(v3);
              const int pix = c * C::PL + (2 * wr + (q >> 1)) * C::P + 2 * xw + (q & 1) + C::PAD;
              hi[pix] = make_float4(h0, h1, h2, h3);
              lo[pix] = make_float4(v0 - h0, v1 - h1, v2 - h2, v3 - h3);
            }
          }
        }
        fence_async_proxy();
        mbar_arrive(&a_full[s]);
      }
    } else
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      const int b = item / C::IPS, rb = item - b * C::IPS;
      const int y0 = rb * C::RI - C::PAD;
      float4 *hi = a_s + s * C::A_STAGE;
      float4 *lo = hi + C::CH * C::PL;
      const float *src = In + (int64_t)b * C::IC * C::H * C::W;
      float v[ITERS][4];
#pragma unroll
      for (int k = 0; k < ITERS; ++k) {
        const int e = ptid + k * kProducers;
        const int x = e % C::W;
        const int t = e / C::W;
        const int sy = t % C::SR, c = t / C::SR;
        const int y = y0 + sy;
        const bool ok = e < TOTAL && y >= 0 && y < C::H;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ic = 4 * c + i;
          v[k][i] = (ok && ic < C::IC) ? __ldg(src + (ic * C::H + y) * C::W + x) : 0.0f;
        }
      }
      mbar_wait(&a_empty[s], ph ^ 1);
#pragma unroll
      for (int k = 0; k < ITERS; ++k) {
        const int e = ptid + k * kProducers;
        if (e < TOTAL) {
          const int x = e % C::W;
          const int t = e / C::W;
          const int sy = t % C::SR, c = t / C::SR;
          const float h0 = tf32_hi(v[k][0]), h1 = tf32_hi(v[k][1]), h2 = tf32_hi(v[k][2]),
                      h3 = tf32_hi(v[k][3]);
          const int pix = c * C::PL + sy * C::P + x + C::PAD;
          hi[pix] = make_float4(h0, h1, h2, h3);
          lo[pix] = make_float4(v[k][0] - h0, v[k][1] - h1, v[k][2] - h2, v[k][3] - h3);
        }
      }
      fence_async_proxy();
      mbar_arrive(&a_full[s]);
    }
  } else if (warp == kMmaWarp) {
    // ---- MMA issuer ----------------------------------------------------------
    const bool leader = elect_one();
#ifdef PB_UMMA_PROF
    long long pt[3] = {0, 0, 0};
    const long long pt_start = clock64();
#endif
    constexpr uint32_t idesc1 = make_idesc(C::MT, C::N1), idesc2 = make_idesc(C::MT, C::N2);
    constexpr uint32_t a_hi32 = (uint32_t)C::P | (1u << 14);  // SBO = row pitch, version 1
    constexpr uint32_t b_hi32 = 8u | (1u << 14);              // SBO = 128 B
    const uint32_t b_lo32 = (smem_u32(b_s) >> 4) | ((uint32_t)C::N1 << 16);  // LBO = N1 rows
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
#ifdef PB_UMMA_PROF
      long long t0_ = clock64();
#endif
      mbar_wait(&a_full[s], ph);
#ifdef PB_UMMA_PROF
      pt[0] += clock64() - t0_; t0_ = clock64();
#endif
      mbar_wait(&d_empty[s], ph ^ 1);
#ifdef PB_UMMA_PROF
      pt[1] += clock64() - t0_; t0_ = clock64();
#endif
      tc_fence_after();
      // Descriptor arithmetic is warp-uniform (uniform datapath); only the
      // tcgen05 instructions themselves are predicated on the elected lane.
      const uint32_t a_stage = (smem_u32(a_s) >> 4) + (uint32_t)(s * C::A_STAGE);
#pragma unroll 1
      for (int tile = 0; tile < C::NT; ++tile) {
        const uint32_t a_hi_base = a_stage + (uint32_t)(tile * 8);
        const uint32_t a_lo_base = a_hi_base + (uint32_t)(C::CH * C::PL);
        const uint32_t d = tmem_base + (uint32_t)(s * C::TMEM_STAGE + tile * C::N1);
#pragma unroll
        for (int kb = 0; kb < C::NKB; ++kb) {
          const KBlock k = C::kblock(kb);
          const uint32_t ak = (uint32_t)k.off | ((uint32_t)k.lbo << 16);
          const uint64_t db = ((uint64_t)b_hi32 << 32) | (uint64_t)(b_lo32 + (uint32_t)(kb * 2 * C::N1));
          const uint64_t da_hi = ((uint64_t)a_hi32 << 32) | (uint64_t)(a_hi_base + ak);
          const uint64_t da_lo = ((uint64_t)a_hi32 << 32) | (uint64_t)(a_lo_base + ak);
          if (leader) {
            umma_tf32(d, da_hi, db, idesc1, kb > 0);
            umma_tf32(d, da_lo, db, idesc2, 1);
          }
        }
      }
      if (leader) {
        umma_commit(&a_empty[s]);
        umma_commit(&d_full[s]);
      }
      __syncwarp();
#ifdef PB_UMMA_PROF
      pt[2] += clock64() - t0_;
#endif
    }
#ifdef PB_UMMA_PROF
    if (blockIdx.x == 0 && lane == 0)
      printf("conv_umma MMA warp (NS=%d BWD=%d ACT=%d): total %lld | wait a_full %lld, wait d_empty %lld, issue %lld, items %d\n",
             (int)C::NS, (int)C::BWD, ACT, clock64() - pt_start, pt[0], pt[1], pt[2], it);
#endif
  } else {
    // ---- epilogue: TMEM -> registers -> (+bias) -> global ---------------------
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      mbar_wait(&d_full[s], ph);
      tc_fence_after();
      const int b = item / C::IPS, rb = item - b * C::IPS;
      // MT=128: TMEM lane = tile row.  MT=64: row i sits in lane 32*(i/16) + i%16.
      const int quad = warp & 3, half = warp >> 2;
      const int row = C::MT == 128 ? quad * 32 + lane : quad * 16 + lane;
      const bool row_ok = C::MT == 128 || lane < 16;
      const int r = rb * C::RI + (row >> 3), xin = row & 7;
      if (C::NS) {
        // stage Q for the whole item, release TMEM, then gather the fx taps
        asm volatile("bar.sync 1, 256;" ::: "memory");  // previous item's gather is done
        if (half == 0) {
#pragma unroll 1
          for (int tile = 0; tile < C::NT; ++tile) {
            float v[C::N1];
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) +
                                   (uint32_t)(s * C::TMEM_STAGE + tile * C::N1);
#pragma unroll
            for (int q = 0; q < C::N1 / 16; ++q) tmem_ld16(taddr + q * 16, v + q * 16);
            tmem_ld_wait();
            float *rec = stg + ((row >> 3) * C::W + tile * 8 + xin) * C::STG_REC;
#pragma unroll
            for (int n = 0; n < C::NOUT; ++n) rec[n] = v[n] + v[C::NP + n];
          }
        }
        tc_fence_before();
        mbar_arrive(&d_empty[s]);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int p = tid; p < C::RI * C::W; p += kEpilogue) {
          const int rl = p / C::W, c = p - rl * C::W;
          float *dst = Out + (((int64_t)b * C::OC) * C::H + rb * C::RI + rl) * C::W + c;
#pragma unroll
          for (int z = 0; z < C::OC; ++z) {
            float acc = 0.0f;
#pragma unroll
            for (int fx = 0; fx < C::F; ++fx) {
              const int xs = c - C::PAD + fx;
              if (xs >= 0 && xs < C::W) acc += stg[(rl * C::W + xs) * C::STG_REC + fx * C::OC + z];
            }
            dst[(int64_t)z * C::H * C::W] = acc;
          }
        }
        continue;
      }
      const bool writer = row_ok && !(lane & 1) && !(lane & 8);
#pragma unroll 1
      for (int tile = 0; tile < C::NT; ++tile) {
        const uint32_t taddr =
            tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(s * C::TMEM_STAGE + tile * C::N1);
        float *dst = Out + (((int64_t)b * C::OC) * C::H + r) * C::W + tile * 8 + xin;
        // the fused activation + pool epilogue belongs to the FORWARD kernel only (in the
        // input-gradient kernel ACT selects the fused producer and Out2 is its output)
        constexpr int EACT = C::BWD ? -1 : ACT;
        float *dst2 = EACT >= 0 ? Out2 + (((int64_t)b * C::OC) * (C::H / 2) + (r >> 1)) * (C::W / 2) +
                                      ((tile * 8 + xin) >> 1)
                                : nullptr;
        if (half == 0) epilogue_tile<C, EACT, 0>(taddr, bias_s, dst, dst2, row_ok, writer);
        else epilogue_tile<C, EACT, 1>(taddr, bias_s, dst, dst2, row_ok, writer);
      }
      tc_fence_before();
      mbar_arrive(&d_empty[s]);
    }
  }

  // ---- teardown ------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(2u * C::TMEM_STAGE)
                 : "memory");
  }
}

template <class C, int ACT = -1>
int launch_umma(pb_context *ctx, const float *In, const float *Wt, float *Out, int64_t batch,
                float *Out2 = nullptr, const float *Aux = nullptr) {
  auto kern = conv_umma_kernel<C, ACT>;
  static bool configured = false;
  if (!configured) {
    PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    configured = true;
  }
  const int64_t n_items = batch * C::IPS;
  const int grid = (int)std::min<int64_t>(n_items, ctx->num_sms);
  const int early_w = ctx->chain == 2;
  PB_CUDA(launch_chained(ctx, kern, dim3(grid), dim3(kThreads), C::SMEM, In, Wt, Out, Out2,
                         (int)batch, early_w, Aux));
  PB_LAUNCHED(ctx);
  return 0;
}

bool umma_enabled() {
  static const bool on = [] {
    const char *e = getenv("PB_DISABLE_UMMA");
    return !(e && e[0] == '1');
  }();
  return on;
}

}  // namespace

// Returns -1 when the shape is outside the tensor-core kernels' envelope.
int conv_evaluate_umma(pb_context *ctx, const pb_layer_desc *l, const float *I, const float *W,
                       float *O, int64_t batch) {
  {
    int rc = conv_evaluate_wt(ctx, l, I, W, O, batch);
    if (rc >= 0) return rc;
    rc = conv_evaluate_rows(ctx, l, I, W, O, batch);
    if (rc >= 0) return rc;
  }
  if (!umma_enabled() || l->stride != 1 || batch > (1 << 24)) return -1;
  if (l->filter_width != 5 || l->filter_height != 5 || l->padding != 2) return -1;
  if (l->width != l->height) return -1;
  const int w = (int)l->width, d = (int)l->depth, f = (int)l->num_filters;
  if (w == 32 && d == 3 && f == 16) return launch_umma<UCfg<3, 16, 32, 32, 5, false>>(ctx, I, W, O, batch);
  if (w == 16 && d == 16 && f == 20) return launch_umma<UCfg<16, 20, 16, 16, 5, false>>(ctx, I, W, O, batch);
  if (w == 8 && d == 20 && f == 20) return launch_umma<UCfg<20, 20, 8, 8, 5, false>>(ctx, I, W, O, batch);
  return -1;
}

template <class C>
static int launch_fused(pb_context *ctx, int act, const float *I, const float *W, float *O,
                        float *O2, int64_t batch) {
  switch (act) {
    case PB_IDENTITY: return launch_umma<C, PB_IDENTITY>(ctx, I, W, O, batch, O2);
    case PB_RELU: return launch_umma<C, PB_RELU>(ctx, I, W, O, batch, O2);
    case PB_LEAKY_RELU: return launch_umma<C, PB_LEAKY_RELU>(ctx, I, W, O, batch, O2);
    case PB_SIGMOID: return launch_umma<C, PB_SIGMOID>(ctx, I, W, O, batch, O2);
  }
  return -1;
}

int conv_act_pool_evaluate_umma(pb_context *ctx, const pb_layer_desc *l, int act,
                                const pb_layer_desc *pool, const float *I, const float *W,
                                float *O, float *O2, int64_t batch) {
  {
    int rc = conv_act_pool_evaluate_wt(ctx, l, act, pool, I, W, O, O2, batch);
    if (rc >= 0) return rc;
    rc = conv_act_pool_evaluate_rows(ctx, l, act, pool, I, W, O, O2, batch);
    if (rc >= 0) return rc;
  }
  if (!umma_enabled() || l->stride != 1 || batch > (1 << 24)) return -1;
  if (l->filter_width != 5 || l->filter_height != 5 || l->padding != 2) return -1;
  if (l->width != l->height) return -1;
  const int w = (int)l->width, d = (int)l->depth, f = (int)l->num_filters;
  if (pool->width != w || pool->height != w || pool->depth != f || pool->out_width * 2 != w ||
      pool->out_height * 2 != w)
    return -1;
  if (w == 32 && d == 3 && f == 16) return launch_fused<UCfg<3, 16, 32, 32, 5, false>>(ctx, act, I, W, O, O2, batch);
  if (w == 16 && d == 16 && f == 20) return launch_fused<UCfg<16, 20, 16, 16, 5, false>>(ctx, act, I, W, O, O2, batch);
  if (w == 8 && d == 20 && f == 20) return launch_fused<UCfg<20, 20, 8, 8, 5, false>>(ctx, act, I, W, O, O2, batch);
  return -1;
}

int conv_input_delta_umma(pb_context *ctx, const pb_layer_desc *l, const float *W, const float *G,
                          float *dI, int64_t batch) {
  {
    const int rc = conv_input_delta_wt(ctx, l, W, G, dI, batch);
    if (rc >= 0) return rc;
  }
  if (!umma_enabled() || l->stride != 1 || batch > (1 << 24)) return -1;
  if (l->filter_width != 5 || l->filter_height != 5 || l->padding != 2) return -1;
  if (l->width != l->height) return -1;
  const int w = (int)l->width, d = (int)l->depth, f = (int)l->num_filters;
  if (w == 32 && d == 3 && f == 16) return launch_umma<UCfg<16, 3, 32, 32, 5, true, true>>(ctx, G, W, dI, batch);
  if (w == 16 && d == 16 && f == 20) return launch_umma<UCfg<20, 16, 16, 16, 5, true>>(ctx, G, W, dI, batch);
  if (w == 8 && d == 20 && f == 20) return launch_umma<UCfg<20, 20, 8, 8, 5, true>>(ctx, G, W, dI, batch);
  return -1;
}

template <class C>
static int launch_fused_bwd(pb_context *ctx, int act, const float *Gp, const float *W, float *dI,
                            float *G_out, const float *O_conv, int64_t batch) {
  switch (act) {
    case PB_IDENTITY: return launch_umma<C, PB_IDENTITY>(ctx, Gp, W, dI, batch, G_out, O_conv);
    case PB_RELU: return launch_umma<C, PB_RELU>(ctx, Gp, W, dI, batch, G_out, O_conv);
    case PB_LEAKY_RELU: return launch_umma<C, PB_LEAKY_RELU>(ctx, Gp, W, dI, batch, G_out, O_conv);
    case PB_SIGMOID: return launch_umma<C, PB_SIGMOID>(ctx, Gp, W, dI, batch, G_out, O_conv);
  }
  return -1;
}

// Input gradient of a convolution -> activation -> exact 2x2 max-pool block in ONE kernel:
// Gp = dE/d(pool output); writes dI = dE/d(convolution input) and G_out = dE/d(convolution
// output) (for the weight-gradient kernel that follows).  -1 outside the envelope.
int conv_act_pool_input_delta_umma(pb_context *ctx, const pb_layer_desc *l, int act,
                                   const pb_layer_desc *pool, const float *W, const float *O_conv,
                                   const float *Gp, float *G_out, float *dI, int64_t batch) {
  {
    const int rc = conv_act_pool_input_delta_wt(ctx, l, act, pool, W, O_conv, Gp, G_out, dI, batch);
    if (rc >= 0) return rc;
  }
  if (!umma_enabled() || l->stride != 1 || batch > (1 << 24)) return -1;
  if (l->filter_width != 5 || l->filter_height != 5 || l->padding != 2) return -1;
  if (l->width != l->height) return -1;
  const int w = (int)l->width, d = (int)l->depth, f = (int)l->num_filters;
  if (pool->width != w || pool->height != w || pool->depth != f || pool->out_width * 2 != w ||
      pool->out_height * 2 != w)
    return -1;
  if ((((uintptr_t)O_conv) | ((uintptr_t)G_out)) & 7) return -1;
  if (w == 32 && d == 3 && f == 16) return launch_fused_bwd<UCfg<16, 3, 32, 32, 5, true, true>>(ctx, act, Gp, W, dI, G_out, O_conv, batch);
  if (w == 16 && d == 16 && f == 20) return launch_fused_bwd<UCfg<20, 16, 16, 16, 5, true>>(ctx, act, Gp, W, dI, G_out, O_conv, batch);
  if (w == 8 && d == 20 && f == 20) return launch_fused_bwd<UCfg<20, 20, 8, 8, 5, true>>(ctx, act, Gp, W, dI, G_out, O_conv, batch);
  return -1;
}

}  // namespace pb
