// 3xTF32 tcgen05 convolution (forward and input gradient) with the FILTERS in tensor memory,
// for the batched 5x5 stride-1 "same" convolutions of the CIFAR network
// (nnet/cifar_test.cc:293-350; semantics nnet/convolution_layer.cc:50-224).
//
//   Out[oc][y][x] = bias_oc + sum_{c,fy,fx} Wk[oc,c,fy,fx] * Ipad[c][y+fy][x+fx]
//
// conv_umma.cu maps pixels to M and the (few) output channels to N: its MMAs are bound by
// fetching the 128-row image operand from shared memory for N <= 48 columns of work.  Here the
// roles are swapped and the fx taps are moved out of the contraction:
//
//   A[(oc,fx)][(fy,c)] = Wk[oc,c,fy,fx]            5*OC <= 100 rows, TENSOR MEMORY, loaded once
//   B[pixel (y,X)][(fy,c)] = Ipad[c][y+fy][X]      shared memory, the image stored ONCE as
//                                                  [4-channel chunk][flat padded pixel][4]
//   D[(oc,fx)][(y,X)] = A * B^T                    N = TR rows x P pixels per tile
//   Out[oc][y][x] = bias_oc + sum_fx D[(oc,fx)][(y, x+fx)]
//
// B is tcgen05's K-major no-swizzle layout with the pixels as rows (8 consecutive pixels =
// 128 B, SBO) and a (fy, chunk) pair per K = 8 step (LBO = their distance): the fy shift is
// the descriptor's start address, no im2col copy exists.  With A in tensor memory nothing is
// re-fetched for it, so the three 3xTF32 products (A_hi*B_hi + A_lo*B_hi + A_hi*B_lo) go into
// ONE accumulator and the MMAs run at the tensor rate (measured with the same operand shapes
// in tools/umma_wgrad_probe.cu).  The fx sum crosses accumulator rows (= tensor-memory
// lanes): the five rows of an output channel sit in ADJACENT lanes of one warp (six channels
// per 32-lane quadrant, two lanes idle) and share the channel's image row - lane fl takes the
// outputs x = fl (mod 5).  For one accumulator column the five lanes then need five different
// rows: ONE indexed shuffle per column, every lane receiving a term of one of its outputs (no
// staging buffer, no dynamic register index; tcgen05.ld cannot read at a lane offset -
// tools/tmem_ld_probe.cu - so the exchange has to be a shuffle).
//
// Input gradient = the same kernel on G with flipped / channel-transposed filters and no
// bias (with pad == filter/2 the reference's centre guard, convolution_layer.cc:209, is the
// identity); its fused variant forms dE/d(conv output) from the pooled gradient and the stored
// convolution output inside the producers, exactly like conv_umma.cu's.
//
// Warp roles (544 threads, one persistent CTA per SM, one item = one sample):
//   warps 0-7   epilogue (tensor-memory lane quadrant = warp & 3; the two warps of a quadrant
//               alternate over the row pairs): TMEM -> registers -> shuffle sum -> bias /
//               activation / 2x2 pool -> global; warps 0-3 also load the filters at start
//   warp  8     MMA issuer, tensor-memory allocator
//   warps 9-16  producers: global CHW fp32 -> hi/lo -> shared memory (two stages)
#include "common.cuh"
#include "kernels.cuh"
#include "umma.cuh"

namespace pb {
namespace {

using namespace umma;

constexpr int kEpi = 256, kMmaWarp = kEpi / 32, kProd = 256;
constexpr int kPerQuad = 6;  // output channels per lane quadrant (5 lanes each, 2 lanes idle)
constexpr int kThreads = kEpi + 32 + kProd;

struct KPair {
  int first, second;  // chunk ids (fy*CH + c4); second == -1: zero-weight dummy
  int off;            // B start offset of `first`, pixels (16-byte units)
  int lbo;            // distance first -> second, pixels
};

template <int IC_, int OC_, int W_, int TR_, bool BWD_>
struct TCfg {
  static constexpr int IC = IC_, OC = OC_, W = W_, H = W_, TR = TR_, F = 5, PAD = 2;
  static constexpr bool BWD = BWD_;
  static constexpr int P = W + F - 1;              // pixel pitch of a staged row: the operand rows
                                                   // are FLAT pixels, any pitch is contiguous
  static constexpr int SR = H + F - 1;             // staged rows: the whole zero-padded sample
  static constexpr int PL = ((SR * P + 7) / 8) * 8; // pixels per chunk plane
  static constexpr int CH = (IC + 3) / 4;          // 4-channel chunks
  static constexpr int NCHUNK = F * CH;            // (fy, chunk)
  static constexpr int KS = (NCHUNK + 1) / 2;      // K = 8 steps
  // A / D row of (oc, fx): lane 32*(oc/6) + 5*(oc%6) + fx
  static constexpr int N = TR * P;                 // accumulator columns of a tile
  static constexpr int TILES = H / TR;
  static constexpr int A_COLS = KS * 8;
  static constexpr int A_COL0 = 2 * N;             // two accumulator stages first
  static constexpr int TMEM_NEED = A_COL0 + 2 * A_COLS;
  static constexpr int TMEM_COLS = TMEM_NEED <= 256 ? 256 : 512;
  static constexpr int A_STAGE = 2 * CH * PL;      // float4 per stage: hi planes, lo planes
  static constexpr int K1 = F * F * (BWD ? OC : IC) + 1;  // reference filter-block stride
  static constexpr int NWT = (BWD ? IC : OC) * K1;  // floats of the layer's filter blocks
  static constexpr size_t SMEM = 16 * (size_t)(2 * A_STAGE) + 32 * 4 + 8 * 8 + 16;
  static_assert(OC <= 4 * kPerQuad && N % 16 == 0 && N <= 256 && H % TR == 0 && TR % 2 == 0, "tile shape");
  static_assert(NWT <= 8 * A_STAGE, "filter scratch");
  static_assert(TMEM_NEED <= 512, "tensor-memory budget");
  static_assert(SMEM <= 227 * 1024, "shared memory");

  static constexpr int chunk_off(int j) { return (j % CH) * PL + (j / CH) * P; }
  static constexpr KPair kpair(int ks) {
    const int j0 = 2 * ks, j1 = 2 * ks + 1;
    if (j1 >= NCHUNK) return KPair{j0, -1, chunk_off(j0), 1};
    const int o0 = chunk_off(j0), o1 = chunk_off(j1);
    return o1 > o0 ? KPair{j0, j1, o0, o1 - o0} : KPair{j1, j0, o1, o0 - o1};
  }
};

__device__ __forceinline__ void tmem_st8u(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ACT >= 0, forward: the epilogue additionally writes Out2 = 2x2 max-pool of act(conv output).
// ACT >= 0, BWD: fused producer - In = dE/d(pool output), Aux = stored convolution output of
// the layer, Out2 receives dE/d(conv output) (see conv_umma.cu).
template <class C, int ACT>
__global__ void __launch_bounds__(kThreads, 1)
conv_wt_kernel(const float *__restrict__ In, const float *__restrict__ Wt, float *__restrict__ Out,
               float *__restrict__ Out2, int batch, int early_w, const float *__restrict__ Aux) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float4 *a_s = reinterpret_cast<float4 *>(smem_raw);                   // [2][A_STAGE] image
  float *bias_s = reinterpret_cast<float *>(a_s + 2 * C::A_STAGE);      // [32]
  uint64_t *bars = reinterpret_cast<uint64_t *>(bias_s + 32);
  uint64_t *a_full = bars, *a_empty = bars + 2, *d_full = bars + 4, *d_empty = bars + 6;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef PB_WT_PROF
  long long pt[6] = {0, 0, 0, 0, 0, 0};
  const long long pt_start = clock64();
#define PROF_T0 const long long pt0_ = clock64();
#define PROF_ADD(i) pt[i] += clock64() - pt0_;
#else
#define PROF_T0
#define PROF_ADD(i)
#endif

  // ---- setup ------------------------------------------------------------------------
  // Chained launch (common.cuh): when the launch before this one does not write this layer's
  // filters (early_w), everything up to the first image load runs ahead of the dependency wait.
  if (!early_w) chain_wait();
  float *w_raw = reinterpret_cast<float *>(a_s);   // filter blocks, reference layout (scratch)
  {
    // every load of the thread in flight before the first store: one L2 round trip, not one per
    // unrolled group (this prologue is ~4 us of a 20-45 us kernel)
    constexpr int WL = (C::NWT + kThreads - 1) / kThreads;
    float wv[WL];
#pragma unroll
    for (int k = 0; k < WL; ++k) {
      const int i = tid + k * kThreads;
      wv[k] = i < C::NWT ? __ldcg(Wt + i) : 0.0f;
    }
#pragma unroll
    for (int k = 0; k < WL; ++k) {
      const int i = tid + k * kThreads;
      if (i < C::NWT) w_raw[i] = wv[k];
    }
  }
  if (tid == 0) {
    mbar_init(&a_full[0], kProd / 32); mbar_init(&a_full[1], kProd / 32);
    mbar_init(&a_empty[0], 1);  mbar_init(&a_empty[1], 1);
    mbar_init(&d_full[0], 1);   mbar_init(&d_full[1], 1);
    mbar_init(&d_empty[0], kEpi / 32); mbar_init(&d_empty[1], kEpi / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_ptr)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp < 4) {
    // filters -> tensor memory: lane = (oc, fx), column = K index (fy, chunk, channel)
    const int g = lane / C::F, fx = lane - g * C::F, oc = warp * kPerQuad + g;
    const bool live = g < kPerQuad && oc < C::OC;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
#pragma unroll
    for (int ks = 0; ks < C::KS; ++ks) {
      const KPair kp = C::kpair(ks);
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = h ? kp.second : kp.first;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float x = 0.0f;
          if (live && j >= 0) {
            const int fy = j / C::CH, ci = 4 * (j % C::CH) + i;
            if (ci < C::IC)
              x = C::BWD ? w_raw[ci * C::K1 + oc * C::F * C::F + (C::F * C::F - 1 - (fy * C::F + fx))]
                         : w_raw[oc * C::K1 + ci * C::F * C::F + fy * C::F + fx];
          }
          const float hx = tf32_hi(x);
          hi[h * 4 + i] = __float_as_uint(hx);
          lo[h * 4 + i] = __float_as_uint(x - hx);
        }
      }
      tmem_st8u(tmem_base + lane_base + C::A_COL0 + ks * 8, hi);
      tmem_st8u(tmem_base + lane_base + C::A_COL0 + C::A_COLS + ks * 8, lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    if (tid < 32) bias_s[tid] = (!C::BWD && tid < C::OC) ? w_raw[tid * C::K1 + C::K1 - 1] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // the image stages: every pad stays zero for the whole kernel, interiors are rewritten
  for (int i = tid; i < 2 * C::A_STAGE; i += kThreads) a_s[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_async_proxy();
  __syncthreads();
  if (early_w) chain_wait();
  chain_release();
#ifdef PB_WT_PROF
  const long long pt_setup = clock64() - pt_start;
#endif

  if (warp > kMmaWarp) {
    // ---- producers: global CHW -> [chunk][flat padded pixel][4] hi / lo ------------------
    const int ptid = tid - (kEpi + 32);
    int it = 0;
    if constexpr (C::BWD && ACT >= 0) {
      // one thread = one 2x2 pool window x one 4-channel chunk (see conv_umma.cu)
      constexpr int ACTV = ACT >= 0 ? ACT : 0;
      constexpr int WW = C::W / 2, WR = C::H / 2;
      constexpr int WTOTAL = C::CH * WR * WW;
      constexpr int WITERS = (WTOTAL + kProd - 1) / kProd;
      for (int item = blockIdx.x; item < batch; item += gridDim.x, ++it) {
        const int s = it & 1, ph = (it >> 1) & 1;
        float4 *hi = a_s + s * C::A_STAGE;
        float4 *lo = hi + C::CH * C::PL;
        const float *o_src = Aux + (int64_t)item * C::IC * C::H * C::W;
        const float *g_src = In + (int64_t)item * C::IC * (C::H / 2) * (C::W / 2);
        float *g_dst = Out2 + (int64_t)item * C::IC * C::H * C::W;
        float2 o0[WITERS][4], o1[WITERS][4];
        float gp[WITERS][4];
#pragma unroll
        for (int k = 0; k < WITERS; ++k) {
          const int e = ptid + k * kProd;
          const int xw = e % WW, t = e / WW;
          const int wr = t % WR, c = t / WR;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int ic = 4 * c + i;
            const bool ok = e < WTOTAL && ic < C::IC;
            const float *po = o_src + ((int64_t)ic * C::H + 2 * wr) * C::W + 2 * xw;
            o0[k][i] = ok ? __ldg(reinterpret_cast<const float2 *>(po)) : make_float2(0.f, 0.f);
            o1[k][i] = ok ? __ldg(reinterpret_cast<const float2 *>(po + C::W)) : make_float2(0.f, 0.f);
            gp[k][i] = ok ? __ldg(g_src + ((int64_t)ic * WR + wr) * WW + xw) : 0.0f;
          }
        }
        mbar_wait(&a_empty[s], ph ^ 1);
#pragma unroll
        for (int k = 0; k < WITERS; ++k) {
          const int e = ptid + k * kProd;
          if (e < WTOTAL) {
            const int xw = e % WW, t = e / WW;
            const int wr = t % WR, c = t / WR;
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
              if (ic < C::IC) {
                float *pg = g_dst + ((int64_t)ic * C::H + 2 * wr) * C::W + 2 * xw;
                *reinterpret_cast<float2 *>(pg) = make_float2(d[0][i], d[1][i]);
                *reinterpret_cast<float2 *>(pg + C::W) = make_float2(d[2][i], d[3][i]);
              }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float h0 = tf32_hi(d[q][0]), h1 = tf32_hi(d[q][1]), h2 = tf32_hi(d[q][2]),
                          h3 = tf32_hi(d[q][3]);
              const int pix = c * C::PL + (2 * wr + (q >> 1) + C::PAD) * C::P + 2 * xw + (q & 1) + C::PAD;
              hi[pix] = make_float4(h0, h1, h2, h3);
              lo[pix] = make_float4(d[q][0] - h0, d[q][1] - h1, d[q][2] - h2, d[q][3] - h3);
            }
          }
        }
        fence_async_proxy();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[s]);
      }
    } else {
      constexpr int TOTAL = C::CH * C::H * C::W;
      constexpr int ITERS = (TOTAL + kProd - 1) / kProd;
      for (int item = blockIdx.x; item < batch; item += gridDim.x, ++it) {
        const int s = it & 1, ph = (it >> 1) & 1;
        float4 *hi = a_s + s * C::A_STAGE;
        float4 *lo = hi + C::CH * C::PL;
        const float *src = In + (int64_t)item * C::IC * C::H * C::W;
        float v[ITERS][4];
#pragma unroll
        for (int k = 0; k < ITERS; ++k) {
          const int e = ptid + k * kProd;
          const int x = e % C::W, t = e / C::W;
          const int y = t % C::H, c = t / C::H;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int ic = 4 * c + i;
            v[k][i] = (e < TOTAL && ic < C::IC) ? __ldg(src + (ic * C::H + y) * C::W + x) : 0.0f;
          }
        }
        { PROF_T0
        mbar_wait(&a_empty[s], ph ^ 1);
        PROF_ADD(0) }
#pragma unroll
        for (int k = 0; k < ITERS; ++k) {
          const int e = ptid + k * kProd;
          if (e < TOTAL) {
            const int x = e % C::W, t = e / C::W;
            const int y = t % C::H, c = t / C::H;
            const float h0 = tf32_hi(v[k][0]), h1 = tf32_hi(v[k][1]), h2 = tf32_hi(v[k][2]),
                        h3 = tf32_hi(v[k][3]);
            const int pix = c * C::PL + (y + C::PAD) * C::P + x + C::PAD;
            hi[pix] = make_float4(h0, h1, h2, h3);
            lo[pix] = make_float4(v[k][0] - h0, v[k][1] - h1, v[k][2] - h2, v[k][3] - h3);
          }
        }
        fence_async_proxy();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[s]);
      }
#ifdef PB_WT_PROF
      if (blockIdx.x == 0 && ptid == 0) printf("producer: total %lld | wait a_empty %lld\n", clock64() - pt_start, pt[0]);
#endif
    }
  } else if (warp == kMmaWarp) {
    // ---- MMA issuer -------------------------------------------------------------------
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc(128, C::N);
    const uint32_t a_hi = tmem_base + C::A_COL0, a_lo = a_hi + C::A_COLS;
    int it = 0, tc = 0;
    for (int item = blockIdx.x; item < batch; item += gridDim.x, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      { PROF_T0
      mbar_wait(&a_full[s], ph);
      PROF_ADD(0) }
      tc_fence_after();
      const uint32_t b_stage = (smem_u32(a_s) >> 4) + (uint32_t)(s * C::A_STAGE);
#pragma unroll 1
      for (int tile = 0; tile < C::TILES; ++tile, ++tc) {
        const int ds = tc & 1, dph = (tc >> 1) & 1;
        { PROF_T0
        mbar_wait(&d_empty[ds], dph ^ 1);
        PROF_ADD(1) }
        tc_fence_after();
        PROF_T0
        const uint32_t d = tmem_base + (uint32_t)(ds * C::N);
        const uint32_t b_tile = b_stage + (uint32_t)(tile * C::TR * C::P);
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < C::KS; ++ks) {
            const KPair kp = C::kpair(ks);
            // start | LBO | SBO = 8 pixels (128 B) | version 1
            const uint64_t bh = ((uint64_t)(8u | (1u << 14)) << 32) | ((uint64_t)kp.lbo << 16) |
                                (uint64_t)((b_tile + (uint32_t)kp.off) & 0x3FFFu);
            const uint64_t bl = ((uint64_t)(8u | (1u << 14)) << 32) | ((uint64_t)kp.lbo << 16) |
                                (uint64_t)((b_tile + (uint32_t)(kp.off + C::CH * C::PL)) & 0x3FFFu);
            umma_tf32_ts(d, a_hi + ks * 8, bh, idesc, ks > 0);
            umma_tf32_ts(d, a_lo + ks * 8, bh, idesc, 1);
            umma_tf32_ts(d, a_hi + ks * 8, bl, idesc, 1);
          }
          umma_commit(&d_full[ds]);
        }
        __syncwarp();
        PROF_ADD(2)
      }
      if (leader) umma_commit(&a_empty[s]);
      __syncwarp();
    }
#ifdef PB_WT_PROF
    if (blockIdx.x == 0 && lane == 0)
      printf("MMA warp: total %lld setup %lld | wait a_full %lld, wait d_empty %lld, issue %lld\n",
             clock64() - pt_start, pt_setup, pt[0], pt[1], pt[2]);
#endif
  } else {
    // ---- epilogue: TMEM -> registers -> shuffle sum over fx -> global -----------------------
    const int quad = warp & 3, sub = warp >> 2;
    const int g = lane / C::F, fl = lane - g * C::F, oc = quad * kPerQuad + g;
    const bool live = g < kPerQuad && oc < C::OC;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const float bias = live ? bias_s[oc] : 0.0f;
    // The five lanes of a channel SHARE its image row: lane fl ends up with the outputs
    // x = fl, fl + 5, fl + 10, ... (slot s: x = fl + 5s).  Column c = 5q + r of the accumulator
    // row fx belongs to output x = c - fx; for a fixed c the five lanes fetch it from the five
    // DIFFERENT rows fx = (r - fl) mod 5 - one indexed shuffle per column in which every lane
    // receives a value it needs (the fx-tree of the first version moved 50 registers per row for
    // one receiving lane in five; the shuffle pipe was the kernel's bound) - and add it to slot q
    // (r >= fl) or q - 1.
    constexpr int NC = C::W + C::F - 1;            // accumulator columns of one image row
    constexpr int LD = ((NC + 7) / 8) * 8;
    constexpr int NS = (NC + C::F - 1) / C::F;     // slots per lane
    static_assert(2 * C::N + LD - C::P <= C::TMEM_COLS, "row window stays inside the allocation");
    int src[C::F];
#pragma unroll
    for (int r = 0; r < C::F; ++r) src[r] = g < kPerQuad ? g * C::F + (r - fl + C::F) % C::F : lane;
    const int pair_src = g < kPerQuad ? (fl < C::F - 1 ? lane + 1 : lane - (C::F - 1)) : lane;
    auto row_fold = [&](uint32_t taddr, float (&acc)[NS]) {
      float v[LD];
#pragma unroll
      for (int c0 = 0; c0 < LD; c0 += 8) tmem_ld8(taddr + c0, &v[c0]);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < NS; ++q) acc[q] = bias;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        constexpr int F = C::F;
        const int q = c / F, r = c % F;
        const float t = __shfl_sync(0xffffffffu, v[c], src[r]);
        if (r >= fl) acc[q] += t;
        else if (q > 0) acc[q - 1] += t;
      }
    };
    int tc = 0;
    for (int item = blockIdx.x; item < batch; item += gridDim.x) {
#pragma unroll 1
      for (int tile = 0; tile < C::TILES; ++tile, ++tc) {
        const int ds = tc & 1, dph = (tc >> 1) & 1;
        { PROF_T0
        mbar_wait(&d_full[ds], dph);
        PROF_ADD(0) }
        tc_fence_after();
        PROF_T0
        const uint32_t d_tile = tmem_base + lane_base + (uint32_t)(ds * C::N);
#pragma unroll 1
        for (int rp = sub; rp < C::TR / 2; rp += 2) {   // row pairs of the tile, alternating warps
          const int y = tile * C::TR + 2 * rp;
          float o0[NS], o1[NS];
          row_fold(d_tile + (uint32_t)((2 * rp) * C::P), o0);
          row_fold(d_tile + (uint32_t)((2 * rp + 1) * C::P), o1);
          float *dst = Out + (((int64_t)item * C::OC + oc) * C::H + y) * C::W + fl;
#pragma unroll
          for (int q = 0; q < NS; ++q) {
            if (live && fl + C::F * q < C::W) {
              dst[C::F * q] = o0[q];
              dst[C::W + C::F * q] = o1[q];
            }
          }
          if constexpr (!C::BWD && ACT >= 0) {
            // 2x2 pool: vertical maximum in the lane, the odd column's from the lane that owns
            // it (x + 1: the next lane's slot q, or - from the last lane - the first lane's q + 1)
            float m[NS], nb[NS];
#pragma unroll
            for (int q = 0; q < NS; ++q) m[q] = fmaxf(o0[q], o1[q]);
#pragma unroll
            for (int q = 0; q < NS; ++q) nb[q] = __shfl_sync(0xffffffffu, m[q], pair_src);
            float *dst2 = Out2 + (((int64_t)item * C::OC + oc) * (C::H / 2) + (y >> 1)) * (C::W / 2);
#pragma unroll
            for (int q = 0; q < NS; ++q) {
              const int x = fl + C::F * q;
              const float up = q + 1 < NS ? nb[q + 1] : 0.0f;
              const float other = fl < C::F - 1 ? nb[q] : up;
              if (live && x < C::W && !(x & 1))
                dst2[x >> 1] = act_fwd<(ACT >= 0 ? ACT : 0)>(fmaxf(m[q], other));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&d_empty[ds]);
        PROF_ADD(1)
      }
    }
#ifdef PB_WT_PROF
    if (blockIdx.x == 0 && tid == 0)
      printf("epilogue: total %lld | wait d_full %lld, rows %lld\n", clock64() - pt_start, pt[0], pt[1]);
#endif
  }

  // ---- teardown -------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

template <class C, int ACT = -1>
int launch_wt(pb_context *ctx, const float *In, const float *Wt, float *Out, int64_t batch,
              float *Out2 = nullptr, const float *Aux = nullptr) {
  auto kern = conv_wt_kernel<C, ACT>;
  static bool configured = false;
  if (!configured) {
    PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    configured = true;
  }
  const int grid = (int)std::min<int64_t>(batch, ctx->num_sms);
  const int early_w = ctx->chain == 2;
  PB_CUDA(launch_chained(ctx, kern, dim3(grid), dim3(kThreads), C::SMEM, In, Wt, Out, Out2,
                         (int)batch, early_w, Aux));
  PB_LAUNCHED(ctx);
  return 0;
}

template <class C>
int launch_wt_act(pb_context *ctx, int act, const float *In, const float *Wt, float *Out,
                  int64_t batch, float *Out2, const float *Aux) {
  switch (act) {
    case PB_IDENTITY: return launch_wt<C, PB_IDENTITY>(ctx, In, Wt, Out, batch, Out2, Aux);
    case PB_RELU: return launch_wt<C, PB_RELU>(ctx, In, Wt, Out, batch, Out2, Aux);
    case PB_LEAKY_RELU: return launch_wt<C, PB_LEAKY_RELU>(ctx, In, Wt, Out, batch, Out2, Aux);
    case PB_SIGMOID: return launch_wt<C, PB_SIGMOID>(ctx, In, Wt, Out, batch, Out2, Aux);
  }
  return -1;
}

bool wt_enabled() {
  static const bool on = [] {
    const char *e = getenv("PB_DISABLE_UMMA"), *f = getenv("PB_CONV_WT");
    return !(e && e[0] == '1') && !(f && f[0] == '0');
  }();
  return on;
}

bool wt_shape(const pb_layer_desc *l, int64_t batch) {
  return wt_enabled() && l->stride == 1 && batch >= 64 && batch <= (1 << 24) && l->filter_width == 5 &&
         l->filter_height == 5 && l->padding == 2 && l->width == l->height;
}

bool pool_matches(const pb_layer_desc *l, const pb_layer_desc *pool) {
  const int64_t w = l->width;
  return pool->width == w && pool->height == w && pool->depth == l->num_filters &&
         pool->out_width * 2 == w && pool->out_height * 2 == w;
}

}  // namespace

// Each returns -1 outside the envelope (conv_umma.cu's kernels then run).
int conv_evaluate_wt(pb_context *ctx, const pb_layer_desc *l, const float *I, const float *W,
                     float *O, int64_t batch) {
  if (!wt_shape(l, batch) || (((uintptr_t)O) & 15)) return -1;
  const int w = (int)l->width, d = (int)l->depth, f = (int)l->num_filters;
  // (the RGB layer's forward pass - K = 15, 80 accumulator rows to fold per 16 outputs, only 648
  // MMA cycles per tile - is bound by the epilogue here: 93 us with the first shuffle fold, 74 us
  // with a quadrant-per-fx row order folded through a shared-memory tile, 50 us (75 us with the
  // fused activation + pool) with the one-shuffle-per-column fold, against conv_umma.cu's 43 us,
  // which keeps it)
  if (w == 16 && d == 16 && f == 20) return launch_wt<TCfg<16, 20, 16, 4, false>>(ctx, I, W, O, batch);
  if (w == 8 && d == 20 && f == 20) return launch_wt<TCfg<20, 20, 8, 4, false>>(ctx, I, W, O, batch);
  return -1;
}

int conv_act_pool_evaluate_wt(pb_context *ctx, const pb_layer_desc *l, int act,
                              const pb_layer_desc *pool, const float *I, const float *W, float *O,
                              float *O2, int64_t batch) {
  if (!wt_shape(l, batch) || !pool_matches(l, pool)) return -1;
  if ((((uintptr_t)O) | ((uintptr_t)O2)) & 15) return -1;
  const int w = (int)l->width, d = (int)l->depth, f = (int)l->num_filters;
  if (w == 16 && d == 16 && f == 20) return launch_wt_act<TCfg<16, 20, 16, 4, false>>(ctx, act, I, W, O, batch, O2, nullptr);
  if (w == 8 && d == 20 && f == 20) return launch_wt_act<TCfg<20, 20, 8, 4, false>>(ctx, act, I, W, O, batch, O2, nullptr);
  return -1;
}

// The RGB layer's input gradient (3 output channels = 15 of 128 accumulator rows) stays with
// conv_umma.cu's N-stacked kernel.
int conv_input_delta_wt(pb_context *ctx, const pb_layer_desc *l, const float *W, const float *G,
                        float *dI, int64_t batch) {
  if (!wt_shape(l, batch) || (((uintptr_t)dI) & 15)) return -1;
  const int w = (int)l->width, d = (int)l->depth, f = (int)l->num_filters;
  if (w == 16 && d == 16 && f == 20) return launch_wt<TCfg<20, 16, 16, 4, true>>(ctx, G, W, dI, batch);
  if (w == 8 && d == 20 && f == 20) return launch_wt<TCfg<20, 20, 8, 4, true>>(ctx, G, W, dI, batch);
  return -1;
}

int conv_act_pool_input_delta_wt(pb_context *ctx, const pb_layer_desc *l, int act,
                                 const pb_layer_desc *pool, const float *W, const float *O_conv,
                                 const float *Gp, float *G_out, float *dI, int64_t batch) {
  if (!wt_shape(l, batch) || !pool_matches(l, pool)) return -1;
  if (((((uintptr_t)O_conv) | ((uintptr_t)G_out)) & 7) || (((uintptr_t)dI) & 15)) return -1;
  const int w = (int)l->width, d = (int)l->depth, f = (int)l->num_filters;
  if (w == 16 && d == 16 && f == 20) return launch_wt_act<TCfg<20, 16, 16, 4, true>>(ctx, act, Gp, W, dI, batch, G_out, O_conv);
  if (w == 8 && d == 20 && f == 20) return launch_wt_act<TCfg<20, 20, 8, 4, true>>(ctx, act, Gp, W, dI, batch, G_out, O_conv);
  return -1;
}

}  // namespace pb
