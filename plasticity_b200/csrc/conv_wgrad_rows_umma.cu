// 3xTF32 tcgen05 weight gradient (+ SGD update) for the network's FIRST convolution (3 input
// channels, 16 filters, 32x32 images, 5x5 stride-1 "same": nnet/cifar_test.cc:293-300; semantics
// nnet/convolution_layer.cc:226-270, update back_prop.kernel.cl:28-44), fused with the backward
// of the activation and the 2x2 max-pool behind it:
//
//   g[oc][c][fy][fx] = sum_{b,y,x} G[b,oc,y,x] * Ipad[b,c,y+fy,x+fx]        bias: sum G[b,oc,.,.]
//   G = dE/d(conv output), formed here from the stored conv output O and the pooled gradient Gp
//       (act_pool_bwd_kernel, elementwise.cu, value for value) - it never travels through HBM.
//
// 1200 outputs against a contraction over 10^6 (sample, pixel) pairs: with (fx, oc) x (fy, c) as
// the free indices (conv_wgrad_umma.cu) the MMA has N = 15 columns, below the ~22 cycle issue
// floor of a tcgen05.mma.  Here the image rows are cut into BANDS of four output rows and the
// row offset becomes a free index on BOTH sides:
//
//   A[(c,Yl,fx)][x] = Ipad[c][4*band + Yl][x + fx]   120 rows (Yl < 8) + a row of ones -> TENSOR MEMORY
//   B[(oc,yl)][x]   = G[oc][4*band + yl][x]          64 rows (yl < 4), K = 32 = one image row
//   D[(c,Yl,fx)][(oc,yl)] += A * B^T                 ONE accumulator for every band of every sample
//   g[oc][c][fy][fx] = sum_yl D[(c, yl+fy, fx)][(oc,yl)]          bias_oc = sum_yl D[ones][(oc,yl)]
//
// 5 of the 8 row offsets Yl - yl are taps, the others are discarded: 62 % of the MMA work is
// useful, N = 64 runs at the tensor rate (32 cycles), twelve MMAs per band, 3072 cycles per
// sample - a quarter of what the mma.sync kernel of conv_wgrad_mma.cu needs.
//
// Warp roles (544 threads, one persistent CTA per SM, items = samples, tiles = bands):
//   warps 0-7    A producers: lane quadrant = warp & 3, the two warps of a quadrant alternate
//                over the bands: 32 conflict-free shared-memory loads per lane (its image row at
//                its fx shift), hi/lo split, tcgen05.st; together they stage the samples
//                (zero-padded, pitches chosen so that lane r reads bank r + x)
//   warp  8      MMA issuer, tensor-memory allocator
//   warps 9-16   gradient formers: O, Gp -> G -> hi / lo -> shared memory, K-major no-swizzle
//                (8 rows x 16 B core matrices; the 4-wide K chunks are 1040 B apart so that the
//                eight lanes of a row store to different banks); four stages
// Epilogue: warps 0-3 move the accumulator to shared memory, all threads fold the row offsets and
// write ONE partial per CTA in the reference filter layout; reduce_partials_update (conv_tiled.cu)
// adds the partials in a fixed order and applies the update (deterministic).
#include <cuda_pipeline.h>

#include "common.cuh"
#include "kernels.cuh"
#include "umma.cuh"

namespace pb {
namespace {

using namespace umma;

struct WCfgR {
  static constexpr int IC = 3, OC = 16, W = 32, H = 32, F = 5, PAD = 2;
  static constexpr int BR = 4;                      // output rows per band
  static constexpr int BANDS = H / BR;
  static constexpr int YL = BR + F - 1;             // input rows per band
  static constexpr int ROWS = IC * YL * F;          // live A rows; row ROWS is the row of ones
  static constexpr int N = OC * BR;                 // (oc, yl)
  static constexpr int KS = W / 8;                  // K = 8 steps per band
  static constexpr int AQ = 2;                      // A-producer warps per lane quadrant
  static constexpr int A_WARPS = 4 * AQ, A_THREADS = 128 * AQ;
  static constexpr int MMA_WARP = A_WARPS;
  static constexpr int G_WARPS = 8, G_THREADS = 32 * G_WARPS, G_WARP0 = A_WARPS + 1;
  static constexpr int THREADS = (A_WARPS + 1 + G_WARPS) * 32;
  // staged sample: [c][36 rows][36 columns], row pitch = 5, plane pitch = 8 (mod 32)
  static constexpr int SP = 37, SR = H + F - 1;
  static constexpr int PLANE = SR * SP + ((8 - (SR * SP) % 32 + 32) % 32);
  static constexpr int IMG = ((IC * PLANE + 3) / 4) * 4;
  // gradient operand: float(n, k) = (k/4)*B_LBO + (n/8)*32 + (n%8)*4 + k%4
  static constexpr int B_LBO = N * 4 + 4;           // floats; + 16 B: bank spread of the stores
  static constexpr int B_SET = (W / 4) * B_LBO;     // hi (or lo) set of one band
  static constexpr int B_STAGE = 2 * B_SET;
  static constexpr int NB = 4;                      // gradient stages
  static constexpr int SLOT_COLS = KS * 16;         // [k step][hi 8 | lo 8]
  static constexpr int NSLOT = 4;
  static constexpr int A_COL0 = N;                  // the accumulator first
  static constexpr int TMEM_COLS = 512;
  static constexpr int K1 = F * F * IC + 1;         // reference filter-block stride
  static constexpr int NW = OC * K1;                // floats per partial
  static constexpr int DP = N + 1;                  // accumulator row pitch in shared memory
  static constexpr int PF = 4;                      // bands in flight per forming thread
  static constexpr int RAW4 = (PF + 2) * G_THREADS * 2 + (PF + 2) * G_THREADS / 2;  // float4 units
  static constexpr size_t SMEM = sizeof(float) * (size_t)(2 * IMG + NB * B_STAGE + 4 * RAW4) + 8 * (2 * NSLOT + 2 * NB + 1) + 16;
  static_assert(ROWS + 1 <= 128 && N % 16 == 0 && A_COL0 + NSLOT * SLOT_COLS <= TMEM_COLS, "tile shape");
  static_assert(SP % 32 == 5 && PLANE % 32 == 8, "conflict-free pitches");
  static_assert(OC * (BR / 2) * (W / 4) == G_THREADS, "one forming item per thread and band");
  static_assert(128 * DP <= NB * B_STAGE, "accumulator staging fits in the gradient stages");
  static_assert(SMEM <= 227 * 1024, "shared memory");
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&h)[8], const uint32_t (&l)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]),
      "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7])
      : "memory");
}
__device__ __forceinline__ float tf32_trunc(float x) {
  return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
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

template <int ACT>
__global__ void __launch_bounds__(WCfgR::THREADS, 1)
conv_wgrad_rows_kernel(const float *__restrict__ In, const float *__restrict__ O,
                       const float *__restrict__ Gp, float *__restrict__ partial, int batch) {
  using C = WCfgR;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float *img_s = reinterpret_cast<float *>(smem_raw);                    // [2][IMG]
  float *b_s = img_s + 2 * C::IMG;                                        // [NB][hi set | lo set]
  float4 *raw_s = reinterpret_cast<float4 *>(b_s + C::NB * C::B_STAGE);   // formers' cp.async ring
  uint64_t *bars = reinterpret_cast<uint64_t *>(raw_s + C::RAW4);
  uint64_t *a_full = bars, *a_empty = bars + C::NSLOT, *b_full = bars + 2 * C::NSLOT,
           *b_empty = bars + 2 * C::NSLOT + C::NB, *d_full = bars + 2 * C::NSLOT + 2 * C::NB;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(d_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef PB_WROWS_PROF
  long long pt_[4] = {0, 0, 0, 0};
  const long long pt_start = clock64();
#define PROF_T0 const long long pt0_ = clock64();
#define PROF_ADD(i) pt_[i] += clock64() - pt0_;
#else
#define PROF_T0
#define PROF_ADD(i)
#endif

  // ---- one-time setup (runs ahead of the dependency wait of a chained launch) ------
  for (int i = tid; i < 2 * C::IMG; i += C::THREADS) img_s[i] = 0.0f;   // pads stay zero
  if (tid == 0) {
    for (int s = 0; s < C::NSLOT; ++s) { mbar_init(&a_full[s], 4); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < C::NB; ++s) { mbar_init(&b_full[s], C::G_WARPS); mbar_init(&b_empty[s], 1); }
    mbar_init(d_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == C::MMA_WARP) {
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
  chain_wait();
  chain_release();

  if (warp < C::A_WARPS) {
    // ---- sample staging + A producers -------------------------------------------------
    const int quad = warp & 3, sub = warp >> 2;
    const int r = quad * 32 + lane;                       // tensor-memory lane = A row = D row
    const int c = r / (C::YL * C::F), yl = (r / C::F) % C::YL, fx = r % C::F;
    const bool live = r < C::ROWS;
    const float fill = r == C::ROWS ? 1.0f : 0.0f;        // the row of ones sums G: the bias gradient
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    float4 px[3];
    auto load_sample = [&](int item) {
      const float4 *src = reinterpret_cast<const float4 *>(In + (int64_t)item * C::IC * C::H * C::W);
#pragma unroll
      for (int k = 0; k < 3; ++k) px[k] = __ldg(src + tid + k * C::A_THREADS);
    };
    auto store_sample = [&](int st) {
      float *dst = img_s + st * C::IMG;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int e = tid + k * C::A_THREADS;
        const int x4 = e % (C::W / 4), rr = e / (C::W / 4);
        const int y = rr % C::H, cc = rr / C::H;
        float *d = dst + cc * C::PLANE + (y + C::PAD) * C::SP + C::PAD + 4 * x4;
        d[0] = px[k].x; d[1] = px[k].y; d[2] = px[k].z; d[3] = px[k].w;
      }
    };
    static_assert(C::IC * C::H * C::W / 4 == 3 * C::A_THREADS, "three pieces per staging thread");
    int it = 0;
    if ((int)blockIdx.x < batch) { load_sample(blockIdx.x); store_sample(0); }
    asm volatile("bar.sync 1, %0;" ::"n"(C::A_THREADS) : "memory");
    for (int item = blockIdx.x; item < batch; item += gridDim.x, ++it) {
      const int st = it & 1;
      const int next = item + gridDim.x;
      if (next < batch) load_sample(next);
      // lane r, K index x: staged row 4*band + Yl, column x + fx
      const float *row0 = img_s + st * C::IMG + (live ? c * C::PLANE + yl * C::SP + fx : 0);
#pragma unroll 1
      for (int band = sub; band < C::BANDS; band += C::AQ) {
        const int tc = it * C::BANDS + band;
        const int slot = tc % C::NSLOT, ph = (tc / C::NSLOT) & 1;
        float v[C::W];
#pragma unroll
        for (int x = 0; x < C::W; ++x) v[x] = live ? row0[band * C::BR * C::SP + x] : fill;
        { PROF_T0
        mbar_wait(&a_empty[slot], ph ^ 1);
        PROF_ADD(0) }
        tc_fence_after();
        const uint32_t col = tmem_base + lane_base + (uint32_t)(C::A_COL0 + slot * C::SLOT_COLS);
#pragma unroll
        for (int ks = 0; ks < C::KS; ++ks) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            hi[j] = __float_as_uint(v[ks * 8 + j]);   // read as tf32: the truncated value
            lo[j] = __float_as_uint(v[ks * 8 + j] - tf32_trunc(v[ks * 8 + j]));
          }
          tmem_st16(col + ks * 16, hi, lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[slot]);
      }
      if (next < batch) store_sample(st ^ 1);   // its readers finished a sample ago
      { PROF_T0
      asm volatile("bar.sync 1, %0;" ::"n"(C::A_THREADS) : "memory");
      PROF_ADD(1) }
    }
#ifdef PB_WROWS_PROF
    if (blockIdx.x == 0 && tid == 0) printf("wrows A producer: total %lld | wait a_empty %lld, sample barrier %lld\n", clock64() - pt_start, pt_[0], pt_[1]);
#endif
  } else if (warp == C::MMA_WARP) {
    // ---- MMA issuer ------------------------------------------------------------------
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc(128, C::N);
    const uint32_t b0 = smem_u32(b_s) >> 4;
    int tc = 0;
    uint32_t first = 0;
    for (int item = blockIdx.x; item < batch; item += gridDim.x) {
#pragma unroll 1
      for (int band = 0; band < C::BANDS; ++band, ++tc) {
        const int slot = tc % C::NSLOT, aph = (tc / C::NSLOT) & 1;
        const int bs = tc % C::NB, bph = (tc / C::NB) & 1;
        { PROF_T0
        mbar_wait(&b_full[bs], bph);
        PROF_ADD(0) }
        { PROF_T0
        mbar_wait(&a_full[slot], aph);
        PROF_ADD(1) }
        tc_fence_after();
        if (leader) {
          const uint32_t a_col = tmem_base + (uint32_t)(C::A_COL0 + slot * C::SLOT_COLS);
          const uint32_t b_hi = b0 + (uint32_t)(bs * C::B_STAGE / 4), b_lo = b_hi + (uint32_t)(C::B_SET / 4);
#pragma unroll
          for (int ks = 0; ks < C::KS; ++ks) {
            // LBO = one 4-wide K chunk, SBO = 8 rows (128 B)
            const uint64_t bh = smem_desc(b_hi + (uint32_t)(ks * 2 * C::B_LBO / 4), C::B_LBO / 4, 8);
            const uint64_t bl = smem_desc(b_lo + (uint32_t)(ks * 2 * C::B_LBO / 4), C::B_LBO / 4, 8);
            const uint32_t ah = a_col + ks * 16, al = ah + 8;
            umma_tf32_ts(tmem_base, ah, bh, idesc, first);
            first = 1;
            umma_tf32_ts(tmem_base, al, bh, idesc, 1);
            umma_tf32_ts(tmem_base, ah, bl, idesc, 1);
          }
          umma_commit(&a_empty[slot]);
          umma_commit(&b_empty[bs]);
        }
        __syncwarp();
      }
    }
    if (leader) umma_commit(d_full);
    __syncwarp();
#ifdef PB_WROWS_PROF
    if (blockIdx.x == 0 && lane == 0) printf("wrows MMA warp: total %lld | wait b_full %lld, wait a_full %lld\n", clock64() - pt_start, pt_[0], pt_[1]);
#endif
  } else {
    // ---- gradient formers -----------------------------------------------------------------
    // a thread owns two pool windows (2 rows x 4 columns) of one filter plane per band
    constexpr int ACTV = ACT >= 0 ? ACT : 0;
    const int pt = tid - C::G_WARP0 * 32;
    const int xq = pt % (C::W / 4), rp = (pt / (C::W / 4)) % (C::BR / 2), oc = pt / ((C::W / 4) * (C::BR / 2));
    const int n0 = oc * C::BR + 2 * rp;                    // operand row of the window's first row
    const int at = xq * C::B_LBO + (n0 >> 3) * 32 + (n0 & 7) * 4;   // n0 is even: n0 + 1 is in the same core matrix
    // the thread's pieces of the next PF bands are in flight (cp.async into its own slots of a
    // small ring: no register cost, no barrier - a thread reads back only what it copied)
    constexpr int PF = C::PF, RS = PF + 2;
    float4 *r0 = raw_s + pt, *r1 = r0 + RS * C::G_THREADS;
    float2 *rg = reinterpret_cast<float2 *>(raw_s + 2 * RS * C::G_THREADS) + pt;
    const int n_mine = (int)blockIdx.x < batch ? (batch - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int T = n_mine * C::BANDS;
    // the copy stream walks (sample, band) with running pointers: no division per band
    const int64_t o_item = (int64_t)gridDim.x * C::OC * C::H * C::W - (int64_t)C::BANDS * C::BR * C::W;
    const int64_t g_item = (int64_t)gridDim.x * C::OC * (C::H / 2) * (C::W / 2) - (int64_t)C::BANDS * (C::BR / 2) * (C::W / 2);
    const float *o_nxt = O + (((int64_t)blockIdx.x * C::OC + oc) * C::H + 2 * rp) * C::W + 4 * xq;
    const float *g_nxt = Gp + (((int64_t)blockIdx.x * C::OC + oc) * (C::H / 2) + rp) * (C::W / 2) + 2 * xq;
    int band_nxt = 0, rs_nxt = 0;
    auto issue = [&](int u) {
      if (u < T) {
        __pipeline_memcpy_async(r0 + rs_nxt * C::G_THREADS, o_nxt, 16);
        __pipeline_memcpy_async(r1 + rs_nxt * C::G_THREADS, o_nxt + C::W, 16);
        __pipeline_memcpy_async(rg + rs_nxt * C::G_THREADS, g_nxt, 8);
        o_nxt += C::BR * C::W;
        g_nxt += (C::BR / 2) * (C::W / 2);
        if (++band_nxt == C::BANDS) { band_nxt = 0; o_nxt += o_item; g_nxt += g_item; }
        if (++rs_nxt == RS) rs_nxt = 0;
      }
      __pipeline_commit();
    };
#pragma unroll
    for (int u = 0; u < PF; ++u) issue(u);
    {
      // two bands per iteration: two independent dependency chains per thread
      auto form = [&](int rs, float4 &h0, float4 &h1, float4 &l0, float4 &l1) {
        const float4 p0 = r0[rs * C::G_THREADS];
        const float4 p1 = r1[rs * C::G_THREADS];
        const float2 gg = rg[rs * C::G_THREADS];
        const float a00 = act_fwd<ACTV>(p0.x), a01 = act_fwd<ACTV>(p0.y), a02 = act_fwd<ACTV>(p0.z),
                    a03 = act_fwd<ACTV>(p0.w);
        const float a10 = act_fwd<ACTV>(p1.x), a11 = act_fwd<ACTV>(p1.y), a12 = act_fwd<ACTV>(p1.z),
                    a13 = act_fwd<ACTV>(p1.w);
        const float m0 = fmaxf(fmaxf(a00, a01), fmaxf(a10, a11));
        const float m1 = fmaxf(fmaxf(a02, a03), fmaxf(a12, a13));
        float4 d0, d1;
        d0.x = act_bwd<ACTV>(p0.x, a00 < m0 ? 0.0f : gg.x);
        d0.y = act_bwd<ACTV>(p0.y, a01 < m0 ? 0.0f : gg.x);
        d0.z = act_bwd<ACTV>(p0.z, a02 < m1 ? 0.0f : gg.y);
        d0.w = act_bwd<ACTV>(p0.w, a03 < m1 ? 0.0f : gg.y);
        d1.x = act_bwd<ACTV>(p1.x, a10 < m0 ? 0.0f : gg.x);
        d1.y = act_bwd<ACTV>(p1.y, a11 < m0 ? 0.0f : gg.x);
        d1.z = act_bwd<ACTV>(p1.z, a12 < m1 ? 0.0f : gg.y);
        d1.w = act_bwd<ACTV>(p1.w, a13 < m1 ? 0.0f : gg.y);
        h0 = make_float4(tf32_trunc(d0.x), tf32_trunc(d0.y), tf32_trunc(d0.z), tf32_trunc(d0.w));
        h1 = make_float4(tf32_trunc(d1.x), tf32_trunc(d1.y), tf32_trunc(d1.z), tf32_trunc(d1.w));
        l0 = make_float4(d0.x - h0.x, d0.y - h0.y, d0.z - h0.z, d0.w - h0.w);
        l1 = make_float4(d1.x - h1.x, d1.y - h1.y, d1.z - h1.z, d1.w - h1.w);
      };
      auto publish = [&](int tc, const float4 &h0, const float4 &h1, const float4 &l0, const float4 &l1) {
        const int bs = tc % C::NB, ph = (tc / C::NB) & 1;
        { PROF_T0
        mbar_wait(&b_empty[bs], ph ^ 1);
        PROF_ADD(1) }
        float *hi = b_s + bs * C::B_STAGE + at, *lo = hi + C::B_SET;
        *reinterpret_cast<float4 *>(hi) = h0;
        *reinterpret_cast<float4 *>(hi + 4) = h1;
        *reinterpret_cast<float4 *>(lo) = l0;
        *reinterpret_cast<float4 *>(lo + 4) = l1;
        fence_async_proxy();
        __syncwarp();
        if (lane == 0) mbar_arrive(&b_full[bs]);
      };
      static_assert(C::BANDS % 2 == 0 && PF % 2 == 0, "bands are formed in pairs");
      static_assert(RS % 2 == 0, "a pair of bands never wraps inside the ring");
      int rs = 0;
#pragma unroll 1
      for (int tc = 0; tc < T; tc += 2) {
        issue(tc + PF);
        issue(tc + PF + 1);
        { PROF_T0
        __pipeline_wait_prior(PF);
        PROF_ADD(0) }
        float4 ha0, ha1, la0, la1, hb0, hb1, lb0, lb1;
        form(rs, ha0, ha1, la0, la1);
        form(rs + 1, hb0, hb1, lb0, lb1);
        rs = rs + 2 == RS ? 0 : rs + 2;
        publish(tc, ha0, ha1, la0, la1);
        publish(tc + 1, hb0, hb1, lb0, lb1);
      }
    }
#ifdef PB_WROWS_PROF
    if (blockIdx.x == 0 && pt == 0) printf("wrows former: total %lld | wait cp.async %lld, wait b_empty %lld\n", clock64() - pt_start, pt_[0], pt_[1]);
#endif
  }

  // ---- epilogue: fold the row offsets, one partial per CTA ---------------------------------
  float *d_s = b_s;   // [128][DP]: every MMA that read the gradient stages has completed by d_full
  if (warp < 4) {
    mbar_wait(d_full, 0);
    tc_fence_after();
    const int r = warp * 32 + lane;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
#pragma unroll 1
    for (int c0 = 0; c0 < C::N; c0 += 16) {
      float v[16];
      tmem_ld16(tmem_base + lane_base + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) d_s[r * C::DP + c0 + j] = v[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  {
    float *out = partial + (int64_t)blockIdx.x * C::NW;
    for (int w = tid; w < C::NW; w += C::THREADS) {
      const int oc = w / C::K1, k = w - oc * C::K1;
      float sum = 0.0f;
      if (k == C::K1 - 1) {
#pragma unroll
        for (int y = 0; y < C::BR; ++y) sum += d_s[C::ROWS * C::DP + oc * C::BR + y];
      } else {
        const int c = k / (C::F * C::F), fy = (k / C::F) % C::F, fx = k % C::F;
#pragma unroll
        for (int y = 0; y < C::BR; ++y) sum += d_s[((c * C::YL + y + fy) * C::F + fx) * C::DP + oc * C::BR + y];
      }
      out[w] = sum;
    }
  }
  if (warp == C::MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

template <int ACT>
int launch_wgrad_rows(pb_context *ctx, const float *I, const float *Wt, const float *O, const float *Gp,
                      float *out, const float *lr, int mode, int batch) {
  using C = WCfgR;
  auto kern = conv_wgrad_rows_kernel<ACT>;
  static bool configured = false;
  if (!configured) {
    PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    configured = true;
  }
  const int grid = std::min(batch, ctx->num_sms);
  if (ensure_workspace(ctx, (size_t)grid * C::NW)) return 1;
  PB_CUDA(launch_chained(ctx, kern, dim3(grid), dim3(C::THREADS), C::SMEM, I, O, Gp, ctx->workspace, batch));
  PB_LAUNCHED(ctx);
  return reduce_partials_update(ctx, ctx->workspace, grid, C::NW, Wt, out, lr, mode);
}

}  // namespace

// The weight gradient of the first (convolution, activation, 2x2 max-pool) block straight from
// the convolution's stored output O and the gradient Gp of the pool's output; the convolution's
// input gradient is NOT computed (the caller has no reader for it).  -1 outside the envelope
// (conv_wgrad_mma.cu's variant then runs).
int conv_act_pool_weight_update_rows(pb_context *ctx, const pb_layer_desc *l, int act,
                                     const pb_layer_desc *pool, const float *I, const float *Wt,
                                     const float *O, const float *Gp, float *out, const float *lr,
                                     int mode, int64_t batch) {
  static const bool on = [] {
    const char *e = getenv("PB_DISABLE_UMMA"), *f = getenv("PB_WGRAD_ROWS");
    return !(e && e[0] == '1') && !(f && f[0] == '0');
  }();
  if (!on || l->stride != 1 || l->filter_width != 5 || l->filter_height != 5 || l->padding != 2 ||
      l->width != 32 || l->height != 32 || l->depth != 3 || l->num_filters != 16 || batch < 64 ||
      batch > (1 << 24))
    return -1;
  if (pool->width != 32 || pool->height != 32 || pool->depth != 16 || pool->out_width != 16 ||
      pool->out_height != 16)
    return -1;
  if ((((uintptr_t)I) & 15) || (((uintptr_t)O) & 15) || (((uintptr_t)Gp) & 7)) return -1;
  switch (act) {
    case PB_IDENTITY: return launch_wgrad_rows<PB_IDENTITY>(ctx, I, Wt, O, Gp, out, lr, mode, (int)batch);
    case PB_RELU: return launch_wgrad_rows<PB_RELU>(ctx, I, Wt, O, Gp, out, lr, mode, (int)batch);
    case PB_LEAKY_RELU: return launch_wgrad_rows<PB_LEAKY_RELU>(ctx, I, Wt, O, Gp, out, lr, mode, (int)batch);
    case PB_SIGMOID: return launch_wgrad_rows<PB_SIGMOID>(ctx, I, Wt, O, Gp, out, lr, mode, (int)batch);
  }
  return -1;
}

}  // namespace pb
