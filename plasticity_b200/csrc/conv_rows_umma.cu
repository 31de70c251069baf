// 3xTF32 tcgen05 forward convolution for the network's FIRST layer (3 input channels, 16
// filters, 32x32 images, 5x5 stride-1 "same": nnet/cifar_test.cc:293-300; semantics
// nnet/convolution_layer.cc:50-131) with the IMAGE in tensor memory and the fy taps summed by the
// tensor core over TILES:
//
//   A[(q,x)][(c,fx)]   = Ipad[c][8q + t][x + fx]      tile t, lane = (row band q, column x)
//   B[(4-fy,oc)][(c,fx)] = Wk[oc,c,fy,fx]             80 x 16 (15 used), shared memory, static
//   D[(q,x)][(j,oc)]  += A_t * B[fy = t - j]^T         for the rows j = t-4 .. t inside the band
//   Out[oc][8q + j][x] = bias_oc + D[(q,x)][(j,oc)]    once tile j + 4 is done
//
// With 3 input channels every other mapping leaves the tensor core a sliver: pixels x (16
// filters) with the 75 taps as K is bound by fetching the 128-row image operand from shared
// memory for 16 columns of work (conv_umma.cu, 43 us), filters x pixels with (fy,c) as K folds
// five accumulator rows per output across lanes (conv_wt_umma.cu, 50 us).  Here
//  * a lane quadrant owns a BAND of eight output rows and walks the twelve padded input rows
//    that feed it, one per tile.  The accumulator of a sample is [output row j][filter] = 128
//    columns; input row t contributes to the rows j = t-4 .. t with the filter rows fy = t - j,
//    i.e. to a WINDOW of the accumulator that slides by one row (16 columns) per tile, against
//    a FIXED filter operand ordered (4 - fy, oc): one MMA per K step with D's address = the
//    window start, N = 16 x (rows of the window inside the band) = 16 .. 80.  The fy sum is
//    the MMA's own accumulation - the epilogue of the first version read 80 columns per tile and
//    added them in registers (40 adds per warp and tile, the kernel's bound);
//  * a row is complete when tile j + 4 is done: the epilogue reads its 16 columns (one x8 load
//    per warp), adds the bias, stores whole 128-byte image rows and writes zeros back - every MMA
//    accumulates, a slot is clean again long before the sample after next reaches it (two
//    accumulator buffers);
//  * the A operand (the 5 fx shifts of 3 channels of one image row, hi and lo) is written to
//    tensor memory with tcgen05.st by the warps that own the quadrant (two, alternating tiles)
//    - 15 conflict-free shared-memory loads per lane and tile, no operand fetch by the MMA;
//  * six MMAs per tile, 2 160 MMA cycles per sample; the MMA warp's tile loop is unrolled (window,
//    instruction descriptor and operand offsets are immediates): with so few MMAs per tile the
//    issuing thread's own instruction stream sets the pace, not the tensor pipe.
// The 12/8 halo redundancy (rows shared by neighbouring bands are multiplied twice) is the price.
//
// Warp roles (544 threads, one persistent CTA per SM, one item = one sample):
//   warps 0-7    epilogue: lane quadrant = warp & 3, filters 8*(warp >> 2) .. +7: bias, store; with
//                ACT >= 0 also the 2x2 max-pool of act(conv output) (the two lanes of a window split
//                the filters: one exchange per filter pair)
//   warp  8      MMA issuer, tensor-memory allocator
//   warps 9-16   A producers (quadrant = warp & 3, tiles alternate between the two warps of a
//                quadrant); together they also stage the samples (cp.async, two stages)
#include <cuda_pipeline.h>

#include "common.cuh"
#include "kernels.cuh"
#include "umma.cuh"

namespace pb {
namespace {

using namespace umma;

template <int IC_, int OC_>
struct RCfg {
  static constexpr int IC = IC_, OC = OC_, W = 32, H = 32, F = 5, PAD = 2;
  static constexpr int BAND = H / 4;               // output rows per lane quadrant
  static constexpr int TILES = BAND + F - 1;       // padded input rows that feed a band
  static constexpr int KU = IC * F;                // (c, fx)
  static constexpr int KS = (KU + 7) / 8;          // K = 8 steps
  static constexpr int N = OC * F;                 // (oc, fy)
  static constexpr int EG = 2;                     // epilogue warps per quadrant
  static constexpr int OPG = OC / EG;              // filters per epilogue warp
  static constexpr int EPI_WARPS = 4 * EG, MMA_WARP = EPI_WARPS, PROD_WARP0 = EPI_WARPS + 1;
  static constexpr int PQ = 2;                     // A-producer warps per lane quadrant (alternating tiles)
  static constexpr int PROD_THREADS = 128 * PQ;
  static constexpr int THREADS = (EPI_WARPS + 1 + 4 * PQ) * 32;
  static constexpr int SP = W + 8;                 // staged row: 4 zeros | W | 4 zeros
  static constexpr int SR = H + F - 1;
  static constexpr int PLANE = SR * SP;
  static constexpr int IMG = IC * PLANE;           // floats per stage
  static constexpr int B_LBO = N * 4;              // floats between 4-wide K chunks
  static constexpr int B_SET = KS * 2 * B_LBO;     // floats of the hi (or lo) filter operand
  static constexpr int SLOT_COLS = KS * 16;        // [k step][hi 8 | lo 8]
  static constexpr int NSLOT = 4;                  // ring of A slots (one tile each)
  static constexpr int DBUF = BAND * OC;           // accumulator columns of one sample: [row j][oc]
  static constexpr int A_COL0 = 2 * DBUF;          // two samples' accumulators first
  static constexpr int TMEM_NEED = A_COL0 + NSLOT * SLOT_COLS;
  static constexpr int TMEM_COLS = TMEM_NEED <= 256 ? 256 : 512;
  static constexpr int K1 = F * F * IC + 1;        // reference filter-block stride
  static constexpr int NWT = OC * K1;
  static constexpr size_t SMEM = sizeof(float) * (size_t)(2 * IMG + 2 * B_SET + 32) + 8 * (2 * NSLOT + 2 * BAND + 2) + 16;
  static_assert(N % 16 == 0 && N <= 256 && OC % EG == 0 && F == 5, "tile shape");
  static_assert(OPG == 8, "an epilogue warp reads its eight filters of a row with one x8 load");
  static_assert(TMEM_NEED <= 512, "tensor-memory budget");
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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
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

// ACT >= 0: Out2 = 2x2 max-pool of act(conv output) as well.
template <class C, int ACT>
__global__ void __launch_bounds__(C::THREADS, 1)
conv_rows_kernel(const float *__restrict__ In, const float *__restrict__ Wt, float *__restrict__ Out,
                 float *__restrict__ Out2, int batch, int early_w) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float *img_s = reinterpret_cast<float *>(smem_raw);                    // [2][IMG]
  float *b_s = img_s + 2 * C::IMG;                                        // hi set, lo set
  float *bias_s = b_s + 2 * C::B_SET;                                     // [32]
  uint64_t *bars = reinterpret_cast<uint64_t *>(bias_s + 32);
  uint64_t *a_full = bars, *a_empty = bars + C::NSLOT, *row_full = bars + 2 * C::NSLOT,   // [2][BAND]
           *buf_empty = bars + 2 * C::NSLOT + 2 * C::BAND;                                  // [2]
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 2 * C::NSLOT + 2 * C::BAND + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef PB_ROWS_PROF
  long long pt[4] = {0, 0, 0, 0};
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
  for (int i = tid; i < 2 * C::IMG; i += C::THREADS) img_s[i] = 0.0f;   // pads stay zero
  for (int i = tid; i < 2 * C::B_SET; i += C::THREADS) b_s[i] = 0.0f;   // K padding
  if (tid == 0) {
    for (int s = 0; s < C::NSLOT; ++s) { mbar_init(&a_full[s], 4); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2 * C::BAND; ++s) mbar_init(&row_full[s], 1);
    for (int s = 0; s < 2; ++s) mbar_init(&buf_empty[s], C::EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == C::MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_ptr)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (!early_w) chain_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp < C::EPI_WARPS) {
    // every accumulator column starts at zero: all MMAs accumulate (a row's slot is zeroed
    // again by the warp that reads it out)
    const uint32_t zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint32_t at = *tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * C::OPG);
#pragma unroll 1
    for (int j = 0; j < 2 * C::BAND; ++j) tmem_st8(at + j * C::OC, zero);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  {
    // filters -> K-major no-swizzle operand: row n = (4 - fy, oc), k = (c, fx);
    // float(n, k) = (k/4)*B_LBO + (n/8)*32 + (n%8)*4 + k%4.  All L2 loads in flight first.
    constexpr int WL = (C::NWT + C::THREADS - 1) / C::THREADS;
    float wv[WL];
#pragma unroll
    for (int k = 0; k < WL; ++k) {
      const int i = tid + k * C::THREADS;
      wv[k] = i < C::NWT ? __ldcg(Wt + i) : 0.0f;
    }
#pragma unroll
    for (int k = 0; k < WL; ++k) {
      const int i = tid + k * C::THREADS;
      if (i < C::NWT) {
        const int oc = i / C::K1, r = i - oc * C::K1;
        if (r == C::K1 - 1) {
          bias_s[oc] = wv[k];
        } else {
          const int c = r / (C::F * C::F), fy = (r / C::F) % C::F, fx = r % C::F;
          const int n = (C::F - 1 - fy) * C::OC + oc, kk = c * C::F + fx;
          const int at = (kk >> 2) * C::B_LBO + (n >> 3) * 32 + (n & 7) * 4 + (kk & 3);
          const float h = tf32_trunc(wv[k]);
          b_s[at] = h;
          b_s[C::B_SET + at] = wv[k] - h;
        }
      }
    }
  }
  fence_async_proxy();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (early_w) chain_wait();
  chain_release();

  if (warp >= C::PROD_WARP0) {
    // ---- A producers ------------------------------------------------------------------
    const int quad = warp & 3;
    const int ptid = tid - C::PROD_WARP0 * 32;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    auto stage = [&](int st, int item) {
      // interior of the padded planes: row y -> staged row y + 2, column x -> x + 4
      float *dst = img_s + st * C::IMG;
      const float *src = In + (int64_t)item * C::IC * C::H * C::W;
      for (int e = ptid; e < C::IC * C::H * C::W / 4; e += C::PROD_THREADS) {
        const int x4 = e % (C::W / 4), r = e / (C::W / 4);
        const int y = r % C::H, c = r / C::H;
        __pipeline_memcpy_async(dst + c * C::PLANE + (y + C::PAD) * C::SP + 4 + 4 * x4, src + 4 * e, 16);
      }
    };
    int it = 0;
    if ((int)blockIdx.x < batch) stage(0, blockIdx.x);
    __pipeline_commit();
    for (int item = blockIdx.x; item < batch; item += gridDim.x, ++it) {
      const int st = it & 1;
      // every producer is done reading the other stage (the sample before this one)
      asm volatile("bar.sync 1, %0;" ::"n"(C::PROD_THREADS) : "memory");
      if (item + (int)gridDim.x < batch) stage(st ^ 1, item + gridDim.x);
      __pipeline_commit();
      __pipeline_wait_prior(1);
      asm volatile("bar.sync 1, %0;" ::"n"(C::PROD_THREADS) : "memory");   // everybody's pieces of this sample landed
      // lane x of quadrant q, tile t: staged row 8q + t, columns x + fx + 2
      const float *row0 = img_s + st * C::IMG + (quad * C::BAND) * C::SP + lane + 2;
      // the shared-memory loads and the hi/lo split of the warp's next tile run between a tile's
      // tcgen05.st and the wait for it
      uint32_t hi[C::KS][8], lo[C::KS][8];
      auto prepare = [&](int t) {
#pragma unroll
        for (int k = 0; k < C::KS * 8; ++k) {
          const float x = k < C::KU ? row0[(k / C::F) * C::PLANE + t * C::SP + (k % C::F)] : 0.0f;
          hi[k >> 3][k & 7] = __float_as_uint(x);   // read as tf32: the truncated value
          lo[k >> 3][k & 7] = __float_as_uint(x - tf32_trunc(x));
        }
      };
      const int sub = ptid >> 7;   // the warps of a quadrant alternate over the tiles
      prepare(sub);
#pragma unroll 1
      for (int t = sub; t < C::TILES; t += C::PQ) {
        const int tc = it * C::TILES + t;
        const int slot = tc % C::NSLOT, ph = (tc / C::NSLOT) & 1;
        { PROF_T0
        mbar_wait(&a_empty[slot], ph ^ 1);
        PROF_ADD(0) }
        tc_fence_after();
        const uint32_t col = tmem_base + lane_base + (uint32_t)(C::A_COL0 + slot * C::SLOT_COLS);
#pragma unroll
        for (int ks = 0; ks < C::KS; ++ks) tmem_st16(col + ks * 16, hi[ks], lo[ks]);
        if (t + C::PQ < C::TILES) prepare(t + C::PQ);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[slot]);
      }
    }
    __pipeline_wait_prior(0);
#ifdef PB_ROWS_PROF
    if (blockIdx.x == 0 && ptid == 0) printf("rows A producer: total %lld | wait a_empty %lld\n", clock64() - pt_start, pt[0]);
#endif
  } else if (warp == C::MMA_WARP) {
    // ---- MMA issuer -------------------------------------------------------------------
    const bool leader = elect_one();
    const uint32_t b_hi = smem_u32(b_s) >> 4, b_lo = smem_u32(b_s + C::B_SET) >> 4;
    int it = 0;
    static_assert(C::TILES % C::NSLOT == 0, "a sample's tiles use the A slots in a fixed pattern");
    for (int item = blockIdx.x; item < batch; item += gridDim.x, ++it) {
      const int buf = it & 1;
      { PROF_T0
      mbar_wait(&buf_empty[buf], ((it >> 1) & 1) ^ 1);   // its rows of two samples ago are read out
      PROF_ADD(0) }
      // unrolled over the twelve tiles: window, instruction descriptor and operand offsets of a
      // tile are immediates - with six MMAs per tile the issuing thread's own instruction
      // stream, not the tensor pipe, sets the pace
#pragma unroll
      for (int t = 0; t < C::TILES; ++t) {
        const int slot = t % C::NSLOT, aph = (it * (C::TILES / C::NSLOT) + t / C::NSLOT) & 1;
        { PROF_T0
        mbar_wait(&a_full[slot], aph);
        PROF_ADD(1) }
        tc_fence_after();
        // input row t feeds output rows j = t-4 .. t (fy = t - j) that lie inside the band:
        // a window of the accumulator [row j][oc] and the matching blocks of the filter operand
        const int j0 = max(0, t - (C::F - 1)), j1 = min(C::BAND - 1, t);
        const int i0 = max(0, (C::F - 1) - t);
        const int n_cols = (j1 - j0 + 1) * C::OC;
        if (leader) {
          const uint32_t idesc = make_idesc(128, n_cols);
          const uint32_t d = tmem_base + (uint32_t)(buf * C::DBUF + j0 * C::OC);
          const uint32_t a_col = tmem_base + (uint32_t)(C::A_COL0 + slot * C::SLOT_COLS);
          const uint32_t b_off = (uint32_t)(i0 * C::OC * 4 / 4);   // 16-byte units: OC rows of 16 B
#pragma unroll
          for (int ks = 0; ks < C::KS; ++ks) {
            // LBO = one 4-wide K chunk (N rows of 16 B), SBO = 8 rows (128 B)
            const uint64_t bh = smem_desc(b_hi + b_off + (uint32_t)(ks * 2 * C::B_LBO / 4), C::B_LBO / 4, 8);
            const uint64_t bl = smem_desc(b_lo + b_off + (uint32_t)(ks * 2 * C::B_LBO / 4), C::B_LBO / 4, 8);
            const uint32_t ah = a_col + ks * 16, al = ah + 8;
            umma_tf32_ts(d, ah, bh, idesc, 1);
            umma_tf32_ts(d, al, bh, idesc, 1);
            umma_tf32_ts(d, ah, bl, idesc, 1);
          }
          umma_commit(&a_empty[slot]);
          if (t >= C::F - 1) umma_commit(&row_full[buf * C::BAND + t - (C::F - 1)]);
        }
        __syncwarp();
      }
    }
#ifdef PB_ROWS_PROF
    if (blockIdx.x == 0 && lane == 0) printf("rows MMA warp: total %lld | wait d_empty %lld, wait a_full %lld\n", clock64() - pt_start, pt[0], pt[1]);
#endif
  } else {
    // ---- epilogue: rolling sums over the tiles of a band ---------------------------------
    const int quad = warp & 3, grp = warp >> 2;
    const uint32_t d_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(grp * C::OPG);
    float bias[C::OPG];
#pragma unroll
    for (int o = 0; o < C::OPG; ++o) bias[o] = bias_s[grp * C::OPG + o];
    const uint32_t zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int it = 0;
    for (int item = blockIdx.x; item < batch; item += gridDim.x, ++it) {
      const int buf = it & 1, par = (it >> 1) & 1;
      float *out = Out + (((int64_t)item * C::OC + grp * C::OPG) * C::H + quad * C::BAND) * C::W + lane;
      float *out2 = Out2 + (((int64_t)item * C::OC + grp * C::OPG) * (C::H / 2) + quad * (C::BAND / 2)) * (C::W / 2) +
                    (lane >> 1) + ((lane & 1) ? (C::OPG / 2) * (C::H / 2) * (C::W / 2) : 0);
      float prev[C::OPG];
#pragma unroll
      for (int o = 0; o < C::OPG; ++o) prev[o] = 0.0f;
#pragma unroll 1
      for (int j = 0; j < C::BAND; ++j) {
        { PROF_T0
        mbar_wait(&row_full[buf * C::BAND + j], par);   // tile j + 4, the row's last, is done
        PROF_ADD(0) }
        tc_fence_after();
        float v[C::OPG];
        const uint32_t at = d_addr + (uint32_t)(buf * C::DBUF + j * C::OC);
        tmem_ld8(at, v);
        tmem_ld_wait();
        tmem_st8(at, zero);   // the slot of row j of the sample after next
        float done[C::OPG];
#pragma unroll
        for (int o = 0; o < C::OPG; ++o) {
          done[o] = v[o] + bias[o];
          out[(o * C::H + j) * C::W] = done[o];
        }
        if constexpr (ACT >= 0) {
          if (j & 1) {   // the second row of a pool window
            // the two lanes of a window split the filters: the even lane finishes filters
            // 0 .. OPG/2-1, the odd lane the other half - one exchange per filter PAIR and
            // half the activations per lane
            const bool odd = lane & 1;
#pragma unroll
            for (int o = 0; o < C::OPG / 2; ++o) {
              const float m0 = fmaxf(prev[o], done[o]);
              const float m1 = fmaxf(prev[o + C::OPG / 2], done[o + C::OPG / 2]);
              const float theirs = __shfl_xor_sync(0xffffffffu, odd ? m0 : m1, 1);
              const float p = act_fwd<(ACT >= 0 ? ACT : 0)>(fmaxf(odd ? m1 : m0, theirs));
              out2[(o * (C::H / 2) + (j >> 1)) * (C::W / 2)] = p;
            }
          }
        }
#pragma unroll
        for (int o = 0; o < C::OPG; ++o) prev[o] = done[o];
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&buf_empty[buf]);
    }
  }

#ifdef PB_ROWS_PROF
  if (blockIdx.x == 0 && tid == 0) printf("rows epilogue: total %lld | wait d_full %lld\n", clock64() - pt_start, pt[0]);
#endif
  // ---- teardown -------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == C::MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

template <class C, int ACT>
int launch_rows(pb_context *ctx, const float *In, const float *Wt, float *Out, float *Out2,
                int64_t batch) {
  auto kern = conv_rows_kernel<C, ACT>;
  static bool configured = false;
  if (!configured) {
    PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    configured = true;
  }
  const int grid = (int)std::min<int64_t>(batch, ctx->num_sms);
  const int early_w = ctx->chain == 2;
  PB_CUDA(launch_chained(ctx, kern, dim3(grid), dim3(C::THREADS), C::SMEM, In, Wt, Out, Out2, (int)batch,
                         early_w));
  PB_LAUNCHED(ctx);
  return 0;
}

bool rows_shape(const pb_layer_desc *l, int64_t batch) {
  static const bool on = [] {
    const char *e = getenv("PB_DISABLE_UMMA"), *f = getenv("PB_CONV_ROWS");
    return !(e && e[0] == '1') && !(f && f[0] == '0');
  }();
  return on && l->stride == 1 && batch >= 64 && batch <= (1 << 24) && l->filter_width == 5 &&
         l->filter_height == 5 && l->padding == 2 && l->width == 32 && l->height == 32 && l->depth == 3 &&
         l->num_filters == 16;
}

}  // namespace

// Each returns -1 outside the envelope (the first layer of the CIFAR network).
int conv_evaluate_rows(pb_context *ctx, const pb_layer_desc *l, const float *I, const float *W, float *O,
                       int64_t batch) {
  if (!rows_shape(l, batch) || (((uintptr_t)I) & 15)) return -1;
  return launch_rows<RCfg<3, 16>, -1>(ctx, I, W, O, nullptr, batch);
}

int conv_act_pool_evaluate_rows(pb_context *ctx, const pb_layer_desc *l, int act, const pb_layer_desc *pool,
                                const float *I, const float *W, float *O, float *O2, int64_t batch) {
  if (!rows_shape(l, batch) || (((uintptr_t)I) & 15)) return -1;
  if (pool->width != 32 || pool->height != 32 || pool->depth != 16 || pool->out_width != 16 ||
      pool->out_height != 16)
    return -1;
  using C = RCfg<3, 16>;
  switch (act) {
    case PB_IDENTITY: return launch_rows<C, PB_IDENTITY>(ctx, I, W, O, O2, batch);
    case PB_RELU: return launch_rows<C, PB_RELU>(ctx, I, W, O, O2, batch);
    case PB_LEAKY_RELU: return launch_rows<C, PB_LEAKY_RELU>(ctx, I, W, O, O2, batch);
    case PB_SIGMOID: return launch_rows<C, PB_SIGMOID>(ctx, I, W, O, O2, batch);
  }
  return -1;
}

}  // namespace pb
