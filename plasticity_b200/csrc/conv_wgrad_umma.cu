// 3xTF32 tcgen05 convolution weight gradient (+ SGD update) for the batched 5x5 stride-1
// "same" convolutions of the CIFAR network (nnet/cifar_test.cc:293-350; semantics
// nnet/convolution_layer.cc:226-270, update back_prop.kernel.cl:28-44):
//
//   g[oc][c][fy][fx] = sum_{b,y,x} G[b,oc,y,x] * In0[b,c,y+fy-2,x+fx-2]     bias: sum G[b,oc,.,.]
//
// The contraction index is (sample, pixel); the free indices are (fx, oc) x (fy, c).  GEMM
// view per image row y of a sample, with X = x + fx + 2 running over the padded row:
//
//   A[(oc,fx)][X] = G[oc][y][X-fx-2]      100 (80) rows     -> TENSOR MEMORY (lane = row)
//   B[(fy,c)][X]  = Ipad[c][y+fy][X]      N = 5*IC rows     -> shared memory, K-major
//   D[(oc,fx)][(fy,c)] += A * B^T         one accumulator for every sample of the CTA
//
// * The fx shift lives in A: the four warps that own the tensor-memory lane quadrants read
//   the gradient row from a zero-padded shared-memory tile at a per-lane offset, split
//   it into hi = v truncated to tf32 and lo = v - hi and write both with tcgen05.st (16
//   columns per K step) - no shifted copies in shared memory, no operand fetch for A at all.
// * The fy shift lives in B's START ADDRESS: the image band is staged once as
//   [4-pixel quad][row Y][channel][4 pixels] (16 bytes per (quad, Y, channel)), i.e. tcgen05's
//   K-major no-swizzle layout with rows n = fy*IC + c contiguous (SBO = 128 B) and the two
//   quads of a K = 8 step one quad plane apart (LBO); image row y starts IC units further.
//   Ipad[c][Y][X] = In[c][Y-2][X-4]: the 4-pixel left pad keeps global float4 loads aligned.
// * 3xTF32 = three MMAs per K step into the SAME accumulator (A_hi*B_hi + A_lo*B_hi +
//   A_hi*B_lo): with A in tensor memory nothing is re-fetched, so no N-stacking.  Measured
//   (tools/umma_wgrad_probe.cu): exactly the tensor time, 40 cycles per M=128 N=80 MMA.
//   The RGB layer (N = 15) would be MMA-issue bound (~22 cycles per MMA); it stacks the hi
//   and lo image rows along N instead ([quad][Y][hi|lo][c], N = 30) and issues two MMAs.
// * MN-major (transposed) tf32 operands read zeros on B200 for every descriptor tried
//   (tools/umma_mn_decode.cu), hence the K-major staging of B.
//
// Warp roles (one persistent CTA per SM, items = (sample, band of RB rows)):
//   warps 0..4*AQ-1  A producers: warp w owns tensor-memory lane quadrant w & 3 and every
//                    AQ-th image row (AQ warps per quadrant hide the shared-memory load ->
//                    tcgen05.st -> wait::st latency chain of a row); together they stage the
//                    next item's gradient tile (register-staged, double buffered); warps 0-3
//                    run the epilogue
//   next warp        MMA issuer, tensor-memory allocator
//   last 4 warps     B producers: global CHW fp32 -> hi/lo -> shared memory (two stages)
// Gradient tile: [oc][row][8 zeros | W | 8 zeros], channel pitch = 5 mod 32 and row r of the
// GEMM = oc*5 + (4 - fx), so that lane r reads bank (r - 4 + X) mod 32: conflict-free.
// The epilogue writes ONE partial per CTA ([column][row], coalesced); a second kernel adds
// the partials in a fixed order and applies the update (deterministic, like the FFMA and
// mma.sync paths).
#include "common.cuh"
#include "kernels.cuh"
#include "umma.cuh"

namespace pb {
namespace {

using namespace umma;

constexpr int kBWarps = 4;

template <int IC_, int OC_, int W_, int RB_, bool STACK_, int AQ_, int NS_>
struct GCfg {
  static constexpr int IC = IC_, OC = OC_, W = W_, H = W_, RB = RB_, F = 5;
  static constexpr bool STACK = STACK_;
  static constexpr int AQ = AQ_;                   // A-producer warps per lane quadrant
  static constexpr int NS = NS_;                   // ring of A slots in tensor memory (one image row each)
  static constexpr int A_WARPS = 4 * AQ, A_THREADS = 128 * AQ;
  static constexpr int MMA_WARP = A_WARPS;
  static constexpr int THREADS = (A_WARPS + 1 + kBWarps) * 32;
  static constexpr int PK = W + 8;                 // K extent of an image row (padded X)
  static constexpr int KR = PK / 8;                // K steps per image row
  static constexpr int NQ = PK / 4;                // quad planes
  static constexpr int HL = STACK ? 2 : 1;
  static constexpr int UN = IC * HL;               // 16-byte units per (quad, Y)
  static constexpr int NC = F * UN;                // accumulator columns in use
  static constexpr int NPAD = ((NC + 15) / 16) * 16;  // MMA N
  static constexpr int YR = RB + F - 1 + (NPAD > NC ? 1 : 0);  // rows per quad plane (+ slack)
  static constexpr int QS = YR * UN;               // units per quad plane
  static constexpr int BSET = NQ * QS;             // units per plane set
  static constexpr int B_STAGE = STACK ? BSET : 2 * BSET;  // hi set, lo set
  static constexpr int GR = W + 16;                // gradient tile row: 8 zeros | W | 8 zeros
  static constexpr int GP = RB * GR + ((37 - (RB * GR) % 32) % 32);  // channel pitch = 5 mod 32
  static constexpr int G_STAGE = OC * GP;          // floats
  static constexpr int ROWS = OC * F;              // live A / D rows: r = oc*F + (F-1-fx)
  static constexpr int SLOT_COLS = KR * 16;        // [k step][hi 8 | lo 8]
  static constexpr int A_COL0 = ((NPAD + 31) / 32) * 32;
  static constexpr int TMEM_NEED = A_COL0 + NS * SLOT_COLS;
  static constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
  static constexpr int IPS = H / RB;               // items per sample
  static constexpr int ROW = NPAD * 128 + 32 * AQ;  // floats per CTA partial: [col][row], bias[AQ][32]
  static constexpr int G_VEC = OC * RB * (W / 4);  // 16-byte pieces of a gradient tile
  static constexpr int G_ITERS = (G_VEC + A_THREADS - 1) / A_THREADS;
  static constexpr int K1 = F * F * IC + 1;        // reference filter-block stride
  static constexpr int B_ITERS = ((RB + F - 1) * IC * (W / 4) + kBWarps * 32 - 1) / (kBWarps * 32);
  static constexpr size_t SMEM = 2 * (size_t)B_STAGE * 16 + 2 * (size_t)G_STAGE * 4 + 32 * 8 + 16;
  static_assert(W % 8 == 0 && H % RB == 0, "tiling");
  static_assert(ROWS <= 128 && NPAD <= 256 && TMEM_NEED <= 512, "tensor-memory budget");
  static_assert(NPAD - NC <= UN, "one slack row covers the rounded-up N");
  static_assert(GP % 32 == 5 && OC <= 32, "gradient tile pitch");
  static_assert(SMEM <= 227 * 1024, "shared memory");
};

// 16 consecutive columns of this warp's lane quadrant: the hi and lo parts of one K step
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&h)[8], const uint32_t (&l)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]),
      "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7])
      : "memory");
}
// x = hi + lo exactly, hi = x with the 13 low mantissa bits cleared: the value a kind::tf32
// MMA reads from x anyway (so x itself is stored as the hi operand where that saves a move)
__device__ __forceinline__ float tf32_trunc(float x) {
  return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc], kind::tf32
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
template <class C>
__global__ void __launch_bounds__(C::THREADS, 1)
conv_wgrad_umma_kernel(const float *__restrict__ In, const float *__restrict__ G,
                       float *__restrict__ partial, int batch) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float4 *b_s = reinterpret_cast<float4 *>(smem_raw);                      // [2][B_STAGE]
  float *g_s = reinterpret_cast<float *>(b_s + 2 * C::B_STAGE);            // [2][G_STAGE]
  uint64_t *bars = reinterpret_cast<uint64_t *>(g_s + 2 * C::G_STAGE);
  constexpr int kSlots = C::NS, kThreads = C::THREADS, kMmaWarp = C::MMA_WARP;
  uint64_t *a_full = bars, *a_empty = bars + kSlots, *b_full = bars + 2 * kSlots,
           *b_empty = bars + 2 * kSlots + 2, *d_full = bars + 2 * kSlots + 4;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 2 * kSlots + 6);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef PB_WGRAD_PROF
  long long pt[6] = {0, 0, 0, 0, 0, 0};
  const long long pt_start = clock64();
#define PROF_T0 const long long pt0_ = clock64();
#define PROF_ADD(i) pt[i] += clock64() - pt0_;
#else
#define PROF_T0
#define PROF_ADD(i)
#endif

  // ---- one-time setup (runs ahead of the dependency wait of a chained launch) ------
  // every pad of both operand images is zero for the whole kernel; interiors are rewritten
  for (int i = tid; i < 2 * C::B_STAGE; i += kThreads) b_s[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = tid; i < 2 * C::G_STAGE; i += kThreads) g_s[i] = 0.0f;
  if (tid == 0) {
    for (int s = 0; s < kSlots; ++s) { mbar_init(&a_full[s], 4); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&b_full[s], kBWarps); mbar_init(&b_empty[s], 1); }
    mbar_init(d_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_ptr)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_proxy();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int n_items = batch * C::IPS;
  chain_wait();
  chain_release();

  if (warp < C::A_WARPS) {
    // ---- gradient tile loader + A producers + epilogue ---------------------------------
    const int quad = warp & 3, sub = warp >> 2;
    const int r = quad * 32 + lane;           // tensor-memory lane = A row = D row
    const int oc = r / C::F, fx = C::F - 1 - (r - oc * C::F);
    const bool live = r < C::ROWS;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    float4 gv[C::G_ITERS];
    auto load_tile = [&](int item) {
      const int b = item / C::IPS, band = item - b * C::IPS;
      const float4 *src = reinterpret_cast<const float4 *>(G + ((int64_t)b * C::OC * C::H + band * C::RB) * C::W);
#pragma unroll
      for (int k = 0; k < C::G_ITERS; ++k) {
        const int e = tid + k * C::A_THREADS;
        const int xq = e % (C::W / 4), t = e / (C::W / 4);
        const int yl = t % C::RB, c = t / C::RB;
        gv[k] = e < C::G_VEC ? __ldg(src + ((int64_t)c * C::H + yl) * (C::W / 4) + xq)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto store_tile = [&](int st) {
      float *dst = g_s + st * C::G_STAGE;
#pragma unroll
      for (int k = 0; k < C::G_ITERS; ++k) {
        const int e = tid + k * C::A_THREADS;
        if (e < C::G_VEC) {
          const int xq = e % (C::W / 4), t = e / (C::W / 4);
          const int yl = t % C::RB, c = t / C::RB;
          float *d = dst + c * C::GP + yl * C::GR + 8 + 4 * xq;
          d[0] = gv[k].x; d[1] = gv[k].y; d[2] = gv[k].z; d[3] = gv[k].w;
        }
      }
    };
    float bias = 0.0f;
    int it = 0;
    if (blockIdx.x < n_items) { load_tile(blockIdx.x); store_tile(0); }
    asm volatile("bar.sync 1, %0;" ::"n"(C::A_THREADS) : "memory");
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int st = it & 1;
      const int next = item + gridDim.x;
      if (next < n_items) load_tile(next);
      // row of this lane: X = 0 reads tile column 8 + (0 - fx - 2)
      const float *grow = g_s + st * C::G_STAGE + (live ? oc : 0) * C::GP + 6 - fx;
#pragma unroll 1
      for (int yl = sub; yl < C::RB; yl += C::AQ) {
        const int rowc = it * C::RB + yl;
        const int slot = rowc % kSlots, ph = (rowc / kSlots) & 1;
        float v[C::PK];
        { PROF_T0
#pragma unroll
        for (int j = 0; j < C::PK; ++j) v[j] = live ? grow[yl * C::GR + j] : 0.0f;
        if (fx == 0) {
          float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < C::PK; ++j) part[j & 3] += v[j];
          bias += (part[0] + part[1]) + (part[2] + part[3]);
        }
        PROF_ADD(1) }
        { PROF_T0
        mbar_wait(&a_empty[slot], ph ^ 1);
        PROF_ADD(2) }
        tc_fence_after();
        PROF_T0
        const uint32_t col = tmem_base + lane_base + (uint32_t)(C::A_COL0 + slot * C::SLOT_COLS);
#pragma unroll
        for (int ks = 0; ks < C::KR; ++ks) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            hi[j] = __float_as_uint(v[ks * 8 + j]);   // read as tf32: the truncated value
            lo[j] = __float_as_uint(v[ks * 8 + j] - tf32_trunc(v[ks * 8 + j]));
          }
          tmem_st16(col + ks * 16, hi, lo);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[slot]);
        PROF_ADD(3)
      }
      if (next < n_items) store_tile(st ^ 1);  // its readers finished an item ago
      { PROF_T0
      asm volatile("bar.sync 1, %0;" ::"n"(C::A_THREADS) : "memory");
      PROF_ADD(0) }
    }
    if (live && fx == 0) partial[(int64_t)blockIdx.x * C::ROW + C::NPAD * 128 + sub * 32 + oc] = bias;
    if (sub == 0) {
    // epilogue: accumulator rows -> this CTA's partial, [column][row] (lanes along rows)
    { PROF_T0
    mbar_wait(d_full, 0);
    PROF_ADD(4) }
#ifdef PB_WGRAD_PROF
    if (blockIdx.x == 0 && tid == 0)
      printf("A producer: total %lld | tile barrier %lld, LDS+bias %lld, wait a_empty %lld, cvt+st+arrive %lld, wait d_full %lld\n",
             clock64() - pt_start, pt[0], pt[1], pt[2], pt[3], pt[4]);
#endif
    tc_fence_after();
    float *out = partial + (int64_t)blockIdx.x * C::ROW;
#pragma unroll 1
    for (int c0 = 0; c0 < C::NPAD; c0 += 16) {
      float v[16];
      tmem_ld16(tmem_base + lane_base + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) out[(c0 + j) * 128 + r] = v[j];
    }
    }
  } else if (warp == kMmaWarp) {
    // ---- MMA issuer ------------------------------------------------------------------
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc(128, C::NPAD);
    // B descriptor: start | LBO = quad plane | SBO = 128 B | version 1
    const uint64_t b_desc0 = ((uint64_t)(8u | (1u << 14)) << 32) | ((uint64_t)C::QS << 16) |
                             (uint64_t)(smem_u32(b_s) >> 4);
    int it = 0, rowc = 0;
    uint32_t first = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int st = it & 1, bph = (it >> 1) & 1;
      { PROF_T0
      mbar_wait(&b_full[st], bph);
      PROF_ADD(0) }
      tc_fence_after();
#pragma unroll 1
      for (int yl = 0; yl < C::RB; ++yl, ++rowc) {
        const int slot = rowc % kSlots, ph = (rowc / kSlots) & 1;
        { PROF_T0
        mbar_wait(&a_full[slot], ph);
        PROF_ADD(1) }
        tc_fence_after();
        PROF_T0
        const uint32_t a_col = tmem_base + (uint32_t)(C::A_COL0 + slot * C::SLOT_COLS);
        const uint64_t b_row = b_desc0 + (uint64_t)(st * C::B_STAGE + yl * C::UN);
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < C::KR; ++ks) {
            const uint64_t bh = b_row + (uint64_t)(2 * ks * C::QS);
            const uint32_t ah = a_col + ks * 16, al = ah + 8;
            umma_tf32_ts(tmem_base, ah, bh, idesc, first);
            first = 1;
            umma_tf32_ts(tmem_base, al, bh, idesc, 1);
            if (!C::STACK) umma_tf32_ts(tmem_base, ah, bh + (uint64_t)C::BSET, idesc, 1);
          }
          umma_commit(&a_empty[slot]);
        }
        __syncwarp();
        PROF_ADD(2)
      }
      if (leader) umma_commit(&b_empty[st]);
      __syncwarp();
    }
    if (leader) umma_commit(d_full);
    __syncwarp();
#ifdef PB_WGRAD_PROF
    if (blockIdx.x == 0 && lane == 0)
      printf("MMA warp: total %lld | wait b_full %lld, wait a_full %lld, issue %lld\n", clock64() - pt_start, pt[0], pt[1], pt[2]);
#endif
  } else {
    // ---- B producers: global CHW -> [quad][Y][hi|lo][channel][4 pixels] ---------------------
    const int ptid = tid - (C::A_WARPS + 1) * 32;
    constexpr int NT = kBWarps * 32;
    constexpr int WQ = C::W / 4;
    constexpr int TOTAL = (C::RB + C::F - 1) * C::IC * WQ;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int st = it & 1, ph = (it >> 1) & 1;
      const int b = item / C::IPS, band = item - b * C::IPS;
      const float *src = In + (int64_t)b * C::IC * C::H * C::W;
      const int y0 = band * C::RB - 2;
      float4 v[C::B_ITERS];
#pragma unroll
      for (int k = 0; k < C::B_ITERS; ++k) {
        const int e = ptid + k * NT;
        const int c = e % C::IC, t = e / C::IC;
        const int qi = t % WQ, yl = t / WQ;
        const int y = y0 + yl;
        v[k] = (e < TOTAL && y >= 0 && y < C::H)
                   ? __ldg(reinterpret_cast<const float4 *>(src + ((int64_t)c * C::H + y) * C::W) + qi)
                   : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      { PROF_T0
      mbar_wait(&b_empty[st], ph ^ 1);
      PROF_ADD(0) }
      float4 *hi = b_s + st * C::B_STAGE;
#pragma unroll
      for (int k = 0; k < C::B_ITERS; ++k) {
        const int e = ptid + k * NT;
        if (e < TOTAL) {
          const int c = e % C::IC, t = e / C::IC;
          const int qi = t % WQ, yl = t / WQ;
          const float h0 = tf32_trunc(v[k].x), h1 = tf32_trunc(v[k].y), h2 = tf32_trunc(v[k].z),
                      h3 = tf32_trunc(v[k].w);
          const int u = (qi + 1) * C::QS + yl * C::UN + c;
          hi[u] = make_float4(h0, h1, h2, h3);
          hi[u + (C::STACK ? C::IC : C::BSET)] =
              make_float4(v[k].x - h0, v[k].y - h1, v[k].z - h2, v[k].w - h3);
        }
      }
      fence_async_proxy();
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_full[st]);
    }
#ifdef PB_WGRAD_PROF
    if (blockIdx.x == 0 && ptid == 0) printf("B producer: total %lld | wait b_empty %lld\n", clock64() - pt_start, pt[0]);
#endif
  }

  // ---- teardown ------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

// partial[p][slot] summed over p = grp, grp + 8, ... in ascending order (20 loads in flight)
__device__ __forceinline__ float sum_partials(const float *__restrict__ partial, int n_partial,
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

// Fixed-order sum of the per-CTA partials + the SGD update.  Block = 32 slots x 8 partial
// groups; slot = col*128 + row over the weight columns (fy, c) [the lo column of a stacked
// layout is added to its hi column], then the bias slots.
template <class C>
__global__ void wgrad_umma_reduce_kernel(const float *__restrict__ partial, int n_partial,
                                         const float *Wt, float *out, const float *__restrict__ lr,
                                         int mode) {
  __shared__ float red[8][32];
  chain_wait();
  chain_release();
  const int lane_ = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int slot = blockIdx.x * 32 + lane_;
  constexpr int WSLOTS = C::F * C::IC * 128;
  int w = -1;
  float part = 0.0f;
  if (slot < WSLOTS) {
    const int row = slot & 127, wc = slot >> 7;   // wc = fy*IC + c
    const int fy = wc / C::IC, c = wc - fy * C::IC;
    const int oc = row / C::F, fx = C::F - 1 - (row - oc * C::F);
    if (row < C::ROWS) {
      w = oc * C::K1 + c * C::F * C::F + fy * C::F + fx;
      const int col = fy * C::UN + c;
      part = sum_partials(partial, n_partial, C::ROW, col * 128 + row, grp);
      if (C::STACK) part += sum_partials(partial, n_partial, C::ROW, (col + C::IC) * 128 + row, grp);
    }
  } else if (slot < WSLOTS + C::OC) {
    const int oc = slot - WSLOTS;
    w = oc * C::K1 + C::K1 - 1;
    for (int h = 0; h < C::AQ; ++h)
      part += sum_partials(partial, n_partial, C::ROW, C::NPAD * 128 + h * 32 + oc, grp);
  }
  red[grp][lane_] = part;
  __syncthreads();
  if (grp != 0 || w < 0) return;
  float acc = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) acc += red[j][lane_];
  apply_update(mode, out, Wt, w, lr[0], acc);
}

template <class C>
int launch(pb_context *ctx, const float *I, const float *Wt, const float *G, float *out,
           const float *lr, int mode, int batch) {
  auto kern = conv_wgrad_umma_kernel<C>;
  static bool configured = false;
  if (!configured) {
    PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    configured = true;
  }
  const int n_items = batch * C::IPS;
  const int grid = std::min(n_items, ctx->num_sms);
  if (ensure_workspace(ctx, (size_t)grid * C::ROW)) return 1;
  PB_CUDA(launch_chained(ctx, kern, dim3(grid), dim3(C::THREADS), C::SMEM, I, G, ctx->workspace, batch));
  PB_LAUNCHED(ctx);
  constexpr int SLOTS = C::F * C::IC * 128 + C::OC;
  PB_CUDA(launch_forked(ctx, wgrad_umma_reduce_kernel<C>, dim3(pb_div_up(SLOTS, 32)), dim3(256), 0,
                         ctx->workspace, grid, Wt, out, lr, mode));
  PB_LAUNCHED(ctx);
  return 0;
}

}  // namespace

// Returns -1 outside the envelope (conv_wgrad_mma.cu / conv_tiled.cu then run).
int conv_weight_update_umma(pb_context *ctx, const pb_layer_desc *l, const float *I, const float *Wt,
                            const float *G, float *out, const float *lr, int mode, int64_t batch) {
  static const bool on = [] {
    const char *e = getenv("PB_DISABLE_UMMA"), *f = getenv("PB_WGRAD_UMMA");
    return !(e && e[0] == '1') && !(f && f[0] == '0');
  }();
  if (!on || l->stride != 1 || l->filter_width != 5 || l->filter_height != 5 || l->padding != 2 ||
      l->width != l->height || batch < 64 || batch > (1 << 22))
    return -1;
  if ((((uintptr_t)I) | ((uintptr_t)G)) & 15) return -1;
  const int w = (int)l->width, d = (int)l->depth, f = (int)l->num_filters;
  if (w == 16 && d == 16 && f == 20) return launch<GCfg<16, 20, 16, 16, false, 2, 8>>(ctx, I, Wt, G, out, lr, mode, (int)batch);
  if (w == 8 && d == 20 && f == 20) return launch<GCfg<20, 20, 8, 8, false, 2, 8>>(ctx, I, Wt, G, out, lr, mode, (int)batch);
  // The RGB layer (3 -> 16 channels) stays with the mma.sync kernel of conv_wgrad_mma.cu (58 us).
  // Both tensor-memory mappings were built and measured on B200 and are bound by PRODUCING the
  // tensor-memory operand with a handful of warps (LDS + split + tcgen05.st + wait::st per image
  // row, ~350 cycles per row and lane quadrant against ~200 cycles of MMA issue):
  //   gradient side in tensor memory (GCfg<3, 16, 32, 16, true, 3, 6>: 80 rows x 40 columns per
  //   image row, hi|lo of the image stacked along N)                                   77 us
  //   image side in tensor memory (75 rows = (c, fy, fx) x 32 columns, conflict-free tile
  //   pitches 5 / 25 mod 32, hi|lo of the gradient stacked along N), 3 / 5 warps per
  //   lane quadrant                                                                66 / 62 us
  return -1;
}

}  // namespace pb
