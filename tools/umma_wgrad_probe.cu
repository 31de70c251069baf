// tcgen05 probe for the convolution weight-gradient formulation (development tool, not
// part of the product library).
//
// Questions answered on a B200 before conv_wgrad_umma.cu relies on them:
//   1. kind::tf32 MMA with the A operand in TENSOR MEMORY (tcgen05.st by the warps that own
//      the lane quadrants; lane = row, column = k);
//   2. (answered: MN-major B with kind::tf32 reads zeros on B200 for every layout tried,
//      tools/umma_mn_decode.cu; the B image is therefore K-major, see probe_kernel)
//         base + (n % 4) * 4 + k * 16 + (n / 4) * SBO            (bytes)
//      (cute's canonical ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO)) with T = 4), with ANY
//      16-byte-multiple SBO, including one that makes the n-groups overlap in memory;
//   3. the whole weight-gradient contraction of one sample of a square same-convolution:
//         A[(fx,oc)][(y,X)] = G[oc][y][X-fx]   (tensor memory, hi and lo parts)
//         B[(fy,c)][(y,X)]  = Ipad[c][y+fy][X] (shared memory, rows [Y][chunk][X][4 ch])
//         D[(fx,oc)][(fy,c)] = sum over (y,X)  = dE/dW[oc][c][fy][fx]
//      with three MMAs per K step (hi*hi + lo*hi + hi*lo) into ONE accumulator;
//   4. cycles per K step of that sequence for the three CIFAR shapes.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_wgrad_probe umma_wgrad_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                       \
    }                                                                                \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// kind::tf32, fp32 accumulate, A K-major (tensor memory), B MN-major (bit 16)
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db,
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
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float *v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
               "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
               "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

constexpr int kDCols = 256;     // accumulator columns reserved
constexpr int kACol = 256;      // A operand columns start here (hi: +0..7, lo: +8..15)

struct Shape {
  int IC, OC, W, H, F;  // F x F filter, pad F/2
  int PK, N, YR;        // derived: K pitch (W + 8), MMA N, shared-memory rows per quad plane
};

// B image (one of hi / lo): [quad q = X/4][row Y][channel c][4 pixels], 16 bytes per (q,Y,c):
// K-major no-swizzle operand with rows n = fy*IC + c contiguous (SBO = 128 B) and the two
// 4-pixel chunks of a K = 8 step one quad plane apart (LBO = plane size).
// Ipad[c][Y][X] = I[c][Y-2][X-4];   A[(fx,oc)][(y,X)] = G[oc][y][X-fx-2].
// mode 1: one sample's weight gradient;  mode 2: timing of the 3-MMA K step
__global__ void __launch_bounds__(128, 1)
probe_kernel(const float *__restrict__ G, const float *__restrict__ I, float *D, long long *cycles,
             Shape s, int mode, int reps, int tm, int nacc) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int NQ = s.PK / 4;
  const int plane_words = NQ * s.YR * s.IC * 4;  // one of hi / lo
  const int QS = s.YR * s.IC * 16;               // bytes per quad plane
  float *b_hi = reinterpret_cast<float *>(smem);
  float *b_lo = b_hi + plane_words;
  for (int i = tid; i < plane_words; i += blockDim.x) {
    const int j = i & 3, c = (i >> 2) % s.IC, Y = ((i >> 2) / s.IC) % s.YR, q = (i >> 2) / (s.IC * s.YR);
    const int y = Y - 2, x = 4 * q + j - 4;
    float v = 0.0f;
    if (y >= 0 && y < s.H && x >= 0 && x < s.W) v = I[(c * s.H + y) * s.W + x];
    const float h = tf32_hi(v);
    b_hi[i] = h;
    b_lo[i] = v - h;
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_s)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t idesc = make_idesc(mode >= 2 ? tm : 128, s.N, 0);
  float *a_s = b_lo + plane_words;  // mode 3: 128 x 8 A image, K-major: [k/4][row][4]
  if (mode == 3) for (int i = tid; i < 2 * 128 * 8; i += blockDim.x) a_s[i] = 0.001f * (float)(i % 7);
  const int row = tid;                    // tensor-memory lane = A row = D row
  const int fx = row / s.OC, oc = row % s.OC;
  const bool row_live = row < s.F * s.OC;
  uint32_t phase = 0;
  const int ksteps_row = s.PK / 8;
  const int total = (mode == 1) ? s.H * ksteps_row : reps;
  long long t0 = 0;
  if (mode >= 2) {
    float a[8];
    for (int j = 0; j < 8; ++j) a[j] = 0.001f * (float)((row + j) % 7);
    tmem_st8(tmem + lane_base + kACol, a);
    tmem_st8(tmem + lane_base + kACol + 8, a);
    tmem_st_wait();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    t0 = clock64();
  }
  if (mode >= 2) {
    // tight issue loop: descriptors precomputed, one elected lane issues reps x 3 MMAs
    if (warp == 0) {
      const uint64_t dbh0 = make_desc(smem_u32(b_hi), QS, 128), dbl0 = make_desc(smem_u32(b_lo), QS, 128);
      const uint64_t dah = make_desc(smem_u32(a_s), 2048, 128), dal = make_desc(smem_u32(a_s) + 4096, 2048, 128);
      if (elect_one()) {
#pragma unroll 4
        for (int ks = 0; ks < reps; ++ks) {
          const uint32_t dcol = tmem + (uint32_t)((ks & (nacc - 1)) * 128);
          const uint64_t step = (uint64_t)((ks & 7) * s.IC);  // next image row: +IC rows of 16 B
          if (mode == 3) {
            umma_ss(dcol, dah, dbh0 + step, idesc, ks >= nacc);
            umma_ss(dcol, dal, dbh0 + step, idesc, 1);
            umma_ss(dcol, dah, dbl0 + step, idesc, 1);
          } else {
            umma_ts(dcol, tmem + kACol, dbh0 + step, idesc, ks >= nacc);
            umma_ts(dcol, tmem + kACol + 8, dbh0 + step, idesc, 1);
            umma_ts(dcol, tmem + kACol, dbl0 + step, idesc, 1);
          }
        }
        umma_commit(&bar);
      }
      __syncwarp();
    }
  } else
  for (int ks = 0; ks < total; ++ks) {
    const int y = (mode == 1) ? ks / ksteps_row : (ks % s.H);
    const int xb = ks % ksteps_row;
    if (mode == 1) {
      float ah[8], al[8];
      for (int j = 0; j < 8; ++j) {
        const int x = xb * 8 + j - fx - 2;
        float v = 0.0f;
        if (row_live && x >= 0 && x < s.W) v = G[(oc * s.H + y) * s.W + x];
        ah[j] = tf32_hi(v);
        al[j] = v - ah[j];
      }
      tmem_st8(tmem + lane_base + kACol, ah);
      tmem_st8(tmem + lane_base + kACol + 8, al);
      tmem_st_wait();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0) {
      // B start: quad plane 2*xb, padded row y (row n = fy*IC + c is n*16 bytes further)
      const uint32_t off = (uint32_t)(2 * xb * QS + y * s.IC * 16);
      const uint64_t dbh = make_desc(smem_u32(b_hi) + off, QS, 128);
      const uint64_t dbl = make_desc(smem_u32(b_lo) + off, QS, 128);
      if (elect_one()) {
        const uint32_t dcol = tmem + (uint32_t)((ks % nacc) * 128);
        if (mode == 3) {
          const uint64_t dah = make_desc(smem_u32(a_s), 2048, 128), dal = make_desc(smem_u32(a_s) + 4096, 2048, 128);
          umma_ss(dcol, dah, dbh, idesc, ks >= nacc);
          umma_ss(dcol, dal, dbh, idesc, 1);
          umma_ss(dcol, dah, dbl, idesc, 1);
        } else {
          umma_ts(dcol, tmem + kACol, dbh, idesc, ks >= nacc);
          umma_ts(dcol, tmem + kACol + 8, dbh, idesc, 1);
          umma_ts(dcol, tmem + kACol, dbl, idesc, 1);
        }
        if (mode == 1 || ks == total - 1) umma_commit(&bar);
      }
      __syncwarp();
    }
    if (mode == 1) {
      mbar_wait(&bar, phase);
      phase ^= 1;
    }
  }
  if (mode >= 2) {
    mbar_wait(&bar, phase);
    phase ^= 1;
    if (tid == 0) cycles[0] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < s.N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + lane_base + c0, v);
    for (int j = 0; j < 16; ++j) D[row * kDCols + c0 + j] = v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u)
                 : "memory");
  }
}

static Shape make_shape(int IC, int OC, int W, int F) {
  Shape s;
  s.IC = IC; s.OC = OC; s.W = W; s.H = W; s.F = F;
  s.PK = W + 8;
  s.N = ((F * IC + 15) / 16) * 16;
  s.YR = W + 4 + 2;  // two slack rows: the MMA's N is rounded up past fy = 4
  return s;
}

int main() {
  float *dG, *dI, *dD;
  long long *dcyc;
  CK(cudaMalloc(&dG, 1 << 20));
  CK(cudaMalloc(&dI, 1 << 20));
  CK(cudaMalloc(&dD, 128 * kDCols * 4));
  CK(cudaMalloc(&dcyc, 8));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  std::vector<float> D(128 * kDCols);
  bool all_ok = true;

  const Shape shapes[3] = {make_shape(16, 20, 16, 5), make_shape(20, 20, 8, 5), make_shape(3, 16, 32, 5)};
  const char *names[3] = {"conv4 16->20 @16", "conv7 20->20 @8", "conv1 3->16 @32"};

  // ---- 3: one sample's weight gradient -------------------------------------------
  for (int si = 0; si < 3; ++si) {
    const Shape s = shapes[si];
    std::vector<float> G(s.OC * s.H * s.W), I(s.IC * s.H * s.W);
    srand(17 + si);
    for (auto &v : G) v = (float)rand() / RAND_MAX - 0.5f;
    for (auto &v : I) v = (float)rand() / RAND_MAX - 0.5f;
    CK(cudaMemcpy(dG, G.data(), G.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dI, I.data(), I.size() * 4, cudaMemcpyHostToDevice));
    const size_t smem = 2 * (size_t)(s.PK / 4) * s.YR * s.IC * 16 + 1024;
    probe_kernel<<<1, 128, smem>>>(dG, dI, dD, dcyc, s, 1, 0, 128, 1);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int oc = 0; oc < s.OC; ++oc)
      for (int c = 0; c < s.IC; ++c)
        for (int fy = 0; fy < s.F; ++fy)
          for (int fx = 0; fx < s.F; ++fx) {
            double ref = 0;
            for (int y = 0; y < s.H; ++y)
              for (int x = 0; x < s.W; ++x) {
                const int iy = y + fy - 2, ix = x + fx - 2;
                if (iy >= 0 && iy < s.H && ix >= 0 && ix < s.W)
                  ref += (double)G[(oc * s.H + y) * s.W + x] * I[(c * s.H + iy) * s.W + ix];
              }
            const double got = D[(fx * s.OC + oc) * kDCols + fy * s.IC + c];
            maxerr = fmax(maxerr, fabs(got - ref));
            maxref = fmax(maxref, fabs(ref));
          }
    const bool ok = maxerr <= 1e-5 * maxref;
    all_ok &= ok;
    printf("wgrad %-18s N=%3d PK=%2d K steps=%3d : %s (max err %.3g of max %.3g = %.2e)\n", names[si], s.N,
           s.PK, s.H * s.PK / 8, ok ? "OK" : "BAD", maxerr, maxref, maxerr / maxref);
    // ---- 4: timing ---------------------------------------------------------------
    const int reps = 4096;
    for (int mode = 2; mode <= 3; ++mode)
      for (int tm = 128; tm >= 64; tm -= 64)
        for (int nacc = 1; nacc <= 2; ++nacc) {
          probe_kernel<<<1, 128, smem + 8192>>>(dG, dI, dD, dcyc, s, mode, reps, tm, nacc);
          CK(cudaDeviceSynchronize());
          long long cyc;
          CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
          printf("  timing: %6.1f cycles per K step (3 MMAs M=%d N=%d, A in %s, %d accumulator%s)\n",
                 (double)cyc / reps, tm, s.N, mode == 2 ? "TMEM" : "smem", nacc, nacc > 1 ? "s" : "");
        }
  }
  printf(all_ok ? "PROBE OK\n" : "PROBE FAILED\n");
  return all_ok ? 0 : 1;
}
