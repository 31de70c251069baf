// tcgen05 probe (development tool, not part of the product library).
//
// Validates on a B200 the shared-memory operand conventions the 3xTF32
// convolution kernels rely on, and measures the MMA issue rate for the small-N
// shapes they use:
//   1. K-major, no-swizzle canonical layout: 16-byte chunk c of row r lives at
//      base + c*LBO + (r/8)*SBO + (r%8)*16;
//   2. a descriptor whose start address is shifted by s*16 bytes addresses the
//      operand shifted by s rows (used for the filter taps of an implicit GEMM);
//   3. SBO different from 128 bytes (row groups = 8-pixel strips of an image
//      with a padded row pitch);
//   4. cycles per tcgen05.mma (kind::tf32, K=8) for M in {64,128}, several N.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                   \
  do {                                                                          \
    cudaError_t e = (x);                                                        \
    if (e != cudaSuccess) {                                                     \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                  \
    }                                                                           \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // layout_type 0: no swizzle
}

__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db,
                                          uint32_t idesc, uint32_t accumulate) {
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
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
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
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

struct Params {
  int M, N, KB;          // KB = number of K=8 blocks
  int lbo_a, sbo_a;      // bytes
  int lbo_b, sbo_b;      // bytes
  int kb_stride_a;       // bytes between consecutive K=8 blocks of A
  int kb_stride_b;
  int shift_rows;        // descriptor start shifted by this many rows (x16 B)
  int a_bytes, b_bytes;  // bytes of A / B images to copy into shared memory
  int reps;              // timing repetitions
  int nacc;              // independent accumulators used round-robin in the timing pass
};

// a_img / b_img are raw shared-memory images prepared on the host.
__global__ void __launch_bounds__(128, 1)
probe_kernel(const float *__restrict__ a_img, const float *__restrict__ b_img, float *D,
             long long *cycles, Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  float *a_s = reinterpret_cast<float *>(smem);
  float *b_s = reinterpret_cast<float *>(smem + ((p.a_bytes + 127) / 128) * 128);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < p.a_bytes / 4; i += blockDim.x) a_s[i] = a_img[i];
  for (int i = tid; i < p.b_bytes / 4; i += blockDim.x) b_s[i] = b_img[i];
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_s)),
                 "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // generic-proxy smem writes -> visible to the async proxy (tensor core reads)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  const uint32_t idesc = make_idesc(p.M, p.N);
  const uint32_t a0 = smem_u32(a_s) + p.shift_rows * 16, b0 = smem_u32(b_s);

  uint32_t phase = 0;
  if (tid == 0) {
    // correctness pass
    for (int kb = 0; kb < p.KB; ++kb)
      umma_tf32(tmem_d, make_desc(a0 + kb * p.kb_stride_a, p.lbo_a, p.sbo_a),
                make_desc(b0 + kb * p.kb_stride_b, p.lbo_b, p.sbo_b), idesc, kb > 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, phase);
  phase ^= 1;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue: warp w reads TMEM lanes 32w..32w+31
  {
    const int row = tid;  // lane == row for M=128; for M=64 dump the raw lanes
    for (int c0 = 0; c0 < p.N; c0 += 8) {
      uint32_t v[8];
      const uint32_t addr = tmem_d + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
            "=r"(v[7])
          : "r"(addr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) D[row * p.N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // timing pass: the whole of warp 0 runs the loop, one elected lane issues
  if (warp == 0) {
    const uint64_t da0 = make_desc(a0, p.lbo_a, p.sbo_a), db0 = make_desc(b0, p.lbo_b, p.sbo_b);
    const uint64_t da_step = (uint64_t)(p.kb_stride_a >> 4), db_step = (uint64_t)(p.kb_stride_b >> 4);
    const long long t0 = clock64();
    for (int r = 0; r < p.reps; ++r) {
      if (elect_one()) {
#pragma unroll 4
        for (int kb = 0; kb < p.KB; ++kb)
          umma_tf32(tmem_d + (uint32_t)((kb % p.nacc) * 64), da0 + kb * da_step, db0 + kb * db_step, idesc, 1);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, phase);
    const long long t1 = clock64();
    if (tid == 0) cycles[0] = t1 - t0;
  } else {
    mbar_wait(&bar, phase);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(256u)
                 : "memory");
  }
}

static int g_nacc = 1;
static float tf32_rand() {
  // small integers / 8: exactly representable in tf32, products and sums exact in fp32
  return (float)((rand() % 33) - 16) / 8.0f;
}


// Lean issue-rate kernel: everything the issuing lane needs is a compile-time
// constant or a loop-invariant register; A/B contents are irrelevant (zeros).
template <int M, int N, int NACC, int KB, bool SAMEA>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long *cycles, int reps) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<float *>(smem)[i] = 0.0f;
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) |
                             ((uint32_t)(M >> 4) << 24);
  if (warp == 0) {
    constexpr uint32_t LBO_A = 2048 + 128, LBO_B = (N / 8) * 128;
    const uint64_t da0 = make_desc(smem_u32(smem), LBO_A, 128);
    const uint64_t db0 = make_desc(smem_u32(smem) + 40 * 1024, LBO_B, 128);
    const bool leader = elect_one();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        const uint64_t da = da0 + (uint64_t)((SAMEA ? 0 : kb) * ((2 * LBO_A) >> 4));
        const uint64_t db = db0 + (uint64_t)((kb % 2) * ((2 * LBO_B) >> 4));
        if (leader) umma_tf32(tmem_d + (uint32_t)((kb % NACC) * 64), da, db, idesc, 1);
      }
    }
    if (leader) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (tid == 0) cycles[0] = t1 - t0;
  } else {
    mbar_wait(&bar, 0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512u)
                 : "memory");
}

template <int M, int N, int NACC, int KB, bool SAMEA>
static void run_rate(const char *name) {
  long long *d_cyc;
  CK(cudaMalloc(&d_cyc, sizeof(long long)));
  auto k = rate_kernel<M, N, NACC, KB, SAMEA>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  long long c[2];
  const int reps[2] = {64, 576};
  for (int i = 0; i < 2; ++i) {
    k<<<1, 128, 64 * 1024>>>(d_cyc, reps[i]);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&c[i], d_cyc, sizeof(long long), cudaMemcpyDeviceToHost));
  }
  printf("rate %-10s M=%3d N=%3d nacc=%d KB=%d sameA=%d : %.2f cyc/MMA\n", name, M, N, NACC, KB,
         (int)SAMEA, (double)(c[1] - c[0]) / (512.0 * KB));
  cudaFree(d_cyc);
}


// Alternating-shape pairs as the convolution kernel issues them: hi pass (N1)
// then lo pass (N2) onto the same accumulator columns.
template <int M, int N1, int N2, int KB>
__global__ void __launch_bounds__(128, 1) pair_kernel(long long *cycles, int reps) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<float *>(smem)[i] = 0.0f;
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  constexpr uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N1 >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  constexpr uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N2 >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  if (warp == 0) {
    constexpr uint32_t LBO_A = 6400, LBO_B = N1 * 16, SBO_A = 320;
    const uint64_t da0 = make_desc(smem_u32(smem), LBO_A, SBO_A);
    const uint64_t dl0 = make_desc(smem_u32(smem) + 30 * 1024, LBO_A, SBO_A);
    const uint64_t db0 = make_desc(smem_u32(smem) + 64 * 1024, LBO_B, 128);
    const bool leader = elect_one();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        const uint64_t da = da0 + (uint64_t)(kb % 5 + (kb / 5) * 20);
        const uint64_t dl = dl0 + (uint64_t)(kb % 5 + (kb / 5) * 20);
        const uint64_t db = db0 + (uint64_t)((kb % 4) * ((2 * LBO_B) >> 4));
        if (leader) {
          umma_tf32(tmem_d, da, db, idesc1, 1);
          umma_tf32(tmem_d, dl, db, idesc2, 1);
        }
      }
    }
    if (leader) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (tid == 0) cycles[0] = t1 - t0;
  } else {
    mbar_wait(&bar, 0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512u) : "memory");
}

template <int M, int N1, int N2, int KB>
static void run_pair(const char *name) {
  long long *d_cyc;
  CK(cudaMalloc(&d_cyc, sizeof(long long)));
  auto k = pair_kernel<M, N1, N2, KB>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  long long c[2];
  const int reps[2] = {16, 144};
  for (int i = 0; i < 2; ++i) {
    k<<<1, 128, 100 * 1024>>>(d_cyc, reps[i]);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&c[i], d_cyc, sizeof(long long), cudaMemcpyDeviceToHost));
  }
  printf("pair %-10s M=%3d N1=%3d N2=%3d KB=%d : %.2f cyc/pair (model %.1f)\n", name, M, N1, N2, KB,
         (double)(c[1] - c[0]) / (128.0 * KB), (M + N1) / 4.0 + (M + N2) / 4.0);
  cudaFree(d_cyc);
}

struct Case {
  const char *name;
  int M, N, KB, sbo_a, shift, rows_a;  // rows_a: rows present in the A image (>= M + shift)
};

static int run_case(const Case &c, int reps, bool verbose, long long *cyc_out = nullptr) {
  const int K = 8 * c.KB;
  const int chunks = 2 * c.KB;
  // A image: chunk-plane layout.  Row r (0..rows_a) chunk ch at
  //   ch*LBO + (r/8)*SBO + (r%8)*16 ; planes sized to hold rows_a rows.
  const int groups_a = (c.rows_a + 7) / 8;
  const int lbo_a = groups_a * c.sbo_a + 128;
  const int a_bytes = chunks * lbo_a;
  const int sbo_b = 128;
  const int lbo_b = (c.N / 8) * sbo_b;
  const int b_bytes = chunks * lbo_b;
  std::vector<float> A((size_t)c.rows_a * K), B((size_t)c.N * K);
  for (auto &v : A) v = tf32_rand();
  for (auto &v : B) v = tf32_rand();
  std::vector<float> a_img(a_bytes / 4, 0.0f), b_img(b_bytes / 4, 0.0f);
  for (int r = 0; r < c.rows_a; ++r)
    for (int k = 0; k < K; ++k) {
      const int ch = k / 4;
      const int off = ch * lbo_a + (r / 8) * c.sbo_a + (r % 8) * 16 + (k % 4) * 4;
      a_img[off / 4] = A[(size_t)r * K + k];
    }
  for (int n = 0; n < c.N; ++n)
    for (int k = 0; k < K; ++k) {
      const int ch = k / 4;
      const int off = ch * lbo_b + (n / 8) * sbo_b + (n % 8) * 16 + (k % 4) * 4;
      b_img[off / 4] = B[(size_t)n * K + k];
    }
  float *d_a, *d_b, *d_D;
  long long *d_cyc;
  CK(cudaMalloc(&d_a, a_bytes));
  CK(cudaMalloc(&d_b, b_bytes));
  CK(cudaMalloc(&d_D, sizeof(float) * 128 * c.N));
  CK(cudaMalloc(&d_cyc, sizeof(long long)));
  CK(cudaMemcpy(d_a, a_img.data(), a_bytes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_b, b_img.data(), b_bytes, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_D, 0, sizeof(float) * 128 * c.N));
  Params p{};
  p.M = c.M; p.N = c.N; p.KB = c.KB;
  p.lbo_a = lbo_a; p.sbo_a = c.sbo_a; p.lbo_b = lbo_b; p.sbo_b = sbo_b;
  p.kb_stride_a = 2 * lbo_a; p.kb_stride_b = 2 * lbo_b;
  p.shift_rows = c.shift; p.a_bytes = a_bytes; p.b_bytes = b_bytes; p.reps = reps; p.nacc = g_nacc;
  const size_t smem = ((a_bytes + 127) / 128) * 128 + b_bytes + 256;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  probe_kernel<<<1, 128, smem>>>(d_a, d_b, d_D, d_cyc, p);
  CK(cudaDeviceSynchronize());
  std::vector<float> D(128 * c.N);
  long long cyc = 0;
  CK(cudaMemcpy(D.data(), d_D, sizeof(float) * 128 * c.N, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost));
  // Reference.  A row for MMA row m: the descriptor start is shifted by
  // `shift` 16-byte units, i.e. row index within group arithmetic:
  //   address(m) = shift*16 + (m/8)*SBO + (m%8)*16.
  // With SBO == 128 this is plain row m+shift; otherwise row (m/8)*8 + m%8 + shift
  // may cross into the pad bytes of the group, which the image holds as rows too
  // only when SBO == 128.  For SBO != 128 the shifted row is looked up by byte address.
  int bad = 0;
  double maxerr = 0;
  if (c.M == 128) {
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < c.N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) {
          const int ch = k / 4;
          const int off = ch * lbo_a + c.shift * 16 + (m / 8) * c.sbo_a + (m % 8) * 16 + (k % 4) * 4;
          ref += (double)a_img[off / 4] * (double)B[(size_t)n * K + k];
        }
        const double err = std::fabs(ref - D[m * c.N + n]);
        if (err > maxerr) maxerr = err;
        if (err > 1e-3) {
          if (bad < 5 && verbose) printf("   mismatch m=%d n=%d got %f want %f\n", m, n, D[m * c.N + n], ref);
          ++bad;
        }
      }
  } else {
    // M=64: report which TMEM lanes hold which rows (first column only)
    std::vector<double> ref(64);
    for (int m = 0; m < 64; ++m) {
      double r = 0;
      for (int k = 0; k < K; ++k) {
        const int ch = k / 4;
        const int off = ch * lbo_a + c.shift * 16 + (m / 8) * c.sbo_a + (m % 8) * 16 + (k % 4) * 4;
        r += (double)a_img[off / 4] * (double)B[k];
      }
      ref[m] = r;
    }
    printf("   M=64 lane->row map (lanes whose col 0 matches a row): ");
    int matched = 0;
    for (int lane = 0; lane < 128; ++lane)
      for (int m = 0; m < 64; ++m)
        if (std::fabs(D[lane * c.N] - ref[m]) < 1e-4 && std::fabs(ref[m]) > 1e-6) {
          if (lane % 16 == 0) printf("[lane %d=row %d] ", lane, m);
          ++matched;
          break;
        }
    printf(" matched=%d\n", matched);
  }
  printf("%-28s M=%3d N=%3d K=%4d sbo_a=%4d shift=%2d : %s (bad=%d maxerr=%.3g)  %.1f cyc/MMA\n",
         c.name, c.M, c.N, K, c.sbo_a, c.shift, bad == 0 ? "OK" : "MISMATCH", bad, maxerr,
         (double)cyc / ((double)reps * c.KB));
  if (cyc_out) *cyc_out = cyc;
  cudaFree(d_a); cudaFree(d_b); cudaFree(d_D); cudaFree(d_cyc);
  return bad;
}

int main() {
  srand(1);
  int bad = 0;
  const Case cases[] = {
      {"basic", 128, 32, 4, 128, 0, 128},
      {"shift1", 128, 32, 4, 128, 1, 136},
      {"shift5", 128, 32, 4, 128, 5, 136},
      {"sbo288", 128, 32, 4, 288, 0, 128},
      {"sbo288+shift3", 128, 32, 4, 288, 3, 128},
      {"sbo160", 128, 32, 4, 160, 0, 128},
      {"n16", 128, 16, 8, 128, 0, 128},
      {"n24", 128, 24, 8, 128, 0, 128},
      {"n48", 128, 48, 8, 128, 0, 128},
      {"n64", 128, 64, 8, 128, 0, 128},
      {"n128", 128, 128, 8, 128, 0, 128},
      {"n32k64", 128, 32, 8, 128, 0, 128},
      {"n256", 128, 256, 8, 128, 0, 128},
      {"m64n32", 64, 32, 8, 128, 0, 128},
      {"m64n64", 64, 64, 8, 128, 0, 128},
  };
  for (int nacc : {1})
  for (const Case &c : cases) {
    if (c.KB != 4) continue;
    if (c.N > 64) continue;
    g_nacc = nacc;
    printf("nacc=%d ", nacc);
    long long c1 = 0, c2 = 0;
    bad += run_case(c, 32, true, &c1);
    bad += run_case(c, 288, false, &c2);
    printf("   -> slope %.2f cyc/MMA, fixed %.0f cyc\n", (double)(c2 - c1) / (256.0 * c.KB),
           (double)c1 - 32.0 * c.KB * (double)(c2 - c1) / (256.0 * c.KB));
  }
  run_pair<128, 48, 32, 25>("conv4fwd");
  run_pair<128, 32, 16, 25>("conv1fwd");
  run_pair<128, 16, 16, 25>("conv1bwd");
  run_pair<64, 48, 24, 25>("conv7");
  run_pair<128, 48, 48, 25>("same48");
  run_rate<128, 16, 1, 8, false>("n16");
  run_rate<128, 32, 1, 8, false>("n32");
  run_rate<128, 48, 1, 8, false>("n48");
  run_rate<128, 64, 1, 8, false>("n64");
  run_rate<128, 128, 1, 8, false>("n128");
  run_rate<128, 32, 2, 8, false>("n32a2");
  run_rate<128, 32, 4, 8, false>("n32a4");
  run_rate<128, 64, 2, 8, false>("n64a2");
  run_rate<128, 64, 4, 8, false>("n64a4");
  run_rate<128, 32, 1, 8, true>("n32sameA");
  run_rate<128, 32, 4, 8, true>("n32a4sameA");
  run_rate<64, 32, 1, 8, false>("m64n32");
  run_rate<64, 32, 4, 8, false>("m64n32a4");
  run_rate<64, 64, 4, 8, false>("m64n64a4");
  printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
  return bad ? 1 : 0;
}
