// tcgen05 probe for the bf16x3 convolution variant of DESIGN.md section 9 (development
// tool, not part of the product library; written without a GPU, to be run first thing when
// one is available).  Checks on a B200 that kind::f16 with bf16 operands follows the same
// shared-memory conventions the 3xTF32 kernels rely on, with 8 elements per 16-byte unit and
// K = 16 per MMA:
//   1. K-major, no-swizzle canonical layout: unit u (8 consecutive k) of row r lives at
//      base + u*LBO + (r/8)*SBO + (r%8)*16;
//   2. a descriptor start shifted by s*16 bytes addresses the operand shifted by s rows;
//   3. SBO different from 128 bytes;
// and measures the issue cost of the three MMAs per K = 16 block the variant would use,
// A0 x [B0|B1|B2] (N = 72 -> 80), A1 x [B0|B1] (N = 48), A2 x [B0] (N = 24 -> 32) at M = 128 (N
// must be a multiple of 16 there), against the model (M+N)/4 cycles each = 136 (the 3xTF32
// pair, N = 48 and 32, costs 168 per 16 K).
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bf16_probe umma_bf16_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstring>
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
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell); layout_type 0: no swizzle
  return d;
}
// instruction descriptor: D = f32 (1 << 4), A = B = bf16 (format 1 at bits 7 and 10), K-major
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
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
  int M, N, KB;          // KB = number of K=16 blocks
  int lbo_a, sbo_a, lbo_b, sbo_b;  // bytes
  int kb_stride_a, kb_stride_b;    // bytes between consecutive K=16 blocks
  int shift_rows;                  // descriptor start of A shifted by this many rows (x16 B)
  int a_bytes, b_bytes;
};

// a_img / b_img: raw shared-memory images prepared on the host.  M = 128 only.
__global__ void __launch_bounds__(128, 1)
probe_kernel(const uint32_t *__restrict__ a_img, const uint32_t *__restrict__ b_img, float *D,
             Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint32_t *a_s = reinterpret_cast<uint32_t *>(smem);
  uint32_t *b_s = reinterpret_cast<uint32_t *>(smem + ((p.a_bytes + 127) / 128) * 128);
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
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  const uint32_t idesc = make_idesc_bf16(p.M, p.N);
  const uint32_t a0 = smem_u32(a_s) + p.shift_rows * 16, b0 = smem_u32(b_s);
  if (tid == 0) {
    for (int kb = 0; kb < p.KB; ++kb)
      umma_bf16(tmem_d, make_desc(a0 + kb * p.kb_stride_a, p.lbo_a, p.sbo_a),
                make_desc(b0 + kb * p.kb_stride_b, p.lbo_b, p.sbo_b), idesc, kb > 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < p.N; c0 += 8) {
    uint32_t v[8];
    const uint32_t addr = tmem_d + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]),
                   "=r"(v[6]), "=r"(v[7])
                 : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[tid * p.N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(256u)
                 : "memory");
}

// Issue cost of the variant's three MMAs per K = 16 block (operands are zeros).
template <int M, int NP, int KB>
__global__ void __launch_bounds__(128, 1) triple_kernel(long long *cycles, int reps) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 1024 / 4; i += blockDim.x) reinterpret_cast<float *>(smem)[i] = 0.0f;
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
  constexpr int N3 = (3 * NP + 15) / 16 * 16, N2 = (2 * NP + 15) / 16 * 16, N1 = (NP + 15) / 16 * 16;
  constexpr uint32_t id3 = make_idesc_bf16(M, N3), id2 = make_idesc_bf16(M, N2),
                     id1 = make_idesc_bf16(M, N1);
  if (warp == 0) {
    // image planes as the convolution kernel lays them out: 8-pixel strips of a padded image
    constexpr uint32_t LBO_A = 6400, SBO_A = 320, LBO_B = N3 * 16;
    const uint64_t da0 = make_desc(smem_u32(smem), LBO_A, SBO_A);
    const uint64_t da1 = make_desc(smem_u32(smem) + 30 * 1024, LBO_A, SBO_A);
    const uint64_t da2 = make_desc(smem_u32(smem) + 60 * 1024, LBO_A, SBO_A);
    const uint64_t db0 = make_desc(smem_u32(smem) + 96 * 1024, LBO_B, 128);
    const bool leader = elect_one();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        const uint64_t sh = (uint64_t)(kb % 5 + (kb / 5) * 20);
        const uint64_t db = db0 + (uint64_t)((kb % 4) * ((2 * LBO_B) >> 4));
        if (leader) {
          umma_bf16(tmem_d, da0 + sh, db, id3, 1);
          umma_bf16(tmem_d, da1 + sh, db, id2, 1);
          umma_bf16(tmem_d, da2 + sh, db, id1, 1);
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512u)
                 : "memory");
}

template <int M, int NP, int KB>
static void run_triple(const char *name) {
  long long *d_cyc;
  CK(cudaMalloc(&d_cyc, sizeof(long long)));
  auto k = triple_kernel<M, NP, KB>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 132 * 1024));
  long long c[2];
  const int reps[2] = {16, 144};
  for (int i = 0; i < 2; ++i) {
    k<<<1, 128, 132 * 1024>>>(d_cyc, reps[i]);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&c[i], d_cyc, sizeof(long long), cudaMemcpyDeviceToHost));
  }
  const int n3 = (3 * NP + 15) / 16 * 16, n2 = (2 * NP + 15) / 16 * 16, n1 = (NP + 15) / 16 * 16;
  printf("triple %-10s M=%3d N=%d/%d/%d KB=%d : %.2f cyc per K=16 block (model %.1f)\n", name, M, n3,
         n2, n1, KB, (double)(c[1] - c[0]) / (128.0 * KB), (M + n3) / 4.0 + (M + n2) / 4.0 + (M + n1) / 4.0);
  cudaFree(d_cyc);
}

struct Case {
  const char *name;
  int N, KB, sbo_a, shift, rows_a;
};

static float small_rand() { return (float)((rand() % 33) - 16) / 8.0f; }  // exact in bf16

static uint16_t bf16_bits(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  return (uint16_t)(u >> 16);  // exact for the values above
}

static int run_case(const Case &c) {
  const int M = 128, K = 16 * c.KB, units = 2 * c.KB;  // a unit = 8 consecutive k = 16 bytes
  const int groups_a = (c.rows_a + 7) / 8;
  const int lbo_a = groups_a * c.sbo_a + 128;
  const int a_bytes = units * lbo_a;
  const int sbo_b = 128, lbo_b = (c.N / 8) * sbo_b, b_bytes = units * lbo_b;
  std::vector<float> A((size_t)c.rows_a * K), B((size_t)c.N * K);
  for (auto &v : A) v = small_rand();
  for (auto &v : B) v = small_rand();
  std::vector<uint16_t> a_img(a_bytes / 2, 0), b_img(b_bytes / 2, 0);
  for (int r = 0; r < c.rows_a; ++r)
    for (int k = 0; k < K; ++k) {
      const int off = (k / 8) * lbo_a + (r / 8) * c.sbo_a + (r % 8) * 16 + (k % 8) * 2;
      a_img[off / 2] = bf16_bits(A[(size_t)r * K + k]);
    }
  for (int n = 0; n < c.N; ++n)
    for (int k = 0; k < K; ++k) {
      const int off = (k / 8) * lbo_b + (n / 8) * sbo_b + (n % 8) * 16 + (k % 8) * 2;
      b_img[off / 2] = bf16_bits(B[(size_t)n * K + k]);
    }
  uint32_t *d_a, *d_b;
  float *d_D;
  CK(cudaMalloc(&d_a, a_bytes));
  CK(cudaMalloc(&d_b, b_bytes));
  CK(cudaMalloc(&d_D, sizeof(float) * 128 * c.N));
  CK(cudaMemcpy(d_a, a_img.data(), a_bytes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_b, b_img.data(), b_bytes, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_D, 0, sizeof(float) * 128 * c.N));
  Params p{};
  p.M = M; p.N = c.N; p.KB = c.KB;
  p.lbo_a = lbo_a; p.sbo_a = c.sbo_a; p.lbo_b = lbo_b; p.sbo_b = sbo_b;
  p.kb_stride_a = 2 * lbo_a; p.kb_stride_b = 2 * lbo_b;
  p.shift_rows = c.shift; p.a_bytes = a_bytes; p.b_bytes = b_bytes;
  const size_t smem = ((a_bytes + 127) / 128) * 128 + b_bytes + 256;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  probe_kernel<<<1, 128, smem>>>(d_a, d_b, d_D, p);
  CK(cudaDeviceSynchronize());
  std::vector<float> D(128 * c.N);
  CK(cudaMemcpy(D.data(), d_D, sizeof(float) * 128 * c.N, cudaMemcpyDeviceToHost));
  int bad = 0;
  double maxerr = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < c.N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) {
        // the row the shifted descriptor addresses, looked up by byte address in the image
        const int off = (k / 8) * lbo_a + c.shift * 16 + (m / 8) * c.sbo_a + (m % 8) * 16 + (k % 8) * 2;
        uint32_t u = (uint32_t)a_img[off / 2] << 16;
        float av;
        memcpy(&av, &u, 4);
        ref += (double)av * (double)B[(size_t)n * K + k];
      }
      const double err = std::fabs(ref - D[m * c.N + n]);
      if (err > maxerr) maxerr = err;
      if (err > 1e-3) {
        if (bad < 5) printf("   mismatch m=%d n=%d got %f want %f\n", m, n, D[m * c.N + n], ref);
        ++bad;
      }
    }
  printf("%-18s N=%3d K=%4d sbo_a=%4d shift=%2d : %s (bad=%d maxerr=%.3g)\n", c.name, c.N, K,
         c.sbo_a, c.shift, bad == 0 ? "OK" : "MISMATCH", bad, maxerr);
  cudaFree(d_a); cudaFree(d_b); cudaFree(d_D);
  return bad;
}

int main() {
  srand(1);
  int bad = 0;
  const Case cases[] = {
      {"basic", 32, 4, 128, 0, 128},       {"shift1", 32, 4, 128, 1, 136},
      {"shift5", 32, 4, 128, 5, 136},      {"sbo288", 32, 4, 288, 0, 128},
      {"sbo288+shift3", 32, 4, 288, 3, 128}, {"n16", 16, 4, 128, 0, 128},
      {"n48", 48, 4, 128, 0, 128},         {"n80", 80, 4, 128, 0, 128},
  };
  for (const Case &c : cases) bad += run_case(c);
  run_triple<128, 24, 25>("conv4 NP=24");
  run_triple<128, 16, 25>("conv1 NP=16");
  printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
  return bad ? 1 : 0;
}
