// Decode which shared-memory word a kind::tf32 MMA reads for B element (n, k) under a
// given descriptor (development tool).  A = identity in tensor memory: D[k][n] = B(n, k).
// B words hold their own index (two passes: low 10 bits, then the rest).
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
struct Cfg { int lbo, sbo, layout, b_mn, a_mn, N, start, a_smem; };
__global__ void __launch_bounds__(128, 1) k(float *D, Cfg c, int pass) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  float *b = reinterpret_cast<float *>(smem);
  const int words = 192 * 1024 / 4;
  for (int i = tid; i < words; i += 128) b[i] = pass == 0 ? (float)(i & 1023) : (float)(i >> 10);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s, lane_base = (uint32_t)(warp * 32) << 16;
  uint32_t a[8];
  for (int j = 0; j < 8; ++j) a[j] = __float_as_uint(tid == j ? 1.0f : 0.0f);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem + lane_base + 256),
               "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0 && elect_one()) {
    uint64_t d = 0;
    d |= (uint64_t)(((smem_u32(b) + c.start) >> 4) & 0x3FFF);
    d |= (uint64_t)((c.lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((c.sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)c.layout << 61;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)c.b_mn << 16) |
                           ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem), "r"(tmem + 256), "l"(d), "r"(idesc), "r"(0) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n\t.reg .pred P1;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DN;\n\tbra WL;\n\tDN:\n\t}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < c.N; c0 += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(tmem + lane_base + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[tid * 256 + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
int main() {
  float *dD; CK(cudaMalloc(&dD, 128 * 256 * 4));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  std::vector<float> D0(128 * 256), D1(128 * 256);
  const Cfg cfgs[] = {
    {2048, 128, 0, 0, 0, 32, 0, 0},     // K-major baseline
    {128, 384, 0, 1, 0, 32, 0, 0},      // MN none: lbo=128 sbo=384
    {384, 128, 0, 1, 0, 32, 0, 0},      // swapped
    {128, 384, 0, 1, 0, 32, 1024, 0},   // start at 1024
    {512, 1024, 0, 1, 0, 32, 0, 0},     // both large
    {384, 256, 6, 1, 0, 32, 0, 0},      // SW32 (layout 6): lbo = n-group, sbo = k-group
    {384, 512, 4, 1, 0, 32, 0, 0},      // SW64 (layout 4)
    {384, 1024, 2, 1, 0, 32, 0, 0},     // SW128 (layout 2)
    {1024, 2048, 2, 1, 0, 64, 0, 0},    // SW128 canonical-ish
  };
  for (const Cfg &c : cfgs) {
    k<<<1, 128, 192 * 1024>>>(dD, c, 0); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D0.data(), dD, D0.size() * 4, cudaMemcpyDeviceToHost));
    k<<<1, 128, 192 * 1024>>>(dD, c, 1); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D1.data(), dD, D1.size() * 4, cudaMemcpyDeviceToHost));
    printf("cfg lbo=%d sbo=%d layout=%d b_mn=%d N=%d start=%d : byte offset read for (n,k)\n", c.lbo, c.sbo, c.layout, c.b_mn, c.N, c.start);
    for (int kk = 0; kk < 8; ++kk) {
      printf("  k=%d:", kk);
      for (int n = 0; n < (c.N < 40 ? c.N : 40); ++n) printf(" %5d", 4 * ((int)D0[kk * 256 + n] + 1024 * (int)D1[kk * 256 + n]));
      printf("\n");
    }
  }
  return 0;
}
