// Two questions about tcgen05.ld.32x32b on sm_100a, answered on the device:
//  1. may the lane field of the address be offset inside the warp's quadrant (thread i then
//     reads lane base + d + i)?  Every lane of the allocation is filled with its own index
//     first; the values read back with d = 0..4 are printed.
//  2. how fast are the loads: cycles for a train of x8 loads from one warp, from the four warps
//     of the four quadrants at once, and from eight warps (two per quadrant).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tmem_ld_probe tmem_ld_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

__global__ void __launch_bounds__(256, 1) probe(int *out, long long *cyc, int max_lane_offset) {
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)),
                 "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_ptr;
  const int quad = warp & 3;
  const uint32_t qbase = base + ((uint32_t)(quad * 32) << 16);
  if (warp < 4) {
    uint32_t v[8];
    for (int c0 = 0; c0 < 64; c0 += 8) {
      for (int i = 0; i < 8; ++i) v[i] = (uint32_t)((quad * 32 + lane) * 1000 + c0 + i);
      st8(qbase + c0, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // 1. lane offsets (quadrant 1 only, offsets 0..4)
  if (warp == 1) {
    for (int d = 0; d <= max_lane_offset; ++d) {
      uint32_t r[8];
      ld8(qbase + ((uint32_t)d << 16) + 8, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      out[d * 32 + lane] = (int)r[1];
    }
  }
  __syncthreads();
  // 2. throughput: n warps each issue 64 x8 loads (16 batches of 4 with one wait)
  for (int mode = 0; mode < 3; ++mode) {
    const int nw = mode == 0 ? 1 : mode == 1 ? 4 : 8;
    __syncthreads();
    if (warp < nw) {
      uint32_t acc = 0;
      const long long t0 = clock64();
      for (int it = 0; it < 16; ++it) {
        uint32_t a[8], b[8], c[8], d[8];
        ld8(qbase + 0, a);
        ld8(qbase + 8, b);
        ld8(qbase + 16, c);
        ld8(qbase + 24, d);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += a[0] + b[1] + c[2] + d[3];
      }
      const long long t1 = clock64();
      if (lane == 0) cyc[mode * 8 + warp] = t1 - t0;
      if (acc == 12345) out[1000] = (int)acc;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(256u) : "memory");
}

int main(int argc, char **argv) {
  int *out;
  long long *cyc;
  cudaMallocManaged(&out, 4096 * sizeof(int));
  cudaMallocManaged(&cyc, 64 * sizeof(long long));
  for (int i = 0; i < 4096; ++i) out[i] = -1;
  for (int i = 0; i < 64; ++i) cyc[i] = 0;
  // answer to 1 on B200: any non-zero lane offset traps with "misaligned address"; pass 4 as
  // the first argument to see it
  const int max_lane_offset = argc > 1 ? atoi(argv[1]) : 0;
  probe<<<1, 256>>>(out, cyc, max_lane_offset);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  for (int d = 0; d <= 4; ++d) {
    printf("lane offset %d, column 9: ", d);
    for (int l = 0; l < 32; ++l) printf("%d ", out[d * 32 + l]);
    printf("\n");
  }
  const char *names[3] = {"1 warp", "4 warps (one per quadrant)", "8 warps (two per quadrant)"};
  for (int m = 0; m < 3; ++m) {
    printf("%s: 64 x8 loads per warp, cycles per warp:", names[m]);
    for (int w = 0; w < 8; ++w)
      if (cyc[m * 8 + w]) printf(" %lld", cyc[m * 8 + w]);
    printf("\n");
  }
  return 0;
}
