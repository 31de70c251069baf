// Multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.
// The reference is single-device (nnet/nnet.h:188-197); the only exchange
// step the path has is the sum of per-rank weight deltas in gradient-sum
// (BatchTrain) training, one all-reduce per step over the flat delta buffer.
// NCCL is bound at run time (dlopen) so the single-GPU path has no link-time
// dependency on it.
//
// For small networks (the CIFAR net has 22 466 weights = 90 KB) the all-reduce is
// pure latency, so the exchange and the weight update are ONE kernel over NVLink
// peer memory behind the gradient pass (p2p_inbox_kernel): every rank pushes its delta
// buffer as (value, step tag) pairs into an inbox slot on every peer with 8-byte remote stores,
// waits for the tags of the peers' pairs and adds the contributions in rank order (bit-identical
// weights on all ranks) straight into W.  Inboxes are double-buffered by step
// parity; peer memory is mapped with CUDA IPC handles exchanged through NCCL at
// the first step.  Large delta buffers (the 4096-wide MLP) exchange layer by layer,
// overlapped with the backward pass (p2p_bucket_kernel), or through ncclAllReduce.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <nccl.h>

#include <vector>

#include "common.cuh"
#include "kernels.cuh"

constexpr int kMaxP2pRanks = 8;
constexpr int64_t kMaxP2pFloats = 1 << 20;   // above this the exchange is bandwidth-bound: NCCL
constexpr int kXCtas = 32;                   // CTAs of the small exchange kernel (one step counter each)
constexpr size_t kP2pHeader = 2048;          // bytes reserved for the flag words [2][kXCtas][ranks]
constexpr int kMaxBuckets = 256;             // layers of a network (one exchange bucket each)
// flag block of the large exchange, uint32 words:
//   [bucket][phase A|B][source rank]  then  [bucket] CTA-arrival counters
constexpr size_t kBigFlagWords = (size_t)kMaxBuckets * 2 * kMaxP2pRanks + kMaxBuckets;

struct pb_comm;
// ranks living in one process (pb_comm_create_local): plain device pointers instead of IPC
struct pb_local_group {
  std::vector<pb_comm *> members;
  int alive = 0;
};

struct pb_comm {
  pb_context *ctx = nullptr;
  pb_local_group *group = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, n_ranks = 1;
  int id = 0;                   // process-wide serial: recorded step graphs are keyed on it
  // peer-memory exchange block: [flags[2][kMaxP2pRanks] | inbox[2][n_ranks][p2p_n]]
  int p2p_state = 0;            // 0 = not tried, 1 = ready, -1 = unavailable (NCCL fallback)
  int64_t p2p_n = 0;
  void *local = nullptr;
  void *peers[kMaxP2pRanks] = {nullptr};
  uint32_t *d_xepoch = nullptr; // [kXCtas] step counters of the small exchange, bumped by its CTAs
  int *d_timeout = nullptr;     // = ctx->fault (mapped host word): a peer never showed up
  // large delta buffers: per-layer all-reduce on a second stream, overlapped with the
  // rest of the backward pass (last layer first)
  cudaStream_t comm_stream = nullptr;
  std::vector<cudaEvent_t> events;
  size_t events_used = 0;
  // large delta buffers over NVLink peer memory: every rank's weight and delta buffers
  // and a flag block are mapped into every process (CUDA IPC)
  int big_state = 0;            // 0 = not tried, 1 = ready, -1 = unavailable (NCCL path)
  float *big_weights = nullptr; // the local buffers the mapping was made for
  float *big_W[kMaxP2pRanks] = {nullptr}, *big_D[kMaxP2pRanks] = {nullptr};
  uint32_t *big_flags[kMaxP2pRanks] = {nullptr};
  uint32_t *big_local_flags = nullptr;
  uint32_t *d_epoch = nullptr;  // step counter of the bucket exchange, bumped ON THE DEVICE
                                // (one tiny kernel per step) so the step can be a CUDA graph
};

namespace pb {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t,
                            ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi *nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    // If the host process already loaded an NCCL (torch bundles one), reuse it.
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.handle = h;
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
      api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
      api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
      if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce)
        api.handle = nullptr;
    }
  }
  return api.handle ? &api : nullptr;
}

#define PB_NCCL(api, expr)                                                         \
  do {                                                                             \
    ncclResult_t r__ = (expr);                                                     \
    if (r__ != ncclSuccess)                                                        \
      return pb::fail(std::string(#expr) + ": " +                                   \
                  ((api)->GetErrorString ? (api)->GetErrorString(r__) : "NCCL error")); \
  } while (0)


struct P2pPeers {
  float *inbox[kMaxP2pRanks];      // peer r's inbox base (floats), mapped into this process
  uint32_t *flags[kMaxP2pRanks];   // peer r's flag words
};

// Waits until *word reaches `epoch`; false (and *fault raised) after 5 s: a rank is missing.
__device__ __forceinline__ bool spin_until(volatile uint32_t *word, uint32_t epoch, int *fault) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while ((int)(*word - epoch) < 0) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 5000000000ull) {
      *(volatile int *)fault = 1;
      __threadfence_system();
      return false;
    }
  }
  return true;
}

// The exchange step of a small network as ONE kernel, launched chained right behind the
// gradient pass on the compute stream.  Every value travels with its own flag: an inbox slot
// holds (value, step tag) pairs, written with ONE 8-byte remote store and read with one 8-byte
// load, so the receiver needs no fence and no separate flag round trip - it spins on the tag of
// the element it wants (the LL protocol of NCCL's small-message paths).  A first version pushed
// plain floats, fenced at system scope, published a flag word and fenced again: the two
// fences cost 20 us per step, eight times the data movement.  kXCtas CTAs; CTA c owns slice c
// of the flat buffer:
//   push   (delta, tag) of its slice into slot [step parity][rank] of EVERY rank's inbox
//          (remote stores over NVLink, posted);
//   wait   until the tags of its slice from every rank are the step's: one NVLink latency;
//   apply  W[slice] += contributions in rank order, read from the local inbox: every rank
//          forms the identical sum, so the weights stay bit-identical with no second barrier
//          and no all-gather.  A rank cannot run two steps ahead of a peer (its next wait needs
//          that peer's next push), so two inbox slots suffice; a stale pair carries an older tag.
// A missing peer raises *fault and the CTA leaves W untouched.
__device__ __forceinline__ void st_pair(uint2 *p, float v, uint32_t tag) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag)
               : "memory");
}
__device__ __forceinline__ uint2 ld_pair(const uint2 *p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256)
p2p_inbox_kernel(P2pPeers peers, const float *__restrict__ deltas, float *W, int64_t n, int rank,
                 int n_ranks, uint32_t *epoch_c, int *fault) {
  // No early chain_release(): a successor made resident now would sit on its shared memory
  // until every peer has arrived - fatal when several ranks share one device.
  chain_wait();
  const int c = blockIdx.x;
  const uint32_t epoch = epoch_c[c], tag = epoch + 1;
  const int slot = epoch & 1;
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t i0 = c * per, i1 = min(n, i0 + per);
  for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const float d = deltas[i];
    for (int r = 0; r < n_ranks; ++r)
      st_pair(reinterpret_cast<uint2 *>(peers.inbox[r]) + ((int64_t)slot * n_ranks + rank) * n + i, d, tag);
  }
  const uint2 *in = reinterpret_cast<const uint2 *>(peers.inbox[rank]) + (int64_t)slot * n_ranks * n;
  // wait for every pair of this thread's elements (5 s: a rank is missing)
  int ok = 1;
  unsigned long long t0 = 0;
  for (int64_t i = i0 + threadIdx.x; i < i1 && ok; i += blockDim.x)
    for (int r = 0; r < n_ranks && ok; ++r) {
      const uint2 *p = in + (int64_t)r * n + i;
      unsigned spins = 0;
      while (ld_pair(p).y != tag) {
        if ((++spins & 1023u) == 0) {
          unsigned long long t1;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
          if (t0 == 0) t0 = t1;
          if (t1 - t0 > 5000000000ull) {
            *(volatile int *)fault = 1;
            __threadfence_system();
            ok = 0;
            break;
          }
        }
      }
    }
  if (!__syncthreads_and(ok)) return;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    float acc = 0.0f;
    for (int r = 0; r < n_ranks; ++r)   // rank order: identical sums everywhere
      acc += __uint_as_float(ld_pair(in + (int64_t)r * n + i).x);
    W[i] += acc;
  }
  if (threadIdx.x == 0) epoch_c[c] = epoch + 1;
}

// Collective: every rank calls it at the same point.  Returns 0 and sets
// comm->p2p_state to 1 (ready) or -1 (all ranks fall back to NCCL).
static int p2p_setup(pb_comm *c, int64_t n) {
  if (c->group) {
    // in-process ranks: the first rank to get here allocates every member's block
    const size_t bytes = kP2pHeader + sizeof(uint2) * (size_t)(2 * c->n_ranks * n);
    int dev0 = 0;
    PB_CUDA(cudaGetDevice(&dev0));
    for (pb_comm *m : c->group->members) {
      PB_CUDA(cudaSetDevice(m->ctx->device));
      PB_CUDA(cudaMalloc(&m->local, bytes));
      PB_CUDA(cudaMemset(m->local, 0, bytes));
      PB_CUDA(cudaMalloc((void **)&m->d_xepoch, sizeof(uint32_t) * kXCtas));
      PB_CUDA(cudaMemset(m->d_xepoch, 0, sizeof(uint32_t) * kXCtas));
      if (ensure_fault_word(m->ctx)) return 1;
      m->d_timeout = m->ctx->fault;
      for (pb_comm *o : c->group->members)
        if (o->ctx->device != m->ctx->device) {
          const cudaError_t e = cudaDeviceEnablePeerAccess(o->ctx->device, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            return fail(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
          cudaGetLastError();
        }
    }
    for (pb_comm *m : c->group->members) {
      for (int r = 0; r < c->n_ranks; ++r) m->peers[r] = c->group->members[r]->local;
      m->p2p_state = 1;
      m->p2p_n = n;
    }
    PB_CUDA(cudaSetDevice(dev0));
    return 0;
  }
  NcclApi *api = nccl_api();
  c->p2p_state = -1;
  if (!api || !api->AllGather || c->n_ranks > kMaxP2pRanks || (n & 3)) return 0;
  const char *off = getenv("PB_DISABLE_P2P");
  int ok = !(off && off[0] == '1');
  cudaStream_t st = c->ctx->stream;
  const size_t bytes = kP2pHeader + sizeof(uint2) * (size_t)(2 * c->n_ranks * n);
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok && cudaMalloc(&c->local, bytes) != cudaSuccess) { ok = 0; c->local = nullptr; }
  if (ok && cudaMemsetAsync(c->local, 0, bytes, st) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&mine, c->local) != cudaSuccess) ok = 0;
  if (ensure_fault_word(c->ctx)) ok = 0;
  c->d_timeout = c->ctx->fault;
  if (ok && cudaMalloc((void **)&c->d_xepoch, sizeof(uint32_t) * kXCtas) != cudaSuccess) ok = 0;
  if (ok && cudaMemsetAsync(c->d_xepoch, 0, sizeof(uint32_t) * kXCtas, st) != cudaSuccess) ok = 0;
  // exchange {ok, handle} with every rank (device-side all-gather through NCCL)
  struct Rec { int ok; int pad; cudaIpcMemHandle_t h; };
  Rec rec; rec.ok = ok; rec.pad = 0; rec.h = mine;
  Rec *d_recs = nullptr;
  std::vector<Rec> recs((size_t)c->n_ranks);
  PB_CUDA(cudaMalloc((void **)&d_recs, sizeof(Rec) * c->n_ranks));
  PB_CUDA(cudaMemcpyAsync(d_recs + c->rank, &rec, sizeof(Rec), cudaMemcpyHostToDevice, st));
  PB_NCCL(api, api->AllGather(d_recs + c->rank, d_recs, sizeof(Rec), ncclChar, c->comm, st));
  PB_CUDA(cudaMemcpyAsync(recs.data(), d_recs, sizeof(Rec) * c->n_ranks, cudaMemcpyDeviceToHost, st));
  PB_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_recs);
  int all_ok = 1;
  for (int r = 0; r < c->n_ranks; ++r) all_ok &= recs[r].ok;
  if (all_ok) {
    for (int r = 0; r < c->n_ranks; ++r) {
      if (r == c->rank) { c->peers[r] = c->local; continue; }
      if (cudaIpcOpenMemHandle(&c->peers[r], recs[r].h, cudaIpcMemLazyEnablePeerAccess) !=
          cudaSuccess) {
        cudaGetLastError();
        all_ok = 0;
        c->peers[r] = nullptr;
      }
    }
  }
  // second round: did everyone manage to map everyone?
  int *d_flag = nullptr;
  PB_CUDA(cudaMalloc((void **)&d_flag, sizeof(int) * c->n_ranks));
  PB_CUDA(cudaMemcpyAsync(d_flag + c->rank, &all_ok, sizeof(int), cudaMemcpyHostToDevice, st));
  PB_NCCL(api, api->AllGather(d_flag + c->rank, d_flag, sizeof(int), ncclChar, c->comm, st));
  std::vector<int> flags((size_t)c->n_ranks);
  PB_CUDA(cudaMemcpyAsync(flags.data(), d_flag, sizeof(int) * c->n_ranks, cudaMemcpyDeviceToHost, st));
  PB_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_flag);
  for (int r = 0; r < c->n_ranks; ++r) all_ok &= flags[r];
  if (all_ok) {
    c->p2p_state = 1;
    c->p2p_n = n;
  }
  return 0;
}

struct BigPeers {
  float *W[kMaxP2pRanks];
  float *D[kMaxP2pRanks];
  uint32_t *flags[kMaxP2pRanks];
};

// One bucket (= one layer's slice [off, off+n) of the flat weight / delta buffers) of the
// data-parallel exchange, FUSED with the weight update, over NVLink peer memory:
//   A. publish "my deltas of this bucket are complete" to every rank, wait for everyone's;
//   B. reduce-scatter by reading: this rank owns 1/R of the bucket, loads that part of
//      every rank's delta buffer (peer loads, summed in rank order), adds it to its weights
//      and STORES the new weights into every rank's weight buffer (the all-gather is of the
//      updated weights, not of the deltas): weights are bit-identical on all ranks;
//   C. the last CTA publishes "my part is written everywhere" and waits for everyone's.
// Small CTAs (256 threads, no shared memory) so they co-reside with the persistent GEMM
// CTAs of the backward pass that is still running on the compute stream.
template <int R, int U>
__device__ __forceinline__ void bucket_reduce_apply(const BigPeers &P, int64_t off, int64_t s0,
                                                    int64_t s1, int rank) {
  // U positions x R ranks of independent 16-byte loads in flight per thread: peer loads
  // cost ~2 us, the exchange is latency-bound without deep memory-level parallelism.
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = s0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < s1;
       i0 += stride * U) {
    float4 v[U][R], w[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < s1) {
#pragma unroll
        for (int q = 0; q < R; ++q) v[u][q] = reinterpret_cast<const float4 *>(P.D[q] + off)[i];
        w[u] = reinterpret_cast<const float4 *>(P.W[rank] + off)[i];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < s1) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < R; ++q) {  // rank order: identical sums everywhere
          acc.x += v[u][q].x; acc.y += v[u][q].y; acc.z += v[u][q].z; acc.w += v[u][q].w;
        }
        float4 x = w[u];
        x.x += acc.x; x.y += acc.y; x.z += acc.z; x.w += acc.w;
#pragma unroll
        for (int q = 0; q < R; ++q) reinterpret_cast<float4 *>(P.W[q] + off)[i] = x;
      }
    }
  }
}

// 128 threads, <= 96 registers: small enough to share an SM with a persistent GEMM CTA
// (544 threads x 96 registers, 194 KiB of shared memory) of the backward pass.
__global__ void __launch_bounds__(128, 5)
p2p_bucket_kernel(BigPeers P, int64_t off, int64_t n, int bucket, int rank, int R,
                  const uint32_t *epoch_p, int *timeout_flag) {
  const uint32_t epoch = *epoch_p;
  int arrived = 1;
  uint32_t *mine = P.flags[rank] + (size_t)bucket * 2 * kMaxP2pRanks;
  uint32_t *counter = P.flags[rank] + (size_t)kMaxBuckets * 2 * kMaxP2pRanks + bucket;
  if (threadIdx.x < R) {
    if (blockIdx.x == 0) {
      __threadfence_system();
      *(volatile uint32_t *)(P.flags[threadIdx.x] + (size_t)bucket * 2 * kMaxP2pRanks + rank) = epoch;
    }
    if (!spin_until(mine + threadIdx.x, epoch, timeout_flag)) arrived = 0;
  }
  // a missing peer: leave W untouched (the raised fault word fails the next host call)
  if (!__syncthreads_and(arrived)) return;
  const int64_t n4 = n >> 2;
  const int64_t per = (n4 + R - 1) / R;
  const int64_t s0 = (int64_t)rank * per, s1 = min(n4, s0 + per);
  switch (R) {
    case 2: bucket_reduce_apply<2, 4>(P, off, s0, s1, rank); break;
    case 3: bucket_reduce_apply<3, 2>(P, off, s0, s1, rank); break;
    case 4: bucket_reduce_apply<4, 2>(P, off, s0, s1, rank); break;
    case 5: bucket_reduce_apply<5, 1>(P, off, s0, s1, rank); break;
    case 6: bucket_reduce_apply<6, 1>(P, off, s0, s1, rank); break;
    case 7: bucket_reduce_apply<7, 1>(P, off, s0, s1, rank); break;
    case 8: bucket_reduce_apply<8, 1>(P, off, s0, s1, rank); break;
    default: bucket_reduce_apply<1, 4>(P, off, s0, s1, rank); break;
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int last;
  if (threadIdx.x == 0) {
    const uint32_t prev = atomicAdd(counter, 1u);
    last = prev == gridDim.x - 1;
    if (last) *counter = 0;
  }
  __syncthreads();
  if (last && threadIdx.x < R) {
    __threadfence_system();
    *(volatile uint32_t *)(P.flags[threadIdx.x] + (size_t)bucket * 2 * kMaxP2pRanks +
                           kMaxP2pRanks + rank) = epoch;
    spin_until(mine + kMaxP2pRanks + threadIdx.x, epoch, timeout_flag);
  }
}

__global__ void bump_epoch_kernel(uint32_t *epoch) { *epoch += 1; }

// Collective: maps every rank's weight / delta buffers and flag block into this process.
static int big_setup(pb_comm *c, float *weights, float *deltas, int n_layers) {
  NcclApi *api = nccl_api();
  c->big_state = -1;
  c->big_weights = weights;
  if (!api || !api->AllGather || c->n_ranks > kMaxP2pRanks || n_layers > kMaxBuckets) return 0;
  const char *off = getenv("PB_DISABLE_P2P");
  int ok = !(off && off[0] == '1');
  cudaStream_t st = c->ctx->stream;
  struct Rec { int ok; int pad; cudaIpcMemHandle_t w, d, f; };
  Rec rec;
  memset(&rec, 0, sizeof(rec));
  if (ok && cudaMalloc((void **)&c->big_local_flags, kBigFlagWords * 4) != cudaSuccess) ok = 0;
  if (ok && cudaMemsetAsync(c->big_local_flags, 0, kBigFlagWords * 4, st) != cudaSuccess) ok = 0;
  if (ok && cudaMalloc((void **)&c->d_epoch, 4) != cudaSuccess) ok = 0;
  if (ok && cudaMemsetAsync(c->d_epoch, 0, 4, st) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&rec.w, weights) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&rec.d, deltas) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&rec.f, c->big_local_flags) != cudaSuccess) ok = 0;
  if (!ok) cudaGetLastError();
  if (ensure_fault_word(c->ctx)) ok = 0;
  c->d_timeout = c->ctx->fault;
  rec.ok = ok;
  Rec *d_recs = nullptr;
  std::vector<Rec> recs((size_t)c->n_ranks);
  PB_CUDA(cudaMalloc((void **)&d_recs, sizeof(Rec) * c->n_ranks));
  PB_CUDA(cudaMemcpyAsync(d_recs + c->rank, &rec, sizeof(Rec), cudaMemcpyHostToDevice, st));
  PB_NCCL(api, api->AllGather(d_recs + c->rank, d_recs, sizeof(Rec), ncclChar, c->comm, st));
  PB_CUDA(cudaMemcpyAsync(recs.data(), d_recs, sizeof(Rec) * c->n_ranks, cudaMemcpyDeviceToHost, st));
  PB_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_recs);
  int all_ok = 1;
  for (int r = 0; r < c->n_ranks; ++r) all_ok &= recs[r].ok;
  if (all_ok) {
    for (int r = 0; r < c->n_ranks; ++r) {
      if (r == c->rank) {
        c->big_W[r] = weights; c->big_D[r] = deltas; c->big_flags[r] = c->big_local_flags;
        continue;
      }
      void *pw = nullptr, *pd = nullptr, *pf = nullptr;
      if (cudaIpcOpenMemHandle(&pw, recs[r].w, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
          cudaIpcOpenMemHandle(&pd, recs[r].d, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
          cudaIpcOpenMemHandle(&pf, recs[r].f, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        all_ok = 0;
      }
      c->big_W[r] = (float *)pw; c->big_D[r] = (float *)pd; c->big_flags[r] = (uint32_t *)pf;
    }
  }
  int *d_flag = nullptr;
  PB_CUDA(cudaMalloc((void **)&d_flag, sizeof(int) * c->n_ranks));
  PB_CUDA(cudaMemcpyAsync(d_flag + c->rank, &all_ok, sizeof(int), cudaMemcpyHostToDevice, st));
  PB_NCCL(api, api->AllGather(d_flag + c->rank, d_flag, sizeof(int), ncclChar, c->comm, st));
  std::vector<int> flags((size_t)c->n_ranks);
  PB_CUDA(cudaMemcpyAsync(flags.data(), d_flag, sizeof(int) * c->n_ranks, cudaMemcpyDeviceToHost, st));
  PB_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_flag);
  for (int r = 0; r < c->n_ranks; ++r) all_ok &= flags[r];
  if (all_ok) c->big_state = 1;
  return 0;
}

static cudaEvent_t next_event(pb_comm *c) {
  if (c->events_used == c->events.size()) {
    cudaEvent_t e = nullptr;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    c->events.push_back(e);
  }
  return c->events[c->events_used++];
}

// delta_hook_fn of the peer-memory path: bucket = layer.
static int p2p_bucket_hook(void *user, int layer, float *d_delta, float *d_weights, int64_t n) {
  (void)d_delta;
  pb_comm *c = (pb_comm *)user;
  cudaEvent_t ev = next_event(c);
  PB_CUDA(cudaEventRecord(ev, c->ctx->stream));
  PB_CUDA(cudaStreamWaitEvent(c->comm_stream, ev, 0));
  BigPeers P;
  for (int r = 0; r < kMaxP2pRanks; ++r) {
    const int q = r < c->n_ranks ? r : c->rank;
    P.W[r] = c->big_W[q]; P.D[r] = c->big_D[q]; P.flags[r] = c->big_flags[q];
  }
  const int64_t off = d_weights - c->big_weights;
  const int64_t n4 = ((n + 3) & ~int64_t(3));  // layer blocks are padded to 4 floats
  const int64_t per4 = ((n4 >> 2) + c->n_ranks - 1) / c->n_ranks;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(c->ctx->num_sms, (per4 + 511) / 512));
  p2p_bucket_kernel<<<grid, 128, 0, c->comm_stream>>>(P, off, n4, layer, c->rank, c->n_ranks,
                                                      c->d_epoch, c->d_timeout);
  PB_LAUNCHED(c->ctx);
  return 0;
}

__global__ void accumulate_slice_kernel(float *__restrict__ w, const float *__restrict__ d,
                                        int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    w[i] += d[i];
}

// delta_hook_fn: layer's slice of the delta buffer is complete on the compute stream ->
// all-reduce it and apply it to the weights on the communication stream.  Every later
// reader of these weights in this step (the layer's input_delta) is queued BEFORE the
// event, the next step waits for the communication stream (pb_nnet_batch_train_dp).
static int overlap_hook(void *user, int layer, float *d_delta, float *d_weights, int64_t n) {
  (void)layer;
  pb_comm *c = (pb_comm *)user;
  NcclApi *api = nccl_api();
  PB_CHECK(api, "pb_nnet_batch_train_dp: NCCL unavailable");
  cudaEvent_t ev = next_event(c);
  PB_CUDA(cudaEventRecord(ev, c->ctx->stream));
  PB_CUDA(cudaStreamWaitEvent(c->comm_stream, ev, 0));
  PB_NCCL(api, api->AllReduce(d_delta, d_delta, (size_t)n, ncclFloat, ncclSum, c->comm,
                              c->comm_stream));
  accumulate_slice_kernel<<<pb_ew_grid(c->ctx, n, 256), 256, 0, c->comm_stream>>>(d_weights,
                                                                                 d_delta, n);
  PB_LAUNCHED(c->ctx);
  return 0;
}

}  // namespace pb

extern "C" {

int pb_comm_unique_id(void *id_bytes) {
  PB_CHECK(id_bytes, "pb_comm_unique_id: null");
  static_assert(sizeof(ncclUniqueId) <= PB_COMM_ID_BYTES, "id size");
  pb::NcclApi *api = pb::nccl_api();
  PB_CHECK(api, "pb_comm_unique_id: libnccl.so.2 not found");
  ncclUniqueId id;
  PB_NCCL(api, api->GetUniqueId(&id));
  memset(id_bytes, 0, PB_COMM_ID_BYTES);
  memcpy(id_bytes, &id, sizeof(id));
  return 0;
}

int pb_comm_create(pb_context *ctx, const void *id_bytes, int rank, int n_ranks,
                   pb_comm **out) {
  PB_CHECK(ctx && id_bytes && out, "pb_comm_create: null");
  PB_CHECK(n_ranks >= 1 && rank >= 0 && rank < n_ranks, "pb_comm_create: bad rank");
  pb::NcclApi *api = pb::nccl_api();
  PB_CHECK(api, "pb_comm_create: libnccl.so.2 not found");
  PB_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof(id));
  pb_comm *c = new pb_comm();
  c->ctx = ctx; c->rank = rank; c->n_ranks = n_ranks;
  static int next_id = 0;
  c->id = ++next_id;  // a graph recorded for a destroyed communicator is never replayed
  ncclResult_t r = api->CommInitRank(&c->comm, n_ranks, id, rank);
  if (r != ncclSuccess) {
    delete c;
    return pb::fail(std::string("ncclCommInitRank: ") +
                    (api->GetErrorString ? api->GetErrorString(r) : "error"));
  }
  *out = c;
  return 0;
}

int pb_comm_create_local(pb_context *const *ctxs, int n_ranks, pb_comm **comms) {
  PB_CHECK(ctxs && comms, "pb_comm_create_local: null");
  PB_CHECK(n_ranks >= 1 && n_ranks <= kMaxP2pRanks, "pb_comm_create_local: 1..8 ranks");
  pb_local_group *g = new pb_local_group();
  static int next_local_id = 1 << 20;
  for (int r = 0; r < n_ranks; ++r) {
    PB_CHECK(ctxs[r], "pb_comm_create_local: null context");
    pb_comm *c = new pb_comm();
    c->ctx = ctxs[r]; c->rank = r; c->n_ranks = n_ranks; c->group = g;
    c->id = ++next_local_id;
    g->members.push_back(c);
    comms[r] = c;
  }
  g->alive = n_ranks;
  return 0;
}

int pb_comm_destroy(pb_comm *comm) {
  if (!comm) return 0;
  if (comm->group) {
    cudaSetDevice(comm->ctx->device);
    cudaStreamSynchronize(comm->ctx->stream);
    if (comm->local) cudaFree(comm->local);
    if (comm->d_xepoch) cudaFree(comm->d_xepoch);
    if (--comm->group->alive == 0) delete comm->group;
    delete comm;
    return 0;
  }
  pb::NcclApi *api = pb::nccl_api();
  if (api && comm->comm) {
    cudaStreamSynchronize(comm->ctx->stream);
    for (int r = 0; r < comm->n_ranks; ++r)
      if (r != comm->rank && comm->peers[r]) cudaIpcCloseMemHandle(comm->peers[r]);
    if (comm->comm_stream) {
      cudaStreamSynchronize(comm->comm_stream);
      cudaStreamDestroy(comm->comm_stream);
    }
    for (cudaEvent_t e : comm->events) cudaEventDestroy(e);
    if (comm->big_state == 1) {
      for (int r = 0; r < comm->n_ranks; ++r) {
        if (r == comm->rank) continue;
        if (comm->big_W[r]) cudaIpcCloseMemHandle(comm->big_W[r]);
        if (comm->big_D[r]) cudaIpcCloseMemHandle(comm->big_D[r]);
        if (comm->big_flags[r]) cudaIpcCloseMemHandle(comm->big_flags[r]);
      }
    }
    if (comm->big_local_flags) cudaFree(comm->big_local_flags);
    if (comm->d_epoch) cudaFree(comm->d_epoch);
    if (comm->local) cudaFree(comm->local);
    if (comm->d_xepoch) cudaFree(comm->d_xepoch);
    api->CommDestroy(comm->comm);
  }
  delete comm;
  return 0;
}

int pb_comm_allreduce_sum(pb_comm *comm, float *d_buf, int64_t n) {
  PB_CHECK(comm && (n == 0 || d_buf), "pb_comm_allreduce_sum: null");
  if (n == 0 || comm->n_ranks == 1) return 0;
  PB_CHECK(!comm->group, "pb_comm_allreduce_sum: not available for in-process rank groups");
  pb::NcclApi *api = pb::nccl_api();
  PB_CHECK(api, "pb_comm_allreduce_sum: NCCL unavailable");
  PB_NCCL(api, api->AllReduce(d_buf, d_buf, (size_t)n, ncclFloat, ncclSum, comm->comm,
                              comm->ctx->stream));
  return 0;
}

int pb_nnet_batch_train_dp(pb_nnet *net, pb_comm *comm, const float *d_ins,
                           const float *d_expected, int64_t local_batch) {
  PB_CHECK(net && comm, "pb_nnet_batch_train_dp: null");
  if (pb::check_fault(comm->ctx)) return 1;
  // one rank: exactly the single-GPU step (update fused into the weight-gradient kernels)
  if (comm->n_ranks == 1) return pb_nnet_batch_train(net, d_ins, d_expected, local_batch);
  float *deltas = nullptr, *W = nullptr;
  int64_t n = 0;
  int32_t n_layers = 0;
  if (pb_nnet_deltas_device(net, &deltas) || pb_nnet_total_weights(net, &n) ||
      pb_nnet_weights_device(net, &W) || pb_nnet_num_layers(net, &n_layers))
    return 1;
  if (n <= kMaxP2pFloats) {
    // Small network (the CIFAR net: 22 466 weights, 90 KB): the exchange is pure latency.
    // The step is the chained gradient pass + ONE exchange-and-update kernel behind it on
    // the same stream, recorded into a CUDA graph like the single-GPU step.
    if (comm->p2p_state == 0 && pb::p2p_setup(comm, n)) return 1;
    if (comm->p2p_state == 1 && comm->p2p_n == n) {
      struct Body {
        pb_nnet *net; pb_comm *comm; const float *ins, *exp; int64_t batch, n; float *W, *deltas;
      } body{net, comm, d_ins, d_expected, local_batch, n, W, deltas};
      auto run = [](void *u) -> int {
        Body *b = (Body *)u;
        pb_comm *c = b->comm;
        if (pb_nnet_batch_gradients(b->net, b->ins, b->exp, b->batch)) return 1;
        pb::P2pPeers peers;
        for (int r = 0; r < kMaxP2pRanks; ++r) {
          char *base = (char *)c->peers[r < c->n_ranks ? r : c->rank];
          peers.flags[r] = (uint32_t *)base;
          peers.inbox[r] = (float *)(base + kP2pHeader);
        }
        c->ctx->chain = c->ctx->chain_last ? 2 : 0;
        const cudaError_t e = pb::launch_chained(c->ctx, pb::p2p_inbox_kernel, dim3(kXCtas), dim3(256),
                                                 0, peers, b->deltas, b->W, b->n, c->rank,
                                                 c->n_ranks, c->d_xepoch, c->d_timeout);
        c->ctx->chain = 0;
        PB_CUDA(e);
        PB_LAUNCHED(c->ctx);
        return 0;
      };
      return pb::nnet_run_graphed(net, d_ins, d_expected, local_batch, 2 + 4 * comm->id, run, &body);
    }
  } else {
    PB_CHECK(!comm->group, "pb_nnet_batch_train_dp: in-process rank groups handle networks of "
                           "up to 2^20 weights");
    // Large delta buffers (the 4096-wide MLP: 134 MB, bandwidth-bound): one exchange per
    // layer, started as soon as that layer's weight gradient is done - last layer first -
    // on a high-priority second stream, so the exchange of layer l overlaps the backward
    // kernels of layers l-1, l-2, ...
    if (!comm->comm_stream) {
      int lo = 0, hi = 0;
      PB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      PB_CUDA(cudaStreamCreateWithPriority(&comm->comm_stream, cudaStreamNonBlocking, hi));
    }
    if (comm->big_state == 0 || (comm->big_state == 1 && comm->big_weights != W)) {
      PB_CHECK(comm->big_state == 0, "pb_nnet_batch_train_dp: one network per communicator");
      if (pb::big_setup(comm, W, deltas, n_layers)) return 1;
    }
    // peer-memory exchange fused with the update when every rank could map every rank;
    // otherwise NCCL all-reduce + accumulate per layer on the same hook
    const bool p2p = comm->big_state == 1;
    struct Body {
      pb_nnet *net; pb_comm *comm; const float *ins, *exp; int64_t batch; bool p2p;
    } body{net, comm, d_ins, d_expected, local_batch, p2p};
    auto run = [](void *u) -> int {
      Body *b = (Body *)u;
      pb_comm *c = b->comm;
      c->events_used = 0;
      if (b->p2p) {
        pb::bump_epoch_kernel<<<1, 1, 0, c->ctx->stream>>>(c->d_epoch);
        PB_LAUNCHED(c->ctx);
      }
      pb::nnet_set_delta_hook(b->net, b->p2p ? pb::p2p_bucket_hook : pb::overlap_hook, c);
      const int rc = pb_nnet_batch_gradients(b->net, b->ins, b->exp, b->batch);
      pb::nnet_set_delta_hook(b->net, nullptr, nullptr);
      if (rc) return 1;
      cudaEvent_t done = pb::next_event(c);
      PB_CUDA(cudaEventRecord(done, c->comm_stream));
      PB_CUDA(cudaStreamWaitEvent(c->ctx->stream, done, 0));
      return 0;
    };
    // the peer-memory step is device-driven end to end (flags, step counter): it is
    // recorded into a CUDA graph like the single-GPU step; the NCCL variant stays eager
    if (p2p) return pb::nnet_run_graphed(net, d_ins, d_expected, local_batch, 1 + 4 * comm->id, run, &body);
    return run(&body);
  }
  // no peer mapping (or PB_DISABLE_P2P=1): NCCL all-reduce of the delta buffer + accumulate
  PB_CHECK(!comm->group, "pb_nnet_batch_train_dp: in-process rank group without peer access");
  if (pb_nnet_batch_gradients(net, d_ins, d_expected, local_batch)) return 1;
  if (pb_comm_allreduce_sum(comm, deltas, n)) return 1;
  return pb_nnet_apply_deltas(net);
}

}  // extern "C"
