// Multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.
// The reference is single-device (nnet/nnet.h:188-197); the only exchange
// step the path has is the sum of per-rank weight deltas in gradient-sum
// (BatchTrain) training, one all-reduce per step over the flat delta buffer.
// NCCL is bound at run time (dlopen) so the single-GPU path has no link-time
// dependency on it.
#include <dlfcn.h>
#include <string.h>
#include <nccl.h>

#include "common.cuh"
#include "kernels.cuh"

struct pb_comm {
  pb_context *ctx = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, n_ranks = 1;
};

namespace pb {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t,
                            ncclComm_t, cudaStream_t) = nullptr;
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
  ncclResult_t r = api->CommInitRank(&c->comm, n_ranks, id, rank);
  if (r != ncclSuccess) {
    delete c;
    return pb::fail(std::string("ncclCommInitRank: ") +
                    (api->GetErrorString ? api->GetErrorString(r) : "error"));
  }
  *out = c;
  return 0;
}

int pb_comm_destroy(pb_comm *comm) {
  if (!comm) return 0;
  pb::NcclApi *api = pb::nccl_api();
  if (api && comm->comm) {
    cudaStreamSynchronize(comm->ctx->stream);
    api->CommDestroy(comm->comm);
  }
  delete comm;
  return 0;
}

int pb_comm_allreduce_sum(pb_comm *comm, float *d_buf, int64_t n) {
  PB_CHECK(comm && (n == 0 || d_buf), "pb_comm_allreduce_sum: null");
  if (n == 0 || comm->n_ranks == 1) return 0;
  pb::NcclApi *api = pb::nccl_api();
  PB_CHECK(api, "pb_comm_allreduce_sum: NCCL unavailable");
  PB_NCCL(api, api->AllReduce(d_buf, d_buf, (size_t)n, ncclFloat, ncclSum, comm->comm,
                              comm->ctx->stream));
  return 0;
}

int pb_nnet_batch_train_dp(pb_nnet *net, pb_comm *comm, const float *d_ins,
                           const float *d_expected, int64_t local_batch) {
  PB_CHECK(net && comm, "pb_nnet_batch_train_dp: null");
  if (pb_nnet_batch_gradients(net, d_ins, d_expected, local_batch)) return 1;
  float *deltas = nullptr;
  int64_t n = 0;
  if (pb_nnet_deltas_device(net, &deltas) || pb_nnet_total_weights(net, &n)) return 1;
  if (pb_comm_allreduce_sum(comm, deltas, n)) return 1;
  return pb_nnet_apply_deltas(net);
}

}  // extern "C"
