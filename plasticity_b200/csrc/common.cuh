// Shared internals of libplasticity_b200: context object, error plumbing,
// launch accounting, reference constants.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "plasticity_b200.h"

struct pb_context {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  uint64_t launches = 0;   // kernels launched by this library on `stream`
  // Staging for double<->float conversion at the host boundary.
  double *stage = nullptr;
  size_t stage_elems = 0;
  // Partial-sum workspace for batched weight gradients: `workspace` is the region the last
  // ensure_workspace() call handed out, inside one allocation (ws_base, ws_cap floats).
  // Outside a forked step every call gets the start of the allocation; inside one
  // (fork_mode) the regions are carved one behind the other, because a weight-gradient
  // kernel's partials are still being read by its reduce kernel on the side stream while
  // the next weight-gradient kernel writes its own.
  float *workspace = nullptr;
  size_t workspace_elems = 0;
  float *ws_base = nullptr;
  size_t ws_cap = 0, ws_off = 0;
  // Fork/join inside a fused training step (pb::launch_forked / pb::join_forked): the small
  // fixed-order reduce + update kernels behind the weight-gradient kernels run on `fork_stream`
  // beside the next layer's kernels instead of in front of them.
  bool fork_mode = false;
  // the caller of the next weight-gradient kernel guarantees that its operands stay untouched
  // until the join (dedicated gradient buffer): the kernel itself may be forked, not only its reduce
  bool fork_operands_stable = false;
  cudaStream_t fork_stream = nullptr;
  static constexpr int kMaxForks = 16;
  cudaEvent_t fork_ev[kMaxForks] = {nullptr}, join_ev[kMaxForks] = {nullptr};
  int forks = 0;
  // Re-tiled convolution weights of the implicit-GEMM kernel (conv_gemm_umma.cu).
  float *aux = nullptr;
  size_t aux_elems = 0;
  // Side streams whose work pb_context_finish must also wait for (result read-back of
  // the host-fed training pipeline).
  std::vector<cudaStream_t> side_streams;
  // bumped whenever workspace / aux are reallocated: captured graphs hold their addresses
  uint64_t alloc_generation = 0;
  // Programmatic dependent launch inside a fused training step (pb::launch_chained):
  // 0 = plain launches; 1 = chained, the kernel's weights may still be being written by
  // the launch before it; 2 = chained and the launch before it does not touch this
  // kernel's weights (its weight prologue may run ahead of the dependency wait).
  int chain = 0;
  bool chain_last = false;  // the gradient pass that just ran was launched chained
  // Set (through a mapped pinned host word) by a data-parallel exchange kernel whose peer
  // never arrived: the kernel skips its update and every later synchronising call fails.
  int *fault = nullptr;
};

struct pb_nnet;

namespace pb {

// Called by the batched gradient pass right after the kernels that complete layer
// `layer`'s slice of the delta buffer have been queued (last chunk only): the data-parallel
// driver starts that slice's all-reduce while the earlier layers' backward still runs.
typedef int (*delta_hook_fn)(void *user, int layer, float *d_delta, float *d_weights, int64_t n);
void nnet_set_delta_hook(pb_nnet *net, delta_hook_fn hook, void *user);

// nnet.cu: run a training-step body eagerly / captured / replayed as a CUDA graph.
int nnet_run_graphed(pb_nnet *net, const float *d_ins, const float *d_expected, int64_t batch,
                     int tag, int (*body)(void *), void *user);

void set_error(const std::string &msg);
int fail(const std::string &msg);
int ensure_stage(pb_context *ctx, size_t elems);
int ensure_workspace(pb_context *ctx, size_t elems);
int ensure_aux(pb_context *ctx, size_t elems);
// non-zero (error set) once a data-parallel exchange kernel reported a missing peer
int check_fault(pb_context *ctx);
int ensure_fault_word(pb_context *ctx);

// The reference prints NumericValue::e with 6 significant digits
// (symbolic/numeric_value.cc:41-53), so its kernels compute pow(2.71828, x).
// pow(2.71828, x) == exp2(x * log2(2.71828)).
constexpr double kRefE = 2.71828;
constexpr float kLog2RefE = 1.442694067955017f;     // fl32(log2(2.71828))
constexpr float kLog2RefE_lo = 2.5012059090556704e-09f;  // log2(2.71828) - hi
constexpr float kLn2 = 0.6931471805599453f;
// std::numeric_limits<double>::epsilon() printed the same way
// (symbolic/symbolic_util.cc:40-42).
constexpr float kRefEps = 2.22045e-16f;

// pow(2.71828, x) in fp32.  x*log2(2.71828) is split into the rounded product
// t and its exact residual r (plus the constant's low word), so the result
// stays accurate to a few ulp for |x| up to the overflow range instead of
// losing |x|*2^-24 in the exponent:  2^(t+r) = 2^t * (1 + r ln2 + ...).
__device__ __forceinline__ float ref_exp(float x) {
  const float t = x * kLog2RefE;
  const float r = isinf(t) ? 0.0f : fmaf(x, kLog2RefE, -t) + x * kLog2RefE_lo;
  return exp2f(t) * fmaf(r, kLn2, 1.0f);
}

}  // namespace pb

#define PB_CUDA(expr)                                                        \
  do {                                                                       \
    cudaError_t e__ = (expr);                                                \
    if (e__ != cudaSuccess)                                                  \
      return pb::fail(std::string(#expr) + ": " + cudaGetErrorString(e__));  \
  } while (0)

#define PB_CHECK(cond, msg)                 \
  do {                                      \
    if (!(cond)) return pb::fail(msg);      \
  } while (0)

// After every kernel launch: count it and surface launch-configuration errors.
#define PB_LAUNCHED(ctx)                                                     \
  do {                                                                       \
    (ctx)->launches++;                                                       \
    cudaError_t e__ = cudaGetLastError();                                    \
    if (e__ != cudaSuccess)                                                  \
      return pb::fail(std::string("kernel launch: ") +                       \
                      cudaGetErrorString(e__));                              \
  } while (0)

namespace pb {

// Programmatic dependent launch (griddepcontrol).  A kernel launched through
// launch_chained with ctx->chain != 0 may START while the kernel queued before it is
// still running; it must execute chain_wait() before it touches anything an earlier
// kernel of the stream wrote (or writes anything an earlier kernel reads).  Every chained
// kernel calls chain_wait() and only then chain_release(), so at most the next kernel's
// prologue overlaps the running one and everything two launches back is complete and
// visible.  Launched the ordinary way both instructions are no-ops.
__device__ __forceinline__ void chain_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void chain_release() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chained(pb_context *ctx, void (*kern)(KArgs...), dim3 grid, dim3 block,
                                  size_t smem, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  if (ctx->chain) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ctx->chain = 2;  // whatever follows is no longer the first launch of the step
  }
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// A kernel that nothing on the compute stream needs before the end of the step (join_forked):
// inside a forked step it is queued on ctx->fork_stream behind everything queued on ctx->stream
// so far - during a stream capture this becomes a branch of the step graph - otherwise it is a
// chained launch on the compute stream like any other.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_forked(pb_context *ctx, void (*kern)(KArgs...), dim3 grid, dim3 block,
                                 size_t smem, Args... args) {
  if (!ctx->fork_mode || ctx->forks >= pb_context::kMaxForks)
    return launch_chained(ctx, kern, grid, block, smem, args...);
  const int i = ctx->forks;
  cudaError_t e = cudaSuccess;
  if (!ctx->fork_stream) e = cudaStreamCreateWithFlags(&ctx->fork_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess && !ctx->fork_ev[i]) e = cudaEventCreateWithFlags(&ctx->fork_ev[i], cudaEventDisableTiming);
  if (e == cudaSuccess && !ctx->join_ev[i]) e = cudaEventCreateWithFlags(&ctx->join_ev[i], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventRecord(ctx->fork_ev[i], ctx->stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->fork_stream, ctx->fork_ev[i], 0);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->fork_stream;
  e = cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
  if (e == cudaSuccess) e = cudaEventRecord(ctx->join_ev[i], ctx->fork_stream);
  ctx->forks = i + 1;
  return e;
}
// the compute stream waits for every forked kernel since the last join
int join_forked(pb_context *ctx);

}  // namespace pb

static inline int pb_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Grid for a grid-stride elementwise kernel: enough CTAs to fill every SM a
// few times over, never more than the work needs.
static inline int pb_ew_grid(const pb_context *ctx, int64_t n, int threads) {
  int64_t want = (n + threads - 1) / threads;
  int64_t cap = (int64_t)ctx->num_sms * 8;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}
