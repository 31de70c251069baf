// Network driver: the body of nnet::Nnet (nnet/nnet.h) over the layer
// launchers.  Evaluate :213-273, Train :369-483, Error :280-314,
// ErrorGradients :316-349, BatchTrain / CalculateGradients :514-669,
// SetLearningParameters :357-361, Xavier init nnet/layer.cc:86-100.
//
// Device-resident state (fp32): all layers' weights concatenated in layer
// order (reference layout per layer), a same-shaped delta buffer for
// gradient-sum training, one activation buffer per layer sized for max_batch
// samples (the reference's eval_layer_outputs_), two ping-pong gradient
// buffers (backprop_gradients_ / next_backprop_gradients_) and a 1-element
// learning-rate buffer (learning_rate_buffer_).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

struct pb_nnet {
  pb_context *ctx = nullptr;
  std::vector<pb_layer_desc> layers;
  int loss = PB_MEAN_SQUARED;
  int64_t max_batch = 1;
  std::vector<int64_t> ni, no, nw, w_off;
  int64_t total_w = 0, max_io = 0;
  float *weights = nullptr, *deltas = nullptr;
  std::vector<float *> outputs;
  float *grad_a = nullptr, *grad_b = nullptr;
  float *grad_c = nullptr;  // dE/d(convolution output) of a fused block (allocated on first use)
  float *grad_z = nullptr;  // dE/d(dense output) of the fused head: its own buffer, so that the
                            // dense weight gradient can run beside the next layer's backward
  float *lr = nullptr;
  float *err = nullptr;  // n_out loss components
  // strict per-sample SGD loop (pb_nnet_train_loop)
  struct Cursor { const float *ins; const float *expected; long long idx; };
  Cursor *cursor = nullptr;
  float *stage_in = nullptr, *stage_exp = nullptr;
  cudaGraphExec_t train_exec = nullptr;
  uint64_t graph_launches = 0;  // kernels inside one replay of train_exec
  // Fused plan of the batched gradient pass (pb_nnet_batch_gradients):
  //   block[l] != 0: layers l (convolution), l+1 (activation), l+2 (exact 2x2 max
  //   pool) run as one forward kernel and one elementwise backward kernel;
  //   alias_input: layer 0 is the Identity input layer (architecture.h:45-47), a
  //   pure copy forward and backward — the next layer reads the caller's buffer.
  //   head: index of the DenseLayer of a trailing dense -> activation -> softmax
  //   chain (or -1): forward, error gradient and backward of those three layers
  //   run as one kernel (head.cu).
  std::vector<char> block;
  int head = -1;
  // pb_nnet_profile_step: CUDA events recorded after each kernel group of a step
  bool profiling = false;
  std::vector<std::pair<std::string, cudaEvent_t>> marks;
  // Host-buffer training (pb_nnet_batch_train_host): two device staging slots
  // filled by a copy stream while the previous step computes.
  struct HostPipe {
    cudaStream_t copy = nullptr;
    float *x[2] = {nullptr, nullptr}, *e[2] = {nullptr, nullptr};
    uint8_t *rec[2] = {nullptr, nullptr};  // raw byte records (pb_nnet_batch_train_host_records)
    // result read-back off the compute stream: the step's outputs are copied to y[s] by a
    // kernel, the D2H DMA of y[s] runs on `out` while the next step computes
    cudaStream_t out = nullptr;
    float *y[2] = {nullptr, nullptr};
    cudaEvent_t y_ready[2] = {nullptr, nullptr};
    cudaEvent_t copied[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr};
    int64_t cap = 0;
    uint64_t calls = 0;
  } pipe;
  bool alias_input = false;
  bool fuse = true;
  pb::delta_hook_fn delta_hook = nullptr;  // data-parallel overlap, see common.cuh
  void *delta_hook_user = nullptr;
  // pb_nnet_batch_train: the in-place step for one (inputs, targets, batch) triple is
  // captured into a CUDA graph on its second use and replayed afterwards (one launch
  // instead of ~18: no host-side launch cost, tighter kernel-to-kernel dependencies)
  struct StepGraph {
    const float *ins = nullptr, *expected = nullptr, *out = nullptr;  // out: last layer's buffer
    int64_t batch = 0;
    int tag = 0;              // which step body (0: in-place step, 1: data-parallel step)
    uint64_t generation = 0;  // ctx->alloc_generation at capture (workspace addresses)
    int uses = 0;
    cudaGraphExec_t exec = nullptr;
    uint64_t launches = 0;
  };
  std::vector<StepGraph> step_graphs;
  size_t step_graph_next = 0;
  bool use_graphs = true;
};

namespace pb {

int layer_check(const pb_layer_desc *l);

void nnet_set_delta_hook(pb_nnet *net, delta_hook_fn hook, void *user) {
  net->delta_hook = hook;
  net->delta_hook_user = user;
}

static int fire_delta_hook(pb_nnet *net, int l, bool last_chunk) {
  if (!net->delta_hook || !last_chunk || net->nw[l] == 0) return 0;
  return net->delta_hook(net->delta_hook_user, l, net->deltas + net->w_off[l],
                         net->weights + net->w_off[l], net->nw[l]);
}

__global__ void xavier_kernel(float *__restrict__ w, int64_t n, float stddev,
                              unsigned long long seed, unsigned long long base) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    // splitmix64 on (seed, global index) -> two uniforms -> Box-Muller.
    unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (base + (unsigned long long)i + 1);
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x = x ^ (x >> 31);
    const float u1 = ((float)(unsigned)(x >> 40) + 1.0f) * (1.0f / 16777217.0f);
    const float u2 = (float)(unsigned)((x >> 8) & 0xFFFFFFu) * (1.0f / 16777216.0f);
    w[i] = stddev * sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
  }
}

// Copies sample `idx` of the resident data set into the fixed staging buffers
// the captured training graph reads, so the graph needs no per-replay update.
__global__ void fetch_sample_kernel(const pb_nnet::Cursor *__restrict__ cur,
                                    float *__restrict__ stage_in,
                                    float *__restrict__ stage_exp, int nin, int nout) {
  const float *src = cur->ins + (int64_t)cur->idx * nin;
  const float *exp = cur->expected + (int64_t)cur->idx * nout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nin + nout;
       i += gridDim.x * blockDim.x) {
    if (i < nin) stage_in[i] = src[i];
    else stage_exp[i - nin] = exp[i - nin];
  }
}

__global__ void advance_cursor_kernel(pb_nnet::Cursor *cur) { cur->idx += 1; }

static void mark(pb_nnet *net, const std::string &label) {
  if (!net->profiling) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, net->ctx->stream);
  net->marks.emplace_back(label, e);
}

static std::string layer_tag(const pb_nnet *net, int l) {
  static const char *names[] = {"activation_layer", "dense_layer", "convolution_layer",
                                "max_pool_layer", "softmax_layer"};
  return std::string(names[net->layers[l].kind]) + "_" + std::to_string(l);
}

// Batched forward of the gradient pass with the fused plan.  layer_in[l] is
// what layer l read (outputs of skipped layers are not materialised).
static int forward_fused(pb_nnet *net, const float *d_in, int64_t batch,
                         std::vector<const float *> &layer_in, std::vector<char> &fused_fwd,
                         bool head_ok) {
  const int n = (int)net->layers.size();
  const float *cur = d_in;
  layer_in.assign(n, nullptr);
  fused_fwd.assign(n, 0);
  for (int l = 0; l < n; ++l) {
    layer_in[l] = cur;
    if (l == 0 && net->alias_input && n > 1) continue;  // identity copy elided
    if (l == net->head && head_ok) break;             // the head kernel does the rest
    const float *W = net->nw[l] ? net->weights + net->w_off[l] : nullptr;
    if (net->block[l]) {
      const int rc = conv_act_pool_evaluate_umma(net->ctx, &net->layers[l],
                                                 net->layers[l + 1].activation,
                                                 &net->layers[l + 2], cur, W, net->outputs[l],
                                                 net->outputs[l + 2], batch);
      if (rc > 0) return 1;
      if (rc == 0) {
        mark(net, layer_tag(net, l) + ":evaluate+activation+pool (tcgen05)");
        fused_fwd[l] = 1;
        layer_in[l + 1] = net->outputs[l];
        layer_in[l + 2] = nullptr;  // activation output not materialised
        cur = net->outputs[l + 2];
        l += 2;
        continue;
      }
    }
    if (pb_layer_evaluate(net->ctx, &net->layers[l], cur, W, net->outputs[l], batch)) return 1;
    mark(net, layer_tag(net, l) + ":evaluate");
    cur = net->outputs[l];
  }
  return 0;
}

// Reverse loop over the fused plan; same per-layer order as backward().
static int backward_fused(pb_nnet *net, int64_t batch, float *g, float *ng, bool first_chunk,
                          const std::vector<const float *> &layer_in,
                          const std::vector<char> &fused_fwd, int top, bool in_place,
                          bool last_chunk) {
  const int n = (int)net->layers.size();
  for (int l = top; l >= 0; --l) {
    const pb_layer_desc *L = &net->layers[l];
    if (l == 0 && net->alias_input && n > 1) break;  // dE/dInput == g, unobserved here
    if (l >= 2 && net->block[l - 2] && L->kind == PB_MAX_POOL && fused_fwd[l - 2]) {
      // The whole block backward in two kernels: the convolution's input-gradient kernel forms
      // dE/d(conv output) from the pooled gradient and the stored convolution output inside
      // its producers (and writes it out once, for the weight gradient that follows).
      const int cl = l - 2;
      if (cl == 1 && net->alias_input) {
        // First convolution of the network: BatchTrain hands nobody dE/dInput (nnet.h:617-622
        // passes a null input_gradients), so only the weight gradient is left to do, and its
        // kernel forms dE/d(conv output) from the pooled gradient by itself.
        float *Wc = net->weights + net->w_off[cl];
        const int m = in_place ? PB_NEW_WEIGHTS : first_chunk ? PB_WEIGHT_DELTAS : PB_ACCUMULATE;
        float *out = in_place ? Wc : net->deltas + net->w_off[cl];
        const float *in = layer_in[cl] ? layer_in[cl] : net->outputs[cl - 1];
        int wrc = conv_act_pool_weight_update_rows(net->ctx, &net->layers[cl],
                                                   net->layers[l - 1].activation, L, in, Wc,
                                                   net->outputs[cl], g, out, net->lr, m, batch);
        if (wrc < 0)
          wrc = conv_act_pool_weight_update_mma(net->ctx, &net->layers[cl],
                                                net->layers[l - 1].activation, L, in, Wc,
                                                net->outputs[cl], g, out, net->lr, m, batch);
        if (wrc > 0) return 1;
        if (wrc == 0) {
          mark(net, layer_tag(net, cl) + ":activation+pool backward+weight_update");
          if (!in_place && fire_delta_hook(net, cl, last_chunk)) return 1;
          break;  // layer 0 is the aliased input
        }
      }
      if (!net->grad_c) {
        if (cudaMalloc((void **)&net->grad_c, sizeof(float) * (size_t)(net->max_io * net->max_batch)) !=
            cudaSuccess) {
          cudaGetLastError();
          net->grad_c = nullptr;
        }
      }
      float *Wc = net->weights + net->w_off[cl];
      const int frc = net->grad_c ? conv_act_pool_input_delta_umma(
                                        net->ctx, &net->layers[cl], net->layers[l - 1].activation, L,
                                        Wc, net->outputs[cl], g, net->grad_c, ng, batch)
                                  : -1;
      if (frc > 0) return 1;
      if (frc == 0) {
        mark(net, layer_tag(net, cl) + ":input_delta+activation+pool backward (tcgen05)");
        const float *in = layer_in[cl] ? layer_in[cl] : net->outputs[cl - 1];
        const int m = in_place ? PB_NEW_WEIGHTS : first_chunk ? PB_WEIGHT_DELTAS : PB_ACCUMULATE;
        float *out = in_place ? Wc : net->deltas + net->w_off[cl];
        if (pb_layer_weight_update(net->ctx, &net->layers[cl], in, Wc, net->grad_c, out, net->lr, m,
                                   batch))
          return 1;
        mark(net, layer_tag(net, cl) + ":weight_update");
        if (!in_place && fire_delta_hook(net, cl, last_chunk)) return 1;
        float *t = g; g = ng; ng = t;
        l = cl;  // pool, activation and convolution are done
        continue;
      }
    }
    if (l >= 2 && net->block[l - 2] && L->kind == PB_MAX_POOL) {
      // pool + activation backward in one pass over the convolution output
      const int rc = act_pool_input_delta(net->ctx, net->layers[l - 1].activation, L,
                                          net->outputs[l - 2], g, ng, batch);
      if (rc > 0) return 1;
      if (rc == 0) {
        mark(net, layer_tag(net, l) + "+" + layer_tag(net, l - 1) + ":input_delta");
        float *t = g; g = ng; ng = t;
        l -= 1;  // the activation layer is done too
        continue;
      }
      if (fused_fwd[l - 2]) {
        // forward skipped the activation output this path needs: recompute it
        if (pb_layer_evaluate(net->ctx, &net->layers[l - 1], net->outputs[l - 2], nullptr,
                              net->outputs[l - 1], batch))
          return 1;
      }
    }
    const float *in = layer_in[l] ? layer_in[l] : net->outputs[l - 1];
    float *W = net->nw[l] ? net->weights + net->w_off[l] : nullptr;
    // the layer behind the aliased input: nobody reads its input gradient (BatchTrain has no
    // input_gradients parameter, nnet.h:617-622)
    const bool unobserved = l == 1 && net->alias_input;
    if (!unobserved) {
      if (pb_layer_input_delta(net->ctx, L, in, W, g, ng, batch)) return 1;
      mark(net, layer_tag(net, l) + ":input_delta");
    }
    if (net->nw[l] > 0) {
      // in_place: W <- W - lr * sum_b grad_b directly (the layer's input_delta, the only
      // later reader of W in this step, is already queued ahead on the in-order stream)
      const int m = in_place ? PB_NEW_WEIGHTS : first_chunk ? PB_WEIGHT_DELTAS : PB_ACCUMULATE;
      float *out = in_place ? W : net->deltas + net->w_off[l];
      if (pb_layer_weight_update(net->ctx, L, in, W, g, out, net->lr, m, batch)) return 1;
      mark(net, layer_tag(net, l) + ":weight_update");
      if (!in_place && fire_delta_hook(net, l, last_chunk)) return 1;
    }
    float *t = g; g = ng; ng = t;
  }
  return 0;
}

static int forward(pb_nnet *net, const float *d_in, int64_t batch) {
  const float *cur = d_in;
  for (size_t l = 0; l < net->layers.size(); ++l) {
    const float *W = net->nw[l] ? net->weights + net->w_off[l] : nullptr;
    // convolution + the ActivationLayer behind it in one kernel (both outputs are
    // materialised: eval_layer_outputs_ keeps every layer's result, nnet.h:81-83)
    if (net->fuse && batch >= 8 && net->layers[l].kind == PB_CONVOLUTION &&
        l + 1 < net->layers.size() && net->layers[l + 1].kind == PB_ACTIVATION &&
        net->layers[l + 1].n == net->no[l]) {
      const int rc = conv_evaluate_gemm_umma(net->ctx, &net->layers[l], cur, W, net->outputs[l],
                                             net->outputs[l + 1], net->layers[l + 1].activation,
                                             batch);
      if (rc > 0) return 1;
      if (rc == 0) {
        mark(net, layer_tag(net, (int)l) + "+" + layer_tag(net, (int)l + 1) + ":evaluate (tcgen05)");
        cur = net->outputs[l + 1];
        l += 1;
        continue;
      }
    }
    if (pb_layer_evaluate(net->ctx, &net->layers[l], cur, W, net->outputs[l], batch)) return 1;
    mark(net, layer_tag(net, (int)l) + ":evaluate");
    cur = net->outputs[l];
  }
  return 0;
}

// Reverse loop of Nnet::Train (nnet/nnet.h:399-460).  `g` holds dE/dO of the
// last layer on entry.  Per layer: input_delta with the pre-update weights,
// then the weight kernel.  Returns the buffer holding dE/dInput in *g_out.
static int backward(pb_nnet *net, const float *d_in, int64_t batch, float *g, float *ng,
                    int mode, bool first_chunk, float **g_out, bool last_chunk = true) {
  for (int l = (int)net->layers.size() - 1; l >= 0; --l) {
    const pb_layer_desc *L = &net->layers[l];
    const float *layer_in = (l > 0) ? net->outputs[l - 1] : d_in;
    float *W = net->nw[l] ? net->weights + net->w_off[l] : nullptr;
    if (pb_layer_input_delta(net->ctx, L, layer_in, W, g, ng, batch)) return 1;
    if (net->nw[l] > 0) {
      // PB_NEW_WEIGHTS updates in place: the weight gradient reads only the
      // layer's inputs and G, and input_delta (which reads W) is already
      // queued ahead of it on the in-order stream — the same ordering the
      // reference gets from DeepClone + handle swap (nnet.h:432-454).
      float *out = (mode == PB_NEW_WEIGHTS) ? W : net->deltas + net->w_off[l];
      const int m = (mode == PB_NEW_WEIGHTS) ? PB_NEW_WEIGHTS
                    : first_chunk            ? PB_WEIGHT_DELTAS
                                             : PB_ACCUMULATE;
      if (pb_layer_weight_update(net->ctx, L, layer_in, W, g, out, net->lr, m, batch)) return 1;
      if (mode != PB_NEW_WEIGHTS && fire_delta_hook(net, l, last_chunk)) return 1;
    }
    float *t = g; g = ng; ng = t;
  }
  *g_out = g;
  return 0;
}

static int train_step(pb_nnet *net, const float *d_in, const float *d_expected,
                      float *d_input_gradients, const float *d_initial_gradients) {
  pb_context *ctx = net->ctx;
  const int64_t n_out = net->no.back();
  if (forward(net, d_in, 1)) return 1;
  if (d_initial_gradients) {
    if (pb_buffer_copy(ctx, net->grad_a, d_initial_gradients, n_out)) return 1;
  } else {
    if (pb_error_gradients(ctx, net->loss, n_out, net->outputs.back(), d_expected,
                           net->grad_a, 1))
      return 1;
  }
  float *g = nullptr;
  if (backward(net, d_in, 1, net->grad_a, net->grad_b, PB_NEW_WEIGHTS, true, &g)) return 1;
  if (d_input_gradients) return pb_buffer_copy(ctx, d_input_gradients, g, net->ni[0]);
  return 0;
}


// Runs one training-step body, recorded into a CUDA graph the second time it is called
// with the same (inputs, targets, output buffer, batch, tag) and replayed afterwards.
// First use: eager (one-off buffers never pay for a capture; the eager run also sizes the
// workspaces).  Second use: eager once more, then the same call sequence is captured —
// capturing executes nothing.  Graphs are dropped when a workspace they captured moved.
// distinct (inputs, targets) pairs with a recorded graph per network: a resident data set of
// up to this many batches replays every step from its second epoch on
constexpr size_t kMaxStepGraphs = 256;

int nnet_run_graphed(pb_nnet *net, const float *d_ins, const float *d_expected, int64_t batch,
                     int tag, int (*body)(void *), void *user) {
  if (!net->use_graphs || net->profiling) return body(user);
  pb_context *ctx = net->ctx;
  pb_nnet::StepGraph *sg = nullptr;
  const float *out_now = net->outputs.back();
  for (auto &g : net->step_graphs)
    if (g.ins == d_ins && g.expected == d_expected && g.out == out_now && g.batch == batch &&
        g.tag == tag)
      sg = &g;
  if (!sg) {
    pb_nnet::StepGraph fresh;
    fresh.ins = d_ins; fresh.expected = d_expected; fresh.out = out_now; fresh.batch = batch;
    fresh.tag = tag;
    if (net->step_graphs.size() < kMaxStepGraphs) {
      net->step_graphs.push_back(fresh);
      sg = &net->step_graphs.back();
    } else {
      // replace, in ring order, an entry that never got a graph (a data set with more
      // batches than slots cycles through here and simply stays eager); recorded graphs
      // are never evicted by one-off keys
      for (size_t tries = 0; tries < kMaxStepGraphs && !sg; ++tries) {
        pb_nnet::StepGraph *cand = &net->step_graphs[net->step_graph_next++ % kMaxStepGraphs];
        if (!cand->exec) sg = cand;
      }
      if (!sg) return body(user);
      *sg = fresh;
    }
  }
  if (sg->exec && sg->generation != ctx->alloc_generation) {
    cudaGraphExecDestroy(sg->exec);  // the workspace moved since the capture
    sg->exec = nullptr;
    sg->uses = 0;
  }
  if (sg->exec) {
    PB_CUDA(cudaGraphLaunch(sg->exec, ctx->stream));
    ctx->launches += sg->launches;
    return 0;
  }
  if (body(user)) return 1;
  if (sg->uses++ == 0) return 0;
  cudaGraph_t graph = nullptr;
  const uint64_t before = ctx->launches;
  if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    net->use_graphs = false;
    return 0;
  }
  const int rc = body(user);
  const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
  sg->launches = ctx->launches - before;
  ctx->launches = before;
  if (rc || e != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    net->use_graphs = false;  // something in the step is not capturable: stay eager
    return 0;
  }
  sg->generation = ctx->alloc_generation;
  const cudaError_t ei = cudaGraphInstantiate(&sg->exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ei != cudaSuccess) {
    sg->exec = nullptr;
    cudaGetLastError();
    net->use_graphs = false;
  }
  return 0;
}

}  // namespace pb

extern "C" {

int pb_nnet_create(pb_context *ctx, const pb_layer_desc *layers, int32_t n_layers,
                   int32_t loss, int64_t max_batch, pb_nnet **out) {
  PB_CHECK(ctx && layers && out, "pb_nnet_create: null");
  PB_CHECK(n_layers > 0, "Invalid dimensions passed to Nnet(): empty architecture");
  PB_CHECK(loss == PB_MEAN_SQUARED || loss == PB_CROSS_ENTROPY,
           "Error: Unknown loss function selected.");
  PB_CHECK(max_batch >= 1, "pb_nnet_create: max_batch must be >= 1");
  PB_CUDA(cudaSetDevice(ctx->device));
  pb_nnet *net = new pb_nnet();
  net->ctx = ctx;
  net->loss = loss;
  net->max_batch = max_batch;
  net->layers.assign(layers, layers + n_layers);
  for (int l = 0; l < n_layers; ++l) {
    if (pb::layer_check(&layers[l])) { delete net; return 1; }
    int64_t a, b, c;
    pb_layer_num_inputs(&layers[l], &a);
    pb_layer_num_outputs(&layers[l], &b);
    pb_layer_num_weights(&layers[l], &c);
    // Architecture::VerifyArchitecture, nnet/architecture.h:16-32: only the
    // flat counts of consecutive layers have to agree.
    if (l > 0 && net->no.back() != a) {
      delete net;
      return pb::fail("Dimension mismatch between output of layer " + std::to_string(l - 1) +
                      " (" + std::to_string(net->no.empty() ? 0 : net->no.back()) +
                      ") and input of layer " + std::to_string(l) + " (" + std::to_string(a) +
                      ").");
    }
    if (a <= 0 || b <= 0) {
      delete net;
      return pb::fail("Layer " + std::to_string(l) + " has zero inputs or outputs");
    }
    net->ni.push_back(a);
    net->no.push_back(b);
    net->nw.push_back(c);
    net->w_off.push_back(net->total_w);
    // Keep every layer's weight block 16-byte aligned inside the flat buffer.
    net->total_w += (c + 3) & ~int64_t(3);
    net->max_io = std::max(net->max_io, std::max(a, b));
  }
  {
    const char *e = getenv("PB_DISABLE_FUSION");
    net->fuse = !(e && e[0] == '1');
    const char *gr = getenv("PB_STEP_GRAPHS");
    net->use_graphs = !(gr && gr[0] == '0');
    net->block.assign(n_layers, 0);
    net->alias_input = net->fuse && layers[0].kind == PB_ACTIVATION &&
                       layers[0].activation == PB_IDENTITY;
    for (int l = 0; net->fuse && l + 2 < n_layers; ++l) {
      const pb_layer_desc &c = layers[l], &a = layers[l + 1], &p = layers[l + 2];
      if (c.kind != PB_CONVOLUTION || a.kind != PB_ACTIVATION || p.kind != PB_MAX_POOL) continue;
      if (p.width != pb::conv_out_w(&c) || p.height != pb::conv_out_h(&c) ||
          p.depth != c.num_filters || p.out_width * 2 != p.width || p.out_height * 2 != p.height)
        continue;
      net->block[l] = 1;
    }
    // trailing dense -> activation -> softmax (h > 0: the dense layer is never layer 0)
    if (net->fuse && n_layers >= 4) {
      const int h = n_layers - 3;
      if (layers[h].kind == PB_DENSE && layers[h + 1].kind == PB_ACTIVATION &&
          layers[h + 2].kind == PB_SOFTMAX && layers[h + 2].n == layers[h].out &&
          layers[h + 1].n == layers[h].out && !net->block[h - 1] && (h < 2 || !net->block[h - 2]))
        net->head = h;
    }
  }
  auto dalloc = [&](float **p, int64_t n) {
    return cudaMalloc((void **)p, sizeof(float) * (size_t)std::max<int64_t>(n, 4));
  };
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = dalloc(&net->weights, net->total_w);
  if (e == cudaSuccess) e = dalloc(&net->deltas, net->total_w);
  if (e == cudaSuccess) e = dalloc(&net->grad_a, net->max_io * max_batch);
  if (e == cudaSuccess) e = dalloc(&net->grad_b, net->max_io * max_batch);
  if (e == cudaSuccess) e = dalloc(&net->lr, 4);
  if (e == cudaSuccess) e = dalloc(&net->err, net->no.back());
  if (e == cudaSuccess) e = dalloc(&net->stage_in, net->ni[0]);
  if (e == cudaSuccess) e = dalloc(&net->stage_exp, net->no.back());
  if (e == cudaSuccess) e = cudaMalloc((void **)&net->cursor, sizeof(pb_nnet::Cursor));
  net->outputs.assign(n_layers, nullptr);
  for (int l = 0; l < n_layers && e == cudaSuccess; ++l)
    e = dalloc(&net->outputs[l], net->no[l] * max_batch);
  if (e == cudaSuccess)
    e = cudaMemsetAsync(net->weights, 0, sizeof(float) * (size_t)std::max<int64_t>(net->total_w, 4), ctx->stream);
  if (e == cudaSuccess)
    e = cudaMemsetAsync(net->deltas, 0, sizeof(float) * (size_t)std::max<int64_t>(net->total_w, 4), ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(net->lr, 0, 4 * sizeof(float), ctx->stream);
  if (e != cudaSuccess) {
    pb_nnet_destroy(net);
    return pb::fail(std::string("pb_nnet_create: ") + cudaGetErrorString(e));
  }
  *out = net;
  return 0;
}

int pb_nnet_destroy(pb_nnet *net) {
  if (!net) return 0;
  cudaSetDevice(net->ctx->device);
  cudaStreamSynchronize(net->ctx->stream);
  if (net->pipe.copy) {
    cudaStreamSynchronize(net->pipe.copy);
    for (int i = 0; i < 2; ++i) {
      cudaFree(net->pipe.x[i]); cudaFree(net->pipe.e[i]); cudaFree(net->pipe.rec[i]);
      cudaFree(net->pipe.y[i]);
      if (net->pipe.y_ready[i]) cudaEventDestroy(net->pipe.y_ready[i]);
      cudaEventDestroy(net->pipe.copied[i]); cudaEventDestroy(net->pipe.consumed[i]);
    }
    cudaStreamDestroy(net->pipe.copy);
    if (net->pipe.out) {
      cudaStreamSynchronize(net->pipe.out);
      auto &ss = net->ctx->side_streams;
      for (size_t i = 0; i < ss.size(); ++i)
        if (ss[i] == net->pipe.out) { ss.erase(ss.begin() + i); break; }
      cudaStreamDestroy(net->pipe.out);
    }
  }
  if (net->train_exec) cudaGraphExecDestroy(net->train_exec);
  for (auto &sg : net->step_graphs)
    if (sg.exec) cudaGraphExecDestroy(sg.exec);
  cudaFree(net->weights); cudaFree(net->deltas);
  cudaFree(net->grad_a); cudaFree(net->grad_b);
  if (net->grad_c) cudaFree(net->grad_c);
  if (net->grad_z) cudaFree(net->grad_z);
  cudaFree(net->lr); cudaFree(net->err);
  cudaFree(net->stage_in); cudaFree(net->stage_exp); cudaFree(net->cursor);
  for (float *p : net->outputs) cudaFree(p);
  delete net;
  return 0;
}

int pb_nnet_num_layers(const pb_nnet *net, int32_t *n) {
  PB_CHECK(net && n, "pb_nnet_num_layers: null");
  *n = (int32_t)net->layers.size();
  return 0;
}

int pb_nnet_total_weights(const pb_nnet *net, int64_t *n) {
  PB_CHECK(net && n, "pb_nnet_total_weights: null");
  *n = net->total_w;
  return 0;
}

int pb_nnet_set_weights_f64(pb_nnet *net, int32_t layer, const double *host, int64_t n) {
  PB_CHECK(net, "pb_nnet_set_weights_f64: null");
  PB_CHECK(layer >= 0 && layer < (int)net->layers.size(), "Too large layer index");
  PB_CHECK(n == net->nw[layer], "layer size mismatch!");  // nnet.h:742-745
  return pb_buffer_write_f64(net->ctx, net->weights + net->w_off[layer], host, (size_t)n);
}

int pb_nnet_get_weights_f64(pb_nnet *net, int32_t layer, double *host, int64_t n) {
  PB_CHECK(net, "pb_nnet_get_weights_f64: null");
  PB_CHECK(layer >= 0 && layer < (int)net->layers.size(), "Too large layer index");
  PB_CHECK(n == net->nw[layer], "layer size mismatch!");
  return pb_buffer_read_f64(net->ctx, net->weights + net->w_off[layer], host, (size_t)n);
}

int pb_nnet_layer_weights_device(pb_nnet *net, int32_t layer, float **dptr) {
  PB_CHECK(net && dptr, "pb_nnet_layer_weights_device: null");
  PB_CHECK(layer >= 0 && layer < (int)net->layers.size(), "Too large layer index");
  *dptr = net->weights + net->w_off[layer];
  return 0;
}

int pb_nnet_weights_device(pb_nnet *net, float **dptr) {
  PB_CHECK(net && dptr, "pb_nnet_weights_device: null");
  *dptr = net->weights;
  return 0;
}

int pb_nnet_deltas_device(pb_nnet *net, float **dptr) {
  PB_CHECK(net && dptr, "pb_nnet_deltas_device: null");
  *dptr = net->deltas;
  return 0;
}

int pb_nnet_init_xavier(pb_nnet *net, uint64_t seed) {
  PB_CHECK(net, "pb_nnet_init_xavier: null");
  pb_context *ctx = net->ctx;
  for (size_t l = 0; l < net->layers.size(); ++l) {
    if (net->nw[l] == 0) continue;
    // stats::Normal(0, 1.0 / num_inputs): variance, nnet/layer.cc:86-88
    const float stddev = (float)std::sqrt(1.0 / (double)net->ni[l]);
    pb::xavier_kernel<<<pb_ew_grid(ctx, net->nw[l], 256), 256, 0, ctx->stream>>>(
        net->weights + net->w_off[l], net->nw[l], stddev, seed,
        (unsigned long long)net->w_off[l]);
    PB_LAUNCHED(ctx);
  }
  return 0;
}

int pb_nnet_init_constant(pb_nnet *net, double value) {
  PB_CHECK(net, "pb_nnet_init_constant: null");
  for (size_t l = 0; l < net->layers.size(); ++l)
    if (net->nw[l] &&
        pb_buffer_fill(net->ctx, net->weights + net->w_off[l], (float)value, (size_t)net->nw[l]))
      return 1;
  return 0;
}

int pb_nnet_set_learning_rate(pb_nnet *net, double learning_rate) {
  PB_CHECK(net, "pb_nnet_set_learning_rate: null");
  const float v = (float)learning_rate;
  return pb_buffer_write_f32(net->ctx, net->lr, &v, 1);
}

int pb_nnet_evaluate(pb_nnet *net, const float *d_in, int64_t batch, float *d_out) {
  PB_CHECK(net && d_in, "pb_nnet_evaluate: null");
  PB_CHECK(batch >= 1 && batch <= net->max_batch,
           "pb_nnet_evaluate: batch exceeds the max_batch the network was created with");
  if (pb::forward(net, d_in, batch)) return 1;
  if (d_out)
    return pb_buffer_copy(net->ctx, d_out, net->outputs.back(), (size_t)(net->no.back() * batch));
  return 0;
}

int pb_nnet_layer_output(pb_nnet *net, int32_t layer, float **dptr) {
  PB_CHECK(net && dptr, "pb_nnet_layer_output: null");
  PB_CHECK(layer >= 0 && layer < (int)net->layers.size(), "Too large layer index");
  *dptr = net->outputs[layer];
  return 0;
}

int pb_nnet_error(pb_nnet *net, const float *d_out, const float *d_expected, double *error) {
  PB_CHECK(net && d_out && d_expected && error, "pb_nnet_error: null");
  const int64_t n = net->no.back();
  if (pb_error_value(net->ctx, net->loss, n, d_out, d_expected, net->err, 1)) return 1;
  std::vector<float> comps((size_t)n);
  if (pb_buffer_read_f32(net->ctx, net->err, comps.data(), (size_t)n)) return 1;
  double e = 0;  // host-side ascending sum, nnet.h:307-311
  for (int64_t i = 0; i < n; ++i) e += (double)comps[i];
  *error = e;
  return 0;
}

int pb_nnet_error_gradients(pb_nnet *net, const float *d_out, const float *d_expected,
                            float *d_gradients, int64_t batch) {
  PB_CHECK(net, "pb_nnet_error_gradients: null");
  return pb_error_gradients(net->ctx, net->loss, net->no.back(), d_out, d_expected,
                            d_gradients, batch);
}

int pb_nnet_train(pb_nnet *net, const float *d_in, const float *d_expected,
                  float *d_input_gradients, const float *d_initial_gradients) {
  PB_CHECK(net && d_in, "pb_nnet_train: null");
  PB_CHECK(d_expected || d_initial_gradients, "pb_nnet_train: need expected or gradients");
  return pb::train_step(net, d_in, d_expected, d_input_gradients, d_initial_gradients);
}

int pb_nnet_train_loop(pb_nnet *net, const float *d_ins, const float *d_expected,
                       int64_t n_samples) {
  PB_CHECK(net && d_ins && d_expected, "pb_nnet_train_loop: null");
  if (n_samples <= 0) return 0;
  pb_context *ctx = net->ctx;
  {
    // small all-dense networks (the circle MLP): the whole loop is one persistent kernel
    const int rc = pb::train_loop_small(ctx, net->layers.data(), (int)net->layers.size(),
                                        net->w_off.data(), net->total_w, net->loss, net->weights,
                                        net->lr, d_ins, d_expected, n_samples);
    if (rc >= 0) return rc;
  }
  if (!net->train_exec) {
    // Capture ONE Train() (fetch sample -> forward -> error gradient ->
    // reverse loop with in-place SGD -> advance cursor) into a CUDA graph.
    cudaGraph_t graph = nullptr;
    const uint64_t before = ctx->launches;
    PB_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    const int nin = (int)net->ni[0], nout = (int)net->no.back();
    pb::fetch_sample_kernel<<<pb_div_up(nin + nout, 256), 256, 0, ctx->stream>>>(
        net->cursor, net->stage_in, net->stage_exp, nin, nout);
    ctx->launches++;
    int rc = pb::train_step(net, net->stage_in, net->stage_exp, nullptr, nullptr);
    pb::advance_cursor_kernel<<<1, 1, 0, ctx->stream>>>(net->cursor);
    ctx->launches++;
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
    net->graph_launches = ctx->launches - before;
    ctx->launches = before;  // captured, not yet executed
    if (rc) { if (graph) cudaGraphDestroy(graph); return 1; }
    if (e != cudaSuccess) return pb::fail(std::string("graph capture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(&net->train_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return pb::fail(std::string("graph instantiate: ") + cudaGetErrorString(e));
  }
  pb_nnet::Cursor cur{d_ins, d_expected, 0};
  PB_CUDA(cudaMemcpyAsync(net->cursor, &cur, sizeof(cur), cudaMemcpyHostToDevice, ctx->stream));
  PB_CUDA(cudaStreamSynchronize(ctx->stream));  // `cur` is a stack object
  for (int64_t i = 0; i < n_samples; ++i) {
    PB_CUDA(cudaGraphLaunch(net->train_exec, ctx->stream));
    ctx->launches += net->graph_launches;
  }
  return 0;
}

// The batched gradient pass.  in_place == false: fills the delta buffer (-lr * sum_b
// grad_b, CalculateGradients); in_place == true (one chunk only): every layer's weight
// kernel writes W - lr * sum_b grad_b straight into W — BatchTrain without the delta
// round trip (nnet.h:652-665 VectorAccumulate fused into the weight kernels).
static int batch_pass_body(pb_nnet *net, const float *d_ins, const float *d_expected,
                           int64_t batch, bool in_place);

static int batch_pass(pb_nnet *net, const float *d_ins, const float *d_expected, int64_t batch,
                      bool in_place) {
  // The in-place single-chunk step launches its kernels chained (programmatic dependent
  // launch, common.cuh): each kernel's setup — filters into shared memory, TMEM allocation,
  // zero fills — overlaps the tail of the kernel before it.  PB_STEP_CHAIN=0 turns it off.
  // The delta-buffer pass (chunked and data-parallel steps) is chained too: the weights do
  // not change inside the pass and the exchange kernels sit behind it (or behind events on
  // their own stream).
  static const bool chain_on = [] {
    const char *e = getenv("PB_STEP_CHAIN");
    return !(e && e[0] == '0');
  }();
  // Forked step (common.cuh launch_forked): the fixed-order reduce + update kernels behind the
  // weight-gradient kernels (4-7 us each, four per step) run on a second stream beside the
  // next layer's kernels and are joined at the end of the pass - a branch of the step graph.
  // Nothing on the compute stream reads what they write before that: a layer's input gradient
  // (the only other reader of its weights) is queued ahead of its weight gradient, and the
  // delta buffer is read only behind the pass.  Not with a per-layer delta hook (the large
  // data-parallel exchange starts a layer's exchange right behind its reduce).
  // PB_STEP_FORK=0 turns it off.
  static const bool fork_on = [] {
    const char *e = getenv("PB_STEP_FORK");
    return !(e && e[0] == '0');
  }();
  (void)in_place;
  pb_context *ctx = net->ctx;
  ctx->chain = (chain_on && net->fuse && !net->profiling) ? 1 : 0;
  ctx->chain_last = ctx->chain != 0;
  // small networks only: the partial-sum regions of a forked pass lie one behind the other
  ctx->fork_mode = fork_on && ctx->chain != 0 && !net->delta_hook && net->total_w <= (1 << 20);
  ctx->ws_off = 0;
  ctx->forks = 0;
  int rc = batch_pass_body(net, d_ins, d_expected, batch, in_place);
  if (pb::join_forked(ctx)) rc = 1;
  ctx->fork_mode = false;
  ctx->chain = 0;
  return rc;
}

static int batch_pass_body(pb_nnet *net, const float *d_ins, const float *d_expected,
                           int64_t batch, bool in_place) {
  pb_context *ctx = net->ctx;
  const int64_t nin = net->ni[0], nout = net->no.back();
  // Batches larger than max_batch run as chunks; every chunk sees the same
  // (batch-start) weights and accumulates into the delta buffer.
  for (int64_t b0 = 0; b0 < batch; b0 += net->max_batch) {
    const int64_t nb = std::min(net->max_batch, batch - b0);
    const float *in = d_ins + b0 * nin;
    if (net->fuse) {
      std::vector<const float *> layer_in;
      std::vector<char> fused_fwd;
      const int n = (int)net->layers.size();
      const int h = net->head;
      if (pb::forward_fused(net, in, nb, layer_in, fused_fwd, h > 0)) return 1;
      int top = n - 1;
      float *g = net->grad_a, *ng = net->grad_b;
      bool head_done = false;
      if (h > 0) {
        // X = what the dense layer reads; Gz (dE/d dense output) goes to `ng`,
        // dE/dX to `g` — then the dense weight gradient, then the loop below.
        const float *X = layer_in[h];
        float *W = net->weights + net->w_off[h];
        if (!net->grad_z) {
          if (cudaMalloc((void **)&net->grad_z, sizeof(float) * (size_t)(net->no[h] * net->max_batch)) !=
              cudaSuccess) {
            cudaGetLastError();
            net->grad_z = nullptr;
          }
        }
        float *gz = net->grad_z ? net->grad_z : ng;
        const int rc = pb::head_forward_backward(
            ctx, &net->layers[h], net->layers[h + 1].activation, net->layers[h + 2].n, net->loss,
            X, W, d_expected + b0 * nout, net->outputs[h], net->outputs[h + 1],
            net->outputs[h + 2], gz, g, nb);
        if (rc > 0) return 1;
        if (rc == 0) {
          pb::mark(net, pb::layer_tag(net, h) + ".." + pb::layer_tag(net, n - 1) +
                            ":evaluate+error_gradients+input_delta (head)");
          const int m = in_place ? PB_NEW_WEIGHTS : b0 == 0 ? PB_WEIGHT_DELTAS : PB_ACCUMULATE;
          // X (a layer output) and gz (its own buffer) stay untouched until the end of the pass
          ctx->fork_operands_stable = gz != ng;
          const int wrc = pb_layer_weight_update(ctx, &net->layers[h], X, W, gz,
                                                 in_place ? W : net->deltas + net->w_off[h], net->lr, m, nb);
          ctx->fork_operands_stable = false;
          if (wrc) return 1;
          pb::mark(net, pb::layer_tag(net, h) + ":weight_update");
          if (!in_place && pb::fire_delta_hook(net, h, b0 + net->max_batch >= batch)) return 1;
          top = h - 1;
          head_done = true;
        } else {
          // outside the head kernel's envelope: finish the forward pass layer by layer
          const float *cur = X;
          for (int l = h; l < n; ++l) {
            layer_in[l] = cur;
            const float *Wl = net->nw[l] ? net->weights + net->w_off[l] : nullptr;
            if (pb_layer_evaluate(ctx, &net->layers[l], cur, Wl, net->outputs[l], nb)) return 1;
            cur = net->outputs[l];
          }
        }
      }
      if (!head_done &&
          pb_error_gradients(ctx, net->loss, nout, net->outputs.back(), d_expected + b0 * nout,
                             net->grad_a, nb))
        return 1;
      if (pb::backward_fused(net, nb, g, ng, b0 == 0, layer_in, fused_fwd, top, in_place,
                             b0 + net->max_batch >= batch))
        return 1;
      // the next chunk overwrites the layer outputs and gradient buffers the forked kernels read
      if (pb::join_forked(ctx)) return 1;
      ctx->ws_off = 0;
      continue;
    }
    if (pb::forward(net, in, nb)) return 1;
    if (pb_error_gradients(ctx, net->loss, nout, net->outputs.back(), d_expected + b0 * nout,
                           net->grad_a, nb))
      return 1;
    float *g = nullptr;
    if (pb::backward(net, in, nb, net->grad_a, net->grad_b,
                     in_place ? PB_NEW_WEIGHTS : PB_WEIGHT_DELTAS, b0 == 0, &g,
                     b0 + net->max_batch >= batch))
      return 1;
  }
  return 0;
}

int pb_nnet_batch_gradients(pb_nnet *net, const float *d_ins, const float *d_expected,
                            int64_t batch) {
  PB_CHECK(net && d_ins && d_expected, "pb_nnet_batch_gradients: null");
  PB_CHECK(batch >= 1, "pb_nnet_batch_gradients: empty batch");
  return batch_pass(net, d_ins, d_expected, batch, false);
}

int pb_nnet_apply_deltas(pb_nnet *net) {
  PB_CHECK(net, "pb_nnet_apply_deltas: null");
  if (pb_vector_accumulate(net->ctx, net->weights, net->deltas, net->total_w)) return 1;
  pb::mark(net, "vector_accumulate");
  return 0;
}

static void profile_begin(pb_nnet *net) {
  for (auto &m : net->marks) cudaEventDestroy(m.second);
  net->marks.clear();
  net->profiling = true;
  pb::mark(net, "start");
}

static int profile_end(pb_nnet *net, int rc, char *out, size_t cap) {
  net->profiling = false;
  if (rc) return 1;
  PB_CUDA(cudaStreamSynchronize(net->ctx->stream));
  std::string text;
  for (size_t i = 1; i < net->marks.size(); ++i) {
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, net->marks[i - 1].second, net->marks[i].second);
    char buf[64];
    snprintf(buf, sizeof(buf), "\t%.6f\n", ms);
    text += net->marks[i].first + buf;
  }
  for (auto &m : net->marks) cudaEventDestroy(m.second);
  net->marks.clear();
  PB_CHECK(text.size() + 1 <= cap, "pb_nnet_profile: output buffer too small");
  memcpy(out, text.c_str(), text.size() + 1);
  return 0;
}

int pb_nnet_profile_step(pb_nnet *net, const float *d_ins, const float *d_expected, int64_t batch,
                         char *out, size_t cap) {
  PB_CHECK(net && out && cap > 2, "pb_nnet_profile_step: null");
  profile_begin(net);
  return profile_end(net, pb_nnet_batch_train(net, d_ins, d_expected, batch), out, cap);
}

int pb_nnet_profile_evaluate(pb_nnet *net, const float *d_in, int64_t batch, char *out,
                             size_t cap) {
  PB_CHECK(net && out && cap > 2, "pb_nnet_profile_evaluate: null");
  profile_begin(net);
  return profile_end(net, pb_nnet_evaluate(net, d_in, batch, nullptr), out, cap);
}

int pb_nnet_batch_train(pb_nnet *net, const float *d_ins, const float *d_expected,
                        int64_t batch) {
  PB_CHECK(net && d_ins && d_expected, "pb_nnet_batch_train: null");
  PB_CHECK(batch >= 1, "pb_nnet_batch_train: empty batch");
  // One chunk: the SGD update is fused into the weight-gradient kernels.  Several chunks
  // must all see the batch-start weights, so they go through the delta buffer.
  if (batch <= net->max_batch && net->fuse) {
    struct Body { pb_nnet *net; const float *ins, *exp; int64_t batch; } body{net, d_ins, d_expected, batch};
    return pb::nnet_run_graphed(
        net, d_ins, d_expected, batch, 0,
        [](void *u) {
          Body *b = (Body *)u;
          return batch_pass(b->net, b->ins, b->exp, b->batch, true);
        },
        &body);
  }
  if (batch_pass(net, d_ins, d_expected, batch, false)) return 1;
  return pb_nnet_apply_deltas(net);
}

// Shared body of the host-fed steps.  h_records != nullptr: byte records, decoded on the
// device into the staging slot; otherwise fp32 inputs and targets.
static int batch_train_host_impl(pb_nnet *net, pb_comm *comm, const float *h_ins,
                                 const float *h_expected, const uint8_t *h_records,
                                 int64_t batch, float *h_outputs) {
  PB_CHECK(batch >= 1 && batch <= net->max_batch,
           "pb_nnet_batch_train_host: batch exceeds the max_batch the network was created with");
  pb_context *ctx = net->ctx;
  PB_CUDA(cudaSetDevice(ctx->device));
  pb_nnet::HostPipe &p = net->pipe;
  const int64_t nin = net->ni[0], nout = net->no.back();
  if (!p.copy) {
    PB_CUDA(cudaStreamCreateWithFlags(&p.copy, cudaStreamNonBlocking));
    p.cap = net->max_batch;
    for (int i = 0; i < 2; ++i) {
      PB_CUDA(cudaMalloc((void **)&p.x[i], sizeof(float) * (size_t)(p.cap * nin)));
      PB_CUDA(cudaMalloc((void **)&p.e[i], sizeof(float) * (size_t)(p.cap * nout)));
      PB_CUDA(cudaEventCreateWithFlags(&p.copied[i], cudaEventDisableTiming));
      PB_CUDA(cudaEventCreateWithFlags(&p.consumed[i], cudaEventDisableTiming));
    }
  }
  const int s = (int)(p.calls & 1);
  if (h_records && !p.rec[s])
    PB_CUDA(cudaMalloc((void **)&p.rec[s], (size_t)(p.cap * (nin + 1))));
  // slot s was last read by the step two calls ago
  if (p.calls >= 2) PB_CUDA(cudaStreamWaitEvent(p.copy, p.consumed[s], 0));
  if (h_records) {
    PB_CUDA(cudaMemcpyAsync(p.rec[s], h_records, (size_t)(batch * (nin + 1)),
                            cudaMemcpyHostToDevice, p.copy));
    // decoded on the copy stream, behind the copy: overlaps the previous step's compute
    if (pb::records_decode_on(ctx, p.copy, p.rec[s], p.x[s], p.e[s], batch, nin, nout)) return 1;
  } else {
    PB_CUDA(cudaMemcpyAsync(p.x[s], h_ins, sizeof(float) * (size_t)(batch * nin),
                            cudaMemcpyHostToDevice, p.copy));
    PB_CUDA(cudaMemcpyAsync(p.e[s], h_expected, sizeof(float) * (size_t)(batch * nout),
                            cudaMemcpyHostToDevice, p.copy));
  }
  PB_CUDA(cudaEventRecord(p.copied[s], p.copy));
  PB_CUDA(cudaStreamWaitEvent(ctx->stream, p.copied[s], 0));
  if (h_outputs && !p.out) {
    PB_CUDA(cudaStreamCreateWithFlags(&p.out, cudaStreamNonBlocking));
    ctx->side_streams.push_back(p.out);
    for (int i = 0; i < 2; ++i) {
      PB_CUDA(cudaMalloc((void **)&p.y[i], sizeof(float) * (size_t)(p.cap * nout)));
      PB_CUDA(cudaEventCreateWithFlags(&p.y_ready[i], cudaEventDisableTiming));
    }
  }
  // The step writes the network's outputs straight into y[s] (the last layer's output
  // buffer is redirected for this call): the D2H DMA then runs on the side stream while
  // the next step computes into the other buffer.  y[s] was last read by the D2H of two
  // calls ago.
  float *saved_out = net->outputs.back();
  if (h_outputs) {
    if (p.calls >= 2) PB_CUDA(cudaStreamWaitEvent(ctx->stream, p.y_ready[s], 0));
    net->outputs.back() = p.y[s];
  }
  const int rc = comm ? pb_nnet_batch_train_dp(net, comm, p.x[s], p.e[s], batch)
                      : pb_nnet_batch_train(net, p.x[s], p.e[s], batch);
  net->outputs.back() = saved_out;
  if (rc) return 1;
  PB_CUDA(cudaEventRecord(p.consumed[s], ctx->stream));
  if (h_outputs) {
    PB_CUDA(cudaStreamWaitEvent(p.out, p.consumed[s], 0));
    PB_CUDA(cudaMemcpyAsync(h_outputs, p.y[s], sizeof(float) * (size_t)(batch * nout),
                            cudaMemcpyDeviceToHost, p.out));
    PB_CUDA(cudaEventRecord(p.y_ready[s], p.out));
  }
  p.calls++;
  return 0;
}

int pb_nnet_batch_train_host(pb_nnet *net, pb_comm *comm, const float *h_ins,
                             const float *h_expected, int64_t batch, float *h_outputs) {
  PB_CHECK(net && h_ins && h_expected, "pb_nnet_batch_train_host: null");
  return batch_train_host_impl(net, comm, h_ins, h_expected, nullptr, batch, h_outputs);
}

int pb_nnet_batch_train_host_records(pb_nnet *net, pb_comm *comm, const uint8_t *h_records,
                                     int64_t batch, float *h_outputs) {
  PB_CHECK(net && h_records, "pb_nnet_batch_train_host_records: null");
  return batch_train_host_impl(net, comm, nullptr, nullptr, h_records, batch, h_outputs);
}

}  // extern "C"
