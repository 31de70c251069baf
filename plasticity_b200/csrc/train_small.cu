// Strict per-sample SGD (the loop of nnet/circle_test.cc:27-38 and nnet/cifar_test.cc:
// 423-428: one Nnet::Train per sample, every step sees the previous step's weights) for
// SMALL all-dense networks, as ONE persistent kernel: a single CTA keeps every layer's
// weights, activations and gradients in shared memory and walks the samples in order.
//
// Per-sample SGD cannot be batched or sharded (step t+1 depends on step t), and launched
// layer by layer it is pure launch latency: the circle MLP (41 parameters) is ~15 dependent
// launches per sample even as a replayed CUDA graph.  Inside one kernel a step costs a
// dozen __syncthreads instead.  Same arithmetic and the same order as the layer kernels:
//   forward   ActivationLayer / DenseLayer / SoftmaxLayer   (nnet.h:234-268)
//   dE/dO     ErrorLayer gradient                            (nnet.h:316-349)
//   reverse   input_delta with the pre-update weights, then the SGD update (nnet.h:399-460)
#include "common.cuh"
#include "kernels.cuh"

namespace pb {
namespace {

constexpr int kMaxLayers = 24;
constexpr int kMaxThreads = 256;

struct SmallNet {
  int n_layers, loss, n_in, n_out, max_io;
  int kind[kMaxLayers], act[kMaxLayers], nin[kMaxLayers], nout[kMaxLayers];
  int w_off[kMaxLayers];    // offset of the layer's weights in the flat weight buffer
  int out_off[kMaxLayers];  // offset of the layer's outputs in the activation arena
  int total_w, total_out;
};

__device__ __forceinline__ float act_fwd_rt(int act, float x) {
  switch (act) {
    case PB_RELU: return act_fwd<PB_RELU>(x);
    case PB_LEAKY_RELU: return act_fwd<PB_LEAKY_RELU>(x);
    case PB_SIGMOID: return act_fwd<PB_SIGMOID>(x);
  }
  return x;
}
__device__ __forceinline__ float act_bwd_rt(int act, float x, float g) {
  switch (act) {
    case PB_RELU: return act_bwd<PB_RELU>(x, g);
    case PB_LEAKY_RELU: return act_bwd<PB_LEAKY_RELU>(x, g);
    case PB_SIGMOID: return act_bwd<PB_SIGMOID>(x, g);
  }
  return act_bwd<PB_IDENTITY>(x, g);
}

__global__ void __launch_bounds__(kMaxThreads, 1)
train_small_kernel(const SmallNet net, float *__restrict__ W_global, const float *__restrict__ lr_p,
                   const float *__restrict__ ins, const float *__restrict__ expected,
                   long long n_samples) {
  extern __shared__ float sm[];
  float *W = sm;                       // [total_w]
  float *x = W + net.total_w;          // [n_in]
  float *outs = x + net.n_in;          // [total_out]
  float *ga = outs + net.total_out;    // [max_io]
  float *gb = ga + net.max_io;         // [max_io]
  const int tid = threadIdx.x, kThreads = blockDim.x;
  for (int i = tid; i < net.total_w; i += kThreads) W[i] = W_global[i];
  const float lr = lr_p[0];
  // the next sample's input / target are fetched into registers while the current one
  // trains (when they fit one value per thread), so no global-memory latency is exposed
  const bool prefetch = net.n_in <= kThreads && net.n_out <= kThreads;
  float x_next = 0.0f, e_next = 0.0f;
  if (prefetch && n_samples > 0) {
    if (tid < net.n_in) x_next = ins[tid];
    if (tid < net.n_out) e_next = expected[tid];
  }
  __syncthreads();

  for (long long s = 0; s < n_samples; ++s) {
    float e_cur = 0.0f;
    if (prefetch) {
      if (tid < net.n_in) x[tid] = x_next;
      e_cur = e_next;
      if (s + 1 < n_samples) {
        if (tid < net.n_in) x_next = ins[(s + 1) * net.n_in + tid];
        if (tid < net.n_out) e_next = expected[(s + 1) * net.n_out + tid];
      }
    } else {
      for (int i = tid; i < net.n_in; i += kThreads) x[i] = ins[s * net.n_in + i];
    }
    __syncthreads();
    // ---- forward ---------------------------------------------------------------
    for (int l = 0; l < net.n_layers; ++l) {
      const float *I = l ? outs + net.out_off[l - 1] : x;
      float *O = outs + net.out_off[l];
      const int nin = net.nin[l], nout = net.nout[l];
      if (net.kind[l] == PB_ACTIVATION) {
        for (int i = tid; i < nout; i += kThreads) O[i] = act_fwd_rt(net.act[l], I[i]);
      } else if (net.kind[l] == PB_DENSE) {
        const float *Wl = W + net.w_off[l];
        for (int o = tid; o < nout; o += kThreads) {
          const float *w = Wl + o * (nin + 1);
          float acc = w[nin];  // the accumulator starts at the bias (dense_layer.cc:12-28)
          for (int i = 0; i < nin; ++i) acc = fmaf(w[i], I[i], acc);
          O[o] = acc;
        }
      } else {  // softmax (softmax_layer.cc:9-59)
        for (int k = tid; k < nout; k += kThreads) {
          float m = -INFINITY;
          for (int i = 0; i < nin; ++i) m = fmaxf(m, I[i]);
          float sum = 0.0f;
          for (int i = 0; i < nin; ++i) sum += ref_exp(I[i] - m);
          O[k] = ref_exp(I[k] - m) / sum;
        }
      }
      __syncthreads();
    }
    // ---- dE/dO (error_layer.cc:81-91) ----------------------------------------------
    {
      const float *O = outs + net.out_off[net.n_layers - 1];
      for (int k = tid; k < net.n_out; k += kThreads) {
        const float e = prefetch ? e_cur : expected[s * net.n_out + k];
        ga[k] = net.loss == PB_MEAN_SQUARED ? O[k] - e : -1.0f * (e * (1.0f / (O[k] + kRefEps)));
      }
      __syncthreads();
    }
    // ---- reverse loop -----------------------------------------------------------
    float *g = ga, *ng = gb;
    for (int l = net.n_layers - 1; l >= 0; --l) {
      const float *I = l ? outs + net.out_off[l - 1] : x;
      const int nin = net.nin[l], nout = net.nout[l];
      if (net.kind[l] == PB_ACTIVATION) {
        for (int i = tid; i < nin; i += kThreads) ng[i] = act_bwd_rt(net.act[l], I[i], g[i]);
      } else if (net.kind[l] == PB_DENSE) {
        float *Wl = W + net.w_off[l];
        for (int i = tid; i < nin; i += kThreads) {  // with the pre-update weights
          float acc = 0.0f;
          for (int o = 0; o < nout; ++o) acc = fmaf(g[o], Wl[o * (nin + 1) + i], acc);
          ng[i] = acc;
        }
        __syncthreads();
        for (int w = tid; w < nout * (nin + 1); w += kThreads) {
          const int o = w / (nin + 1), i = w - o * (nin + 1);
          Wl[w] = Wl[w] - lr * (i < nin ? g[o] * I[i] : g[o]);
        }
      } else {  // softmax: dI[k] = sum_i (O_i * (delta_ik - O_k)) * G_i, ascending i
        const float *O = outs + net.out_off[l];
        for (int k = tid; k < nin; k += kThreads) {
          const float ok = O[k];
          float acc = 0.0f;
          for (int i = 0; i < nout; ++i) acc += (O[i] * (((i == k) ? 1.0f : 0.0f) - ok)) * g[i];
          ng[k] = acc;
        }
      }
      __syncthreads();
      float *t = g; g = ng; ng = t;
    }
  }
  for (int i = tid; i < net.total_w; i += kThreads) W_global[i] = W[i];
}

}  // namespace

// Returns -1 outside the envelope (layers other than activation / dense / softmax, more
// than kMaxLayers layers, or state that does not fit one CTA's shared memory).
// w_off: offsets of the layers' weight blocks inside `weights` (the network's flat buffer).
int train_loop_small(pb_context *ctx, const pb_layer_desc *layers, int n_layers,
                     const int64_t *w_off, int64_t total_w, int loss, float *weights,
                     const float *lr, const float *d_ins, const float *d_expected,
                     int64_t n_samples) {
  static const bool off = [] {
    const char *e = getenv("PB_TRAIN_SMALL");
    return e && e[0] == '0';
  }();
  if (off || n_layers > kMaxLayers || n_layers < 1 || total_w > (1 << 16)) return -1;
  SmallNet net{};
  net.n_layers = n_layers;
  net.loss = loss;
  net.total_w = (int)total_w;
  int64_t out_total = 0, max_io = 1;
  for (int l = 0; l < n_layers; ++l) {
    const pb_layer_desc &d = layers[l];
    int64_t a = 0, b = 0;
    if (d.kind != PB_ACTIVATION && d.kind != PB_DENSE && d.kind != PB_SOFTMAX) return -1;
    if (pb_layer_num_inputs(&d, &a) || pb_layer_num_outputs(&d, &b)) return -1;
    if (a > 4096 || b > 4096) return -1;
    net.kind[l] = d.kind;
    net.act[l] = d.activation;
    net.nin[l] = (int)a;
    net.nout[l] = (int)b;
    net.w_off[l] = (int)w_off[l];
    net.out_off[l] = (int)out_total;
    out_total += b;
    max_io = std::max(max_io, std::max(a, b));
  }
  net.n_in = net.nin[0];
  net.n_out = net.nout[n_layers - 1];
  net.max_io = (int)max_io;
  net.total_out = (int)out_total;
  const size_t smem = sizeof(float) * (size_t)(total_w + net.n_in + out_total + 2 * max_io);
  if (smem > 200 * 1024) return -1;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    PB_CUDA(cudaFuncSetAttribute(train_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    configured = smem;
  }
  // one warp per 32 values of the widest layer: narrow networks synchronise a single warp
  const int threads = (int)std::min<int64_t>(kMaxThreads, ((max_io + 31) / 32) * 32);
  train_small_kernel<<<1, threads, smem, ctx->stream>>>(net, weights, lr, d_ins, d_expected,
                                                        (long long)n_samples);
  PB_LAUNCHED(ctx);
  return 0;
}

}  // namespace pb
