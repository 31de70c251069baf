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
  int wt_off[kMaxLayers];   // one-warp variant: offset of the layer's transposed weight block
  int wt_total;
};

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

// ---- one-warp variant: every layer at most 32 values wide --------------------------------
// Lane j owns value j of every layer, so a step needs no CTA barrier at all (__syncwarp
// only).  Dense weights live in shared memory TRANSPOSED with pitch 33 —
// wT[i * 33 + o] = W[o][i], row nin = the biases — so that both the forward product (fixed
// input i, lanes = outputs) and the input gradient (fixed output o, lanes = inputs) read
// conflict-free, and the activations of a layer are read as broadcasts.  The gradient vector
// lives in one register per lane.  Same arithmetic and accumulation order as the kernel
// above (ascending-index FMA chains; the softmax denominator is summed in ascending order
// from the per-lane exponentials).
constexpr int kWarpPitch = 33;

__global__ void __launch_bounds__(32, 1)
train_warp_kernel(const SmallNet net, float *__restrict__ W_global, const float *__restrict__ lr_p,
                  const float *__restrict__ ins, const float *__restrict__ expected,
                  long long n_samples) {
  extern __shared__ float sm[];
  float *acts = sm + net.wt_total;                 // [n_layers + 1][32]: acts[0] = the sample
  float *gs = acts + (net.n_layers + 1) * 32;      // [32] broadcast buffer
  const int lane = threadIdx.x;
  for (int l = 0; l < net.n_layers; ++l) {
    if (net.kind[l] != PB_DENSE) continue;
    const int ld = net.nin[l] + 1;
    for (int idx = lane; idx < net.nout[l] * ld; idx += 32) {
      const int o = idx / ld, i = idx - o * ld;
      sm[net.wt_off[l] + i * kWarpPitch + o] = W_global[net.w_off[l] + idx];
    }
  }
  const float lr = lr_p[0];
  float x_next = 0.0f, e_next = 0.0f;
  if (n_samples > 0) {
    if (lane < net.n_in) x_next = ins[lane];
    if (lane < net.n_out) e_next = expected[lane];
  }
  __syncwarp();

  for (long long s = 0; s < n_samples; ++s) {
    if (lane < net.n_in) acts[lane] = x_next;
    const float e_cur = e_next;
    if (s + 1 < n_samples) {
      if (lane < net.n_in) x_next = ins[(s + 1) * net.n_in + lane];
      if (lane < net.n_out) e_next = expected[(s + 1) * net.n_out + lane];
    }
    __syncwarp();
    // ---- forward ---------------------------------------------------------------
    for (int l = 0; l < net.n_layers; ++l) {
      const float *I = acts + l * 32;
      float *O = acts + (l + 1) * 32;
      const int nin = net.nin[l], nout = net.nout[l];
      if (net.kind[l] == PB_ACTIVATION) {
        if (lane < nout) O[lane] = act_fwd_rt(net.act[l], I[lane]);
      } else if (net.kind[l] == PB_DENSE) {
        if (lane < nout) {
          const float *w = sm + net.wt_off[l] + lane;
          float acc = w[nin * kWarpPitch];  // the accumulator starts at the bias
          for (int i = 0; i < nin; ++i) acc = fmaf(w[i * kWarpPitch], I[i], acc);
          O[lane] = acc;
        }
      } else {  // softmax
        float m = lane < nin ? I[lane] : -INFINITY;
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        const float p = lane < nin ? ref_exp(I[lane] - m) : 0.0f;
        gs[lane] = p;
        __syncwarp();
        float sum = 0.0f;
        for (int i = 0; i < nin; ++i) sum += gs[i];
        if (lane < nout) O[lane] = p / sum;
      }
      __syncwarp();
    }
    // ---- dE/dO (error_layer.cc:81-91): lane k holds the gradient of output k -----------
    float g = 0.0f;
    if (lane < net.n_out) {
      const float o = acts[net.n_layers * 32 + lane];
      g = net.loss == PB_MEAN_SQUARED ? o - e_cur : -1.0f * (e_cur * (1.0f / (o + kRefEps)));
    }
    // ---- reverse loop -----------------------------------------------------------
    for (int l = net.n_layers - 1; l >= 0; --l) {
      const float *I = acts + l * 32;
      const int nin = net.nin[l], nout = net.nout[l];
      if (net.kind[l] == PB_ACTIVATION) {
        g = lane < nin ? act_bwd_rt(net.act[l], I[lane], g) : 0.0f;
      } else if (net.kind[l] == PB_DENSE) {
        gs[lane] = lane < nout ? g : 0.0f;
        __syncwarp();
        float ng = 0.0f;
        if (lane < nin) {  // with the pre-update weights
          const float *w = sm + net.wt_off[l] + lane * kWarpPitch;
          for (int o = 0; o < nout; ++o) ng = fmaf(gs[o], w[o], ng);
        }
        __syncwarp();
        if (lane < nout) {
          float *w = sm + net.wt_off[l] + lane;
          for (int i = 0; i < nin; ++i) w[i * kWarpPitch] = w[i * kWarpPitch] - lr * (g * I[i]);
          w[nin * kWarpPitch] = w[nin * kWarpPitch] - lr * g;
        }
        g = ng;
        __syncwarp();
      } else {  // softmax: dI[k] = sum_i (O_i * (delta_ik - O_k)) * G_i, ascending i
        const float *O = I + 32;
        gs[lane] = lane < nout ? g : 0.0f;
        __syncwarp();
        float acc = 0.0f;
        if (lane < nin) {
          const float ok = O[lane];
          for (int i = 0; i < nout; ++i) acc += (O[i] * (((i == lane) ? 1.0f : 0.0f) - ok)) * gs[i];
        }
        g = acc;
        __syncwarp();
      }
    }
  }
  for (int l = 0; l < net.n_layers; ++l) {
    if (net.kind[l] != PB_DENSE) continue;
    const int ld = net.nin[l] + 1;
    for (int idx = lane; idx < net.nout[l] * ld; idx += 32) {
      const int o = idx / ld, i = idx - o * ld;
      W_global[net.w_off[l] + idx] = sm[net.wt_off[l] + i * kWarpPitch + o];
    }
  }
}

}  // namespace

// Returns -1 outside the envelope (layers other than activation / dense / softmax, more
// than kMaxLayers layers, or state that does not fit one CTA's shared memory).
// w_off: offsets of the layers' weight blocks inside `weights` (the network's flat buffer).
int train_loop_small(pb_context *ctx, const pb_layer_desc *layers, int n_layers,
                     const int64_t *w_off, int64_t total_w, int loss, float *weights,
                     const float *lr, const float *d_ins, const float *d_expected,
                     int64_t n_samples) {
  const char *small_env = getenv("PB_TRAIN_SMALL");  // read per call: one getenv per loop
  const bool off = small_env && small_env[0] == '0';
  if (off || n_layers > kMaxLayers || n_layers < 1 || total_w > (1 << 16)) return -1;
  SmallNet net{};
  net.n_layers = n_layers;
  net.loss = loss;
  net.total_w = (int)total_w;
  int64_t out_total = 0, max_io = 1, wt_total = 0;
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
    net.wt_off[l] = (int)wt_total;
    if (d.kind == PB_DENSE) wt_total += (a + 1) * kWarpPitch;
    out_total += b;
    max_io = std::max(max_io, std::max(a, b));
  }
  net.n_in = net.nin[0];
  net.n_out = net.nout[n_layers - 1];
  net.max_io = (int)max_io;
  net.total_out = (int)out_total;
  net.wt_total = (int)wt_total;
  // networks no wider than a warp run the one-warp kernel
  if (max_io <= 32) {
    const size_t wsmem = sizeof(float) * (size_t)(wt_total + (n_layers + 1) * 32 + 32);
    if (wsmem <= 200 * 1024) {
      static size_t wconfigured = 0;
      if (wsmem > 48 * 1024 && wsmem > wconfigured) {
        PB_CUDA(cudaFuncSetAttribute(train_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)wsmem));
        wconfigured = wsmem;
      }
      train_warp_kernel<<<1, 32, wsmem, ctx->stream>>>(net, weights, lr, d_ins, d_expected,
                                                       (long long)n_samples);
      PB_LAUNCHED(ctx);
      return 0;
    }
  }
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
