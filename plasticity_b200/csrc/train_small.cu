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
#include <type_traits>

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
// Lane j owns value j of every layer.  The value a layer has just produced stays in a
// REGISTER and is handed to the next layer with shuffles (a dense product is one
// LDS + SHFL + FMA per input); the per-layer activations are also stored to shared memory,
// but only their own lane ever reads them back (the backward pass), so the forward pass has
// no synchronisation at all.  Dense weights live in shared memory TRANSPOSED with pitch 33 —
// wT[i * 33 + o] = W[o][i], row nin = the biases — so that the forward product and the
// update (lane = output o, own column) and the input gradient (lane = input i, own row) all
// read conflict-free; the only cross-lane hand-over through memory is update -> next
// sample's input gradient, one __syncwarp per dense layer and sample.  Same arithmetic and
// accumulation order as the kernel above (ascending-index FMA chains; the softmax denominator
// is summed in ascending order).
constexpr int kWarpPitch = 33;
constexpr unsigned kFull = 0xffffffffu;

// acc = bias + sum_i w[i][lane] * x_i (ascending i), x_i = value i of `cur`; eight inputs at a
// time: all loads and shuffles of a batch are issued before its FMA chain, so a single
// in-order warp pays the shared-memory / shuffle latency once per eight inputs, not per input
__device__ __forceinline__ float warp_dense_fwd(const float *w, int nin, float cur, float acc, bool live) {
  for (int i0 = 0; i0 < nin; i0 += 8) {
    float wv[8], xv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = min(i0 + k, nin - 1);
      wv[k] = w[i * kWarpPitch];
      xv[k] = __shfl_sync(kFull, cur, i);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (live && i0 + k < nin) acc = fmaf(wv[k], xv[k], acc);
  }
  return acc;
}

// NL > 0: the layer loops are unrolled for exactly NL layers (every per-layer parameter is a
// constant-bank operand instead of a dependent indexed load); NL == 0: any layer count.
template <int NL>
__global__ void __launch_bounds__(32, 1)
train_warp_kernel(const SmallNet net, float *__restrict__ W_global, const float *__restrict__ lr_p,
                  const float *__restrict__ ins, const float *__restrict__ expected,
                  long long n_samples) {
  extern __shared__ float sm[];
  const int n_layers = NL ? NL : net.n_layers;
  float *acts = sm + net.wt_total;                 // [n_layers + 1][32]: acts[0] = the sample
  const int lane = threadIdx.x;
  for (int l = 0; l < n_layers; ++l) {
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
    float cur = x_next;                             // value `lane` of the current layer's input
    const float e_cur = e_next;
    acts[lane] = cur;
    if (s + 1 < n_samples) {
      if (lane < net.n_in) x_next = ins[(s + 1) * net.n_in + lane];
      if (lane < net.n_out) e_next = expected[(s + 1) * net.n_out + lane];
    }
    // ---- forward ---------------------------------------------------------------
#pragma unroll
    for (int l = 0; l < n_layers; ++l) {
      const int nin = net.nin[l], nout = net.nout[l];
      float out = 0.0f;
      if (net.kind[l] == PB_ACTIVATION) {
        out = lane < nout ? act_fwd_rt(net.act[l], cur) : 0.0f;
      } else if (net.kind[l] == PB_DENSE) {
        const float *w = sm + net.wt_off[l] + lane;
        // the accumulator starts at the bias
        out = warp_dense_fwd(w, nin, cur, lane < nout ? w[nin * kWarpPitch] : 0.0f, lane < nout);
      } else {  // softmax
        float m = lane < nin ? cur : -INFINITY;
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
        const float p = lane < nin ? ref_exp(cur - m) : 0.0f;
        float sum = 0.0f;
        for (int i = 0; i < nin; ++i) sum += __shfl_sync(kFull, p, i);
        out = lane < nout ? p / sum : 0.0f;
      }
      cur = out;
      acts[(l + 1) * 32 + lane] = out;              // read back by this lane only
    }
    // ---- dE/dO (error_layer.cc:81-91): lane k holds the gradient of output k -----------
    float g = 0.0f;
    if (lane < net.n_out)
      g = net.loss == PB_MEAN_SQUARED ? cur - e_cur : -1.0f * (e_cur * (1.0f / (cur + kRefEps)));
    // ---- reverse loop -----------------------------------------------------------
#pragma unroll
    for (int l = n_layers - 1; l >= 0; --l) {
      const float in = acts[l * 32 + lane];         // this lane's input value of layer l
      const int nin = net.nin[l], nout = net.nout[l];
      if (net.kind[l] == PB_ACTIVATION) {
        g = lane < nin ? act_bwd_rt(net.act[l], in, g) : 0.0f;
      } else if (net.kind[l] == PB_DENSE) {
        float ng = 0.0f;
        // row = this lane as an input (lanes past nin read row 0, their sum is dropped)
        const float *wr = sm + net.wt_off[l] + (lane < nin ? lane : 0) * kWarpPitch;
        for (int o0 = 0; o0 < nout; o0 += 8) {       // with the pre-update weights
          float wv[8], gv[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int o = min(o0 + k, nout - 1);
            wv[k] = wr[o];
            gv[k] = __shfl_sync(kFull, g, o);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (o0 + k < nout) ng = fmaf(gv[k], wv[k], ng);
        }
        if (lane >= nin) ng = 0.0f;
        __syncwarp();                                // every row is read before any column moves
        float *w = sm + net.wt_off[l] + (lane < nout ? lane : 0);   // column = this lane as an output
        const float bias_old = w[nin * kWarpPitch];
        for (int i0 = 0; i0 < nin; i0 += 8) {
          float wv[8], xv[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int i = min(i0 + k, nin - 1);
            wv[k] = w[i * kWarpPitch];
            xv[k] = __shfl_sync(kFull, in, i);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (lane < nout && i0 + k < nin) w[(i0 + k) * kWarpPitch] = wv[k] - lr * (g * xv[k]);
        }
        if (lane < nout) w[nin * kWarpPitch] = bias_old - lr * g;
        __syncwarp();                                // columns visible to the next row reads
        g = ng;
      } else {  // softmax: dI[k] = sum_i (O_i * (delta_ik - O_k)) * G_i, ascending i
        const float ok = acts[(l + 1) * 32 + lane];
        float acc = 0.0f;
        for (int i = 0; i < nout; ++i) {
          const float oi = __shfl_sync(kFull, ok, i), gi = __shfl_sync(kFull, g, i);
          acc += (oi * (((i == lane) ? 1.0f : 0.0f) - ok)) * gi;
        }
        g = lane < nin ? acc : 0.0f;
      }
    }
  }
  __syncwarp();
  for (int l = 0; l < n_layers; ++l) {
    if (net.kind[l] != PB_DENSE) continue;
    const int ld = net.nin[l] + 1;
    for (int idx = lane; idx < net.nout[l] * ld; idx += 32) {
      const int o = idx / ld, i = idx - o * ld;
      W_global[net.w_off[l] + idx] = sm[net.wt_off[l] + i * kWarpPitch + o];
    }
  }
}

// ---- register-resident variant: [input activation] (dense, activation) x ND, all <= 32 wide ----
// The chain above still pays a shared-memory round trip per weight and sample (update ->
// next sample's product) and a __syncwarp per dense layer.  For the networks the reference's
// AddDenseLayer chain builds (nnet/circle_test.cc:19-22: Identity input, then dense + activation
// pairs) every weight lives in REGISTERS, twice: lane o holds row o of W (the product of the
// forward pass and the update read it: out_o = b_o + sum_i W[o][i] x_i), lane i holds column i
// (the input gradient reads it: ng_i = sum_o g_o W[o][i]).  Both copies get the same update
// (w - lr * (g_o * x_i), the same expression, bit-identical), one with x_i broadcast by a
// shuffle, the other with g_o.  Nothing goes through memory inside a sample; the dependent
// chain of a sample is shuffle -> FMA chain -> activation per layer.  Same arithmetic and
// accumulation order as the kernels above.  The input gradient of the first dense layer has no
// reader in a training loop and is not formed.
template <class F>
__device__ __forceinline__ void by_size(int n, F &&f) {
  if (n <= 4) f(std::integral_constant<int, 4>{});
  else if (n <= 8) f(std::integral_constant<int, 8>{});
  else if (n <= 16) f(std::integral_constant<int, 16>{});
  else f(std::integral_constant<int, 32>{});
}

template <int ND>
__global__ void __launch_bounds__(32, 1)
train_regs_kernel(const SmallNet net, float *__restrict__ W_global, const float *__restrict__ lr_p,
                  const float *__restrict__ ins, const float *__restrict__ expected,
                  long long n_samples) {
  const int lane = threadIdx.x;
  float wr[ND][32], br[ND], wc[ND][32];
  int nin[ND], nout[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    const int l = 2 * d + 1;
    nin[d] = net.nin[l];
    nout[d] = net.nout[l];
    const float *base = W_global + net.w_off[l];
    const int ld = nin[d] + 1;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      wr[d][i] = (lane < nout[d] && i < nin[d]) ? base[lane * ld + i] : 0.0f;
      wc[d][i] = (lane < nin[d] && i < nout[d]) ? base[i * ld + lane] : 0.0f;
    }
    br[d] = lane < nout[d] ? base[lane * ld + nin[d]] : 0.0f;
  }
  const float lr = lr_p[0];
  float x_next = 0.0f, e_next = 0.0f;
  if (n_samples > 0) {
    if (lane < net.n_in) x_next = ins[lane];
    if (lane < net.n_out) e_next = expected[lane];
  }
  for (long long s = 0; s < n_samples; ++s) {
    const float x_cur = x_next, e_cur = e_next;
    if (s + 1 < n_samples) {
      if (lane < net.n_in) x_next = ins[(s + 1) * net.n_in + lane];
      if (lane < net.n_out) e_next = expected[(s + 1) * net.n_out + lane];
    }
    // ---- forward: a[d] = input of dense layer d, z[d] = its output ----------------------
    // The loops over a layer's inputs / outputs are unrolled for the next power of two (4, 8,
    // 16, 32) with predicated bodies: static register indices, no per-iteration branch (a
    // single warp pays ~15 cycles for every one).
    float a[ND + 1], z[ND], pz[ND];
    a[0] = lane < net.n_in ? act_fwd_rt(net.act[0], x_cur) : 0.0f;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      float acc = br[d];
      by_size(nin[d], [&](auto N) {
#pragma unroll
        for (int i = 0; i < decltype(N)::value; ++i) {
          const float xi = __shfl_sync(kFull, a[d], i);
          if (i < nin[d]) acc = fmaf(wr[d][i], xi, acc);
        }
      });
      z[d] = acc;
      const int act = net.act[2 * d + 2];
      float out;
      if (act == PB_SIGMOID) {   // p = 2.71828^-z is kept for the derivative (the same value)
        pz[d] = ref_exp(-acc);
        out = 1.0f / (1.0f + pz[d]);
      } else {
        pz[d] = 0.0f;
        out = act_fwd_rt(act, acc);
      }
      a[d + 1] = lane < nout[d] ? out : 0.0f;
    }
    // ---- dE/dO (error_layer.cc:81-91) ------------------------------------------------------
    float g = 0.0f;
    if (lane < net.n_out)
      g = net.loss == PB_MEAN_SQUARED ? a[ND] - e_cur : -1.0f * (e_cur * (1.0f / (a[ND] + kRefEps)));
    // ---- reverse loop ----------------------------------------------------------------------
#pragma unroll
    for (int d = ND - 1; d >= 0; --d) {
      const int act = net.act[2 * d + 2];
      float gz;
      if (act == PB_SIGMOID) {
        const float q = 1.0f + pz[d];
        gz = g * (pz[d] / (q * q));
      } else {
        gz = act_bwd_rt(act, z[d], g);
      }
      g = lane < nout[d] ? gz : 0.0f;   // lane = output o
      float ng = 0.0f;
      if (d > 0) {   // with the pre-update weights; lane = input i
        by_size(nout[d], [&](auto N) {
#pragma unroll
          for (int o = 0; o < decltype(N)::value; ++o) {
            const float go = __shfl_sync(kFull, g, o);
            if (o < nout[d]) ng = fmaf(go, wc[d][o], ng);
          }
        });
        if (lane >= nin[d]) ng = 0.0f;
      }
      const float x_own = lane < nin[d] ? a[d] : 0.0f;
      by_size(nin[d], [&](auto N) {
#pragma unroll
        for (int i = 0; i < decltype(N)::value; ++i) {
          const float xi = __shfl_sync(kFull, a[d], i);
          if (i < nin[d]) wr[d][i] = wr[d][i] - lr * (g * xi);
        }
      });
      br[d] = br[d] - lr * g;
      by_size(nout[d], [&](auto N) {
#pragma unroll
        for (int o = 0; o < decltype(N)::value; ++o) {
          const float go = __shfl_sync(kFull, g, o);
          if (o < nout[d]) wc[d][o] = wc[d][o] - lr * (go * x_own);
        }
      });
      g = ng;
    }
  }
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    float *base = W_global + net.w_off[2 * d + 1];
    const int ld = nin[d] + 1;
    if (lane < nout[d]) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < nin[d]) base[lane * ld + i] = wr[d][i];
      base[lane * ld + nin[d]] = br[d];
    }
  }
}

typedef void (*warp_kernel_fn)(const SmallNet, float *, const float *, const float *, const float *,
                               long long);
warp_kernel_fn warp_kernel_for(int n_layers) {
  switch (n_layers) {
    case 1: return train_warp_kernel<1>;
    case 2: return train_warp_kernel<2>;
    case 3: return train_warp_kernel<3>;
    case 4: return train_warp_kernel<4>;
    case 5: return train_warp_kernel<5>;
    case 6: return train_warp_kernel<6>;
    case 7: return train_warp_kernel<7>;
    case 8: return train_warp_kernel<8>;
  }
  return train_warp_kernel<0>;
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
  // Identity/activation input, then (dense, activation) pairs, nothing wider than a warp:
  // every weight in registers
  if (max_io <= 32 && (n_layers == 3 || n_layers == 5 || n_layers == 7)) {
    const char *regs_env = getenv("PB_TRAIN_REGS");
    bool fits = !(regs_env && regs_env[0] == '0') && net.kind[0] == PB_ACTIVATION;
    for (int l = 1; l < n_layers && fits; ++l)
      fits = net.kind[l] == ((l & 1) ? PB_DENSE : PB_ACTIVATION);
    if (fits) {
      warp_kernel_fn kern = n_layers == 3 ? train_regs_kernel<1>
                            : n_layers == 5 ? train_regs_kernel<2> : train_regs_kernel<3>;
      kern<<<1, 32, 0, ctx->stream>>>(net, weights, lr, d_ins, d_expected, (long long)n_samples);
      PB_LAUNCHED(ctx);
      return 0;
    }
  }
  // networks no wider than a warp run the one-warp kernel
  if (max_io <= 32) {
    const size_t wsmem = sizeof(float) * (size_t)(wt_total + (n_layers + 1) * 32);
    if (wsmem <= 200 * 1024) {
      warp_kernel_fn kern = warp_kernel_for(n_layers);
      if (wsmem > 48 * 1024)
        PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
      kern<<<1, 32, wsmem, ctx->stream>>>(net, weights, lr, d_ins, d_expected, (long long)n_samples);
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
