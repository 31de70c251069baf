// Strict per-sample SGD (the loop of nnet/cifar_test.cc:423-428: one Nnet::Train per
// sample, every step sees the previous step's weights) for ANY layer list, as ONE
// cooperative persistent kernel: the whole grid walks the samples in order, a layer's
// forward / input-gradient / weight-update is a grid-stride phase and phases are separated
// by grid-wide barriers instead of kernel boundaries.
//
// Launched layer by layer a CIFAR sample is ~31 dependent launches (~250 us as a replayed
// CUDA graph, nnet.cu); here it is ~25 grid barriers.  Weights stay in global memory (L2
// resident: 90 KB for the CIFAR network), activations in the network's per-layer output
// buffers, the gradient vector ping-pongs between two buffers.  Same arithmetic as the
// layer kernels of conv.cu / dense.cu / elementwise.cu at batch 1:
//   forward   nnet.h:234-268      dE/dO  nnet.h:316-349 (error_layer.cc:81-91)
//   reverse   input_delta with the pre-update weights, then W <- W - lr * g (nnet.h:399-460)
//
// OPT-IN (PB_TRAIN_COOP=1) until it has been through the GPU parity tests
// (tests/test_gpu_parity.py::test_train_loop_cooperative_kernel, PB_EXPERIMENTAL=1).
//
// Every buffer here is rewritten by other CTAs between barriers: no __restrict__ / __ldg on
// anything but the sample stream, so no load takes the non-coherent path.
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace pb {
namespace {

constexpr int kCoopMaxLayers = 24;
constexpr int kCoopThreads = 256;
constexpr int kCoopMaxSoftmax = 1024;

struct CoopLayer {
  int kind, act, nin, nout, nw, w_off;
  // convolution / pool geometry (ConvGeom of conv.cu; pool: W, H, D, ow, oh)
  int W, H, D, fw, fh, s, p, F, ow, oh, K1;
};

struct CoopNet {
  int n_layers, loss, n_in, n_out;
  CoopLayer L[kCoopMaxLayers];
  float *out[kCoopMaxLayers];  // layer outputs of the current sample
  float *g[2];                 // gradient ping-pong
};

struct Who {
  int tid, nthreads;    // over the whole grid
  int warp, nwarps, lane;
};

// ---- forward phases ----------------------------------------------------------------

__device__ void fwd_activation(const CoopLayer &L, const float *I, float *O, const Who &w) {
  for (int t = w.tid; t < L.nout; t += w.nthreads) O[t] = act_fwd_rt(L.act, I[t]);
}

// one warp per output: lanes stride over the weight row (dense.cu dense_fwd_gemv_kernel)
__device__ void fwd_dense(const CoopLayer &L, const float *I, float *O, const float *Wt,
                          const Who &w) {
  const int ld = L.nin + 1;
  for (int o = w.warp; o < L.nout; o += w.nwarps) {
    const float *row = Wt + (int64_t)o * ld;
    float acc = 0.0f;
    for (int i = w.lane; i < L.nin; i += 32) acc = fmaf(row[i], I[i], acc);
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (w.lane == 0) O[o] = row[L.nin] + acc;
  }
}

// one thread per output (conv.cu conv_fwd_general_kernel)
__device__ void fwd_conv(const CoopLayer &L, const float *I, float *O, const float *Wt,
                         const Who &w) {
  const int oplane = L.ow * L.oh, iplane = L.W * L.H, fplane = L.fw * L.fh;
  for (int t = w.tid; t < oplane * L.F; t += w.nthreads) {
    const int f = t / oplane, rem = t - f * oplane;
    const int r = rem / L.ow, c = rem - r * L.ow;
    const int y0 = r * L.s - L.p, x0 = c * L.s - L.p;
    const float *wf = Wt + (int64_t)f * L.K1;
    float acc = wf[L.K1 - 1];
    for (int z = 0; z < L.D; ++z)
      for (int y = 0; y < L.fh; ++y) {
        const int iy = y0 + y;
        if (iy < 0 || iy >= L.H) continue;
        for (int x = 0; x < L.fw; ++x) {
          const int ix = x0 + x;
          if (ix < 0 || ix >= L.W) continue;
          acc = fmaf(wf[fplane * z + y * L.fw + x], I[iplane * z + iy * L.W + ix], acc);
        }
      }
    O[t] = acc;
  }
}

// one thread per output, strict '>' scan (elementwise.cu pool_fwd_kernel)
__device__ void fwd_pool(const CoopLayer &L, const float *I, float *O, const Who &w) {
  const int gw = L.W / L.ow, gh = L.H / L.oh, oplane = L.ow * L.oh;
  for (int t = w.tid; t < oplane * L.D; t += w.nthreads) {
    const int z = t / oplane, rem = t - z * oplane;
    const int orow = rem / L.ow, ocol = rem - orow * L.ow;
    const float *src = I + (int64_t)z * L.W * L.H + (orow * gh) * L.W + ocol * gw;
    float m = -INFINITY;
    for (int r = 0; r < gh; ++r)
      for (int c = 0; c < gw; ++c) {
        const float v = src[r * L.W + c];
        if (v > m) m = v;
      }
    O[t] = m;
  }
}

// CTA 0 only; the denominator is summed in ascending order (softmax_layer.cc:9-40)
__device__ void fwd_softmax(const CoopLayer &L, const float *I, float *O, float *ps) {
  if (blockIdx.x != 0) return;
  const int n = L.nin;
  float m = -INFINITY;
  for (int i = 0; i < n; ++i) m = fmaxf(m, I[i]);
  for (int i = threadIdx.x; i < n; i += blockDim.x) ps[i] = ref_exp(I[i] - m);
  __syncthreads();
  float sum = 0.0f;
  for (int i = 0; i < n; ++i) sum += ps[i];
  for (int k = threadIdx.x; k < n; k += blockDim.x) O[k] = ps[k] / sum;
  __syncthreads();  // ps is reused by the next sample
}

// ---- reverse phases ----------------------------------------------------------------

__device__ void bwd_activation(const CoopLayer &L, const float *I, const float *G, float *dI,
                               const Who &w) {
  for (int t = w.tid; t < L.nin; t += w.nthreads) dI[t] = act_bwd_rt(L.act, I[t], G[t]);
}

// one thread per input, ascending outputs (dense.cu dense_bwd_gemv_kernel)
__device__ void bwd_dense(const CoopLayer &L, const float *G, float *dI, const float *Wt,
                          const Who &w) {
  const int ld = L.nin + 1;
  for (int i = w.tid; i < L.nin; i += w.nthreads) {
    float acc = 0.0f;
    for (int o = 0; o < L.nout; ++o) acc = fmaf(G[o], Wt[(int64_t)o * ld + i], acc);
    dI[i] = acc;
  }
}

// one thread per weight (dense.cu dense_wupd_small_kernel at batch 1, PB_NEW_WEIGHTS)
__device__ void upd_dense(const CoopLayer &L, const float *I, const float *G, float *Wt, float lr,
                          const Who &w) {
  const int ld = L.nin + 1;
  for (int k = w.tid; k < L.nw; k += w.nthreads) {
    const int node = k / ld, edge = k - node * ld;
    float g = 0.0f;
    g += (edge < L.nin) ? G[node] * I[edge] : G[node];
    Wt[k] = Wt[k] - lr * g;
  }
}

// one thread per input element, centre guard included (conv.cu conv_bwd_data_general_kernel)
__device__ void bwd_conv(const CoopLayer &L, const float *G, float *dI, const float *Wt,
                         const Who &w) {
  const int oplane = L.ow * L.oh, iplane = L.W * L.H, fplane = L.fw * L.fh;
  const int hh = L.fh / 2, hw = L.fw / 2;
  for (int t = w.tid; t < iplane * L.D; t += w.nthreads) {
    const int z = t / iplane, rem = t - z * iplane;
    const int r = rem / L.W, c = rem - r * L.W;
    int d_lo = (r + L.p - L.fh + 1);
    d_lo = d_lo <= 0 ? 0 : (d_lo + L.s - 1) / L.s;
    int d_hi = (r + L.p) / L.s;
    if (d_hi > L.oh - 1) d_hi = L.oh - 1;
    int k_lo = (c + L.p - L.fw + 1);
    k_lo = k_lo <= 0 ? 0 : (k_lo + L.s - 1) / L.s;
    int k_hi = (c + L.p) / L.s;
    if (k_hi > L.ow - 1) k_hi = L.ow - 1;
    float acc = 0.0f;
    for (int d = d_lo; d <= d_hi; ++d) {
      const int cy = d * L.s - L.p + hh;  // centre guard (reference quirk)
      if (cy < 0 || cy >= L.H) continue;
      const int srow = r + L.p - d * L.s;
      for (int k = k_lo; k <= k_hi; ++k) {
        const int cx = k * L.s - L.p + hw;
        if (cx < 0 || cx >= L.W) continue;
        const int scol = c + L.p - k * L.s;
        const float *wp = Wt + fplane * z + srow * L.fw + scol;
        const float *gp = G + d * L.ow + k;
        for (int f = 0; f < L.F; ++f)
          acc = fmaf(wp[(int64_t)f * L.K1], gp[(int64_t)f * oplane], acc);
      }
    }
    dI[t] = acc;
  }
}

// one warp per weight: lanes stride over the output plane, shuffle sum, lane 0 updates
__device__ void upd_conv(const CoopLayer &L, const float *I, const float *G, float *Wt, float lr,
                         const Who &w) {
  const int oplane = L.ow * L.oh, iplane = L.W * L.H, fplane = L.fw * L.fh;
  for (int k = w.warp; k < L.nw; k += w.nwarps) {
    const int f = k / L.K1, off = k - f * L.K1;
    const bool is_bias = off == L.K1 - 1;
    const int wz = off / fplane, wrem = off - wz * fplane;
    const int wy = wrem / L.fw, wx = wrem - wy * L.fw;
    float acc = 0.0f;
    for (int pix = w.lane; pix < oplane; pix += 32) {
      const float gv = G[(int64_t)f * oplane + pix];
      float in = 1.0f;
      if (!is_bias) {
        const int oy = pix / L.ow, ox = pix - oy * L.ow;
        const int iy = oy * L.s - L.p + wy, ix = ox * L.s - L.p + wx;
        in = (iy >= 0 && iy < L.H && ix >= 0 && ix < L.W) ? I[(int64_t)wz * iplane + iy * L.W + ix]
                                                         : 0.0f;
      }
      acc = fmaf(gv, in, acc);
    }
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (w.lane == 0) Wt[k] = Wt[k] - lr * acc;
  }
}

// one thread per input: every tied maximum receives the gradient (pool_bwd_kernel)
__device__ void bwd_pool(const CoopLayer &L, const float *I, const float *G, float *dI,
                         const Who &w) {
  const int gw = L.W / L.ow, gh = L.H / L.oh, iplane = L.W * L.H;
  for (int t = w.tid; t < iplane * L.D; t += w.nthreads) {
    const int z = t / iplane, rem = t - z * iplane;
    const int r = rem / L.W, c = rem - r * L.W;
    const int orow = r / gh, ocol = c / gw;
    float out = 0.0f;
    if (orow < L.oh && ocol < L.ow) {
      const float self = I[t];
      const float *win = I + (int64_t)z * iplane + (orow * gh) * L.W + ocol * gw;
      bool beaten = false;
      for (int gr = 0; gr < gh; ++gr)
        for (int gc = 0; gc < gw; ++gc) beaten |= (win[gr * L.W + gc] > self);
      if (!beaten) out = G[(int64_t)z * L.ow * L.oh + orow * L.ow + ocol];
    }
    dI[t] = out;
  }
}

// CTA 0 only: dI[k] = sum_i (O_i * (delta_ik - O_k)) * G_i, ascending i (softmax_layer.cc:42-59)
__device__ void bwd_softmax(const CoopLayer &L, const float *O, const float *G, float *dI) {
  if (blockIdx.x != 0) return;
  const int n = L.nin;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const float ok = O[k];
    float acc = 0.0f;
    for (int i = 0; i < n; ++i) acc += (O[i] * (((i == k) ? 1.0f : 0.0f) - ok)) * G[i];
    dI[k] = acc;
  }
}

__global__ void __launch_bounds__(kCoopThreads)
train_coop_kernel(const __grid_constant__ CoopNet net, float *weights, const float *lr_p,
                  const float *__restrict__ ins, const float *__restrict__ expected,
                  long long n_samples) {
  __shared__ float ps[kCoopMaxSoftmax];
  cg::grid_group grid = cg::this_grid();
  Who w;
  w.tid = blockIdx.x * blockDim.x + threadIdx.x;
  w.nthreads = gridDim.x * blockDim.x;
  w.warp = w.tid >> 5;
  w.nwarps = w.nthreads >> 5;
  w.lane = threadIdx.x & 31;
  const float lr = lr_p[0];
  const int n = net.n_layers;

  for (long long s = 0; s < n_samples; ++s) {
    const float *x = ins + s * net.n_in;
    const float *e = expected + s * net.n_out;
    // ---- forward ---------------------------------------------------------------
    for (int l = 0; l < n; ++l) {
      const CoopLayer &L = net.L[l];
      const float *I = l ? net.out[l - 1] : x;
      float *O = net.out[l];
      const float *Wt = weights + L.w_off;
      switch (L.kind) {
        case PB_ACTIVATION: fwd_activation(L, I, O, w); break;
        case PB_DENSE: fwd_dense(L, I, O, Wt, w); break;
        case PB_CONVOLUTION: fwd_conv(L, I, O, Wt, w); break;
        case PB_MAX_POOL: fwd_pool(L, I, O, w); break;
        default: fwd_softmax(L, I, O, ps); break;
      }
      grid.sync();
    }
    // ---- dE/dO (error_layer.cc:81-91) ----------------------------------------------
    float *g = net.g[0], *ng = net.g[1];
    {
      const float *O = net.out[n - 1];
      for (int k = w.tid; k < net.n_out; k += w.nthreads)
        g[k] = net.loss == PB_MEAN_SQUARED ? O[k] - e[k]
                                           : -1.0f * (e[k] * (1.0f / (O[k] + kRefEps)));
      grid.sync();
    }
    // ---- reverse loop: input_delta with the pre-update weights, then the update ----------
    for (int l = n - 1; l >= 0; --l) {
      const CoopLayer &L = net.L[l];
      const float *I = l ? net.out[l - 1] : x;
      float *Wt = weights + L.w_off;
      switch (L.kind) {
        case PB_ACTIVATION: bwd_activation(L, I, g, ng, w); break;
        case PB_DENSE: bwd_dense(L, g, ng, Wt, w); break;
        case PB_CONVOLUTION: bwd_conv(L, g, ng, Wt, w); break;
        case PB_MAX_POOL: bwd_pool(L, I, g, ng, w); break;
        default: bwd_softmax(L, net.out[l], g, ng); break;
      }
      grid.sync();
      if (L.nw > 0) {
        if (L.kind == PB_DENSE) upd_dense(L, I, g, Wt, lr, w);
        else upd_conv(L, I, g, Wt, lr, w);
        grid.sync();
      }
      float *t = g; g = ng; ng = t;
    }
  }
}

}  // namespace

// Returns -1 when the path is not selected or the network is outside its envelope.
// outputs[l]: per-layer output buffers (>= the layer's output count); grad_a / grad_b: two
// buffers of max(layer width) floats; w_off: offsets of the layers' weight blocks in `weights`.
int train_loop_coop(pb_context *ctx, const pb_layer_desc *layers, int n_layers,
                    const int64_t *w_off, int loss, float *weights, const float *lr,
                    float *const *outputs, float *grad_a, float *grad_b, const float *d_ins,
                    const float *d_expected, int64_t n_samples) {
  const char *env = getenv("PB_TRAIN_COOP");
  if (!(env && env[0] == '1')) return -1;
  if (n_layers < 1 || n_layers > kCoopMaxLayers) return -1;
  CoopNet net{};
  net.n_layers = n_layers;
  net.loss = loss;
  for (int l = 0; l < n_layers; ++l) {
    const pb_layer_desc &d = layers[l];
    CoopLayer &L = net.L[l];
    int64_t a = 0, b = 0, c = 0;
    if (pb_layer_num_inputs(&d, &a) || pb_layer_num_outputs(&d, &b) ||
        pb_layer_num_weights(&d, &c))
      return -1;
    if (a >= (1 << 30) || b >= (1 << 30) || c >= (1 << 30) || w_off[l] >= (1ll << 31)) return -1;
    L.kind = d.kind;
    L.act = d.activation;
    L.nin = (int)a;
    L.nout = (int)b;
    L.nw = (int)c;
    L.w_off = (int)w_off[l];
    switch (d.kind) {
      case PB_ACTIVATION:
      case PB_DENSE:
        break;
      case PB_SOFTMAX:
        if (a > kCoopMaxSoftmax) return -1;
        break;
      case PB_CONVOLUTION:
        L.W = (int)d.width; L.H = (int)d.height; L.D = (int)d.depth;
        L.fw = (int)d.filter_width; L.fh = (int)d.filter_height;
        L.s = (int)d.stride; L.p = (int)d.padding; L.F = (int)d.num_filters;
        L.ow = (int)conv_out_w(&d); L.oh = (int)conv_out_h(&d);
        L.K1 = L.fw * L.fh * L.D + 1;
        break;
      case PB_MAX_POOL:
        L.W = (int)d.width; L.H = (int)d.height; L.D = (int)d.depth;
        L.ow = (int)d.out_width; L.oh = (int)d.out_height;
        break;
      default:
        return -1;
    }
    net.out[l] = outputs[l];
  }
  net.n_in = net.L[0].nin;
  net.n_out = net.L[n_layers - 1].nout;
  net.g[0] = grad_a;
  net.g[1] = grad_b;

  int coop = 0;
  PB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
  if (!coop) return -1;
  int per_sm = 0;
  PB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, train_coop_kernel, kCoopThreads,
                                                        0));
  if (per_sm < 1) return -1;
  // one CTA per SM: the barrier cost grows with the CTA count, and 148 x 256 threads cover
  // the widest per-sample phase of the networks this path is for
  const dim3 grid_dim((unsigned)ctx->num_sms), block_dim(kCoopThreads);
  long long n = (long long)n_samples;
  void *args[] = {(void *)&net, (void *)&weights, (void *)&lr, (void *)&d_ins,
                  (void *)&d_expected, (void *)&n};
  PB_CUDA(cudaLaunchCooperativeKernel((const void *)train_coop_kernel, grid_dim, block_dim, args, 0,
                                      ctx->stream));
  PB_LAUNCHED(ctx);
  return 0;
}

}  // namespace pb
