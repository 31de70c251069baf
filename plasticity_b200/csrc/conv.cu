// ConvolutionLayer kernels (nnet/convolution_layer.cc:36-270).
//
//   evaluate     O[b,f,r,c]  = bias_f + sum_{z,y,x} W[f,z,y,x] * I0[b,z,r*s-p+y,c*s-p+x]
//   input_delta  dI[b,z,r,c] = sum_{f,d,k} W[f,z,r-(d*s-p),c-(k*s-p)] * G'[b,f,d,k]
//   weight grad  g[f,z,y,x]  = sum_b sum_{oy,ox} G[b,f,oy,ox] * I0[b,z,oy*s-p+y,ox*s-p+x]
//                g[f,bias]   = sum_b sum_{oy,ox} G[b,f,oy,ox]
//
// I0 is the zero-padded input.  G' is G with the reference's centre guard
// applied: input_delta only accepts an output (d,k) whose convolution centre
// (d*s-p+fh/2, k*s-p+fw/2) lies inside the input image
// (convolution_layer.cc:209) — identical to plain G when pad == filter/2.
// Layouts: volumes planar CHW per sample, weights [f][z][y][x] + bias with
// block stride fw*fh*fd+1 (nnet/symbol_generator.h:229-249).
//
// This file holds the GENERAL kernels (any stride / padding / filter shape):
// one thread per output element, used for shapes the tiled kernels in
// conv_tiled.cu do not cover.
#include "common.cuh"
#include "kernels.cuh"

namespace pb {

struct ConvGeom {
  int W, H, D;        // input volume
  int fw, fh;         // filter extent (depth == D)
  int s, p, F;        // stride, padding, filters
  int ow, oh;         // output extent
  int K1;             // fw*fh*D + 1: filter block stride incl. bias
};

static ConvGeom make_geom(const pb_layer_desc *l) {
  ConvGeom g;
  g.W = (int)l->width; g.H = (int)l->height; g.D = (int)l->depth;
  g.fw = (int)l->filter_width; g.fh = (int)l->filter_height;
  g.s = (int)l->stride; g.p = (int)l->padding; g.F = (int)l->num_filters;
  g.ow = (int)conv_out_w(l); g.oh = (int)conv_out_h(l);
  g.K1 = g.fw * g.fh * g.D + 1;
  return g;
}

__global__ void conv_fwd_general_kernel(const float *__restrict__ I,
                                        const float *__restrict__ Wt,
                                        float *__restrict__ O, ConvGeom g, int64_t total) {
  const int oplane = g.ow * g.oh, iplane = g.W * g.H, fplane = g.fw * g.fh;
  const int64_t per_out = (int64_t)oplane * g.F, per_in = (int64_t)iplane * g.D;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / per_out;
    const int idx = (int)(t - b * per_out);
    const int f = idx / oplane, rem = idx - f * oplane;
    const int r = rem / g.ow, c = rem - r * g.ow;
    const int y0 = r * g.s - g.p, x0 = c * g.s - g.p;
    const float *w = Wt + (int64_t)f * g.K1;
    const float *in = I + b * per_in;
    float acc = w[g.K1 - 1];
    for (int z = 0; z < g.D; ++z)
      for (int y = 0; y < g.fh; ++y) {
        const int iy = y0 + y;
        if (iy < 0 || iy >= g.H) continue;
        for (int x = 0; x < g.fw; ++x) {
          const int ix = x0 + x;
          if (ix < 0 || ix >= g.W) continue;
          acc = fmaf(w[fplane * z + y * g.fw + x], in[iplane * z + iy * g.W + ix], acc);
        }
      }
    O[t] = acc;
  }
}

__global__ void conv_bwd_data_general_kernel(const float *__restrict__ Wt,
                                             const float *__restrict__ G,
                                             float *__restrict__ dI, ConvGeom g,
                                             int64_t total) {
  const int oplane = g.ow * g.oh, iplane = g.W * g.H, fplane = g.fw * g.fh;
  const int64_t per_out = (int64_t)oplane * g.F, per_in = (int64_t)iplane * g.D;
  const int hh = g.fh / 2, hw = g.fw / 2;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / per_in;
    const int idx = (int)(t - b * per_in);
    const int z = idx / iplane, rem = idx - z * iplane;
    const int r = rem / g.W, c = rem - r * g.W;
    // Output rows d whose filter covers input row r: 0 <= r+p-d*s < fh.
    int d_lo = (r + g.p - g.fh + 1);
    d_lo = d_lo <= 0 ? 0 : (d_lo + g.s - 1) / g.s;
    int d_hi = (r + g.p) / g.s;
    if (d_hi > g.oh - 1) d_hi = g.oh - 1;
    int k_lo = (c + g.p - g.fw + 1);
    k_lo = k_lo <= 0 ? 0 : (k_lo + g.s - 1) / g.s;
    int k_hi = (c + g.p) / g.s;
    if (k_hi > g.ow - 1) k_hi = g.ow - 1;
    const float *gb = G + b * per_out;
    float acc = 0.0f;
    for (int d = d_lo; d <= d_hi; ++d) {
      const int cy = d * g.s - g.p + hh;          // centre guard (reference quirk)
      if (cy < 0 || cy >= g.H) continue;
      const int srow = r + g.p - d * g.s;
      for (int k = k_lo; k <= k_hi; ++k) {
        const int cx = k * g.s - g.p + hw;
        if (cx < 0 || cx >= g.W) continue;
        const int scol = c + g.p - k * g.s;
        const float *w = Wt + fplane * z + srow * g.fw + scol;
        const float *gp = gb + d * g.ow + k;
        for (int f = 0; f < g.F; ++f)
          acc = fmaf(w[(int64_t)f * g.K1], gp[(int64_t)f * oplane], acc);
      }
    }
    dI[t] = acc;
  }
}

// One CTA per weight: threads stride over (sample, output pixel), block
// reduction, thread 0 applies the SGD update.  Deterministic.
__global__ void __launch_bounds__(256)
conv_wgrad_general_kernel(const float *__restrict__ I, const float *Wt,
                          const float *__restrict__ G, float *out,
                          const float *__restrict__ lr, ConvGeom g, int64_t batch,
                          int mode) {
  const int oplane = g.ow * g.oh, iplane = g.W * g.H, fplane = g.fw * g.fh;
  const int64_t per_out = (int64_t)oplane * g.F, per_in = (int64_t)iplane * g.D;
  const int w = blockIdx.x;
  const int f = w / g.K1, off = w - f * g.K1;
  const bool is_bias = off == g.K1 - 1;
  const int wz = off / fplane, wrem = off - wz * fplane;
  const int wy = wrem / g.fw, wx = wrem - wy * g.fw;
  float acc = 0.0f;
  const int64_t total = batch * oplane;
  for (int64_t t = threadIdx.x; t < total; t += blockDim.x) {
    const int64_t b = t / oplane;
    const int pix = (int)(t - b * oplane);
    const float gv = G[b * per_out + (int64_t)f * oplane + pix];
    float in = 1.0f;
    if (!is_bias) {
      const int oy = pix / g.ow, ox = pix - oy * g.ow;
      const int iy = oy * g.s - g.p + wy, ix = ox * g.s - g.p + wx;
      in = (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W)
               ? I[b * per_in + (int64_t)wz * iplane + iy * g.W + ix]
               : 0.0f;
    }
    acc = fmaf(gv, in, acc);
  }
  __shared__ float red[8];
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float sum = 0.0f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) sum += red[i];
    apply_update(mode, out, Wt, w, lr[0], sum);
  }
}

constexpr int64_t kTiledMinBatch = 32;

int conv_check(const pb_layer_desc *l) {
  PB_CHECK(l->filter_depth == l->depth,
           "Convolution layer input depth != filter depth");  // convolution_layer.cc:12-16
  PB_CHECK(l->stride >= 1 && l->num_filters >= 1 && l->filter_width >= 1 &&
               l->filter_height >= 1,
           "Convolution layer: bad filter parameters");
  PB_CHECK(l->width + 2 * l->padding >= l->filter_width &&
               l->height + 2 * l->padding >= l->filter_height,
           "Convolution layer: filter larger than padded input");
  return 0;
}

int conv_evaluate_general(pb_context *ctx, const pb_layer_desc *l, const float *I,
                          const float *W, float *O, int64_t batch) {
  const ConvGeom g = make_geom(l);
  const int64_t total = batch * g.ow * g.oh * g.F;
  if (total == 0) return 0;
  conv_fwd_general_kernel<<<pb_ew_grid(ctx, total, 128), 128, 0, ctx->stream>>>(I, W, O, g,
                                                                               total);
  PB_LAUNCHED(ctx);
  return 0;
}

int conv_input_delta_general(pb_context *ctx, const pb_layer_desc *l, const float *W,
                             const float *G, float *dI, int64_t batch) {
  const ConvGeom g = make_geom(l);
  const int64_t total = batch * g.W * g.H * g.D;
  if (total == 0) return 0;
  conv_bwd_data_general_kernel<<<pb_ew_grid(ctx, total, 128), 128, 0, ctx->stream>>>(
      W, G, dI, g, total);
  PB_LAUNCHED(ctx);
  return 0;
}

int conv_weight_update_general(pb_context *ctx, const pb_layer_desc *l, const float *I,
                               const float *W, const float *G, float *out, const float *lr,
                               int mode, int64_t batch) {
  const ConvGeom g = make_geom(l);
  const int nw = g.K1 * g.F;
  if (nw == 0) return 0;
  conv_wgrad_general_kernel<<<nw, 256, 0, ctx->stream>>>(I, W, G, out, lr, g, batch, mode);
  PB_LAUNCHED(ctx);
  return 0;
}

int conv_evaluate(pb_context *ctx, const pb_layer_desc *l, const float *I, const float *W,
                  float *O, int64_t batch) {
  if (conv_check(l)) return 1;
  // Small batches (per-sample SGD) are latency-bound: the one-thread-per-
  // output kernel spreads a single sample over many more CTAs.
  if (batch >= kTiledMinBatch) {
    int rc = conv_evaluate_umma(ctx, l, I, W, O, batch);
    if (rc >= 0) return rc;
    rc = conv_evaluate_gemm_umma(ctx, l, I, W, O, nullptr, PB_IDENTITY, batch);
    if (rc >= 0) return rc;
    rc = conv_evaluate_tiled(ctx, l, I, W, O, batch);
    if (rc >= 0) return rc;
  }
  return conv_evaluate_general(ctx, l, I, W, O, batch);
}

int conv_input_delta(pb_context *ctx, const pb_layer_desc *l, const float *W, const float *G,
                     float *dI, int64_t batch) {
  if (conv_check(l)) return 1;
  if (batch >= kTiledMinBatch) {
    int rc = conv_input_delta_umma(ctx, l, W, G, dI, batch);
    if (rc >= 0) return rc;
    rc = conv_input_delta_gemm_umma(ctx, l, W, G, dI, batch);
    if (rc >= 0) return rc;
    rc = conv_input_delta_tiled(ctx, l, W, G, dI, batch);
    if (rc >= 0) return rc;
  }
  return conv_input_delta_general(ctx, l, W, G, dI, batch);
}

int conv_weight_update(pb_context *ctx, const pb_layer_desc *l, const float *I,
                       const float *W, const float *G, float *out, const float *lr,
                       int mode, int64_t batch) {
  if (conv_check(l)) return 1;
  if (batch >= kTiledMinBatch) {
    int rc = conv_weight_update_umma(ctx, l, I, W, G, out, lr, mode, batch);
    if (rc >= 0) return rc;
    rc = conv_weight_update_mma(ctx, l, I, W, G, out, lr, mode, batch);
    if (rc >= 0) return rc;
    rc = conv_weight_update_gemm_umma(ctx, l, I, W, G, out, lr, mode, batch);
    if (rc >= 0) return rc;
    rc = conv_weight_update_tiled(ctx, l, I, W, G, out, lr, mode, batch);
    if (rc >= 0) return rc;
  }
  return conv_weight_update_general(ctx, l, I, W, G, out, lr, mode, batch);
}

}  // namespace pb
