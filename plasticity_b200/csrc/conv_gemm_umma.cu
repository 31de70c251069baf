// General 3xTF32 tcgen05 implicit-GEMM convolution, forward
// (nnet/convolution_layer.cc:36-96), for the large batched shapes of the YOLOv1
// network (nnet/yolov1.cc:7-231: 7x7/2, 3x3, 1x1, 3x3/2 filters, 3..1024
// channels) — any filter size, stride and padding:
//
//   Out[b,oc,r,c] = bias_oc + sum_{ic,fy,fx} Wk[oc,ic,fy,fx] * In0[b,ic,r*s-p+fy,c*s-p+fx]
//
// GEMM view on the skeleton of umma_gemm.cuh: M = (sample, output pixel), N = output
// channel, K = (tap, input channel) with the channel running fastest.  Activations
// stay in the reference's planar CHW layout: for a fixed k the 128 pixels of an M
// tile are (mostly) consecutive addresses, so the A producer is a gather with lanes
// along the pixels and a pointer that walks the channel planes — the im2col matrix
// never exists.  Zero padding is a predicate on the tap.  The filters are re-tiled
// once per call into [oc][tap][ic] (K-major, rows padded to the K stage) by a small
// pre-pass, so the B producer is the plain aligned float4 path.  The epilogue adds the
// bias and writes [b][oc][pixel] — lanes along the pixels, coalesced — and can apply
// the following ActivationLayer in the same pass (second output).
//
// Input gradient (nnet/convolution_layer.cc:100-224) on the same policy, `bwd` set:
//   dI[b,c,y,x] = sum_{oc,fy,fx} Wk[oc,c,fy,fx] * G[b,oc,(y+p-fy)/s,(x+p-fx)/s]
// (terms whose quotient is not integral or out of range vanish).  M = (sample, INPUT pixel),
// N = input channel, K = (tap, output channel): the gather reads the gradient planes at the
// transposed coordinates and the filters are re-tiled as [c][tap][oc].  Envelope: odd filters
// with pad == filter/2 (every YOLOv1 layer), where the reference's centre guard (:209) is the
// range check above; output-channel counts that are multiples of 16.
#include "umma_gemm.cuh"

namespace pb {
namespace {

using namespace ugemm;

constexpr int kMaxTaps = 128;

struct ConvArgs : GemmCore {
  const float *In;    // [batch][C][H][W]
  const float *Wk;    // re-tiled filters [OC][Kp], k' = tap*ICP + ic
  const float *bias;  // [OC], gathered from the reference layout by the pre-pass
  float *Out;         // [batch][OC][OH][OW]
  float *Out2;        // act(Out), nullable
  int act;
  int C, H, W, OC, OH, OW, FH, FW, stride, pad;
  int ICP;            // channels per tap in K order: 4, 8 or a multiple of 16
  int taps, Kp;
  // K order.  0: k' = tap*ICP + ic (tap_yx[tap] = fy << 8 | fx).  1 ("rows", few input
  // channels and a wide filter — the RGB 7x7 layer): k' = (fy*C + ic)*8 + fx, one filter
  // ROW of one channel per 8 k values (tap_yx[fy*C + ic] = fy << 8 | ic): a thread reads 8
  // consecutive input floats per row instead of gathering taps one by one.
  int korder;
  // bwd: the gathered tensor is the gradient (C = its channels = the layer's filters, H x W =
  // the layer's OUTPUT extent), OH x OW = the layer's input extent, OC = its input channels
  int bwd;
  uint16_t tap_yx[kMaxTaps];
};

// Input gradient: Wk[c][tap*ICP + oc] = Wref[oc][c][tap]; bias[] = 0.
__global__ void retile_filters_bwd_kernel(const float *__restrict__ Wref, float *__restrict__ Wk,
                                          float *__restrict__ bias, int OC, int C, int taps,
                                          int ICP, int Kp) {
  const int64_t total = (int64_t)C * Kp;
  const int64_t kref1 = (int64_t)C * taps + 1;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < C;
       t += (int64_t)gridDim.x * blockDim.x)
    bias[t] = 0.0f;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = t / Kp;
    const int k = (int)(t - c * Kp);
    const int tap = k / ICP, oc = k - tap * ICP;
    Wk[t] = (tap < taps && oc < OC) ? Wref[oc * kref1 + c * taps + tap] : 0.0f;
  }
}

// Wk[oc][tap*ICP + ic] = Wref[oc][ic][tap], zero in the padding.
__global__ void retile_filters_kernel(const float *__restrict__ Wref, float *__restrict__ Wk,
                                      float *__restrict__ bias, int OC, int C, int taps, int ICP,
                                      int Kp, int korder, int FW) {
  const int64_t total = (int64_t)OC * Kp;
  const int64_t kref1 = (int64_t)C * taps + 1;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < OC;
       t += (int64_t)gridDim.x * blockDim.x)
    bias[t] = Wref[t * kref1 + kref1 - 1];
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t oc = t / Kp;
    const int k = (int)(t - oc * Kp);
    if (korder == 1) {
      const int row = k >> 3, fx = k & 7;     // row = fy*C + ic
      const int fy = row / C, ic = row - fy * C;
      Wk[t] = (fy * FW < taps && fx < FW) ? Wref[oc * kref1 + (int64_t)ic * taps + fy * FW + fx] : 0.0f;
      continue;
    }
    const int tap = k / ICP, ic = k - tap * ICP;
    Wk[t] = (tap < taps && ic < C) ? Wref[oc * kref1 + (int64_t)ic * taps + tap] : 0.0f;
  }
}

struct ConvPolicy {
  using Args = ConvArgs;
  static constexpr bool kEpiTranspose = false;
  static constexpr int kEpiCols = 16;

  // One output pixel per thread (lanes along M), 16 k values of the stage each.
  struct A {
    float v[16];
    int64_t base_off;  // element offset of (sample, channel 0, iy0, ix0); may point before the plane
    int iy0, ix0;
    bool m_ok;
    __device__ __forceinline__ void init(const Args &) {}
    __device__ __forceinline__ void tile(const Args &g, int tm, int p) {
      const int m = tm * BM + p;
      m_ok = m < g.M;
      const int mm = m_ok ? m : 0;
      const int ohw = g.OH * g.OW;
      const int b = mm / ohw, pix = mm - b * ohw;
      const int r = pix / g.OW, c = pix - r * g.OW;
      if (g.bwd) {
        iy0 = r + g.pad;   // gradient row = (iy0 - fy) / stride
        ix0 = c + g.pad;
        base_off = (int64_t)b * g.C * g.H * g.W;
        return;
      }
      iy0 = r * g.stride - g.pad;
      ix0 = c * g.stride - g.pad;
      base_off = (int64_t)b * g.C * g.H * g.W + (int64_t)iy0 * g.W + ix0;
    }
    __device__ __forceinline__ void load(const Args &g, int ks, int) {
      const int64_t hw = (int64_t)g.H * g.W;
      if (g.bwd) {
        // one tap per stage, 16 consecutive gradient channels (ICP is a multiple of 16)
        const int per = g.ICP >> 4;
        const int tap = ks / per, icb = ks - tap * per;
        const int yx = g.tap_yx[min(tap, kMaxTaps - 1)];
        const int dy = iy0 - (yx >> 8), dx = ix0 - (yx & 255);
        const int qy = dy / g.stride, qx = dx / g.stride;
        const bool ok = m_ok && tap < g.taps && dy >= 0 && dx >= 0 && qy * g.stride == dy &&
                        qx * g.stride == dx && qy < g.H && qx < g.W;
        const float *q = g.In + base_off + (int64_t)(icb * 16) * hw + (int64_t)qy * g.W + qx;
        if (ok) {
#pragma unroll
          for (int kk = 0; kk < 16; ++kk, q += hw) v[kk] = __ldg(q);
        } else {
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) v[kk] = 0.0f;
        }
        return;
      }
      if (g.korder == 1) {
        // two filter rows (fy, ic) per stage, 8 consecutive input columns each
        const int n_rows = g.FH * g.C;
        const bool cols_in = ix0 >= 0 && ix0 + 7 < g.W;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int row = 2 * ks + h;
          const int yc = g.tap_yx[min(row, kMaxTaps - 1)];
          const int fy = yc >> 8, ic = yc & 255;
          const bool ok = m_ok && row < n_rows && (unsigned)(iy0 + fy) < (unsigned)g.H;
          const float *q = g.In + base_off + (int64_t)ic * hw + fy * g.W;
          if (ok && cols_in) {
#pragma unroll
            for (int fx = 0; fx < 8; ++fx) v[8 * h + fx] = fx < g.FW ? __ldg(q + fx) : 0.0f;
          } else {
#pragma unroll
            for (int fx = 0; fx < 8; ++fx)
              v[8 * h + fx] = (ok && fx < g.FW && (unsigned)(ix0 + fx) < (unsigned)g.W) ? __ldg(q + fx) : 0.0f;
          }
        }
      } else if (g.ICP >= 16) {
        // one tap per stage, 16 consecutive channels
        const int per = g.ICP >> 4;
        const int tap = ks / per, icb = ks - tap * per;
        const int yx = g.tap_yx[min(tap, kMaxTaps - 1)];
        const int fy = yx >> 8, fx = yx & 255;
        const bool ok = m_ok && tap < g.taps && (unsigned)(iy0 + fy) < (unsigned)g.H &&
                        (unsigned)(ix0 + fx) < (unsigned)g.W;
        const float *q = g.In + base_off + (int64_t)(icb * 16) * hw + fy * g.W + fx;
        if (ok) {
#pragma unroll
          for (int kk = 0; kk < 16; ++kk, q += hw) v[kk] = __ldg(q);
        } else {
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) v[kk] = 0.0f;
        }
      } else {
        // few input channels (the RGB layer): 16 / ICP taps per stage
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int kq = ks * 16 + 4 * c;
          const int tap = g.ICP == 4 ? kq >> 2 : kq >> 3, ic0 = kq & (g.ICP - 1);
          const int yx = g.tap_yx[min(tap, kMaxTaps - 1)];
          const int fy = yx >> 8, fx = yx & 255;
          const bool ok = m_ok && tap < g.taps && (unsigned)(iy0 + fy) < (unsigned)g.H &&
                          (unsigned)(ix0 + fx) < (unsigned)g.W;
          const float *q = g.In + base_off + (int64_t)ic0 * hw + fy * g.W + fx;
#pragma unroll
          for (int i = 0; i < 4; ++i) v[4 * c + i] = (ok && ic0 + i < g.C) ? __ldg(q + i * hw) : 0.0f;
        }
      }
    }
    __device__ __forceinline__ void store(float4 *hi, int p) const { store_row16<LA>(v, hi, p); }
  };

  static constexpr int kMinBN = 64;
  template <int BN_T>
  struct B : MatrixOperand<OP_KV, BN_T, BN_T + 2> {
    __device__ __forceinline__ void init(const Args &g) {
      this->setup(g.Wk, g.Kp, g.N, g.Kp, g.mma_n);
    }
    __device__ __forceinline__ void tile(const Args &, int t, int) { this->set_tile(t); }
    __device__ __forceinline__ void load(const Args &, int ks, int p) { this->load_stage(ks, p); }
  };

  static __device__ __forceinline__ void store1(const Args &g, int64_t m, int n, float acc) {
    const int ohw = g.OH * g.OW;
    const int64_t b = m / ohw;
    const int pix = (int)(m - b * ohw);
    const float val = acc + __ldg(g.bias + n);
    const int64_t idx = (b * g.OC + n) * ohw + pix;
    g.Out[idx] = val;
    if (g.Out2) {
      float a = val;
      if (g.act == PB_RELU) a = act_fwd<PB_RELU>(val);
      else if (g.act == PB_LEAKY_RELU) a = act_fwd<PB_LEAKY_RELU>(val);
      else if (g.act == PB_SIGMOID) a = act_fwd<PB_SIGMOID>(val);
      g.Out2[idx] = a;
    }
  }

  // 16 output channels of one pixel: lanes are consecutive pixels, so every store
  // instruction of the warp writes one 128-byte line of one channel plane.
  template <int ACT>
  static __device__ __forceinline__ void store_act(float *o2, int ohw, const float (&val)[16],
                                                   int nv) {
#pragma unroll
    for (int j = 0; j < 16; ++j, o2 += ohw)
      if (j < nv) *o2 = act_fwd<ACT>(val[j]);
  }

  static __device__ __forceinline__ void store_cols(const Args &g, int row0, int lane, int n0,
                                                    const float (&v)[16], int nv, float *) {
    const int m = row0 + lane;
    if (m >= g.M) return;
    const int ohw = g.OH * g.OW;
    const int b = m / ohw, pix = m - b * ohw;
    const int64_t idx = ((int64_t)b * g.OC + n0) * ohw + pix;
    float val[16];
    if (nv == 16) {
      const float4 *bp = reinterpret_cast<const float4 *>(g.bias + n0);  // n0 % 16 == 0
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 bq = __ldg(bp + q);
        val[4 * q] = v[4 * q] + bq.x; val[4 * q + 1] = v[4 * q + 1] + bq.y;
        val[4 * q + 2] = v[4 * q + 2] + bq.z; val[4 * q + 3] = v[4 * q + 3] + bq.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) val[j] = j < nv ? v[j] + __ldg(g.bias + n0 + j) : 0.0f;
    }
    float *o = g.Out + idx;
#pragma unroll
    for (int j = 0; j < 16; ++j, o += ohw)
      if (j < nv) *o = val[j];
    if (g.Out2) {
      float *o2 = g.Out2 + idx;
      switch (g.act) {
        case PB_RELU: store_act<PB_RELU>(o2, ohw, val, nv); break;
        case PB_LEAKY_RELU: store_act<PB_LEAKY_RELU>(o2, ohw, val, nv); break;
        case PB_SIGMOID: store_act<PB_SIGMOID>(o2, ohw, val, nv); break;
        default: store_act<PB_IDENTITY>(o2, ohw, val, nv); break;
      }
    }
  }
};

}  // namespace

// Returns -1 outside the envelope (then the FFMA kernels run).  O2 != nullptr:
// additionally writes act(O) (the following ActivationLayer's output).
int conv_evaluate_gemm_umma(pb_context *ctx, const pb_layer_desc *l, const float *I,
                            const float *W, float *O, float *O2, int act, int64_t batch) {
  if (!umma_gemm_enabled()) return -1;
  const int64_t C = l->depth, OC = l->num_filters;
  const int64_t OW = conv_out_w(l), OH = conv_out_h(l);
  const int64_t taps = l->filter_width * l->filter_height;
  if (C > 8 && C % 16 != 0) return -1;
  const int64_t ICP = C <= 4 ? 4 : (C <= 8 ? 8 : C);
  // "rows" K order when it is the shorter K: few channels under a wide (<= 8) filter
  const int korder = (l->filter_width <= 8 && l->filter_height * C <= kMaxTaps &&
                      l->filter_height * C * 8 < taps * ICP) ? 1 : 0;
  const int64_t K = korder == 1 ? l->filter_height * C * 8 : taps * ICP;
  const int64_t M = batch * OH * OW;
  if (M <= 0 || OC <= 0) return 0;
  // worth a tensor-core launch?  (small products are latency-bound either way)
  if (M * OC * K < ((int64_t)1 << 26) || M >= ((int64_t)1 << 31) - BM || OC * (K + BK) >= ((int64_t)1 << 31))
    return -1;
  const int Kp = pb_div_up(K, BK) * BK;
  if (taps > kMaxTaps || l->filter_width > 255 || l->filter_height > 255) return -1;
  const size_t bias_off = ((size_t)OC * Kp + 3) & ~(size_t)3;  // keeps the bias 16-byte aligned
  if (ensure_aux(ctx, bias_off + (size_t)OC + 16)) return 1;
  retile_filters_kernel<<<pb_ew_grid(ctx, OC * Kp, 256), 256, 0, ctx->stream>>>(
      W, ctx->aux, ctx->aux + bias_off, (int)OC, (int)C, (int)taps, (int)ICP, Kp, korder,
      (int)l->filter_width);
  PB_LAUNCHED(ctx);
  ConvArgs g{};
  g.In = I; g.Wk = ctx->aux; g.bias = ctx->aux + bias_off; g.Out = O; g.Out2 = O2; g.act = act;
  g.korder = korder;
  if (korder == 1) {
    for (int r = 0; r < (int)(l->filter_height * C); ++r)
      g.tap_yx[r] = (uint16_t)(((r / C) << 8) | (r % C));
  } else {
    for (int t = 0; t < (int)taps; ++t)
      g.tap_yx[t] = (uint16_t)(((t / l->filter_width) << 8) | (t % l->filter_width));
  }
  g.C = (int)C; g.H = (int)l->height; g.W = (int)l->width; g.OC = (int)OC;
  g.OH = (int)OH; g.OW = (int)OW; g.FH = (int)l->filter_height; g.FW = (int)l->filter_width;
  g.stride = (int)l->stride; g.pad = (int)l->padding;
  g.ICP = (int)ICP; g.taps = (int)taps; g.Kp = Kp;
  return launch_gemm<ConvPolicy>(ctx, g, (int)M, (int)OC, Kp);
}

// Input gradient on the same kernel; -1 outside the envelope.
int conv_input_delta_gemm_umma(pb_context *ctx, const pb_layer_desc *l, const float *W,
                               const float *G, float *dI, int64_t batch) {
  if (!umma_gemm_enabled()) return -1;
  const int64_t C = l->depth, OC = l->num_filters;
  const int64_t OW = conv_out_w(l), OH = conv_out_h(l);
  const int64_t taps = l->filter_width * l->filter_height;
  if (OC % 16 != 0 || !(l->filter_width & 1) || !(l->filter_height & 1) ||
      l->padding != l->filter_width / 2 || l->padding != l->filter_height / 2)
    return -1;
  if (taps > kMaxTaps || l->filter_width > 255 || l->filter_height > 255) return -1;
  const int64_t K = taps * OC, M = batch * l->height * l->width;
  if (M <= 0 || C <= 0) return 0;
  if (M * C * K < ((int64_t)1 << 26) || M >= ((int64_t)1 << 31) - BM || C * (K + BK) >= ((int64_t)1 << 31))
    return -1;
  const int Kp = pb_div_up(K, BK) * BK;
  const size_t bias_off = ((size_t)C * Kp + 3) & ~(size_t)3;
  if (ensure_aux(ctx, bias_off + (size_t)C + 16)) return 1;
  retile_filters_bwd_kernel<<<pb_ew_grid(ctx, C * Kp, 256), 256, 0, ctx->stream>>>(
      W, ctx->aux, ctx->aux + bias_off, (int)OC, (int)C, (int)taps, (int)OC, Kp);
  PB_LAUNCHED(ctx);
  ConvArgs g{};
  g.In = G; g.Wk = ctx->aux; g.bias = ctx->aux + bias_off; g.Out = dI; g.Out2 = nullptr;
  g.act = PB_IDENTITY; g.korder = 0; g.bwd = 1;
  for (int t = 0; t < (int)taps; ++t)
    g.tap_yx[t] = (uint16_t)(((t / l->filter_width) << 8) | (t % l->filter_width));
  g.C = (int)OC; g.H = (int)OH; g.W = (int)OW;            // the gathered tensor: the gradient
  g.OC = (int)C; g.OH = (int)l->height; g.OW = (int)l->width;
  g.FH = (int)l->filter_height; g.FW = (int)l->filter_width;
  g.stride = (int)l->stride; g.pad = (int)l->padding;
  g.ICP = (int)OC; g.taps = (int)taps; g.Kp = Kp;
  return launch_gemm<ConvPolicy>(ctx, g, (int)M, (int)C, Kp);
}

}  // namespace pb
