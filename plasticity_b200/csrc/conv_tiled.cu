// Register-tiled FP32 FFMA convolution kernels for stride-1 layers whose
// output plane fits one CTA (the CIFAR network of nnet/cifar_test.cc:293-350:
// 32x32, 16x16 and 8x8 planes, 5x5 filters).  Shapes outside that envelope
// fall back to the general kernels in conv.cu.
//
// Forward and input-gradient share one kernel: the input gradient of a
// stride-1 convolution is itself a stride-1 correlation of the (zero-padded)
// output gradient with the flipped, channel-transposed filters,
//   dI[z,r,c] = sum_f sum_{y',x'} W[f,z,fh-1-y',fw-1-x'] * G'[f, r-p'+y', c-p'+x'],
// with p' = f-1-p.  G' is G with the reference's centre guard applied on load
// (convolution_layer.cc:209; a no-op when pad == filter/2).
//
// Kernel shape (both): a CTA owns NS samples; each thread owns PIX consecutive
// output pixels of one row for ALL output channels (FT accumulators per
// pixel).  Per input channel and filter row a thread reads PIX+FW-1 inputs
// from shared memory as float4 and slides the filter across them in
// registers; weights are broadcast float4 reads ([ic][fy][fx][oc] in shared
// memory).  ~FT*PIX*FW FFMA per (2 + FW*FT/4) LDS.128: FFMA-issue bound.
//
// The weight-gradient kernel is an implicit GEMM with the reduction over
// (sample, output pixel): a CTA walks its share of the batch one sample at a
// time (input tile with zero halo + gradient tile staged in shared memory,
// double-buffered with cp.async); a thread owns one (channel, filter-row)
// pair x TF filters x FW taps = TF*FW accumulators and slides along the
// output row 4 pixels at a time.  Per-CTA partial sums go to a workspace; a
// second kernel reduces them in a fixed order (deterministic) and applies the
// SGD update.
#include <cuda_pipeline.h>

#include "common.cuh"
#include "kernels.cuh"

namespace pb {

// ---------------------------------------------------------------------------
// forward / backward-data
// ---------------------------------------------------------------------------
struct S1Geom {
  int IC, IH, IW;     // input channels / extent   (fwd: D,H,W     bwd: F,oh,ow)
  int OC, OH, OW;     // output channels / extent  (fwd: F,oh,ow   bwd: D,H,W)
  int padY, padX;     // (fwd: p,p                 bwd: fh-1-p, fw-1-p)
  int K1;             // reference filter-block stride fw*fh*D+1
  int NS, ZC;         // samples per CTA, input channels per shared-memory chunk
  int IHs, IWs;       // shared-memory tile extent (halo included, IWs % 4 == 0)
  int XG, TPS;        // pixel groups per row, threads per sample
  int guardLoY, guardHiY, guardLoX, guardHiX;  // bwd: input rows/cols that pass the centre guard
  int batch;
};

template <int FW, int FH, int FT, int PIX, bool BWD>
__global__ void __launch_bounds__(256, 2)
conv_s1_kernel(const float *__restrict__ In, const float *__restrict__ Wt,
               float *__restrict__ Out, const S1Geom g) {
  extern __shared__ __align__(16) float smem[];
  constexpr int NLOAD = (PIX + FW - 1 + 3) / 4;
  float *w_s = smem;                                   // [ZC][FH][FW][FT]
  float *in_s = smem + g.ZC * FH * FW * FT;            // [NS][ZC][IHs][IWs]
  const int tid = threadIdx.x;
  const int s = tid / g.TPS, rem = tid - s * g.TPS;
  const int r = rem / g.XG, x0 = (rem - r * g.XG) * PIX;
  const int64_t b0 = (int64_t)blockIdx.x * g.NS;
  const bool active = s < g.NS && (b0 + s) < g.batch;
  const int in_plane = g.IH * g.IW;
  const int tile_plane = g.IHs * g.IWs;

  float acc[PIX][FT];
#pragma unroll
  for (int f = 0; f < FT; ++f) {
    float bias = 0.0f;
    if (!BWD && f < g.OC) bias = Wt[(int64_t)f * g.K1 + g.K1 - 1];
#pragma unroll
    for (int p = 0; p < PIX; ++p) acc[p][f] = bias;
  }

  for (int c0 = 0; c0 < g.IC; c0 += g.ZC) {
    const int zc_n = min(g.ZC, g.IC - c0);
    __syncthreads();
    // weights of this channel chunk, transposed to [ic][fy][fx][oc].  Threads
    // walk the reference layout in memory order (taps of one (oc, ic) pair are
    // contiguous) so the global reads coalesce.
    {
      constexpr int TAPS = FH * FW;
      const int per_oc = zc_n * TAPS;
      for (int e = tid; e < per_oc * FT; e += blockDim.x) {
        const int f = e / per_oc, t = e - f * per_oc;
        const int zc = t / TAPS, tap = t - zc * TAPS;
        float v = 0.0f;
        if (f < g.OC) {
          if (!BWD) v = Wt[(int64_t)f * g.K1 + (c0 + zc) * TAPS + tap];
          else v = Wt[(int64_t)(c0 + zc) * g.K1 + f * TAPS + (TAPS - 1 - tap)];
        }
        w_s[(zc * TAPS + tap) * FT + f] = v;
      }
    }
    // input tiles with zero halo: one warp per tile row, lanes along x
    {
      const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
      const int rows = g.NS * zc_n * g.IHs;
      for (int row = warp; row < rows; row += nwarps) {
        const int y = row % g.IHs;
        const int t = row / g.IHs;
        const int zc = t % zc_n, ss = t / zc_n;
        const int iy = y - g.padY;
        const int64_t b = b0 + ss;
        bool row_ok = b < g.batch && iy >= 0 && iy < g.IH;
        if (BWD) row_ok = row_ok && iy >= g.guardLoY && iy <= g.guardHiY;
        const float *src = In + ((b * g.IC + c0 + zc) * in_plane + iy * g.IW);
        float *dst = in_s + (ss * g.ZC + zc) * tile_plane + y * g.IWs;
        int lo = 0, hi = g.IW - 1;
        if (BWD) { lo = g.guardLoX; hi = g.guardHiX; }
        for (int x = lane; x < g.IWs; x += 32) {
          const int ix = x - g.padX;
          dst[x] = (row_ok && ix >= lo && ix <= hi) ? src[ix] : 0.0f;
        }
      }
    }
    __syncthreads();
    if (active) {
      for (int zc = 0; zc < zc_n; ++zc) {
        const float *tile = in_s + (s * g.ZC + zc) * tile_plane + r * g.IWs + x0;
        const float *wz = w_s + zc * FH * FW * FT;
#pragma unroll 1
        for (int fy = 0; fy < FH; ++fy) {
          float in[NLOAD * 4];
#pragma unroll
          for (int q = 0; q < NLOAD; ++q) {
            const float4 v = *reinterpret_cast<const float4 *>(tile + fy * g.IWs + q * 4);
            in[q * 4 + 0] = v.x; in[q * 4 + 1] = v.y; in[q * 4 + 2] = v.z; in[q * 4 + 3] = v.w;
          }
#pragma unroll
          for (int fx = 0; fx < FW; ++fx) {
#pragma unroll
            for (int f4 = 0; f4 < FT / 4; ++f4) {
              const float4 w = *reinterpret_cast<const float4 *>(wz + (fy * FW + fx) * FT + f4 * 4);
#pragma unroll
              for (int p = 0; p < PIX; ++p) {
                acc[p][f4 * 4 + 0] = fmaf(in[p + fx], w.x, acc[p][f4 * 4 + 0]);
                acc[p][f4 * 4 + 1] = fmaf(in[p + fx], w.y, acc[p][f4 * 4 + 1]);
                acc[p][f4 * 4 + 2] = fmaf(in[p + fx], w.z, acc[p][f4 * 4 + 2]);
                acc[p][f4 * 4 + 3] = fmaf(in[p + fx], w.w, acc[p][f4 * 4 + 3]);
              }
            }
          }
        }
      }
    }
  }
  if (!active) return;
  const int64_t b = b0 + s;
  const int out_plane = g.OH * g.OW;
#pragma unroll
  for (int f = 0; f < FT; ++f) {
    if (f >= g.OC) break;
    float *dst = Out + (b * g.OC + f) * out_plane + r * g.OW + x0;
#pragma unroll
    for (int q = 0; q < PIX / 4; ++q)
      *reinterpret_cast<float4 *>(dst + q * 4) =
          make_float4(acc[q * 4 + 0][f], acc[q * 4 + 1][f], acc[q * 4 + 2][f], acc[q * 4 + 3][f]);
  }
}

template <int FW, int FH, int FT, int PIX, bool BWD>
static int launch_s1(pb_context *ctx, const float *In, const float *Wt, float *Out, S1Geom g) {
  constexpr int NLOAD = (PIX + FW - 1 + 3) / 4;
  g.XG = g.OW / PIX;
  g.TPS = g.OH * g.XG;
  g.NS = std::max(1, 256 / g.TPS);
  g.IHs = g.OH + FH - 1;
  g.IWs = g.OW - PIX + 4 * NLOAD;
  const int threads = ((g.NS * g.TPS + 31) / 32) * 32;
  // Largest channel chunk that keeps two CTAs resident per SM (<= ~100 KB each).
  const size_t budget = 100 * 1024;
  int zc = g.IC;
  auto bytes = [&](int z) {
    return sizeof(float) * ((size_t)z * FH * FW * FT + (size_t)g.NS * z * g.IHs * g.IWs);
  };
  while (zc > 1 && bytes(zc) > budget) zc = (zc + 1) / 2;
  if (bytes(zc) > 200 * 1024) return -1;
  g.ZC = zc;
  const size_t smem = bytes(zc);
  auto kern = conv_s1_kernel<FW, FH, FT, PIX, BWD>;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = 200 * 1024;
  }
  const int grid = (g.batch + g.NS - 1) / g.NS;
  kern<<<grid, threads, smem, ctx->stream>>>(In, Wt, Out, g);
  PB_LAUNCHED(ctx);
  return 0;
}

// Returns -1 when the shape is outside the tiled kernels' envelope.
template <bool BWD>
static int dispatch_s1(pb_context *ctx, const pb_layer_desc *l, const float *In, const float *Wt,
                       float *Out, int64_t batch) {
  const int fw = (int)l->filter_width, fh = (int)l->filter_height, p = (int)l->padding;
  const int ow = (int)conv_out_w(l), oh = (int)conv_out_h(l);
  if (l->stride != 1 || batch > INT32_MAX) return -1;
  S1Geom g{};
  g.K1 = fw * fh * (int)l->depth + 1;
  g.batch = (int)batch;
  if (!BWD) {
    g.IC = (int)l->depth; g.IH = (int)l->height; g.IW = (int)l->width;
    g.OC = (int)l->num_filters; g.OH = oh; g.OW = ow;
    g.padY = p; g.padX = p;
  } else {
    g.IC = (int)l->num_filters; g.IH = oh; g.IW = ow;
    g.OC = (int)l->depth; g.OH = (int)l->height; g.OW = (int)l->width;
    g.padY = fh - 1 - p; g.padX = fw - 1 - p;
    if (g.padY < 0 || g.padX < 0) return -1;
    // centre guard: output (d,k) counts only if 0 <= d-p+fh/2 < H and 0 <= k-p+fw/2 < W
    g.guardLoY = std::max(0, p - fh / 2);
    g.guardHiY = std::min(oh - 1, (int)l->height - 1 + p - fh / 2);
    g.guardLoX = std::max(0, p - fw / 2);
    g.guardHiX = std::min(ow - 1, (int)l->width - 1 + p - fw / 2);
  }
  if (g.OW % 4 != 0 || g.OH * (g.OW / 4) > 256 || g.OC > 20) return -1;
  const bool pix8 = g.OC <= 4 && g.OW % 8 == 0;
#define PB_S1(FWV, FHV)                                                                        \
  if (fw == FWV && fh == FHV) {                                                                \
    if (pix8) return launch_s1<FWV, FHV, 4, 8, BWD>(ctx, In, Wt, Out, g);                      \
    if (g.OC <= 4) return launch_s1<FWV, FHV, 4, 4, BWD>(ctx, In, Wt, Out, g);                 \
    if (g.OC <= 8) return launch_s1<FWV, FHV, 8, 4, BWD>(ctx, In, Wt, Out, g);                 \
    if (g.OC <= 16) return launch_s1<FWV, FHV, 16, 4, BWD>(ctx, In, Wt, Out, g);               \
    return launch_s1<FWV, FHV, 20, 4, BWD>(ctx, In, Wt, Out, g);                               \
  }
  PB_S1(5, 5)
  PB_S1(3, 3)
#undef PB_S1
  return -1;
}

int conv_evaluate_tiled(pb_context *ctx, const pb_layer_desc *l, const float *I, const float *W,
                        float *O, int64_t batch) {
  return dispatch_s1<false>(ctx, l, I, W, O, batch);
}

int conv_input_delta_tiled(pb_context *ctx, const pb_layer_desc *l, const float *W,
                           const float *G, float *dI, int64_t batch) {
  return dispatch_s1<true>(ctx, l, G, W, dI, batch);
}

// ---------------------------------------------------------------------------
// weight gradient
// ---------------------------------------------------------------------------
struct WgGeom {
  int D, H, W, F, oh, ow, p, K1;
  int IHs, IWs;      // input tile with halo (IWs % 4 == 0)
  int P;             // (channel, filter-row) pairs + 1 bias pair
  int RC, rows_per_chunk;   // row chunks per f-group
  int U;             // threads per f-group (multiple of 32)
  int NFG;           // f-groups
  int in_elems, g_elems, ones_off, stage_elems;
  int batch, samples_per_cta;
};

// One pipeline stage of shared memory: [input tile | ones plane | gradient tile]
// OWC: compile-time output width (0 = runtime): the pixel loop of a row unrolls
// fully, shared-memory offsets become immediates and the loads of the next
// 4-pixel step are issued under the FFMAs of the current one.
template <int FW, int FH, int TF, int OWC, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
conv_wgrad_s1_kernel(const float *__restrict__ In, const float *__restrict__ G,
                     float *__restrict__ partial, const WgGeom g) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int fg = tid / g.U, u = tid - fg * g.U;
  const int rc = u / g.P, pr = u - rc * g.P;       // row chunk, pair
  const bool worker = fg < g.NFG && rc < g.RC;
  const bool bias_pair = pr == g.P - 1;
  const int z = pr / FH, fy = pr - z * FH;
  const int in_plane = g.H * g.W, out_plane = g.oh * g.ow;
  const int b_begin = blockIdx.x * g.samples_per_cta;
  const int b_end = min(g.batch, b_begin + g.samples_per_cta);

  // zero both stages' input tiles (halo stays zero), fill the ones planes
  for (int e = tid; e < 2 * g.stage_elems; e += blockDim.x) {
    const int o = e % g.stage_elems;
    smem[e] = (o >= g.ones_off && o < g.ones_off + g.oh * g.IWs + 8) ? 1.0f : 0.0f;
  }
  __syncthreads();

  auto issue_load = [&](int b, int stage) {
    float *st = smem + stage * g.stage_elems;
    const float *src = In + (int64_t)b * g.D * in_plane;
    for (int e = tid; e < g.D * in_plane; e += blockDim.x) {
      const int x = e % g.W;
      int t = e / g.W;
      const int y = t % g.H;
      const int zz = t / g.H;
      __pipeline_memcpy_async(st + (zz * g.IHs + y + g.p) * g.IWs + x + g.p, src + e, 4);
    }
    float *gs = st + g.ones_off + g.oh * g.IWs + 8;
    const float4 *gsrc = reinterpret_cast<const float4 *>(G + (int64_t)b * g.F * out_plane);
    for (int e = tid; e < g.F * out_plane / 4; e += blockDim.x)
      __pipeline_memcpy_async(reinterpret_cast<float4 *>(gs) + e, gsrc + e, 16);
  };

  float acc[TF][FW];
#pragma unroll
  for (int f = 0; f < TF; ++f)
#pragma unroll
    for (int x = 0; x < FW; ++x) acc[f][x] = 0.0f;

  if (b_begin < b_end) issue_load(b_begin, 0);
  __pipeline_commit();
  for (int b = b_begin; b < b_end; ++b) {
    const int stage = (b - b_begin) & 1;
    if (b + 1 < b_end) issue_load(b + 1, stage ^ 1);
    __pipeline_commit();
    __pipeline_wait_prior(1);
    __syncthreads();
    if (worker) {
      const float *st = smem + stage * g.stage_elems;
      const float *in_base = bias_pair ? st + g.ones_off : st + (z * g.IHs + fy) * g.IWs;
      const float *gs = st + g.ones_off + g.oh * g.IWs + 8 + (fg * TF) * out_plane;
      const int row0 = rc * g.rows_per_chunk;
      const int row1 = min(g.oh, row0 + g.rows_per_chunk);
      const int ow = OWC ? OWC : g.ow;
      for (int oy = row0; oy < row1; ++oy) {
        const float *irow = in_base + oy * g.IWs;
        const float *grow = gs + oy * ow;
        float4 lo = *reinterpret_cast<const float4 *>(irow);
#pragma unroll
        for (int ox = 0; ox < ow; ox += 4) {
          const float4 hi = *reinterpret_cast<const float4 *>(irow + ox + 4);
          const float in[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
          for (int f = 0; f < TF; ++f) {
            const float4 gv = *reinterpret_cast<const float4 *>(grow + f * out_plane + ox);
            const float gq[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
              for (int x = 0; x < FW; ++x) acc[f][x] = fmaf(gq[q], in[q + x], acc[f][x]);
          }
          lo = hi;
        }
      }
    }
    __syncthreads();
  }
  __pipeline_wait_prior(0);

  // reduce the row chunks of each (f-group, pair) through shared memory
  __syncthreads();
  float *red = smem;  // [NFG][RC][P][TF*FW]
  if (worker) {
    float *dst = red + ((fg * g.RC + rc) * g.P + pr) * (TF * FW);
#pragma unroll
    for (int f = 0; f < TF; ++f)
#pragma unroll
      for (int x = 0; x < FW; ++x) dst[f * FW + x] = acc[f][x];
  }
  __syncthreads();
  float *out = partial + (int64_t)blockIdx.x * g.F * g.K1;
  for (int e = tid; e < g.NFG * g.P * TF * FW; e += blockDim.x) {
    const int x = e % FW;
    int t = e / FW;
    const int f_in = t % TF; t /= TF;
    const int pp = t % g.P;
    const int fgi = t / g.P;
    const int f = fgi * TF + f_in;
    if (f >= g.F) continue;
    float sum = 0.0f;
    for (int c = 0; c < g.RC; ++c) sum += red[((fgi * g.RC + c) * g.P + pp) * (TF * FW) + f_in * FW + x];
    if (pp == g.P - 1) {
      if (x == 0) out[f * g.K1 + g.K1 - 1] = sum;   // bias: every tap saw the same ones
    } else {
      const int zz = pp / FH, yy = pp - zz * FH;
      out[f * g.K1 + (zz * FH + yy) * FW + x] = sum;
    }
  }
}

// Deterministic reduction of the per-CTA partial sums + SGD update.  A CTA owns
// 32 weights; its 8 warps each sum every 8th partial (coalesced rows), then one
// warp adds the 8 sub-sums in a fixed order.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float *__restrict__ partial, int n_partial, int nw, const float *Wt,
                    float *out, const float *__restrict__ lr, int mode) {
  __shared__ float red[8][33];
  chain_wait();  // chained launch, common.cuh
  chain_release();
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int w = blockIdx.x * 32 + lane;
  float s0 = 0.0f, s1 = 0.0f;
  if (w < nw) {
    int i = g;
    for (; i + 8 < n_partial; i += 16) {
      s0 += partial[(int64_t)i * nw + w];
      s1 += partial[(int64_t)(i + 8) * nw + w];
    }
    if (i < n_partial) s0 += partial[(int64_t)i * nw + w];
  }
  red[g][lane] = s0 + s1;
  __syncthreads();
  if (g == 0 && w < nw) {
    float sum = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += red[k][lane];
    apply_update(mode, out, Wt, w, lr[0], sum);
  }
}

int reduce_partials_update(pb_context *ctx, const float *partial, int n_partial, int nw,
                           const float *Wt, float *out, const float *lr, int mode) {
  PB_CUDA(launch_forked(ctx, wgrad_reduce_kernel, dim3(pb_div_up(nw, 32)), dim3(256), 0, partial,
                         n_partial, nw, Wt, out, lr, mode));
  PB_LAUNCHED(ctx);
  return 0;
}

template <int FW, int FH, int TF, int OWC = 0, int MAXT = 1024>
static int launch_wgrad(pb_context *ctx, const float *I, const float *Wt, const float *G,
                        float *out, const float *lr, int mode, WgGeom g) {
  g.NFG = (g.F + TF - 1) / TF;
  g.P = g.D * FH + 1;
  // thread budget per f-group
  const int max_u = (MAXT / g.NFG) / 32 * 32;
  if (max_u < 32 || g.P > max_u) return -1;
  g.RC = std::max(1, std::min(g.oh, max_u / g.P));
  g.rows_per_chunk = (g.oh + g.RC - 1) / g.RC;
  g.RC = (g.oh + g.rows_per_chunk - 1) / g.rows_per_chunk;
  g.U = ((g.RC * g.P + 31) / 32) * 32;
  const int threads = g.U * g.NFG;
  g.IHs = g.H + 2 * g.p;
  g.IWs = ((g.ow + 4 + 3) / 4) * 4;          // reads reach column ow+3+4
  if (g.IWs < g.W + 2 * g.p) g.IWs = ((g.W + 2 * g.p + 3) / 4) * 4;
  g.in_elems = g.D * g.IHs * g.IWs;
  // the F gradient planes are padded up to NFG*TF planes (zero-filled once)
  g.g_elems = g.NFG * TF * g.oh * g.ow;
  g.ones_off = g.in_elems;
  g.stage_elems = ((g.in_elems + g.oh * g.IWs + 8 + g.g_elems + 3) / 4) * 4;
  const size_t red_elems = (size_t)g.NFG * g.RC * g.P * TF * FW;
  const size_t smem = sizeof(float) * std::max<size_t>(2 * (size_t)g.stage_elems, red_elems);
  if (smem > 220 * 1024) return -1;
  if ((g.F * g.oh * g.ow) % 4 != 0 || g.ow % 4 != 0) return -1;
  // last input row read: fy + oh - 1 <= IHs - 1  (stride 1: oh = H + 2p - FH + 1)
  if (FH - 1 + g.oh - 1 > g.IHs - 1) return -1;
  const int n_cta = std::min(g.batch, ctx->num_sms);
  g.samples_per_cta = (g.batch + n_cta - 1) / n_cta;
  const int grid = (g.batch + g.samples_per_cta - 1) / g.samples_per_cta;
  const int nw = g.F * g.K1;
  if (ensure_workspace(ctx, (size_t)grid * nw)) return 1;
  auto kern = conv_wgrad_s1_kernel<FW, FH, TF, OWC, MAXT>;
  PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  kern<<<grid, threads, smem, ctx->stream>>>(I, G, ctx->workspace, g);
  PB_LAUNCHED(ctx);
  return reduce_partials_update(ctx, ctx->workspace, grid, nw, Wt, out, lr, mode);
}


// ---------------------------------------------------------------------------
// Warp-specialised weight gradient (square "same" stride-1 layers, S = W = H =
// ow = oh known at compile time).  Same per-thread register tile as
// conv_wgrad_s1_kernel, different orchestration:
//   warp 0        producer: cp.async of the next sample's input rows (8 B) and
//                 gradient planes (16 B) into an NST-stage ring; completion is
//                 signalled on an mbarrier (cp.async.mbarrier.arrive.noinc);
//   warps 1..     consumers: (f-group, row-chunk, pair) units laid out densely
//                 over the consumer threads; wait(full[s]) -> FFMA -> one
//                 arrive(empty[s]) per warp.
// There is no CTA-wide barrier inside the sample loop, so warps drift and the
// FFMA pipe stays fed while the ring refills.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ws_smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void ws_mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ws_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ws_mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ws_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ws_mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WS_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra WS_DONE;\n\t"
      "bra WS_WAIT;\n\t"
      "WS_DONE:\n\t"
      "}\n" ::"r"(ws_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void ws_cp_async8(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(ws_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void ws_cp_async16(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ws_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void ws_cp_async_arrive(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(ws_smem_u32(bar)) : "memory");
}

constexpr int kWsMaxStages = 8;

template <int FW, int FH, int TF, int S, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
conv_wgrad_ws_kernel(const float *__restrict__ In, const float *__restrict__ G,
                     float *__restrict__ partial, const WgGeom g, const int nstage) {
  extern __shared__ __align__(16) float smem[];
  __shared__ uint64_t full[kWsMaxStages], empty[kWsMaxStages];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_cons_warps = (blockDim.x >> 5) - 1;
  const int b_begin = blockIdx.x * g.samples_per_cta;
  const int b_end = min(g.batch, b_begin + g.samples_per_cta);
  constexpr int PLANE = S * S;

  // zero every stage's input tile (the halo stays zero), fill the ones planes
  for (int e = tid; e < nstage * g.stage_elems; e += blockDim.x) {
    const int o = e % g.stage_elems;
    smem[e] = (o >= g.ones_off && o < g.ones_off + S * g.IWs + 8) ? 1.0f : 0.0f;
  }
  if (tid == 0) {
    for (int s = 0; s < nstage; ++s) {
      ws_mbar_init(&full[s], 32);
      ws_mbar_init(&empty[s], n_cons_warps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  float acc[TF][FW];
#pragma unroll
  for (int f = 0; f < TF; ++f)
#pragma unroll
    for (int x = 0; x < FW; ++x) acc[f][x] = 0.0f;

  // consumer unit of this thread (valid for warps >= 1)
  const int u = tid - 32;
  const int per_fg = g.RC * g.P;
  const int fg = u >= 0 ? u / per_fg : 0;
  const int r2 = u >= 0 ? u - fg * per_fg : 0;
  const int rc = r2 / g.P, pr = r2 - rc * g.P;
  const bool worker = u >= 0 && fg < g.NFG;
  const bool bias_pair = pr == g.P - 1;
  const int z = pr / FH, fy = pr - z * FH;

  if (warp == 0) {
    // ---- producer ----------------------------------------------------------
    int it = 0;
    for (int b = b_begin; b < b_end; ++b, ++it) {
      const int s = it % nstage, ph = (it / nstage) & 1;
      ws_mbar_wait(&empty[s], ph ^ 1);
      float *st = smem + s * g.stage_elems;
      const float *src = In + (int64_t)b * g.D * PLANE;
      for (int e = lane; e < g.D * PLANE / 2; e += 32) {   // float2 chunks of the interior
        const int x2 = e % (S / 2);
        const int t = e / (S / 2);
        const int y = t % S, zz = t / S;
        ws_cp_async8(st + (zz * g.IHs + y + g.p) * g.IWs + 2 * x2 + g.p, src + 2 * e);
      }
      float *gs = st + g.ones_off + S * g.IWs + 8;
      const float *gsrc = G + (int64_t)b * g.F * PLANE;
      for (int e = lane; e < g.F * PLANE / 4; e += 32) ws_cp_async16(gs + 4 * e, gsrc + 4 * e);
      ws_cp_async_arrive(&full[s]);
    }
  } else {
    // ---- consumers ---------------------------------------------------------
    int it = 0;
    for (int b = b_begin; b < b_end; ++b, ++it) {
      const int s = it % nstage, ph = (it / nstage) & 1;
      ws_mbar_wait(&full[s], ph);
      if (worker) {
        const float *st = smem + s * g.stage_elems;
        const float *in_base = bias_pair ? st + g.ones_off : st + (z * g.IHs + fy) * g.IWs;
        const float *gs = st + g.ones_off + S * g.IWs + 8 + (fg * TF) * PLANE;
        const int row0 = rc * g.rows_per_chunk;
        const int row1 = min(S, row0 + g.rows_per_chunk);
        for (int oy = row0; oy < row1; ++oy) {
          const float *irow = in_base + oy * g.IWs;
          const float *grow = gs + oy * S;
          float4 lo = *reinterpret_cast<const float4 *>(irow);
#pragma unroll
          for (int ox = 0; ox < S; ox += 4) {
            const float4 hi = *reinterpret_cast<const float4 *>(irow + ox + 4);
            const float in[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
            for (int f = 0; f < TF; ++f) {
              const float4 gv = *reinterpret_cast<const float4 *>(grow + f * PLANE + ox);
              const float gq[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
              for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int x = 0; x < FW; ++x) acc[f][x] = fmaf(gq[q], in[q + x], acc[f][x]);
            }
            lo = hi;
          }
        }
      }
      __syncwarp();
      if (lane == 0) ws_mbar_arrive(&empty[s]);
    }
  }

  // reduce the row chunks of each (f-group, pair) through shared memory
  __syncthreads();
  float *red = smem;  // [NFG][RC][P][TF*FW]
  if (worker) {
    float *dst = red + ((fg * g.RC + rc) * g.P + pr) * (TF * FW);
#pragma unroll
    for (int f = 0; f < TF; ++f)
#pragma unroll
      for (int x = 0; x < FW; ++x) dst[f * FW + x] = acc[f][x];
  }
  __syncthreads();
  float *out = partial + (int64_t)blockIdx.x * g.F * g.K1;
  for (int e = tid; e < g.NFG * g.P * TF * FW; e += blockDim.x) {
    const int x = e % FW;
    int t = e / FW;
    const int f_in = t % TF; t /= TF;
    const int pp = t % g.P;
    const int fgi = t / g.P;
    const int f = fgi * TF + f_in;
    if (f >= g.F) continue;
    float sum = 0.0f;
    for (int c = 0; c < g.RC; ++c) sum += red[((fgi * g.RC + c) * g.P + pp) * (TF * FW) + f_in * FW + x];
    if (pp == g.P - 1) {
      if (x == 0) out[f * g.K1 + g.K1 - 1] = sum;   // bias: every tap saw the same ones
    } else {
      const int zz = pp / FH, yy = pp - zz * FH;
      out[f * g.K1 + (zz * FH + yy) * FW + x] = sum;
    }
  }
}

// MAXT bounds the block size so ptxas may use 65536/MAXT registers per thread
// (the accumulator tile spills at the 64 registers a 1024-thread block allows).
template <int FW, int FH, int TF, int S, int MAXT>
static int launch_wgrad_ws(pb_context *ctx, const float *I, const float *Wt, const float *G,
                           float *out, const float *lr, int mode, WgGeom g) {
  if (g.W != S || g.H != S || g.ow != S || g.oh != S || (S & 3)) return -1;
  g.NFG = (g.F + TF - 1) / TF;
  g.P = g.D * FH + 1;
  // densest row-chunk split that fits 31 consumer warps
  const int max_units = MAXT - 32;
  if (g.NFG * g.P > max_units) return -1;
  g.RC = std::max(1, std::min(S, max_units / (g.NFG * g.P)));
  g.rows_per_chunk = (S + g.RC - 1) / g.RC;
  g.RC = (S + g.rows_per_chunk - 1) / g.rows_per_chunk;
  const int units = g.NFG * g.RC * g.P;
  const int threads = 32 + ((units + 31) / 32) * 32;
  g.IHs = S + 2 * g.p;
  g.IWs = ((S + 4 + 3) / 4) * 4;
  if (g.IWs < S + 2 * g.p) g.IWs = ((S + 2 * g.p + 3) / 4) * 4;
  g.in_elems = g.D * g.IHs * g.IWs;
  g.g_elems = g.NFG * TF * S * S;
  g.ones_off = g.in_elems;
  g.stage_elems = ((g.in_elems + S * g.IWs + 8 + g.g_elems + 3) / 4) * 4;
  if ((g.ones_off + S * g.IWs + 8) % 4 != 0 || (g.p & 1)) return -1;  // cp.async alignment
  const size_t red_elems = (size_t)g.NFG * g.RC * g.P * TF * FW;
  const size_t budget = 220 * 1024;
  int nstage = (int)std::min<size_t>(kWsMaxStages, budget / (sizeof(float) * g.stage_elems));
  if (nstage < 2) return -1;
  const size_t smem = sizeof(float) * std::max<size_t>((size_t)nstage * g.stage_elems, red_elems);
  if (smem > budget) return -1;
  if (FH - 1 + S - 1 > g.IHs - 1) return -1;
  const int n_cta = std::min(g.batch, ctx->num_sms);
  g.samples_per_cta = (g.batch + n_cta - 1) / n_cta;
  const int grid = (g.batch + g.samples_per_cta - 1) / g.samples_per_cta;
  const int nw = g.F * g.K1;
  if (ensure_workspace(ctx, (size_t)grid * nw)) return 1;
  auto kern = conv_wgrad_ws_kernel<FW, FH, TF, S, MAXT>;
  PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
  kern<<<grid, threads, smem, ctx->stream>>>(I, G, ctx->workspace, g, nstage);
  PB_LAUNCHED(ctx);
  return reduce_partials_update(ctx, ctx->workspace, grid, nw, Wt, out, lr, mode);
}

int conv_weight_update_tiled(pb_context *ctx, const pb_layer_desc *l, const float *I,
                             const float *Wt, const float *G, float *out, const float *lr,
                             int mode, int64_t batch) {
  if (l->stride != 1 || batch > INT32_MAX || batch < 1) return -1;
  WgGeom g{};
  g.D = (int)l->depth; g.H = (int)l->height; g.W = (int)l->width;
  g.F = (int)l->num_filters; g.oh = (int)conv_out_h(l); g.ow = (int)conv_out_w(l);
  g.p = (int)l->padding;
  const int fw = (int)l->filter_width, fh = (int)l->filter_height;
  g.K1 = fw * fh * g.D + 1;
  g.batch = (int)batch;
  if (g.F > 32) return -1;
  if (fw == 5 && fh == 5) {
    if (g.W == g.ow && g.H == g.oh && g.ow == g.oh) {
      int rc = -1;
      if (g.ow == 32) rc = launch_wgrad_ws<5, 5, 4, 32, 768>(ctx, I, Wt, G, out, lr, mode, g);
      else if (g.ow == 16) rc = launch_wgrad_ws<5, 5, 4, 16, 896>(ctx, I, Wt, G, out, lr, mode, g);
      else if (g.ow == 8) rc = launch_wgrad_ws<5, 5, 4, 8, 576>(ctx, I, Wt, G, out, lr, mode, g);
      if (rc < 0) {  // shape does not fit the tuned block size: generic budget
        if (g.ow == 32) rc = launch_wgrad_ws<5, 5, 4, 32, 1024>(ctx, I, Wt, G, out, lr, mode, g);
        else if (g.ow == 16) rc = launch_wgrad_ws<5, 5, 4, 16, 1024>(ctx, I, Wt, G, out, lr, mode, g);
        else if (g.ow == 8) rc = launch_wgrad_ws<5, 5, 4, 8, 1024>(ctx, I, Wt, G, out, lr, mode, g);
      }
      if (rc >= 0) return rc;
    }
    if (g.ow == 32) return launch_wgrad<5, 5, 4, 32>(ctx, I, Wt, G, out, lr, mode, g);
    if (g.ow == 16) return launch_wgrad<5, 5, 4, 16>(ctx, I, Wt, G, out, lr, mode, g);
    if (g.ow == 8) return launch_wgrad<5, 5, 4, 8>(ctx, I, Wt, G, out, lr, mode, g);
    return launch_wgrad<5, 5, 4>(ctx, I, Wt, G, out, lr, mode, g);
  }
  if (fw == 3 && fh == 3) return launch_wgrad<3, 3, 4>(ctx, I, Wt, G, out, lr, mode, g);
  return -1;
}

}  // namespace pb
