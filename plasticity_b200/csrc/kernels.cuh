// Internal launch functions shared between translation units.
#pragma once
#include <algorithm>

#include "common.cuh"

namespace pb {

// elementwise.cu
int activation_evaluate(pb_context *ctx, int act, const float *I, float *O, int64_t n);
int activation_input_delta(pb_context *ctx, int act, const float *I, const float *G,
                           float *dI, int64_t n);
int pool_evaluate(pb_context *ctx, const pb_layer_desc *l, const float *I, float *O,
                  int64_t batch);
int pool_input_delta(pb_context *ctx, const pb_layer_desc *l, const float *I,
                     const float *G, float *dI, int64_t batch);
int softmax_evaluate(pb_context *ctx, int64_t n, const float *I, float *O, int64_t batch);
int softmax_input_delta(pb_context *ctx, int64_t n, const float *I, const float *G,
                        float *dI, int64_t batch);

int act_pool_input_delta(pb_context *ctx, int act, const pb_layer_desc *pool, const float *O,
                         const float *Gp, float *dO, int64_t batch);

// dense.cu
int dense_evaluate(pb_context *ctx, const pb_layer_desc *l, const float *I, const float *W,
                   float *O, int64_t batch);
int dense_input_delta(pb_context *ctx, const pb_layer_desc *l, const float *W,
                      const float *G, float *dI, int64_t batch);
int dense_weight_update(pb_context *ctx, const pb_layer_desc *l, const float *I,
                        const float *W, const float *G, float *out, const float *lr,
                        int mode, int64_t batch);

// dense_umma.cu: 3xTF32 tcgen05 GEMM for large batched dense products; -1 when
// the product is too small to be worth a tensor-core launch (or PB_DISABLE_UMMA=1).
int dense_evaluate_umma(pb_context *ctx, const pb_layer_desc *l, const float *I, const float *W,
                        float *O, int64_t batch);
int dense_input_delta_umma(pb_context *ctx, const pb_layer_desc *l, const float *W, const float *G,
                           float *dI, int64_t batch);
int dense_weight_update_umma(pb_context *ctx, const pb_layer_desc *l, const float *I,
                             const float *W, const float *G, float *out, const float *lr,
                             int mode, int64_t batch);

// head.cu: fused dense -> activation -> softmax -> error-gradient head (forward and
// backward in one launch) and the skinny dense weight gradient; -1 outside the envelope.
int head_forward_backward(pb_context *ctx, const pb_layer_desc *dense, int act, int64_t softmax_n,
                          int loss, const float *X, const float *W, const float *E, float *Z,
                          float *A, float *Y, float *Gz, float *dX, int64_t batch);
int dense_weight_update_skinny(pb_context *ctx, const pb_layer_desc *l, const float *X,
                               const float *W, const float *G, float *out, const float *lr,
                               int mode, int64_t batch);
// conv_tiled.cu: out (=|+=) [W] - lr * sum_p partial[p][w], fixed summation order.
int reduce_partials_update(pb_context *ctx, const float *partial, int n_partial, int nw,
                           const float *Wt, float *out, const float *lr, int mode);

// conv.cu
int conv_evaluate(pb_context *ctx, const pb_layer_desc *l, const float *I, const float *W,
                  float *O, int64_t batch);
int conv_input_delta(pb_context *ctx, const pb_layer_desc *l, const float *W,
                     const float *G, float *dI, int64_t batch);
int conv_weight_update(pb_context *ctx, const pb_layer_desc *l, const float *I,
                       const float *W, const float *G, float *out, const float *lr,
                       int mode, int64_t batch);

// conv_tiled.cu: register-tiled stride-1 kernels; return -1 when the shape is
// outside their envelope (the caller then uses the general kernels).
int conv_evaluate_tiled(pb_context *ctx, const pb_layer_desc *l, const float *I, const float *W,
                        float *O, int64_t batch);
int conv_input_delta_tiled(pb_context *ctx, const pb_layer_desc *l, const float *W,
                           const float *G, float *dI, int64_t batch);
int conv_weight_update_tiled(pb_context *ctx, const pb_layer_desc *l, const float *I,
                             const float *W, const float *G, float *out, const float *lr,
                             int mode, int64_t batch);

// conv_wgrad_mma.cu: 3xTF32 mma.sync weight gradient for the 5x5 stride-1 layers whose
// channel counts fill an MMA tile; -1 outside the envelope.
// conv_wgrad_rows_umma.cu: the same for the first layer of the CIFAR network on tcgen05 (bands
// of image rows, the row offset as a free index of both operands); tried first.
int conv_act_pool_weight_update_rows(pb_context *ctx, const pb_layer_desc *l, int act,
                                     const pb_layer_desc *pool, const float *I, const float *Wt,
                                     const float *O, const float *Gp, float *out, const float *lr,
                                     int mode, int64_t batch);
int conv_act_pool_weight_update_mma(pb_context *ctx, const pb_layer_desc *l, int act,
                                    const pb_layer_desc *pool, const float *I, const float *Wt,
                                    const float *O, const float *Gp, float *out, const float *lr,
                                    int mode, int64_t batch);
int conv_weight_update_mma(pb_context *ctx, const pb_layer_desc *l, const float *I, const float *Wt,
                           const float *G, float *out, const float *lr, int mode, int64_t batch);

// conv_wgrad_umma.cu: 3xTF32 tcgen05 weight gradient (gradient operand in tensor memory) for
// the same layers; tried first, -1 outside the envelope.
int conv_weight_update_umma(pb_context *ctx, const pb_layer_desc *l, const float *I, const float *Wt,
                            const float *G, float *out, const float *lr, int mode, int64_t batch);

// conv_umma.cu: 3xTF32 tcgen05 implicit-GEMM kernels for the batched CIFAR
// shapes; return -1 outside their envelope.
int conv_evaluate_umma(pb_context *ctx, const pb_layer_desc *l, const float *I, const float *W,
                       float *O, int64_t batch);
int conv_input_delta_umma(pb_context *ctx, const pb_layer_desc *l, const float *W, const float *G,
                          float *dI, int64_t batch);
// convolution + activation + 2x2 max pool in one kernel: writes the convolution
// output (pre-activation, needed by back-prop) and the pooled output.
int conv_act_pool_evaluate_umma(pb_context *ctx, const pb_layer_desc *conv, int act,
                                const pb_layer_desc *pool, const float *I, const float *W,
                                float *O_conv, float *O_pool, int64_t batch);

// the backward of such a block: pool + activation backward formed inside the producers of the
// input-gradient kernel; also writes G_out = dE/d(convolution output) for the weight gradient.
int conv_act_pool_input_delta_umma(pb_context *ctx, const pb_layer_desc *conv, int act,
                                   const pb_layer_desc *pool, const float *W, const float *O_conv,
                                   const float *Gp, float *G_out, float *dI, int64_t batch);

// conv_wt_umma.cu: the same four entry points with the filters in tensor memory (pixels are
// the MMA's N, the fx taps are summed in the epilogue); tried first, -1 outside the envelope.
int conv_evaluate_wt(pb_context *ctx, const pb_layer_desc *l, const float *I, const float *W,
                     float *O, int64_t batch);
int conv_act_pool_evaluate_wt(pb_context *ctx, const pb_layer_desc *conv, int act,
                              const pb_layer_desc *pool, const float *I, const float *W,
                              float *O_conv, float *O_pool, int64_t batch);
int conv_input_delta_wt(pb_context *ctx, const pb_layer_desc *l, const float *W, const float *G,
                        float *dI, int64_t batch);
int conv_act_pool_input_delta_wt(pb_context *ctx, const pb_layer_desc *conv, int act,
                                 const pb_layer_desc *pool, const float *W, const float *O_conv,
                                 const float *Gp, float *G_out, float *dI, int64_t batch);

// conv_rows_umma.cu: the forward pass of the network's first (RGB) layer with the image rows in
// tensor memory and the fy taps summed over tiles in registers; tried first, -1 outside the
// envelope.
int conv_evaluate_rows(pb_context *ctx, const pb_layer_desc *l, const float *I, const float *W,
                       float *O, int64_t batch);
int conv_act_pool_evaluate_rows(pb_context *ctx, const pb_layer_desc *conv, int act,
                                const pb_layer_desc *pool, const float *I, const float *W,
                                float *O_conv, float *O_pool, int64_t batch);

// conv_gemm_umma.cu: general implicit-GEMM tensor-core convolution (forward) for large
// batched shapes, any filter / stride / padding; O2 != nullptr additionally receives
// act(O), the output of a following ActivationLayer.  -1 outside the envelope.
int conv_evaluate_gemm_umma(pb_context *ctx, const pb_layer_desc *l, const float *I,
                            const float *W, float *O, float *O2, int act, int64_t batch);

// the same skeleton for the backward pass of any convolution: input gradient (conv_gemm_umma.cu,
// odd filters with pad == filter/2, filter counts that are multiples of 16) and weight gradient
// + SGD update (dense_umma.cu, any shape); -1 outside the envelope.
int conv_input_delta_gemm_umma(pb_context *ctx, const pb_layer_desc *l, const float *W,
                               const float *G, float *dI, int64_t batch);
int conv_weight_update_gemm_umma(pb_context *ctx, const pb_layer_desc *l, const float *I,
                                 const float *Wt, const float *G, float *out, const float *lr,
                                 int mode, int64_t batch);

// train_small.cu: n per-sample SGD steps of a small all-dense network as ONE persistent
// single-CTA kernel (weights, activations and gradients in shared memory); -1 outside the
// envelope (the graph-replayed layer-by-layer loop then runs).
int train_loop_small(pb_context *ctx, const pb_layer_desc *layers, int n_layers,
                     const int64_t *w_off, int64_t total_w, int loss, float *weights,
                     const float *lr, const float *d_ins, const float *d_expected,
                     int64_t n_samples);
// records.cu: pb_records_decode on an explicit stream.
int records_decode_on(pb_context *ctx, cudaStream_t stream, const uint8_t *d_records,
                      float *d_inputs, float *d_expected, int64_t n_records, int64_t sample_size,
                      int64_t n_classes);

// Output extent of a convolution, nnet/convolution_layer.cc:36-44.
static inline int64_t conv_out_w(const pb_layer_desc *l) {
  return (l->width - l->filter_width + 2 * l->padding) / l->stride + 1;
}
static inline int64_t conv_out_h(const pb_layer_desc *l) {
  return (l->height - l->filter_height + 2 * l->padding) / l->stride + 1;
}

// out = W - lr*g | -lr*g | out - lr*g   (back_prop.kernel.cl:28-44 + combine)
__device__ __forceinline__ void apply_update(int mode, float *out, const float *W,
                                             int64_t idx, float lr, float g) {
  if (mode == PB_NEW_WEIGHTS) out[idx] = W[idx] - lr * g;
  else if (mode == PB_WEIGHT_DELTAS) out[idx] = -lr * g;
  else out[idx] += -lr * g;
}

// ---- activation functions (symbolic/symbolic_util.cc:10-28) ---------------------
template <int ACT>
__device__ __forceinline__ float act_fwd(float x) {
  if (ACT == PB_RELU) return (x >= 0.0f) ? x : 0.0f;
  if (ACT == PB_LEAKY_RELU) {
    // x / 10 as reciprocal multiply + one FMA correction step (the correctly
    // rounded quotient for normal operands) instead of the generic division.
    const float q = x * 0.1f;
    return (x >= 0.0f) ? x : fmaf(fmaf(q, -10.0f, x), 0.1f, q);
  }
  if (ACT == PB_SIGMOID) return 1.0f / (1.0f + ref_exp(-x));
  return x;
}

// d/dx as the reference's symbolic Derive() prints it: relu (x>=0 ? 1 : 0),
// leaky (x>=0 ? 1 : 10/100), sigmoid p/((1+p)(1+p)) with p = 2.71828^-x
// recomputed from the INPUT (not the output).
template <int ACT>
__device__ __forceinline__ float act_bwd(float x, float g) {
  if (ACT == PB_RELU) return g * ((x >= 0.0f) ? 1.0f : 0.0f);
  if (ACT == PB_LEAKY_RELU) return g * ((x >= 0.0f) ? 1.0f : 0.1f);
  if (ACT == PB_SIGMOID) {
    const float p = ref_exp(-x);
    const float q = 1.0f + p;
    return g * (p / (q * q));
  }
  return g * 1.0f;
}

// The same with the activation kind as a run-time value (persistent per-sample kernels).
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

}  // namespace pb
