// The per-kernel "operator" entry points: what Nnet crosses per launch in the
// reference (kernel name + NDRange + positional buffers; nnet/nnet.h:243-253,
// 406-418, 435-447) becomes (layer descriptor, device pointers, batch).
#include "common.cuh"
#include "kernels.cuh"

namespace pb {

int layer_check(const pb_layer_desc *l) {
  PB_CHECK(l, "layer: null descriptor");
  switch (l->kind) {
    case PB_ACTIVATION:
      PB_CHECK(l->activation >= PB_IDENTITY && l->activation <= PB_SIGMOID,
               "ActivationLayer: only symbolic::Identity/Relu/LeakyRelu/Sigmoid are supported "
               "(no code generator, no fallback)");
      PB_CHECK(l->n >= 0, "ActivationLayer: negative size");
      return 0;
    case PB_SOFTMAX:
      PB_CHECK(l->n >= 0, "SoftmaxLayer: negative size");
      return 0;
    case PB_DENSE:
      PB_CHECK(l->in >= 0 && l->out >= 0, "DenseLayer: negative dimensions");
      return 0;
    case PB_CONVOLUTION:
      PB_CHECK(l->filter_depth == l->depth,
               "Convolution layer input depth != filter depth. Error!");
      PB_CHECK(l->stride >= 1, "Convolution layer: stride must be >= 1");
      return 0;
    case PB_MAX_POOL:
      PB_CHECK(l->out_width >= 1 && l->out_height >= 1 && l->out_width <= l->width &&
                   l->out_height <= l->height,
               "MaxPoolLayer constructed with more outputs than inputs?");
      return 0;
    default:
      return fail("layer: unknown kind");
  }
}

}  // namespace pb

extern "C" {

int pb_layer_num_inputs(const pb_layer_desc *l, int64_t *n) {
  PB_CHECK(l && n, "pb_layer_num_inputs: null");
  switch (l->kind) {
    case PB_ACTIVATION:
    case PB_SOFTMAX: *n = l->n; return 0;
    case PB_DENSE: *n = l->in; return 0;
    case PB_CONVOLUTION:
    case PB_MAX_POOL: *n = l->width * l->height * l->depth; return 0;
  }
  return pb::fail("pb_layer_num_inputs: unknown layer kind");
}

int pb_layer_num_outputs(const pb_layer_desc *l, int64_t *n) {
  PB_CHECK(l && n, "pb_layer_num_outputs: null");
  switch (l->kind) {
    case PB_ACTIVATION:
    case PB_SOFTMAX: *n = l->n; return 0;
    case PB_DENSE: *n = l->out; return 0;
    case PB_CONVOLUTION:
      *n = pb::conv_out_w(l) * pb::conv_out_h(l) * l->num_filters;
      return 0;
    case PB_MAX_POOL: *n = l->out_width * l->out_height * l->depth; return 0;
  }
  return pb::fail("pb_layer_num_outputs: unknown layer kind");
}

int pb_layer_num_weights(const pb_layer_desc *l, int64_t *n) {
  PB_CHECK(l && n, "pb_layer_num_weights: null");
  switch (l->kind) {
    case PB_DENSE: *n = l->out * (l->in + 1); return 0;
    case PB_CONVOLUTION:
      *n = l->num_filters * (l->filter_width * l->filter_height * l->filter_depth + 1);
      return 0;
    case PB_ACTIVATION:
    case PB_SOFTMAX:
    case PB_MAX_POOL: *n = 0; return 0;
  }
  return pb::fail("pb_layer_num_weights: unknown layer kind");
}

int pb_layer_evaluate(pb_context *ctx, const pb_layer_desc *l, const float *I,
                      const float *W, float *O, int64_t batch) {
  PB_CHECK(ctx && I && O, "pb_layer_evaluate: null");
  PB_CHECK(batch >= 0, "pb_layer_evaluate: negative batch");
  if (pb::layer_check(l)) return 1;
  switch (l->kind) {
    case PB_ACTIVATION: return pb::activation_evaluate(ctx, l->activation, I, O, l->n * batch);
    case PB_SOFTMAX: return pb::softmax_evaluate(ctx, l->n, I, O, batch);
    case PB_MAX_POOL: return pb::pool_evaluate(ctx, l, I, O, batch);
    case PB_DENSE:
      PB_CHECK(W, "pb_layer_evaluate: dense layer needs weights");
      return pb::dense_evaluate(ctx, l, I, W, O, batch);
    case PB_CONVOLUTION:
      PB_CHECK(W, "pb_layer_evaluate: convolution layer needs weights");
      return pb::conv_evaluate(ctx, l, I, W, O, batch);
  }
  return 1;
}

int pb_layer_input_delta(pb_context *ctx, const pb_layer_desc *l, const float *I,
                         const float *W, const float *G, float *dI, int64_t batch) {
  PB_CHECK(ctx && G && dI, "pb_layer_input_delta: null");
  PB_CHECK(batch >= 0, "pb_layer_input_delta: negative batch");
  if (pb::layer_check(l)) return 1;
  switch (l->kind) {
    case PB_ACTIVATION:
      PB_CHECK(I, "pb_layer_input_delta: activation layer needs its inputs");
      return pb::activation_input_delta(ctx, l->activation, I, G, dI, l->n * batch);
    case PB_SOFTMAX:
      PB_CHECK(I, "pb_layer_input_delta: softmax layer needs its inputs");
      return pb::softmax_input_delta(ctx, l->n, I, G, dI, batch);
    case PB_MAX_POOL:
      PB_CHECK(I, "pb_layer_input_delta: max-pool layer needs its inputs");
      return pb::pool_input_delta(ctx, l, I, G, dI, batch);
    case PB_DENSE:
      PB_CHECK(W, "pb_layer_input_delta: dense layer needs weights");
      return pb::dense_input_delta(ctx, l, W, G, dI, batch);
    case PB_CONVOLUTION:
      PB_CHECK(W, "pb_layer_input_delta: convolution layer needs weights");
      return pb::conv_input_delta(ctx, l, W, G, dI, batch);
  }
  return 1;
}

int pb_layer_weight_update(pb_context *ctx, const pb_layer_desc *l, const float *I,
                           const float *W, const float *G, float *out, const float *lr,
                           int mode, int64_t batch) {
  PB_CHECK(ctx, "pb_layer_weight_update: null ctx");
  if (pb::layer_check(l)) return 1;
  PB_CHECK(mode >= PB_NEW_WEIGHTS && mode <= PB_ACCUMULATE,
           "pb_layer_weight_update: bad mode");
  if (l->kind != PB_DENSE && l->kind != PB_CONVOLUTION) return 0;  // no weights
  PB_CHECK(I && G && out && lr, "pb_layer_weight_update: null buffer");
  PB_CHECK(mode != PB_NEW_WEIGHTS || W, "pb_layer_weight_update: NEW_WEIGHTS needs W");
  if (l->kind == PB_DENSE) return pb::dense_weight_update(ctx, l, I, W, G, out, lr, mode, batch);
  return pb::conv_weight_update(ctx, l, I, W, G, out, lr, mode, batch);
}

}  // extern "C"
