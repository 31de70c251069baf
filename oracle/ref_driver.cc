// TEST INFRASTRUCTURE — not part of the product path.
//
// Host loop that plays nnet::Nnet's role over the reference-generated kernels
// (see ref_gen.cc): same launch order as Nnet::Evaluate (nnet/nnet.h:234-268),
// Nnet::Train (nnet/nnet.h:369-483), Nnet::Error (:280-314) and the intended
// semantics of Nnet::BatchTrain (:624-669; W += sum over the batch of
// weight_deltas computed at the batch-start weights, example order).
// All buffers are flat double arrays: weights = layers' weight vectors
// concatenated in layer order, layer_out = layers' outputs concatenated.
#include <cstdlib>
#include <cstring>
#include <vector>

#ifdef REF_FP32  // fp32 variant of the checker (ref_shim.h): every buffer is float
#define double float
#endif

extern "C" {
int ref_parallel_threshold = 256;

int ref_num_layers(void);
int ref_loss(void);
void ref_layer_info(int l, long *ni, long *no, long *nw, const char **name);
void ref_evaluate(int l, double *I, double *W, double *O);
void ref_input_delta(int l, double *I, double *W, double *G, double *dI);
void ref_new_weights(int l, double *I, double *W, double *G, double *Wn,
                     double *lr);
void ref_weight_deltas(int l, double *I, double *W, double *G, double *dW,
                       double *lr);
void ref_error_value(double *O, double *E, double *c);
void ref_error_gradients(double *O, double *E, double *g);
void ref_vector_accumulate(double *a, double *b, long n);
}

namespace {

struct Shape {
  int layers;
  std::vector<long> ni, no, nw, w_off, o_off;
  long total_w = 0, total_o = 0, max_io = 0;
  Shape() {
    layers = ref_num_layers();
    ni.resize(layers), no.resize(layers), nw.resize(layers);
    w_off.resize(layers), o_off.resize(layers);
    for (int l = 0; l < layers; ++l) {
      const char *name;
      ref_layer_info(l, &ni[l], &no[l], &nw[l], &name);
      w_off[l] = total_w;
      o_off[l] = total_o;
      total_w += nw[l];
      total_o += no[l];
      if (ni[l] > max_io) max_io = ni[l];
      if (no[l] > max_io) max_io = no[l];
    }
  }
};

const Shape &shape() {
  static Shape s;
  return s;
}

void Forward(const double *in, double *weights, double *layer_out) {
  const Shape &s = shape();
  double *cur = const_cast<double *>(in);
  for (int l = 0; l < s.layers; ++l) {
    double *o = layer_out + s.o_off[l];
    ref_evaluate(l, cur, weights + s.w_off[l], o);
    cur = o;
  }
}

// Reverse loop of Nnet::Train. grad_io: on entry dE/dO of the last layer, on
// exit dE/dI of layer 0 (first num_inputs(0) entries).  update_mode 0 writes
// new weights in place (via a scratch copy, like the DeepClone + handle swap
// at nnet.h:432-454); update_mode 1 writes weight_deltas into `deltas`.
void Backward(const double *in, double *weights, double *layer_out,
              double *grad_io, double lr, int update_mode, double *deltas) {
  const Shape &s = shape();
  std::vector<double> next(s.max_io), scratch;
  double *g = grad_io;
  double *ng = next.data();
  double lr_buf[1] = {lr};
  for (int l = s.layers - 1; l >= 0; --l) {
    double *layer_in = (l > 0) ? layer_out + s.o_off[l - 1]
                               : const_cast<double *>(in);
    double *W = weights + s.w_off[l];
    ref_input_delta(l, layer_in, W, g, ng);
    if (s.nw[l] > 0) {
      if (update_mode == 0) {
        scratch.assign(W, W + s.nw[l]);
        ref_new_weights(l, layer_in, W, g, scratch.data(), lr_buf);
        std::memcpy(W, scratch.data(), sizeof(double) * s.nw[l]);
      } else {
        ref_weight_deltas(l, layer_in, W, g, deltas + s.w_off[l], lr_buf);
      }
    }
    double *t = g;
    g = ng;
    ng = t;
  }
  if (g != grad_io) std::memcpy(grad_io, g, sizeof(double) * s.ni[0]);
}

}  // namespace

extern "C" {

long ref_total_weights(void) { return shape().total_w; }
long ref_total_outputs(void) { return shape().total_o; }
long ref_max_io(void) { return shape().max_io; }

void ref_net_evaluate(const double *in, double *weights, double *layer_out) {
  Forward(in, weights, layer_out);
}

double ref_net_error(const double *out, const double *expected) {
  const Shape &s = shape();
  const long n = s.no[s.layers - 1];
  std::vector<double> comps(n);
  ref_error_value(const_cast<double *>(out), const_cast<double *>(expected),
                  comps.data());
  double e = 0;  // host-side ascending sum, nnet.h:307-311
  for (long i = 0; i < n; ++i) e += comps[i];
  return e;
}

// One Nnet::Train call.  initial_gradients != NULL selects the 4-argument
// overload (nnet.h:369); otherwise error_gradients seeds back-prop (:481).
// input_gradients (nullable) receives dE/dI of layer 0.  forwards = 2
// reproduces the as-shipped double forward pass (:384 and :477); the second
// pass recomputes identical values, so results do not depend on it.
void ref_net_train(const double *in, const double *expected, double *weights,
                   double lr, double *layer_out, double *input_gradients,
                   const double *initial_gradients, int forwards) {
  const Shape &s = shape();
  std::vector<double> grad(s.max_io, 0.0);
  for (int f = 0; f < (forwards < 1 ? 1 : forwards); ++f)
    Forward(in, weights, layer_out);
  const long nout = s.no[s.layers - 1];
  if (initial_gradients) {
    std::memcpy(grad.data(), initial_gradients, sizeof(double) * nout);
  } else {
    ref_error_gradients(layer_out + s.o_off[s.layers - 1],
                        const_cast<double *>(expected), grad.data());
  }
  Backward(in, weights, layer_out, grad.data(), lr, 0, nullptr);
  if (input_gradients)
    std::memcpy(input_gradients, grad.data(), sizeof(double) * s.ni[0]);
}

// n consecutive per-sample SGD steps (the loop of cifar_test.cc:423-428).
void ref_net_train_loop(const double *ins, const double *expected, long n,
                        double *weights, double lr, int forwards) {
  const Shape &s = shape();
  std::vector<double> layer_out(s.total_o);
  const long nin = s.ni[0], nout = s.no[s.layers - 1];
  for (long b = 0; b < n; ++b)
    ref_net_train(ins + b * nin, expected + b * nout, weights, lr,
                  layer_out.data(), nullptr, nullptr, forwards);
}

// Intended Nnet::BatchTrain semantics (nnet.h:624-669): every example's
// weight_deltas (= -lr * gradient) are computed at the batch-start weights,
// then accumulated into the weights example by example with
// vector_accumulate.  sum_deltas (nullable, total_weights) receives the sum.
void ref_net_batch_train(const double *ins, const double *expected, long batch,
                         double *weights, double lr, double *sum_deltas) {
  const Shape &s = shape();
  const long nin = s.ni[0], nout = s.no[s.layers - 1];
  std::vector<double> layer_out(s.total_o), grad(s.max_io);
  std::vector<double> frozen(weights, weights + s.total_w), delta(s.total_w, 0.0);
  if (sum_deltas) std::memset(sum_deltas, 0, sizeof(double) * s.total_w);
  for (long b = 0; b < batch; ++b) {
    Forward(ins + b * nin, frozen.data(), layer_out.data());
    ref_error_gradients(layer_out.data() + s.o_off[s.layers - 1],
                        const_cast<double *>(expected + b * nout), grad.data());
    Backward(ins + b * nin, frozen.data(), layer_out.data(), grad.data(), lr, 1,
             delta.data());
    for (int l = 0; l < s.layers; ++l) {
      if (s.nw[l] == 0) continue;
      ref_vector_accumulate(weights + s.w_off[l], delta.data() + s.w_off[l],
                            s.nw[l]);
      if (sum_deltas)
        ref_vector_accumulate(sum_deltas + s.w_off[l], delta.data() + s.w_off[l],
                              s.nw[l]);
    }
  }
}


// Same semantics as ref_net_batch_train with the examples spread over host
// threads — the reference's BatchTrain gives every example its own command
// queue (nnet.h:639-648), so examples are the natural unit of host
// parallelism.  Work-items inside a kernel then run serially on the calling
// thread (nested OpenMP regions are inactive).  Per-thread delta sums are
// combined at the end, so the summation order differs from example order by
// rounding only.  Used for CPU-baseline timing.
void ref_net_batch_train_mt(const double *ins, const double *expected, long batch,
                            double *weights, double lr) {
  const Shape &s = shape();
  const long nin = s.ni[0], nout = s.no[s.layers - 1];
  std::vector<double> frozen(weights, weights + s.total_w);
  std::vector<double> total(s.total_w, 0.0);
#pragma omp parallel
  {
    std::vector<double> layer_out(s.total_o), grad(s.max_io), delta(s.total_w, 0.0),
        sum(s.total_w, 0.0);
#pragma omp for schedule(dynamic, 1)
    for (long b = 0; b < batch; ++b) {
      Forward(ins + b * nin, frozen.data(), layer_out.data());
      ref_error_gradients(layer_out.data() + s.o_off[s.layers - 1],
                          const_cast<double *>(expected + b * nout), grad.data());
      Backward(ins + b * nin, frozen.data(), layer_out.data(), grad.data(), lr, 1,
               delta.data());
      for (long i = 0; i < s.total_w; ++i) sum[i] += delta[i];
    }
#pragma omp critical
    for (long i = 0; i < s.total_w; ++i) total[i] += sum[i];
  }
  for (long i = 0; i < s.total_w; ++i) weights[i] += total[i];
}

}  // extern "C"
