/* TEST INFRASTRUCTURE — see plasticity_oracle.h.  Plain C, double precision.
 *
 * Each function restates what the reference's generated OpenCL kernel computes
 * for ONE work-item, in the same operation order, then loops over work-items
 * (the enqueueNDRangeKernel of nnet/nnet.h:251, :416, :445).  Compile with
 * -ffp-contract=off so a*b+c is not fused behind the reference's back.
 */
#include "plasticity_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* symbolic/numeric_value.cc:41-53 prints NumericValue::e with 6 significant
 * digits, so every generated exp is pow(2.71828, x) (expression.cc:671-673). */
#define PO_E 2.71828
/* std::numeric_limits<double>::epsilon() printed the same way
 * (symbolic_util.cc:40-42). */
#define PO_EPS 2.22045e-16

#define PO_FOR(i, n)                                                        \
  for (long po_n__ = (n), po_go__ = 1; po_go__; po_go__ = 0)                \
    _Pragma("omp parallel for schedule(static) if (po_n__ >= 256)")         \
    for (long i = 0; i < po_n__; ++i)

void po_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---- dimensions -------------------------------------------------------- */

/* nnet/convolution_layer.cc:36-44 */
static long conv_ow(const po_layer *l) {
  return (l->W - l->fw + 2 * l->pad) / l->stride + 1;
}
static long conv_oh(const po_layer *l) {
  return (l->H - l->fh + 2 * l->pad) / l->stride + 1;
}

long po_num_inputs(const po_layer *l) {
  switch (l->kind) {
    case PO_ACT:
    case PO_SOFTMAX: return l->n;
    case PO_DENSE: return l->in;
    case PO_CONV:
    case PO_POOL: return l->W * l->H * l->D;
  }
  return -1;
}

long po_num_outputs(const po_layer *l) {
  switch (l->kind) {
    case PO_ACT:
    case PO_SOFTMAX: return l->n;
    case PO_DENSE: return l->out;
    case PO_CONV: return conv_ow(l) * conv_oh(l) * l->nf;
    case PO_POOL: return l->ow * l->oh * l->D; /* max_pool_layer.h:21-29 */
  }
  return -1;
}

long po_num_weights(const po_layer *l) {
  switch (l->kind) {
    case PO_DENSE: return l->out * (l->in + 1); /* symbol_generator.h:58-62 */
    case PO_CONV:  /* symbol_generator.h:229-232 */
      return l->nf * (l->fw * l->fh * l->fd + 1);
    default: return 0;
  }
}

/* ---- activation functions: symbolic/symbolic_util.cc:10-28 -------------- */

static double act_fwd(int act, double x) {
  switch (act) {
    case PO_RELU: return (x >= 0.0) ? x : 0.0;
    case PO_LEAKY: return (x >= 0.0) ? x : x / 10;
    case PO_SIGMOID: return 1.0 / (1.0 + pow(PO_E, -1.0 * x));
    default: return x;
  }
}

/* Derive() of the above as printed into input_delta_activation_layer_*:
 * relu   -> (x >= 0) ? 1 : 0
 * leaky  -> (x >= 0) ? 1 : (1*10 + (x*0)*-1) / (10*10)
 * sigmoid-> (0*(1+p) + (1*(0 + (1*p)*((-1*1) + x*0)))*-1) / ((1+p)*(1+p)),
 *           p = pow(2.71828, -x); the "1*" is log(|e|) of the un-truncated e
 *           (expression.cc:657-668), i.e. the multiplier is exactly 1. */
static double act_deriv(int act, double x) {
  switch (act) {
    case PO_RELU: return (x >= 0.0) ? 1.0 : 0.0;
    case PO_LEAKY:
      return (x >= 0.0) ? 1.0 : ((1.0 * 10) + ((x * 0.0) * -1)) / (10 * 10);
    case PO_SIGMOID: {
      double p = pow(PO_E, -1.0 * x);
      double num = (0.0 * (1.0 + p)) +
                   ((1.0 * (0.0 + ((1.0 * p) * ((-1.0 * 1.0) + (x * 0.0))))) * -1);
      return num / ((1.0 + p) * (1.0 + p));
    }
    default: return 1.0;
  }
}

/* ---- dense: nnet/dense_layer.cc:12-55 ----------------------------------- */

static void dense_eval(const po_layer *l, const double *I, const double *W,
                       double *O) {
  const long in = l->in, ld = l->in + 1;
  PO_FOR(o, l->out) {
    double acc = W[o * ld + in]; /* bias first, dense_layer.cc:16-18 */
    for (long i = 0; i < in; ++i) acc += W[o * ld + i] * I[i];
    O[o] = acc;
  }
}

static void dense_input_delta(const po_layer *l, const double *W,
                              const double *G, double *dI) {
  const long ld = l->in + 1;
  PO_FOR(i, l->in) {
    double acc = 0.0; /* dense_layer.cc:34-39: one term per output, ascending */
    for (long o = 0; o < l->out; ++o) acc += G[o] * W[o * ld + i];
    dI[i] = acc;
  }
}

static double dense_wgrad(const po_layer *l, const double *I, const double *G,
                          long w) {
  const long ld = l->in + 1;
  const long node = w / ld, edge = w % ld; /* dense_layer.cc:47-48 */
  return (edge >= 0 && !(edge >= l->in)) ? G[node] * I[edge] : G[node];
}

/* ---- volumes: nnet/symbol_generator.cc:22-26 (planar, row-major) -------- */

static double vol_checked(const double *I, long W, long H, long D, long row,
                          long col, long z) {
  /* InputVolumeSymbolGenerator::BoundsCheckedI, symbol_generator.h:199-211 */
  if (z >= 0 && z < D && col >= 0 && col < W && row >= 0 && row < H)
    return I[W * H * z + row * W + col];
  return 0.0;
}

/* ---- convolution: nnet/convolution_layer.cc:50-270 ---------------------- */

static void conv_eval(const po_layer *l, const double *I, const double *W,
                      double *O) {
  const long ow = conv_ow(l), oh = conv_oh(l), plane = ow * oh;
  const long K1 = l->fw * l->fh * l->fd + 1;
  const long hh = l->fh / 2, hw = l->fw / 2;
  PO_FOR(idx, plane * l->nf) {
    const long f = idx / plane, rem = idx - f * plane;
    const long r = rem / ow, c = rem % ow;
    const long conv_row = r * l->stride - l->pad + hh; /* :100-107 */
    const long conv_col = c * l->stride - l->pad + hw;
    double acc = W[f * K1 + (K1 - 1)];
    for (long fx = 0; fx < l->fw; ++fx)     /* loop nest order :82-91 */
      for (long fy = 0; fy < l->fh; ++fy)
        for (long fz = 0; fz < l->fd; ++fz)
          acc += W[f * K1 + l->fw * l->fh * fz + fy * l->fw + fx] *
                 vol_checked(I, l->W, l->H, l->D, conv_row + fy - hh,
                             conv_col + fx - hw, fz);
    O[idx] = acc;
  }
}

static void conv_input_delta(const po_layer *l, const double *W,
                             const double *G, double *dI) {
  const long ow = conv_ow(l), oh = conv_oh(l), oplane = ow * oh;
  const long K1 = l->fw * l->fh * l->fd + 1, s = l->stride, p = l->pad;
  const long hh = l->fh / 2, hw = l->fw / 2, iplane = l->W * l->H;
  PO_FOR(idx, iplane * l->D) {
    const long z = idx / iplane, rem = idx - z * iplane;
    const long r = rem / l->W, c = rem % l->W;
    /* GetOutputCoordinates of the two net corners, :141-160; C division
     * truncates toward zero exactly as in the emitted int arithmetic. */
    const long d_lo = ((r - hh) + p - hh) / s, d_hi = ((r + hh) + p - hh) / s;
    const long k_lo = ((c - hw) + p - hw) / s, k_hi = ((c + hw) + p - hw) / s;
    double grad = 0;
    for (long d = d_lo; d <= d_hi; ++d)
      for (long k = k_lo; k <= k_hi; ++k)
        for (long f = 0; f < l->nf; ++f) {
          const long in_d = d * s - p + hh, in_k = k * s - p + hw;
          /* The reference guards the convolution CENTRE against the input
           * extent (:209), not the output index against the output extent. */
          if (!(in_d >= 0 && in_d < l->H && in_k >= 0 && in_k < l->W)) continue;
          /* The reference reads GRADIENT[...] unchecked here; outside the
           * output plane that is an out-of-bounds / wrong-row read.  Restated
           * as "contributes nothing" — identical wherever the reference is
           * defined (always for pad == filter/2, odd filters). */
          if (!(d >= 0 && d < oh && k >= 0 && k < ow)) continue;
          const long srow = r - in_d + hh, scol = c - in_k + hw;
          double w = 0; /* ConvSymbolGenerator::BoundsCheckedW :306-319 */
          if (z >= 0 && z < l->fd && scol >= 0 && scol < l->fw && srow >= 0 &&
              srow < l->fh)
            w = W[f * K1 + l->fw * l->fh * z + srow * l->fw + scol];
          grad += w * G[oplane * f + d * ow + k];
        }
    dI[idx] = grad;
  }
}

static double conv_wgrad(const po_layer *l, const double *I, const double *G,
                         long w) {
  const long ow = conv_ow(l), oh = conv_oh(l), oplane = ow * oh;
  const long K1 = l->fw * l->fh * l->fd + 1, fplane = l->fw * l->fh;
  const long hh = l->fh / 2, hw = l->fw / 2;
  const long f = w / K1, off = w % K1;
  const int is_bias = (off >= K1 - 1);
  const long wz = off / fplane, wrem = off - wz * fplane;
  const long wy = wrem / l->fw, wx = wrem % l->fw;
  double grad = 0;
  /* :260-265 iterates out_x < output_height (outer), out_y < output_width
   * (inner): swapped bounds, harmless when the output is square.  Restated
   * with the bounds the index algebra needs (cols < ow, rows < oh); equal to
   * the reference for every square output, which is every shipped config. */
  for (long ox = 0; ox < ow; ++ox)
    for (long oy = 0; oy < oh; ++oy) {
      const long in_y = oy * l->stride - l->pad + hh;
      const long in_x = ox * l->stride - l->pad + hw;
      const double in =
          is_bias ? 1.0
                  : vol_checked(I, l->W, l->H, l->D, in_y + wy - hh,
                                in_x + wx - hw, wz);
      grad += G[oplane * f + oy * ow + ox] * in;
    }
  return grad;
}

/* ---- max pool: nnet/max_pool_layer.cc:40-141 ---------------------------- */

static void pool_eval(const po_layer *l, const double *I, double *O) {
  const long gw = l->W / l->ow, gh = l->H / l->oh, oplane = l->ow * l->oh;
  PO_FOR(idx, oplane * l->D) {
    const long z = idx / oplane, rem = idx - z * oplane;
    const long orow = rem / l->ow, ocol = rem % l->ow;
    double max = -INFINITY;
    for (long gr = 0; gr < gh; ++gr)
      for (long gc = 0; gc < gw; ++gc) {
        double v = vol_checked(I, l->W, l->H, l->D, orow * gh + gr,
                               ocol * gw + gc, z);
        if (v > max) max = v; /* strict >: first maximum wins */
      }
    O[idx] = max;
  }
}

static void pool_input_delta(const po_layer *l, const double *I,
                             const double *G, double *dI) {
  const long gw = l->W / l->ow, gh = l->H / l->oh, iplane = l->W * l->H;
  PO_FOR(idx, iplane * l->D) {
    const long z = idx / iplane, rem = idx - z * iplane;
    const long r = rem / l->W, c = rem % l->W;
    const long orow = r / gh, ocol = c / gw;
    double out = 0.0;
    /* Inputs past the last full window (dims that do not divide; the
     * reference only warns, max_pool_layer.cc:26-32) feed no output. */
    if (orow < l->oh && ocol < l->ow) {
      int beaten = 0; /* :126-140: any strictly larger value in the window */
      for (long gr = 0; gr < gh && !beaten; ++gr)
        for (long gc = 0; gc < gw; ++gc)
          if (I[iplane * z + (orow * gh + gr) * l->W + ocol * gw + gc] >
              I[idx]) {
            beaten = 1;
            break;
          }
      if (!beaten) out = G[l->ow * l->oh * z + orow * l->ow + ocol];
    }
    dI[idx] = out;
  }
}

/* ---- softmax: nnet/softmax_layer.cc:9-59 -------------------------------- */

static double softmax_max(const double *I, long n) {
  double max = -INFINITY;
  for (long i = 0; i < n; ++i)
    if (I[i] > max) max = I[i];
  return max;
}

static double softmax_sum(const double *I, long n, double max) {
  double s = 0.0; /* expsum starts at NumericValue(0.0), ascending */
  for (long i = 0; i < n; ++i) s = s + pow(PO_E, I[i] + (max * -1));
  return s;
}

static void softmax_eval(const po_layer *l, const double *I, double *O) {
  const long n = l->n;
  const double max = softmax_max(I, n), s = softmax_sum(I, n, max);
  for (long k = 0; k < n; ++k) O[k] = pow(PO_E, I[k] + (max * -1)) / s;
}

static void softmax_input_delta(const po_layer *l, const double *I,
                                const double *G, double *dI) {
  const long n = l->n;
  const double max = softmax_max(I, n), s = softmax_sum(I, n, max);
  double *O = (double *)malloc(sizeof(double) * (size_t)n);
  for (long k = 0; k < n; ++k) O[k] = pow(PO_E, I[k] + (max * -1)) / s;
  for (long k = 0; k < n; ++k) {
    double acc = 0.0; /* :52-57 */
    for (long i = 0; i < n; ++i)
      acc += (O[i] * (((i == k) ? 1.0 : 0.0) + (O[k] * -1))) * G[i];
    dI[k] = acc;
  }
  free(O);
}

/* ---- public per-kernel entry points ------------------------------------- */

void po_evaluate(const po_layer *l, const double *I, const double *W,
                 double *O) {
  switch (l->kind) {
    case PO_ACT: {
      const int act = l->act;
      PO_FOR(i, l->n) O[i] = act_fwd(act, I[i]);
      break;
    }
    case PO_DENSE: dense_eval(l, I, W, O); break;
    case PO_CONV: conv_eval(l, I, W, O); break;
    case PO_POOL: pool_eval(l, I, O); break;
    case PO_SOFTMAX: softmax_eval(l, I, O); break;
  }
}

void po_input_delta(const po_layer *l, const double *I, const double *W,
                    const double *G, double *dI) {
  switch (l->kind) {
    case PO_ACT: {
      const int act = l->act; /* activation_layer.cc:13-22 */
      PO_FOR(i, l->n) dI[i] = G[i] * act_deriv(act, I[i]);
      break;
    }
    case PO_DENSE: dense_input_delta(l, W, G, dI); break;
    case PO_CONV: conv_input_delta(l, W, G, dI); break;
    case PO_POOL: pool_input_delta(l, I, G, dI); break;
    case PO_SOFTMAX: softmax_input_delta(l, I, G, dI); break;
  }
}

void po_weight_update(const po_layer *l, const double *I, const double *W,
                      const double *G, double lr, int mode, double *out) {
  const long nw = po_num_weights(l);
  if (l->kind == PO_DENSE) {
    PO_FOR(w, nw) {
      const double g = dense_wgrad(l, I, G, w);
      out[w] = mode == 0 ? W[w] - lr * g : -lr * g;
    }
  } else if (l->kind == PO_CONV) {
    PO_FOR(w, nw) {
      const double g = conv_wgrad(l, I, G, w);
      out[w] = mode == 0 ? W[w] - lr * g : -lr * g;
    }
  }
}

/* nnet/error_layer.cc:81-91 and the Derive() text of each. */
void po_error_value(int loss, long n, const double *O, const double *E,
                    double *components) {
  for (long i = 0; i < n; ++i) {
    if (loss == PO_MSE) {
      const double d = E[i] + (O[i] * -1);
      components[i] = (d * d) / 2;
    } else {
      components[i] = -1.0 * (E[i] * (log(O[i] + PO_EPS) / log(PO_E)));
    }
  }
}

void po_error_gradients(int loss, long n, const double *O, const double *E,
                        double *gradients) {
  for (long i = 0; i < n; ++i) {
    if (loss == PO_MSE) {
      /* ((d*-1 + d*-1)*2 + ((d*d)*0)*-1) / (2*2),  d = E-O  ==  O-E */
      const double d = E[i] + (O[i] * -1);
      const double m = d * (0.0 + ((O[i] * 0.0) + (-1 * 1.0)));
      gradients[i] = (((m + m) * 2) + (((d * d) * 0.0) * -1)) / (2 * 2);
    } else {
      /* -1*( E*((1/((O+eps)*1))*(1+0)) + (log(..)/log(e'))*0 ) + (E*log..)*0 */
      const double lg = log(O[i] + PO_EPS) / log(PO_E);
      gradients[i] =
          (-1.0 * ((E[i] * ((1.0 / ((O[i] + PO_EPS) * 1.0)) * (1.0 + 0.0))) +
                   (lg * 0.0))) +
          ((E[i] * lg) * 0.0);
    }
  }
}

void po_vector_accumulate(double *a, const double *b, long n) {
  PO_FOR(i, n) a[i] += b[i]; /* combine.kernel.cl:2-5 */
}

/* ---- network drivers (nnet/nnet.h order) -------------------------------- */

long po_net_total_weights(const po_layer *layers, int n_layers) {
  long t = 0;
  for (int l = 0; l < n_layers; ++l) t += po_num_weights(&layers[l]);
  return t;
}

long po_net_total_outputs(const po_layer *layers, int n_layers) {
  long t = 0;
  for (int l = 0; l < n_layers; ++l) t += po_num_outputs(&layers[l]);
  return t;
}

static long net_max_io(const po_layer *layers, int n_layers) {
  long m = 0;
  for (int l = 0; l < n_layers; ++l) {
    long a = po_num_inputs(&layers[l]), b = po_num_outputs(&layers[l]);
    if (a > m) m = a;
    if (b > m) m = b;
  }
  return m;
}

/* Nnet::Evaluate, nnet.h:234-268 */
void po_net_evaluate(const po_layer *layers, int n_layers, const double *in,
                     const double *weights, double *layer_out) {
  const double *cur = in;
  long woff = 0, ooff = 0;
  for (int l = 0; l < n_layers; ++l) {
    po_evaluate(&layers[l], cur, weights + woff, layer_out + ooff);
    cur = layer_out + ooff;
    woff += po_num_weights(&layers[l]);
    ooff += po_num_outputs(&layers[l]);
  }
}

/* Nnet::Error, nnet.h:280-314: per-output components, host-side ascending sum */
double po_net_error(const po_layer *layers, int n_layers, int loss,
                    const double *out, const double *expected) {
  const long n = po_num_outputs(&layers[n_layers - 1]);
  double *c = (double *)malloc(sizeof(double) * (size_t)n);
  po_error_value(loss, n, out, expected, c);
  double e = 0;
  for (long i = 0; i < n; ++i) e += c[i];
  free(c);
  return e;
}

/* Reverse loop of Nnet::Train, nnet.h:399-460.  mode as po_weight_update;
 * mode 0 updates `weights` in place, mode 1 writes into `deltas`. */
static void net_backward(const po_layer *layers, int n_layers, const double *in,
                         double *weights, const double *layer_out, double *grad,
                         double lr, int mode, double *deltas) {
  const long max_io = net_max_io(layers, n_layers);
  double *next = (double *)malloc(sizeof(double) * (size_t)max_io);
  double *g = grad, *ng = next;
  long woff = po_net_total_weights(layers, n_layers);
  long ooff = po_net_total_outputs(layers, n_layers);
  for (int l = n_layers - 1; l >= 0; --l) {
    const po_layer *L = &layers[l];
    const long nw = po_num_weights(L);
    woff -= nw;
    ooff -= po_num_outputs(L);
    const double *lin = (l > 0) ? layer_out + (ooff - po_num_outputs(&layers[l - 1])) : in;
    /* input gradient first, with the pre-update weights (:406-418) */
    po_input_delta(L, lin, weights + woff, g, ng);
    if (nw > 0) {
      if (mode == 0) {
        double *nwts = (double *)malloc(sizeof(double) * (size_t)nw);
        po_weight_update(L, lin, weights + woff, g, lr, 0, nwts);
        memcpy(weights + woff, nwts, sizeof(double) * (size_t)nw);
        free(nwts);
      } else {
        po_weight_update(L, lin, weights + woff, g, lr, 1, deltas + woff);
      }
    }
    double *t = g; g = ng; ng = t;
  }
  if (g != grad)
    memcpy(grad, g, sizeof(double) * (size_t)po_num_inputs(&layers[0]));
  free(next);
}

void po_net_train(const po_layer *layers, int n_layers, int loss,
                  const double *in, const double *expected, double *weights,
                  double lr, double *layer_out, double *input_gradients,
                  const double *initial_gradients) {
  const long max_io = net_max_io(layers, n_layers);
  const long nout = po_num_outputs(&layers[n_layers - 1]);
  double *grad = (double *)calloc((size_t)max_io, sizeof(double));
  po_net_evaluate(layers, n_layers, in, weights, layer_out);
  if (initial_gradients) {
    memcpy(grad, initial_gradients, sizeof(double) * (size_t)nout);
  } else {
    const long last = po_net_total_outputs(layers, n_layers) - nout;
    po_error_gradients(loss, nout, layer_out + last, expected, grad);
  }
  net_backward(layers, n_layers, in, weights, layer_out, grad, lr, 0, NULL);
  if (input_gradients)
    memcpy(input_gradients, grad,
           sizeof(double) * (size_t)po_num_inputs(&layers[0]));
  free(grad);
}

void po_net_train_loop(const po_layer *layers, int n_layers, int loss,
                       const double *ins, const double *expected, long n,
                       double *weights, double lr) {
  const long nin = po_num_inputs(&layers[0]);
  const long nout = po_num_outputs(&layers[n_layers - 1]);
  double *lo = (double *)malloc(
      sizeof(double) * (size_t)po_net_total_outputs(layers, n_layers));
  for (long b = 0; b < n; ++b)
    po_net_train(layers, n_layers, loss, ins + b * nin, expected + b * nout,
                 weights, lr, lo, NULL, NULL);
  free(lo);
}

/* Intended Nnet::BatchTrain (nnet.h:624-669): deltas at the batch-start
 * weights, accumulated example by example. */
void po_net_batch_train(const po_layer *layers, int n_layers, int loss,
                        const double *ins, const double *expected, long batch,
                        double *weights, double lr, double *sum_deltas) {
  const long nin = po_num_inputs(&layers[0]);
  const long nout = po_num_outputs(&layers[n_layers - 1]);
  const long tw = po_net_total_weights(layers, n_layers);
  const long to = po_net_total_outputs(layers, n_layers);
  const long max_io = net_max_io(layers, n_layers);
  double *frozen = (double *)malloc(sizeof(double) * (size_t)tw);
  double *delta = (double *)calloc((size_t)tw, sizeof(double));
  double *lo = (double *)malloc(sizeof(double) * (size_t)to);
  double *grad = (double *)malloc(sizeof(double) * (size_t)max_io);
  memcpy(frozen, weights, sizeof(double) * (size_t)tw);
  if (sum_deltas) memset(sum_deltas, 0, sizeof(double) * (size_t)tw);
  for (long b = 0; b < batch; ++b) {
    po_net_evaluate(layers, n_layers, ins + b * nin, frozen, lo);
    po_error_gradients(loss, nout, lo + (to - nout), expected + b * nout, grad);
    net_backward(layers, n_layers, ins + b * nin, frozen, lo, grad, lr, 1,
                 delta);
    po_vector_accumulate(weights, delta, tw);
    if (sum_deltas) po_vector_accumulate(sum_deltas, delta, tw);
  }
  free(frozen); free(delta); free(lo); free(grad);
}

/* po_net_batch_train with examples spread over host threads (see
 * oracle/ref_driver.cc: ref_net_batch_train_mt). CPU-baseline timing only. */
void po_net_batch_train_mt(const po_layer *layers, int n_layers, int loss,
                           const double *ins, const double *expected, long batch,
                           double *weights, double lr) {
  const long nin = po_num_inputs(&layers[0]);
  const long nout = po_num_outputs(&layers[n_layers - 1]);
  const long tw = po_net_total_weights(layers, n_layers);
  const long to = po_net_total_outputs(layers, n_layers);
  const long max_io = net_max_io(layers, n_layers);
  double *frozen = (double *)malloc(sizeof(double) * (size_t)tw);
  double *total = (double *)calloc((size_t)tw, sizeof(double));
  memcpy(frozen, weights, sizeof(double) * (size_t)tw);
#pragma omp parallel
  {
    double *delta = (double *)calloc((size_t)tw, sizeof(double));
    double *sum = (double *)calloc((size_t)tw, sizeof(double));
    double *lo = (double *)malloc(sizeof(double) * (size_t)to);
    double *grad = (double *)malloc(sizeof(double) * (size_t)max_io);
#pragma omp for schedule(dynamic, 1)
    for (long b = 0; b < batch; ++b) {
      po_net_evaluate(layers, n_layers, ins + b * nin, frozen, lo);
      po_error_gradients(loss, nout, lo + (to - nout), expected + b * nout, grad);
      net_backward(layers, n_layers, ins + b * nin, frozen, lo, grad, lr, 1, delta);
      for (long i = 0; i < tw; ++i) sum[i] += delta[i];
    }
#pragma omp critical
    for (long i = 0; i < tw; ++i) total[i] += sum[i];
    free(delta); free(sum); free(lo); free(grad);
  }
  for (long i = 0; i < tw; ++i) weights[i] += total[i];
  free(frozen); free(total);
}

void po_net_evaluate_batch(const po_layer *layers, int n_layers,
                           const double *ins, long batch, const double *weights,
                           double *outs) {
  const long nin = po_num_inputs(&layers[0]);
  const long nout = po_num_outputs(&layers[n_layers - 1]);
  const long to = po_net_total_outputs(layers, n_layers);
  double *lo = (double *)malloc(sizeof(double) * (size_t)to);
  for (long b = 0; b < batch; ++b) {
    po_net_evaluate(layers, n_layers, ins + b * nin, weights, lo);
    memcpy(outs + b * nout, lo + (to - nout), sizeof(double) * (size_t)nout);
  }
  free(lo);
}
