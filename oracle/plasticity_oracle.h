/* TEST INFRASTRUCTURE — CPU restatement of the reference's layer kernels.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product (plasticity_b200/) never
 * links, imports or executes anything under oracle/.
 *
 * Parity pin: every function here is checked in tests/test_oracle.py against
 *  (1) the reference's own generator output compiled for the host
 *      (oracle/_ref/libref_*.so, built by oracle/Makefile from /root/reference),
 *  (2) the known answers in the reference's nnet/nnet_test.cc (cited per test),
 *  (3) committed fixtures under tests/golden/ produced by (1).
 *
 * Arithmetic is `double`, as in the reference (nnet/nnet.h:27), and keeps the
 * reference's evaluation order and printed constants (2.71828, 2.22045e-16).
 */
#ifndef PLASTICITY_ORACLE_H_
#define PLASTICITY_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

enum { PO_ACT = 0, PO_DENSE = 1, PO_CONV = 2, PO_POOL = 3, PO_SOFTMAX = 4 };
enum { PO_IDENTITY = 0, PO_RELU = 1, PO_LEAKY = 2, PO_SIGMOID = 3 };
enum { PO_MSE = 0, PO_CE = 1 }; /* nnet/layer_dimensions.h:8-11 */

typedef struct {
  int kind;              /* PO_* layer kind */
  int act;               /* PO_ACT: activation function */
  long n;                /* PO_ACT / PO_SOFTMAX: width */
  long in, out;          /* PO_DENSE */
  long W, H, D;          /* PO_CONV / PO_POOL input volume (width,height,depth) */
  long fw, fh, fd;       /* PO_CONV filter width,height,depth */
  long stride, pad, nf;  /* PO_CONV */
  long ow, oh;           /* PO_POOL output area */
} po_layer;

long po_num_inputs(const po_layer *l);
long po_num_outputs(const po_layer *l);
long po_num_weights(const po_layer *l);

/* One kernel == one call; each loops over the reference's work-items. */
void po_evaluate(const po_layer *l, const double *I, const double *W, double *O);
void po_input_delta(const po_layer *l, const double *I, const double *W,
                    const double *G, double *dI);
/* mode 0: out[w] = W[w] - lr*g[w] (new_weights_*);
 * mode 1: out[w] = -lr*g[w]       (weight_deltas_*)  back_prop.kernel.cl:28-44 */
void po_weight_update(const po_layer *l, const double *I, const double *W,
                      const double *G, double lr, int mode, double *out);
void po_error_value(int loss, long n, const double *O, const double *E,
                    double *components);
void po_error_gradients(int loss, long n, const double *O, const double *E,
                        double *gradients);
void po_vector_accumulate(double *a, const double *b, long n);

/* Network drivers; flat buffers exactly like oracle/ref_driver.cc. */
long po_net_total_weights(const po_layer *layers, int n_layers);
long po_net_total_outputs(const po_layer *layers, int n_layers);
void po_net_evaluate(const po_layer *layers, int n_layers, const double *in,
                     const double *weights, double *layer_out);
double po_net_error(const po_layer *layers, int n_layers, int loss,
                    const double *out, const double *expected);
void po_net_train(const po_layer *layers, int n_layers, int loss,
                  const double *in, const double *expected, double *weights,
                  double lr, double *layer_out, double *input_gradients,
                  const double *initial_gradients);
void po_net_train_loop(const po_layer *layers, int n_layers, int loss,
                       const double *ins, const double *expected, long n,
                       double *weights, double lr);
void po_net_batch_train(const po_layer *layers, int n_layers, int loss,
                        const double *ins, const double *expected, long batch,
                        double *weights, double lr, double *sum_deltas);
void po_net_batch_train_mt(const po_layer *layers, int n_layers, int loss,
                           const double *ins, const double *expected, long batch,
                           double *weights, double lr);
/* Batched forward of `batch` independent samples (batched inference). */
void po_net_evaluate_batch(const po_layer *layers, int n_layers,
                           const double *ins, long batch, const double *weights,
                           double *outs);
void po_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif /* PLASTICITY_ORACLE_H_ */
