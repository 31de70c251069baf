/* plasticity_b200 — C ABI of the B200-native layer forward / back-prop /
 * weight-update path of jsharf/plasticity.
 *
 * This header is the drop-in boundary.  Every entry point names the reference
 * interface it replaces (file:line relative to the reference checkout).  The
 * reference reaches its device through cl::Context / cl::CommandQueue /
 * cl::Buffer / cl::Kernel (OpenCL C++ bindings via clutil); a maintainer binds
 * these functions instead.  Signatures are plain C:
 * pointers, sizes, PODs.  No torch, no C++ types.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure;
 *    pb_last_error() returns the message (the reference prints + exit(1),
 *    compute/cl_buffer.h:19-38 — the C++ shim in plasticity_b200/cpp does the
 *    same with this string);
 *  - device memory holds fp32 (`float*` device pointers); host-visible values
 *    are `double`, as in the reference (nnet/nnet.h:27);
 *  - `batch` samples are laid out sample-major: sample b of a layer with n
 *    values starts at element b*n.  batch == 1 is exactly the reference's
 *    one-ClBuffer-per-sample case;
 *  - volumes are planar CHW, idx = W*H*z + row*W + col
 *    (nnet/symbol_generator.cc:22-26); dense weights are row-major
 *    [out][in+1], bias last (nnet/symbol_generator.h:58-86,141-156); conv
 *    weights are [filter][z][row][col] + bias, block stride fw*fh*fd+1
 *    (nnet/symbol_generator.h:229-249,339-354);
 *  - all launches go to the context's stream, asynchronously; reads and
 *    pb_context_finish() synchronise (compute/cl_buffer.cc:64,95 are blocking;
 *    nnet/nnet.h:270 queue->finish()).
 */
#ifndef PLASTICITY_B200_H_
#define PLASTICITY_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB_ABI_VERSION 1

/* ------------------------------------------------------------------------
 * Context = cl::Context + in-order cl::CommandQueue on device 0
 * (nnet/nnet.h:188-197 SelectDevice, :898-907 CompileCl, :498-502
 * MakeCommandQueue; compute/command_queue.h:7-11 CommandQueue::finish).
 * One context owns one CUDA stream.
 * ---------------------------------------------------------------------- */
typedef struct pb_context pb_context;

int pb_abi_version(void);
const char *pb_last_error(void);
int pb_device_count(int *count);
int pb_context_create(int device, pb_context **ctx);
int pb_context_destroy(pb_context *ctx);
int pb_context_finish(pb_context *ctx);            /* CommandQueue::finish() */
int pb_context_device(const pb_context *ctx, int *device);
/* cudaStream_t of the context, as an opaque pointer (for event timing). */
int pb_context_stream(const pb_context *ctx, void **stream);
/* Number of kernels this library launched on the context since creation. */
int pb_context_launch_count(const pb_context *ctx, uint64_t *count);

/* ------------------------------------------------------------------------
 * Buffers = cl::Buffer behind compute::ClBuffer (compute/cl_buffer.h:42-187).
 *   alloc        cl::Buffer(context, CL_MEM_READ_WRITE, n*sizeof)  cl_buffer.cc:108-123
 *   write_f64    enqueueWriteBuffer(CL_TRUE)  MoveToGpu            cl_buffer.cc:124-126
 *   read_f64     enqueueReadBuffer(CL_TRUE)   MoveToCpu            cl_buffer.cc:91-97
 *   copy         enqueueCopyBuffer            DeepClone/copy-ctor  cl_buffer.h:52-65,94-111
 * The f64 variants convert double<->float on the device (host side stays
 * `double` like the reference's std::vector<double> cpu_buffer_).
 * ---------------------------------------------------------------------- */
int pb_buffer_alloc(pb_context *ctx, size_t n, float **dptr);
int pb_buffer_free(pb_context *ctx, float *dptr);
int pb_buffer_write_f64(pb_context *ctx, float *dst, const double *host, size_t n);
int pb_buffer_read_f64(pb_context *ctx, const float *src, double *host, size_t n);
int pb_buffer_write_f32(pb_context *ctx, float *dst, const float *host, size_t n);
int pb_buffer_read_f32(pb_context *ctx, const float *src, float *host, size_t n);
int pb_buffer_copy(pb_context *ctx, float *dst, const float *src, size_t n);
int pb_buffer_fill(pb_context *ctx, float *dst, float value, size_t n);
/* Pinned host staging (cudaHostAlloc) for asynchronous transfers. */
int pb_host_alloc(size_t bytes, void **ptr);
int pb_host_free(void *ptr);
/* Asynchronous variants on the context stream; host memory must be pinned. */
int pb_buffer_write_f32_async(pb_context *ctx, float *dst, const float *host, size_t n);
int pb_buffer_read_f32_async(pb_context *ctx, const float *src, float *host, size_t n);

/* ------------------------------------------------------------------------
 * Layer descriptors (nnet/layer_dimensions.h:13-51).
 * ---------------------------------------------------------------------- */
enum pb_layer_kind {
  PB_ACTIVATION = 0, /* nnet/activation_layer.{h,cc}  */
  PB_DENSE = 1,      /* nnet/dense_layer.{h,cc}       */
  PB_CONVOLUTION = 2,/* nnet/convolution_layer.{h,cc} */
  PB_MAX_POOL = 3,   /* nnet/max_pool_layer.{h,cc}    */
  PB_SOFTMAX = 4     /* nnet/softmax_layer.{h,cc}     */
};

/* symbolic/symbolic_util.h:10-15.  The reference accepts any
 * std::function<Expression(Expression)> (nnet/layer_impl.h:21-22); without a
 * code generator only these four built-ins exist and anything else is
 * refused (no fallback). */
enum pb_activation {
  PB_IDENTITY = 0,
  PB_RELU = 1,
  PB_LEAKY_RELU = 2,
  PB_SIGMOID = 3
};

enum pb_loss { PB_MEAN_SQUARED = 0, PB_CROSS_ENTROPY = 1 }; /* layer_dimensions.h:8-11 */

/* back_prop.kernel.cl:28-44 defines two weight kernels; PB_ACCUMULATE is the
 * fusion of weight_deltas_* with vector_accumulate (combine.kernel.cl:2-5). */
enum pb_update_mode {
  PB_NEW_WEIGHTS = 0,   /* out[w]  = W[w] - lr * sum_b g_b[w]  (new_weights_*)   */
  PB_WEIGHT_DELTAS = 1, /* out[w]  =       - lr * sum_b g_b[w] (weight_deltas_*) */
  PB_ACCUMULATE = 2     /* out[w] +=       - lr * sum_b g_b[w]                   */
};

typedef struct {
  int32_t kind;        /* pb_layer_kind */
  int32_t activation;  /* pb_activation, PB_ACTIVATION only */
  int64_t n;           /* PB_ACTIVATION / PB_SOFTMAX width */
  int64_t in, out;     /* PB_DENSE: Dimensions{num_inputs,num_outputs} */
  int64_t width, height, depth;    /* VolumeDimensions (conv, pool input) */
  int64_t filter_width, filter_height, filter_depth; /* FilterParams */
  int64_t stride, padding, num_filters;              /* FilterParams */
  int64_t out_width, out_height;   /* AreaDimensions (pool output) */
} pb_layer_desc;

/* LayerImpl::GetDimensions / weights().size() (nnet/layer_impl.h:24-43). */
int pb_layer_num_inputs(const pb_layer_desc *l, int64_t *n);
int pb_layer_num_outputs(const pb_layer_desc *l, int64_t *n);
int pb_layer_num_weights(const pb_layer_desc *l, int64_t *n);

/* ------------------------------------------------------------------------
 * Per-kernel launchers: the "operator" interface crossed per launch in the
 * reference (kernel name + NDRange + positional buffers, nnet/nnet.h:243-253,
 * 406-418, 435-447).  `lr` is a 1-element device buffer exactly like
 * learning_rate_buffer_ (nnet/nnet.h:91,441).
 *
 *   pb_layer_evaluate      evaluate_<i>(I, W, O)               evaluate.kernel.cl:8-11
 *   pb_layer_input_delta   input_delta_<type>_<i>(I, W, G, dI) back_prop.kernel.cl:46-54
 *   pb_layer_weight_update new_weights_/weight_deltas_<type>_<i>(I, W, G, out, lr)
 *                                                              back_prop.kernel.cl:28-44
 * With batch > 1 the weight kernels use the gradient SUM over the batch
 * (Nnet::BatchTrain semantics, nnet/nnet.h:624-669).  `W` may be NULL for
 * weightless layers.  For PB_NEW_WEIGHTS `out` may alias `W`.
 * ---------------------------------------------------------------------- */
int pb_layer_evaluate(pb_context *ctx, const pb_layer_desc *l, const float *I,
                      const float *W, float *O, int64_t batch);
int pb_layer_input_delta(pb_context *ctx, const pb_layer_desc *l, const float *I,
                         const float *W, const float *G, float *dI, int64_t batch);
int pb_layer_weight_update(pb_context *ctx, const pb_layer_desc *l, const float *I,
                           const float *W, const float *G, float *out,
                           const float *lr, int mode, int64_t batch);

/*   pb_error_value      error_value(O, E, components)     error.kernel.cl:9-14
 *   pb_error_gradients  error_gradients(O, E, gradients)  error.kernel.cl:16-21
 *   pb_vector_accumulate vector_accumulate(a, b)          combine.kernel.cl:2-5 */
int pb_error_value(pb_context *ctx, int loss, int64_t n, const float *O,
                   const float *E, float *components, int64_t batch);
int pb_error_gradients(pb_context *ctx, int loss, int64_t n, const float *O,
                       const float *E, float *gradients, int64_t batch);
int pb_vector_accumulate(pb_context *ctx, float *a, const float *b, int64_t n);

/* ------------------------------------------------------------------------
 * Network driver = the body of nnet::Nnet (nnet/nnet.h) over the launchers.
 * ---------------------------------------------------------------------- */
typedef struct pb_nnet pb_nnet;

/* Nnet::Nnet(model, init, loss), nnet/nnet.h:52-100.  Weights start at zero
 * (NoWeightInit); Xavier / InitToOne are applied by the host mirror through
 * pb_nnet_set_weights_f64, or on the device by pb_nnet_init_xavier.
 * max_batch sizes the per-layer activation / gradient scratch (the
 * reference's eval_layer_outputs_, backprop_gradients_, :81-97). */
int pb_nnet_create(pb_context *ctx, const pb_layer_desc *layers, int32_t n_layers,
                   int32_t loss, int64_t max_batch, pb_nnet **net);
int pb_nnet_destroy(pb_nnet *net);
int pb_nnet_num_layers(const pb_nnet *net, int32_t *n);
int pb_nnet_total_weights(const pb_nnet *net, int64_t *n);

/* Layer::weight_buffer() MoveToGpu / MoveToCpu, Nnet::GetWeight, Layer::W
 * (nnet/nnet.h:199-202, nnet/layer.h:74-83).  Reference layout. */
int pb_nnet_set_weights_f64(pb_nnet *net, int32_t layer, const double *host, int64_t n);
int pb_nnet_get_weights_f64(pb_nnet *net, int32_t layer, double *host, int64_t n);
/* Device pointer to one layer's weights (reference layout, fp32). */
int pb_nnet_layer_weights_device(pb_nnet *net, int32_t layer, float **dptr);
/* All layers' weights in layer order, each block padded to a multiple of 4
 * floats (pb_nnet_total_weights counts the padding). Device pointer, fp32. */
int pb_nnet_weights_device(pb_nnet *net, float **dptr);
/* Layer::XavierInitializeWeights (nnet/layer.cc:86-100): N(0, 1/num_inputs),
 * generated on the device from `seed` (the reference seeds from
 * random_device, stats/normal.h:11,37 — not reproducible). */
int pb_nnet_init_xavier(pb_nnet *net, uint64_t seed);
int pb_nnet_init_constant(pb_nnet *net, double value); /* InitToOne, nnet.h:852-856 */

/* Nnet::SetLearningParameters, nnet/nnet.h:357-361 */
int pb_nnet_set_learning_rate(pb_nnet *net, double learning_rate);

/* Nnet::Evaluate, nnet/nnet.h:213-273.  d_in: batch*num_inputs floats on the
 * device; d_out (nullable): batch*num_outputs.  Layer outputs stay resident
 * and can be fetched with pb_nnet_layer_output (eval_layer_outputs_). */
int pb_nnet_evaluate(pb_nnet *net, const float *d_in, int64_t batch, float *d_out);
int pb_nnet_layer_output(pb_nnet *net, int32_t layer, float **dptr);

/* Nnet::Error, nnet/nnet.h:280-314: sum of per-output loss components of ONE
 * sample (summed in double on the host like the reference). */
int pb_nnet_error(pb_nnet *net, const float *d_out, const float *d_expected,
                  double *error);
/* Nnet::ErrorGradients, nnet/nnet.h:316-349 */
int pb_nnet_error_gradients(pb_nnet *net, const float *d_out, const float *d_expected,
                            float *d_gradients, int64_t batch);

/* Nnet::Train, nnet/nnet.h:351-355,369-483: ONE per-sample SGD step.
 * d_initial_gradients != NULL selects the 4-argument overload (custom
 * back-prop seed); d_input_gradients (nullable) receives dE/dInput. */
int pb_nnet_train(pb_nnet *net, const float *d_in, const float *d_expected,
                  float *d_input_gradients, const float *d_initial_gradients);
/* The loop of nnet/cifar_test.cc:423-428: n consecutive Train() calls over
 * samples stored sample-major on the device (strict per-sample SGD, CUDA
 * graph replay per sample). */
int pb_nnet_train_loop(pb_nnet *net, const float *d_ins, const float *d_expected,
                       int64_t n_samples);

/* Nnet::BatchTrain, nnet/nnet.h:617-669: W <- W - lr * sum_b grad_b(W_old).
 *   pb_nnet_batch_gradients  CalculateGradients over the batch, :514-615:
 *                            fills the flat delta buffer (all layers
 *                            concatenated) with -lr * sum_b grad_b;
 *   pb_nnet_apply_deltas     VectorAccumulate per layer, :652-665;
 *   pb_nnet_batch_train      both.
 * Data-parallel training all-reduces the delta buffer between the two. */
int pb_nnet_batch_gradients(pb_nnet *net, const float *d_ins, const float *d_expected,
                            int64_t batch);
int pb_nnet_deltas_device(pb_nnet *net, float **dptr);
int pb_nnet_apply_deltas(pb_nnet *net);
int pb_nnet_batch_train(pb_nnet *net, const float *d_ins, const float *d_expected,
                        int64_t batch);

/* ------------------------------------------------------------------------
 * Multi-GPU (no counterpart in the reference, which is single-device:
 * nnet/nnet.h:188-197).  One process per GPU; the id is created on rank 0
 * and distributed by the host (torch.distributed / MPI / a file).
 * ---------------------------------------------------------------------- */
typedef struct pb_comm pb_comm;
#define PB_COMM_ID_BYTES 128
int pb_comm_unique_id(void *id_bytes /* PB_COMM_ID_BYTES */);
int pb_comm_create(pb_context *ctx, const void *id_bytes, int rank, int n_ranks,
                   pb_comm **comm);
/* n_ranks communicators inside ONE process (contexts on the same or on different devices),
 * no NCCL: rank r of the group is comms[r].  The ranks' steps are queued from one host thread,
 * rank after rank, and meet on the device; supports the peer-memory exchange of networks up
 * to 2^20 weights.  What the multi-rank parity test uses on a single-GPU box (ranks that share
 * a device wait for each other ON the device: run with CUDA_DEVICE_MAX_CONNECTIONS >= the
 * number of streams in use so that no two of them share a hardware queue). */
int pb_comm_create_local(pb_context *const *ctxs, int n_ranks, pb_comm **comms /* [n_ranks] */);
int pb_comm_destroy(pb_comm *comm);
int pb_comm_allreduce_sum(pb_comm *comm, float *d_buf, int64_t n);
/* The data-parallel gradient-sum step on this rank's shard: pb_nnet_batch_gradients, the
 * exchange of the summed deltas and the update, W += sum over ranks, bit-identical on every
 * rank.  Fails (and keeps failing) once a peer did not arrive within 5 s. */
int pb_nnet_batch_train_dp(pb_nnet *net, pb_comm *comm, const float *d_ins,
                           const float *d_expected, int64_t local_batch);

/* The same step fed from HOST memory (pinned: pb_host_alloc), asynchronously:
 * the inputs of call i+1 are copied to the device by a second stream while call
 * i computes (two staging slots), and the network outputs of the step (before
 * the update, batch * num_outputs floats) are copied back to h_outputs
 * (nullable).  Returns after enqueuing; pb_context_finish() waits for all.  This
 * is what replaces the per-sample MakeBuffer + MoveToGpu + Train loop of
 * nnet/cifar_test.cc:423-428 for a host-resident data set.  comm != NULL: the
 * data-parallel step (pb_nnet_batch_train_dp) on this rank's shard. */
int pb_nnet_batch_train_host(pb_nnet *net, pb_comm *comm, const float *h_ins,
                             const float *h_expected, int64_t local_batch, float *h_outputs);

/* Labelled byte records (the CIFAR-10 binary format of nnet/cifar_test.cc:191-205: one
 * label byte + sample_size pixel bytes) decoded on the device into network inputs
 * (signed-char pixels / their L2 norm, Sample::NormalizedInput, cifar_test.cc:66-81) and
 * one-hot targets (Sample::OneHotEncodedOutput, :83-97).  Bit-identical to normalising in
 * double on the host and converting to float. */
int pb_records_decode(pb_context *ctx, const uint8_t *d_records, float *d_inputs,
                      float *d_expected, int64_t n_records, int64_t sample_size,
                      int64_t n_classes);
/* pb_nnet_batch_train_host fed with such records from pinned host memory: the data set
 * crosses PCIe as bytes (sample_size + 1 per sample) and is decoded where it is consumed.
 * sample_size = the network's input size, n_classes = its output size. */
int pb_nnet_batch_train_host_records(pb_nnet *net, pb_comm *comm, const uint8_t *h_records,
                                     int64_t local_batch, float *h_outputs);

/* One pb_nnet_batch_train step with a CUDA event after every kernel group; writes
 * "label<TAB>milliseconds<NL>" lines (the kernels the step really runs: fused
 * convolution blocks, the fused head, ...) into `out`.  Measurement only. */
int pb_nnet_profile_step(pb_nnet *net, const float *d_ins, const float *d_expected,
                         int64_t batch, char *out, size_t cap);

/* The same for one batched pb_nnet_evaluate (forward only). */
int pb_nnet_profile_evaluate(pb_nnet *net, const float *d_in, int64_t batch, char *out,
                             size_t cap);

/* ------------------------------------------------------------------------
 * Measurement helpers (bench.py): peak FP32 FFMA throughput and HBM copy
 * bandwidth of this device, measured with CUDA events.
 * ---------------------------------------------------------------------- */
int pb_measure_fp32_tflops(pb_context *ctx, double *tflops);
int pb_measure_copy_gbs(pb_context *ctx, size_t bytes, double *gbs);

#ifdef __cplusplus
}
#endif
#endif /* PLASTICITY_B200_H_ */
