// Own test of the C++ host mirror (plasticity_b200/cpp): drives the reference's
// public API shape — Architecture builder, Nnet, ClBuffer, WeightNumber, the
// JSON weight file — on cuda:0 and checks the known answers the reference's
// nnet/nnet_test.cc pins (cited per case).  Exit code 0 = all checks passed.
#include <cmath>
#include <cstdio>
#include <set>

#include "compute/cuda_command_queue.h"
#include "nnet/nnet.h"
#include "nnet/yolov1.h"

static int failures = 0;
#define EXPECT_NEAR(a, b, eps)                                                            \
  do {                                                                                    \
    const double a__ = (a), b__ = (b);                                                    \
    if (!(std::fabs(a__ - b__) <= (eps) * std::max(1.0, std::fabs(b__)))) {               \
      std::printf("FAIL %s:%d  %s = %.9g, want %.9g\n", __FILE__, __LINE__, #a, a__, b__); \
      ++failures;                                                                         \
    }                                                                                     \
  } while (0)
#define EXPECT_TRUE(c)                                                    \
  do {                                                                    \
    if (!(c)) {                                                           \
      std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #c);            \
      ++failures;                                                         \
    }                                                                     \
  } while (0)

using namespace nnet;

// dense 3->3 + ReLU, nnet_test.cc:17-70
static void OneLayerRelu() {
  Architecture model(3);
  model.AddDenseLayer(3, symbolic::Relu);
  Nnet net(model, Nnet::NoWeightInit, CrossEntropy);
  DenseSymbolGenerator s(Dimensions{3, 3});
  const double w[3][3] = {{0.1, 0.3, 0.4}, {0.2, 0.2, 0.3}, {0.3, 0.7, 0.9}};
  for (int o = 0; o < 3; ++o) {
    for (int i = 0; i < 3; ++i) net.GetWeight(1, s.WeightNumber(o, i)) = w[o][i];
    net.GetWeight(1, s.WeightNumber(o)) = 1;
  }
  auto in = net.MakeBuffer({0.1, 0.2, 0.7});
  auto out = net.Evaluate(in);
  out->MoveToCpu();
  EXPECT_TRUE(out->size() == 3);
  EXPECT_NEAR(out->at(0), 1.35, 1e-3);
  EXPECT_NEAR(out->at(1), 1.27, 1e-3);
  EXPECT_NEAR(out->at(2), 1.8, 1e-3);
}

// 2-2-2 sigmoid MSE, one SGD step at lr 0.5, nnet_test.cc:194-274
static void MazurStep() {
  Architecture model(2);
  model.AddDenseLayer(2, symbolic::Sigmoid);
  model.AddDenseLayer(2, symbolic::Sigmoid);
  DenseSymbolGenerator s(Dimensions{2, 2});
  model.layers[1].W(s.WeightNumber(0, 0)) = 0.15; model.layers[1].W(s.WeightNumber(0, 1)) = 0.2;
  model.layers[1].W(s.WeightNumber(0)) = 0.35;
  model.layers[1].W(s.WeightNumber(1, 0)) = 0.25; model.layers[1].W(s.WeightNumber(1, 1)) = 0.3;
  model.layers[1].W(s.WeightNumber(1)) = 0.35;
  model.layers[3].W(s.WeightNumber(0, 0)) = 0.4; model.layers[3].W(s.WeightNumber(0, 1)) = 0.45;
  model.layers[3].W(s.WeightNumber(0)) = 0.6;
  model.layers[3].W(s.WeightNumber(1, 0)) = 0.5; model.layers[3].W(s.WeightNumber(1, 1)) = 0.55;
  model.layers[3].W(s.WeightNumber(1)) = 0.6;
  Nnet net(model, Nnet::NoWeightInit, MeanSquared);
  auto in = net.MakeBuffer({0.05, 0.10});
  auto expected = net.MakeBuffer({0.01, 0.99});
  auto out = net.Evaluate(in);
  EXPECT_NEAR(net.Error(out, expected), 0.298371109, 1e-5);
  out->MoveToCpu();
  EXPECT_NEAR(out->at(0), 0.75136507, 1e-5);
  EXPECT_NEAR(out->at(1), 0.772928465, 1e-5);
  net.SetLearningParameters(Nnet::LearningParameters{0.5});
  net.Train(in, expected);
  Architecture &m = net.model();
  EXPECT_NEAR(m.layers[1].W(s.WeightNumber(0, 0)), 0.149780716, 2e-5);
  EXPECT_NEAR(m.layers[1].W(s.WeightNumber(1, 1)), 0.29950229, 2e-5);
  EXPECT_NEAR(m.layers[3].W(s.WeightNumber(0, 0)), 0.35891648, 2e-5);
  EXPECT_NEAR(m.layers[3].W(s.WeightNumber(1, 1)), 0.561370121, 2e-5);
}

// conv weight layout, nnet_test.cc:486-581
static void ConvLayout() {
  ConvSymbolGenerator s(VolumeDimensions{5, 5, 3}, FilterParams{3, 3, 3, 2, 1, 2});
  EXPECT_TRUE(s.WeightNumber(0, 0, 0, 0) == 0);
  EXPECT_TRUE(s.WeightNumber(0, 2, 2, 2) == 26);
  EXPECT_TRUE(s.WeightNumber(0) == 27);
  EXPECT_TRUE(s.WeightNumber(1, 0, 0, 0) == 28);
  EXPECT_TRUE(s.WeightNumber(1) == 55);
}

// JSON round trip, nnet_test.cc:1159-1215; BatchTrain + per-sample loop smoke.
static void SaveLoadAndBatch() {
  Architecture model(3);
  model.AddDenseLayer(9, symbolic::Relu);
  model.AddConvolutionLayer({3, 3, 1}, {2, 2, 1, 1, 0, 2});
  model.AddDenseLayer(9, symbolic::Relu);
  model.AddSoftmaxLayer(9);
  Nnet a(model, Nnet::Xavier, CrossEntropy, 8);
  const std::string saved = a.WeightsToString();
  Nnet b(model, Nnet::NoWeightInit, CrossEntropy, 8);
  EXPECT_TRUE(b.LoadWeightsFromString(saved));
  EXPECT_TRUE(b.WeightsToString() == saved);
  EXPECT_TRUE(!b.LoadWeightsFromString("{\"version\":\"nope\"}"));
  // the same minibatch step on both copies gives the same weights
  std::vector<std::unique_ptr<compute::ClBuffer>> ins, outs;
  for (int k = 0; k < 6; ++k) {
    ins.push_back(a.MakeBuffer({0.1 * k, -0.2 + 0.05 * k, 0.3}));
    std::vector<double> e(9, 0.0);
    e[k % 9] = 1.0;
    outs.push_back(a.MakeBuffer(e));
  }
  a.SetLearningParameters(Nnet::LearningParameters{0.01});
  b.SetLearningParameters(Nnet::LearningParameters{0.01});
  a.BatchTrain(ins, outs, std::set<int>{0, 1, 2, 3, 4, 5});
  b.BatchTrain(ins, outs, std::set<int>{0, 1, 2, 3, 4, 5});
  EXPECT_TRUE(a.WeightsToString() == b.WeightsToString());
  EXPECT_TRUE(a.WeightsToString() != saved);
}

// nnet/yolov1.{h,cc}: the repaired layer list chains at the reference's size, and a
// scaled-down instance evaluates a batch, trains one sample and moves its outputs.
static void Yolo() {
  Architecture full = YoloArchitecture();
  EXPECT_TRUE(full.VerifyArchitecture());
  EXPECT_TRUE(full.layers.size() == 55);
  EXPECT_TRUE(full.input_size() == 448 * 448 * 3 && full.output_size() == 1470);
  size_t params = 0;
  for (auto &l : full.layers) params += l.weight_buffer().size();
  EXPECT_TRUE(params == 266721790);
  Yolov1 yolo(/*max_batch=*/4, /*input_width=*/64, /*dense_width=*/64, /*outputs=*/30);
  EXPECT_TRUE(yolo.SetParams(0.01));
  std::vector<double> px(4 * 64 * 64 * 3);
  for (size_t i = 0; i < px.size(); ++i) px[i] = (double)((i * 2654435761u) % 1000) / 1000.0;
  auto images = yolo.net().MakeBuffer(px);
  auto out = yolo.Evaluate(images);
  out->MoveToCpu();
  EXPECT_TRUE(out->size() == 4 * 30);
  for (size_t i = 0; i < out->size(); ++i) EXPECT_TRUE(out->at(i) > 0.0 && out->at(i) < 1.0);  // sigmoid
  std::vector<double> one(px.begin(), px.begin() + 64 * 64 * 3), target(30, 0.25);
  auto image = yolo.net().MakeBuffer(one);
  auto expected = yolo.net().MakeBuffer(target);
  auto before = yolo.Evaluate(image);
  const double e0 = yolo.net().Error(before, expected);
  for (int k = 0; k < 5; ++k) yolo.TrainSample(image, expected);
  auto after = yolo.Evaluate(image);
  EXPECT_TRUE(yolo.net().Error(after, expected) < e0);
}

// nnet/layer_dimensions.cc:12-54
static void WorkgroupSizes() {
  EXPECT_TRUE(CalculateWorkgroupSize(0) == 1);
  EXPECT_TRUE(CalculateWorkgroupSize(100) == 100);
  EXPECT_TRUE(CalculateWorkgroupSize(128) == 128);
  EXPECT_TRUE(CalculateWorkgroupSize(16384) == 32);   // one warp per group on powers of two
  EXPECT_TRUE(CalculateWorkgroupSize(129) == 43);     // smallest divisor >= 32
  EXPECT_TRUE(CalculateWorkgroupSize(131) == 1);      // prime
  EXPECT_TRUE(CalculateWorkgroupSize(192) == 32);
  EXPECT_TRUE(CalculateWorkgroupSize(2 * 97) == 1);   // 194: no divisor in [32, 97)
  Architecture model(3072);
  model.AddConvolutionLayer({32, 32, 3}, {5, 5, 3, 1, 2, 16});
  EXPECT_TRUE(model.layers[1].eval_workgroup_size() == 32);           // 16384 outputs
  EXPECT_TRUE(model.layers[1].weight_train_workgroup_size() == 32);   // 1216 = 32 * 38
  EXPECT_TRUE(model.layers[1].bp_train_workgroup_size() == 32);       // 3072 inputs
}

// compute::CommandQueue (compute/command_queue.h:7-11) over CUDA: the reference's kernel
// names + positional buffers reproduce Nnet::Evaluate and one Nnet::Train step of the
// 2-2-2 sigmoid / MSE network (nnet_test.cc:194-274), launch by launch in nnet.h's order.
static void KernelQueue() {
  Architecture model(2);
  model.AddDenseLayer(2, symbolic::Sigmoid);
  model.AddDenseLayer(2, symbolic::Sigmoid);
  DenseSymbolGenerator s(Dimensions{2, 2});
  const double w1[2][3] = {{0.15, 0.2, 0.35}, {0.25, 0.3, 0.35}};
  const double w3[2][3] = {{0.4, 0.45, 0.6}, {0.5, 0.55, 0.6}};
  for (int o = 0; o < 2; ++o) {
    for (int i = 0; i < 2; ++i) {
      model.layers[1].W(s.WeightNumber(o, i)) = w1[o][i];
      model.layers[3].W(s.WeightNumber(o, i)) = w3[o][i];
    }
    model.layers[1].W(s.WeightNumber(o)) = w1[o][2];
    model.layers[3].W(s.WeightNumber(o)) = w3[o][2];
  }
  Nnet net(model, Nnet::NoWeightInit, MeanSquared);
  net.SetLearningParameters(Nnet::LearningParameters{0.5});
  pb_context *ctx = net.context();
  std::vector<pb_layer_desc> descs;
  std::vector<std::string> types;
  for (auto &l : model.layers) {
    descs.push_back(l.desc());
    types.push_back(l.layer_type());
  }
  compute::CudaCommandQueue queue(ctx, descs, types, PB_MEAN_SQUARED);
  const size_t n = model.layers.size();
  // per-layer device buffers: outputs, weights, gradients
  std::vector<compute::ClBuffer> out(n), W(n);
  compute::ClBuffer in(std::vector<double>{0.05, 0.1}), expected(std::vector<double>{0.01, 0.99});
  compute::ClBuffer lr(std::vector<double>{0.5});
  for (compute::ClBuffer *b : {&in, &expected, &lr}) b->RegisterClBackend(ctx);
  for (size_t l = 0; l < n; ++l) {
    out[l] = compute::ClBuffer(model.layers[l].GetDimensions().num_outputs, ctx);
    out[l].MoveToGpu();
    model.layers[l].weight_buffer().MoveToCpu();
    W[l] = compute::ClBuffer(model.layers[l].weight_buffer());  // private copy of the weights
    W[l].RegisterClBackend(ctx);
  }
  auto args = [](std::initializer_list<compute::ClBuffer *> bufs) {
    auto v = std::make_unique<std::vector<compute::ClBuffer>>(bufs.size());
    size_t k = 0;
    for (compute::ClBuffer *b : bufs) (*v)[k++] = *b;  // operator= aliases the device handle
    return v;
  };
  // forward, nnet.h:234-268
  for (size_t l = 0; l < n; ++l) {
    compute::ClBuffer &I = l ? out[l - 1] : in;
    queue.enqueueKernel(model.layers[l].EvaluateKernelName(),
                        (int)model.layers[l].GetDimensions().num_outputs,
                        (int)model.layers[l].eval_workgroup_size(), args({&I, &W[l], &out[l]}));
  }
  queue.finish();
  auto net_in = net.MakeBuffer({0.05, 0.1});
  auto net_expected = net.MakeBuffer({0.01, 0.99});
  auto want = net.Evaluate(net_in);
  want->MoveToCpu();
  compute::ClBuffer got(out[n - 1]);  // device copy, then read back
  got.MoveToCpu();
  EXPECT_NEAR(got[0], want->at(0), 1e-6);
  EXPECT_NEAR(got[1], want->at(1), 1e-6);
  EXPECT_NEAR(got[0], 0.75136507, 1e-5);
  // error gradients + reverse loop, nnet.h:316-349, 399-460
  compute::ClBuffer g(2, ctx), ng(2, ctx);
  g.MoveToGpu();
  ng.MoveToGpu();
  queue.enqueueKernel("error_gradients", 2, 2, args({&out[n - 1], &expected, &g}));
  for (size_t l = n; l-- > 0;) {
    Layer &L = model.layers[l];
    compute::ClBuffer &I = l ? out[l - 1] : in;
    queue.enqueueKernel(L.InputGradientKernelName(), (int)L.GetDimensions().num_inputs,
                        (int)L.bp_train_workgroup_size(), args({&I, &W[l], &g, &ng}));
    if (W[l].size() > 0) {
      compute::ClBuffer new_w(W[l].size(), ctx);
      new_w.MoveToGpu();
      queue.enqueueKernel(L.WeightUpdateKernelName(), (int)W[l].size(),
                          (int)L.weight_train_workgroup_size(), args({&I, &W[l], &g, &new_w, &lr}));
      W[l] = new_w;  // handle swap, nnet.h:454
    }
    compute::ClBuffer t;
    t = g; g = ng; ng = t;
  }
  queue.finish();
  net.Train(net_in, net_expected);
  for (size_t l : {size_t(1), size_t(3)}) {
    compute::ClBuffer mine(W[l]);
    mine.MoveToCpu();
    for (size_t k = 0; k < mine.size(); ++k) EXPECT_NEAR(mine[k], net.GetWeight(l, k), 1e-6);
  }
  compute::ClBuffer w3_new(W[3]);
  w3_new.MoveToCpu();
  EXPECT_NEAR(w3_new[s.WeightNumber(0, 0)], 0.35891648, 1e-5);  // nnet_test.cc:260-273
}

// The cases of the reference's compute/cl_buffer_test.cc:48-227, restated for the mirror:
// the cl::Buffer a caller assigns through `*buf.gpu_buffer() = cl_buffer` becomes a device
// allocation the buffer adopts (operator= aliases the handle, cl_buffer.h:166-179).
static void ClBufferCases() {
  pb_context *ctx = nullptr;
  PB_CHECK_OK(pb_context_create(0, &ctx));
  auto fresh = [&] {
    compute::ClBuffer b(5, ctx);
    for (size_t i = 0; i < 5; ++i) b[i] = i * 2 + 1;
    return b;
  };
  auto device_values = [&](double scale) {  // 5 doubles written straight to a device buffer
    float *p = nullptr;
    PB_CHECK_OK(pb_buffer_alloc(ctx, 5, &p));
    double v[5];
    for (size_t i = 0; i < 5; ++i) v[i] = i * scale;
    PB_CHECK_OK(pb_buffer_write_f64(ctx, p, v, 5));
    return compute::ClBuffer(ctx, p, 5);
  };
  {  // "CPU-only test" :48-63
    compute::ClBuffer buf = fresh();
    for (size_t i = 0; i < 5; ++i) EXPECT_TRUE(buf[i] == i * 2 + 1);
  }
  {  // "Moving to opencl and back doesnt kill values!" :78-84
    compute::ClBuffer buf = fresh();
    buf.MoveToGpu();
    EXPECT_TRUE(buf.GetBufferLocation() == compute::ClBuffer::GPU && buf.size() == 5);
    buf.MoveToCpu();
    for (size_t i = 0; i < 5; ++i) EXPECT_TRUE(buf[i] == i * 2 + 1);
  }
  {  // "DeepClone in GPU mode doesn't ruin everything" :86-94
    compute::ClBuffer buf = fresh();
    buf.MoveToGpu();
    auto buf2 = buf.DeepClone();
    EXPECT_TRUE(buf2.gpu_buffer() != buf.gpu_buffer());  // a copy, not an alias
    buf2.MoveToCpu();
    buf.MoveToCpu();
    for (size_t i = 0; i < 5; ++i) EXPECT_TRUE(buf[i] == buf2[i]);
  }
  {  // "Assigning via opencl doesn't kill values" :96-126
    compute::ClBuffer buf = fresh();
    buf.MoveToGpu();
    buf = device_values(10);
    buf.MoveToCpu();
    for (size_t i = 0; i < 5; ++i) EXPECT_TRUE(buf[i] == i * 10);
  }
  {  // "opencl copy constructor" :128-184: 300 device-side copies, each moved both ways
    compute::ClBuffer buf = fresh();
    buf.MoveToGpu();
    buf = device_values(10);
    compute::ClBuffer buf2(buf);
    for (size_t i = 0; i < 100; ++i) {
      compute::ClBuffer buf3(buf2);
      buf3.MoveToCpu();
      buf3.MoveToGpu();
      compute::ClBuffer buf4(buf3);
      buf4.MoveToCpu();
      buf4.MoveToGpu();
      compute::ClBuffer buf5(buf4);
      buf5.MoveToCpu();
      buf5.MoveToGpu();
    }
    buf2.MoveToCpu();
    for (size_t i = 0; i < 5; ++i) {
      EXPECT_TRUE(buf2[i] == i * 10);
      buf2[i] = i * 2 + 1;
    }
    buf2.MoveToGpu();
    buf.MoveToCpu();  // the original is untouched by what happened to its copies
    for (size_t i = 0; i < 5; ++i) EXPECT_TRUE(buf[i] == i * 10);
  }
  {  // "Assigning via cpu doesn't kill opencl values" :186-226
    compute::ClBuffer buf = fresh();
    buf.MoveToGpu();
    buf.MoveToGpu();
    buf = device_values(10);
    buf.MoveToCpu();
    buf.MoveToGpu();
    buf.MoveToCpu();
    for (size_t i = 0; i < 5; ++i) buf[i] = i * 2 + 1;
    for (size_t i = 0; i < 100; ++i) {
      buf.MoveToCpu();
      buf.MoveToGpu();
      buf.MoveToCpu();
    }
    for (size_t i = 0; i < 5; ++i) EXPECT_TRUE(buf[i] == i * 2 + 1);
  }
  {  // empty buffers and resize through the host, cl_buffer.cc:14-51
    compute::ClBuffer buf(ctx);
    buf.MoveToGpu();
    EXPECT_TRUE(buf.size() == 0);
    buf.resize(3, 7.0);
    EXPECT_TRUE(buf.GetBufferLocation() == compute::ClBuffer::GPU && buf.size() == 3);
    buf.MoveToCpu();
    EXPECT_TRUE(buf[0] == 7.0 && buf[2] == 7.0);
  }
  PB_CHECK_OK(pb_context_destroy(ctx));
}

int main() {
  ClBufferCases();
  KernelQueue();
  WorkgroupSizes();
  Yolo();
  OneLayerRelu();
  MazurStep();
  ConvLayout();
  SaveLoadAndBatch();
  std::printf(failures ? "mirror_test: %d FAILED\n" : "mirror_test: all passed\n", failures);
  return failures ? 1 : 0;
}
