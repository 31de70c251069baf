// CIFAR-10 training driver over the B200 layer path: what nnet/cifar_test.cc does
// (binary loader :191-205, signed-byte L2 normalisation :66-81, one-hot targets
// :83-97, the 13-layer model :293-350, lr 1e-4, per-sample SGD loop :419-428 with a
// progress / rate print every 5 000 samples :437-447, example accuracy :221-243,
// weight file load / save :207-216, 352-367, `--short`), written against the same
// nnet::Architecture / nnet::Nnet / compute::ClBuffer interface.  Not carried over:
// the JPEG dump of sorted images (:99-186, 456-479; needs libjpeg-turbo, SURVEY
// section 2 "out of scope").
//
//   cifar_train [weights.json] [--short] [--data DIR] [--epochs N] [--batch B]
//               [--synthetic N] [--examples N] [--seed S] [--lr RATE]
//
//   --data DIR       directory holding data_batch_{1..5}.bin and test_batch.bin
//                    (default nnet/data/cifar-10-batches-bin, as in the reference)
//   --synthetic N    no data set on this machine: N records in the CIFAR binary format
//                    (1 label byte + 3072 pixel bytes) whose class is readable from the
//                    pixels (a class-coloured blob on noise), generated in memory
//   --batch B        B == 0 (default): strict per-sample SGD, Nnet::Train order (device
//                    loop, CUDA-graph replay per sample); B > 0: gradient-sum minibatches
//                    of B samples (Nnet::BatchTrain semantics)
//   --epochs N       default 100000 like the reference; --short = evaluate only (0 epochs)
//   --lr RATE        learning rate (default 0.0001, cifar_test.cc:414)
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "nnet/layer_dimensions.h"
#include "nnet/nnet.h"

constexpr size_t kSampleSize = 32 * 32 * 3;
constexpr size_t kOutputSize = 10;
constexpr size_t kRecordSize = kSampleSize + 1;  // one label byte + the pixels

static const char *kLabels[10] = {"airplane", "automobile", "bird", "cat", "deer",
                                  "dog", "frog", "horse", "ship", "truck"};

struct Sample {
  char label;
  char pixels[kSampleSize];  // read as signed char, like the reference (:55-62)
};

static void LoadCifarFile(const std::string &file, std::vector<Sample> *out) {
  std::ifstream f(file.c_str(), std::ios::binary);
  std::stringstream buffer;
  buffer << f.rdbuf();
  const std::string bytes = buffer.str();
  std::cout << "Loading file " << file << "..." << std::endl;
  for (size_t at = 0; at + kRecordSize <= bytes.size(); at += kRecordSize) {
    Sample s;
    s.label = bytes[at];
    memcpy(s.pixels, &bytes[at + 1], kSampleSize);
    out->push_back(s);
  }
}

// Records with a learnable rule: class c paints a 12x12 blob whose position and colour
// depend on c over uniform noise.
static void MakeSynthetic(size_t n, uint64_t seed, std::vector<Sample> *out) {
  std::mt19937_64 rng(seed);
  for (size_t i = 0; i < n; ++i) {
    Sample s;
    const int c = (int)(rng() % 10);
    s.label = (char)c;
    for (size_t k = 0; k < kSampleSize; ++k) s.pixels[k] = (char)((int)(rng() % 64) - 32);
    const int x0 = (c % 5) * 4, y0 = (c / 5) * 16 + 2, plane = c % 3;
    for (int y = 0; y < 12; ++y)
      for (int x = 0; x < 12; ++x)
        s.pixels[plane * 1024 + (y0 + y) * 32 + x0 + x] = (char)(100 + (int)(rng() % 27));
    out->push_back(s);
  }
}

static void Normalize(const Sample &s, double *dst) {  // cifar_test.cc:66-81
  double norm = 0;
  for (size_t i = 0; i < kSampleSize; ++i) norm += (double)s.pixels[i] * s.pixels[i];
  norm = std::sqrt(norm);
  if (norm == 0) norm = 1;
  for (size_t i = 0; i < kSampleSize; ++i) dst[i] = (double)s.pixels[i] / norm;
}

static int ArgMax(const compute::ClBuffer &out, size_t offset) {
  int best = 0;
  for (size_t i = 1; i < kOutputSize; ++i)
    if (out[offset + i] > out[offset + best]) best = (int)i;
  return best;
}

// Accuracy over `count` examples starting at `first`, evaluated in batches.
static double Accuracy(nnet::Nnet *net, const std::vector<Sample> &samples, size_t first,
                       size_t count, size_t max_batch, bool print) {
  size_t correct = 0;
  for (size_t b0 = 0; b0 < count; b0 += max_batch) {
    const size_t nb = std::min(max_batch, count - b0);
    std::vector<double> x(nb * kSampleSize);
    for (size_t i = 0; i < nb; ++i) Normalize(samples[(first + b0 + i) % samples.size()], &x[i * kSampleSize]);
    auto in = net->MakeBuffer(x);
    auto out = net->Evaluate(in);
    out->MoveToCpu();
    for (size_t i = 0; i < nb; ++i) {
      const Sample &s = samples[(first + b0 + i) % samples.size()];
      const int got = ArgMax(*out, i * kOutputSize);
      if (print && b0 + i < 5)
        std::cout << "Actual Answer: " << kLabels[(int)s.label % 10]
                  << "\nnnet output: " << kLabels[got] << std::endl;
      correct += got == (int)s.label;
    }
  }
  return 100.0 * (double)correct / (double)count;
}

int main(int argc, char *argv[]) {
  bool only_one_epoch = false;
  std::string weight_file_path, data_dir = "nnet/data/cifar-10-batches-bin";
  size_t epochs = 100000, batch = 0, synthetic = 0, examples = 1000;
  uint64_t seed = 1;
  double learning_rate = 0.0001;
  for (int i = 1; i < argc; ++i) {
    const std::string arg = argv[i];
    auto value = [&](size_t *dst) { if (i + 1 < argc) *dst = (size_t)std::strtoull(argv[++i], nullptr, 10); };
    if (arg == "--short") only_one_epoch = true;
    else if (arg == "--data" && i + 1 < argc) data_dir = argv[++i];
    else if (arg == "--epochs") value(&epochs);
    else if (arg == "--batch") value(&batch);
    else if (arg == "--synthetic") value(&synthetic);
    else if (arg == "--examples") value(&examples);
    else if (arg == "--lr" && i + 1 < argc) learning_rate = std::strtod(argv[++i], nullptr);
    else if (arg == "--seed") {  // synthetic records AND the Xavier draw (PLASTICITY_SEED)
      size_t s = 1;
      value(&s);
      seed = s;
      setenv("PLASTICITY_SEED", std::to_string(s).c_str(), 1);
    }
    else if (arg.rfind("--", 0) != 0) weight_file_path = arg;
  }
  if (only_one_epoch) epochs = 0;  // cifar_test.cc:262-264
  std::cout << "Number of Epochs: " << epochs << std::endl;

  // cifar_test.cc:293-350, http://cs231n.github.io/convolutional-networks/
  nnet::Architecture model(kSampleSize);
  model.AddConvolutionLayer({32, 32, 3}, {5, 5, 3, 1, 2, 16})
      .AddMaxPoolLayer({32, 32, 16}, {16, 16})
      .AddConvolutionLayer({16, 16, 16}, {5, 5, 16, 1, 2, 20})
      .AddMaxPoolLayer({16, 16, 20}, {8, 8})
      .AddConvolutionLayer({8, 8, 20}, {5, 5, 20, 1, 2, 20})
      .AddMaxPoolLayer({8, 8, 20}, {4, 4})
      .AddDenseLayer(10, symbolic::Identity)
      .AddSoftmaxLayer(10);
  const size_t max_batch = std::max<size_t>(batch, 500);
  nnet::Nnet test_net(model, nnet::Nnet::Xavier, nnet::CrossEntropy, max_batch);

  if (!weight_file_path.empty()) {
    std::ifstream weight_file(weight_file_path.c_str());
    std::stringstream buffer;
    buffer << weight_file.rdbuf();
    const std::string weight_string = buffer.str();
    if (weight_string.empty()) {
      std::cout << "Provided weight file is empty. Initializing with random weights" << std::endl;
    } else {
      std::cout << "Loading weights from file " << weight_file_path << std::endl;
      if (!test_net.LoadWeightsFromString(weight_string)) {
        std::cerr << "Failed to load weights from file: " << weight_file_path << std::endl;
        return 1;
      }
    }
  }

  std::vector<Sample> samples, test_batch;
  if (synthetic > 0) {
    MakeSynthetic(synthetic, seed, &samples);
    MakeSynthetic(std::max<size_t>(synthetic / 5, 100), seed + 1, &test_batch);
  } else {
    LoadCifarFile(data_dir + "/test_batch.bin", &test_batch);
    std::cout << "Reading in training files..." << std::endl;
    for (int k = 1; k <= 5; ++k)
      LoadCifarFile(data_dir + "/data_batch_" + std::to_string(k) + ".bin", &samples);
  }
  std::cout << "Loaded " << samples.size() << " Samples from disk!" << std::endl;
  if (samples.empty()) {
    std::cerr << "No samples found, quitting!" << std::endl;
    return -1;
  }

  // The whole training set goes to the device once, samples back to back (the reference
  // keeps one ClBuffer per sample, cifar_test.cc:393-412).
  std::cout << "Moving samples to GPU..." << std::endl;
  std::vector<double> x(samples.size() * kSampleSize), e(samples.size() * kOutputSize, 0.0);
  for (size_t i = 0; i < samples.size(); ++i) {
    Normalize(samples[i], &x[i * kSampleSize]);
    e[i * kOutputSize + ((int)samples[i].label % 10)] = 1.0;
  }
  auto inputs = test_net.MakeBuffer(x);
  auto outputs = test_net.MakeBuffer(e);
  inputs->MoveToGpu();
  outputs->MoveToGpu();
  std::cout << "Entire dataset of " << samples.size() << " examples is now stored on GPU!"
            << std::endl;

  test_net.SetLearningParameters({.learning_rate = learning_rate});

  size_t samples_so_far = 0, next_report = 5000, next_save = 100000;
  auto start = std::chrono::high_resolution_clock::now();
  size_t since = 0;
  for (size_t epoch = 1; epoch <= epochs; ++epoch) {
    const size_t step = batch ? batch : std::min<size_t>(samples.size(), 5000);
    for (size_t i = 0; i < samples.size(); i += step) {
      const size_t nb = std::min(step, samples.size() - i);
      auto xp = compute::ClBuffer::View(test_net.SharedContext(),
                                        inputs->gpu_buffer() + i * kSampleSize, nb * kSampleSize);
      auto ep = compute::ClBuffer::View(test_net.SharedContext(),
                                        outputs->gpu_buffer() + i * kOutputSize, nb * kOutputSize);
      if (batch) test_net.BatchTrain(xp, ep, nb);  // minibatch, W -= lr * sum_b grad_b
      else test_net.TrainLoop(xp, ep, nb);          // SGD, one Train() per sample
      samples_so_far += nb;
      since += nb;
      if (samples_so_far >= next_save) {
        next_save += 100000;
        std::cout << "Examples correct: " << Accuracy(&test_net, samples, 0, std::min(examples, samples.size()), max_batch, false) << "%" << std::endl;
        if (!weight_file_path.empty()) {
          std::cout << "Saving weights to file: " << weight_file_path << std::endl;
          std::ofstream(weight_file_path) << test_net.WeightsToString();
        }
      }
      if (samples_so_far >= next_report) {
        next_report += 5000;
        PB_CHECK_OK(pb_context_finish(test_net.SharedContext()));
        auto end = std::chrono::high_resolution_clock::now();
        const double ms = std::chrono::duration<double, std::milli>(end - start).count();
        std::cout << "Progress: " << samples_so_far - 1 << " / " << epochs * samples.size() << std::endl;
        std::cout << "Epoch " << epoch << std::endl;
        std::cout << "Training rate (samples per second): " << 1000.0 * (double)since / ms << std::endl;
        start = std::chrono::high_resolution_clock::now();
        since = 0;
      }
    }
  }
  std::cout << "Training completed!" << std::endl;
  std::cout << "Trained over " << samples.size() << " Samples!" << std::endl;
  const size_t n_ex = std::min(examples, test_batch.size());
  std::cout << "Example network outputs: " << std::endl;
  const double acc = Accuracy(&test_net, test_batch, 0, n_ex, max_batch, true);
  std::cout << "Examples correct: " << (size_t)std::llround(acc * n_ex / 100.0) << " / " << n_ex
            << " (" << acc << "%)" << std::endl;
  if (!weight_file_path.empty() && epochs > 0) {
    std::cout << "Saving weights to file: " << weight_file_path << std::endl;
    std::ofstream(weight_file_path) << test_net.WeightsToString();
  }
  return 0;
}
