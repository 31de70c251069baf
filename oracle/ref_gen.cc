// TEST INFRASTRUCTURE — not part of the product path.
//
// ref_gen: links the reference's OWN kernel generator (symbolic/, codegen/,
// nnet/*_layer.cc compiled in place from /root/reference by oracle/Makefile)
// and emits, for one network spec, the same OpenCL C text that
// nnet::Layer::GenerateEvaluationKernel / GenerateTrainingKernels
// (nnet/layer.cc:140-238) and ErrorLayer::GenerateErrorKernels
// (nnet/error_layer.cc:36-56) would hand to clutil::Compile, followed by a
// dispatch table so a host loop can play the role of enqueueNDRangeKernel.
//
// nnet/layer.cc itself cannot be compiled here (it includes nnet/nnet.h ->
// clutil + OpenCL, both absent), so the template substitution of layer.cc is
// restated below; every LayerImpl / ErrorLayer / template file used is the
// reference's.  Output goes to oracle/_ref/ only (git-ignored).
//
// Spec file: one directive per line
//   loss mse|ce
//   act <n> identity|relu|leaky|sigmoid
//   dense <in> <out>
//   conv <W> <H> <D> <fw> <fh> <fd> <stride> <pad> <nfilters>
//   pool <W> <H> <D> <ow> <oh>
//   softmax <n>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "codegen/codegen.h"
#include "nnet/activation_layer.h"
#include "nnet/convolution_layer.h"
#include "nnet/dense_layer.h"
#include "nnet/error_layer.h"
#include "nnet/layer_dimensions.h"
#include "nnet/layer_impl.h"
#include "nnet/max_pool_layer.h"
#include "nnet/softmax_layer.h"
#include "symbolic/expression.h"
#include "symbolic/symbolic_util.h"

namespace {

std::string Slurp(const std::string &path) {
  std::ifstream in(path);
  if (!in.is_open()) {
    std::cerr << "ref_gen: cannot open " << path << std::endl;
    std::exit(2);
  }
  std::stringstream ss;
  ss << in.rdbuf();
  return ss.str();
}

void ReplaceOnce(std::string *text, const std::string &key,
                 const std::string &value) {
  size_t at = text->find(key);
  if (at == std::string::npos) {
    std::cerr << "ref_gen: placeholder " << key << " missing" << std::endl;
    std::exit(2);
  }
  text->replace(at, key.size(), value);
}

void ReplaceAll(std::string *text, const std::string &key,
                const std::string &value) {
  size_t at = 0;
  while ((at = text->find(key, at)) != std::string::npos) {
    text->replace(at, key.size(), value);
    at += value.size();
  }
}

struct Entry {
  std::unique_ptr<nnet::LayerImpl> impl;
  std::string suffix;  // "<layer_type>_<index>", nnet/layer.h:99-101
};

nnet::LayerImpl::ActivationFunctionType ActivationByName(
    const std::string &name) {
  if (name == "identity") return symbolic::Identity;
  if (name == "relu") return symbolic::Relu;
  if (name == "leaky") return symbolic::LeakyRelu;
  if (name == "sigmoid") return symbolic::Sigmoid;
  std::cerr << "ref_gen: unknown activation " << name << std::endl;
  std::exit(2);
}

}  // namespace

int main(int argc, char **argv) {
  if (argc != 4) {
    std::cerr << "usage: ref_gen <reference_root> <spec> <out.cc>" << std::endl;
    return 2;
  }
  const std::string root = argv[1];
  std::ifstream spec(argv[2]);
  if (!spec.is_open()) {
    std::cerr << "ref_gen: cannot open spec " << argv[2] << std::endl;
    return 2;
  }

  nnet::LossFunction loss = nnet::MeanSquared;
  std::vector<Entry> layers;
  std::string line;
  while (std::getline(spec, line)) {
    std::istringstream ls(line);
    std::string kind;
    if (!(ls >> kind) || kind[0] == '#') continue;
    const size_t index = layers.size();
    Entry e;
    if (kind == "loss") {
      std::string which;
      ls >> which;
      loss = (which == "ce") ? nnet::CrossEntropy : nnet::MeanSquared;
      continue;
    } else if (kind == "act") {
      size_t n;
      std::string fn;
      ls >> n >> fn;
      e.impl = std::make_unique<nnet::ActivationLayer>(n, ActivationByName(fn),
                                                       index);
    } else if (kind == "dense") {
      size_t in, out;
      ls >> in >> out;
      e.impl =
          std::make_unique<nnet::DenseLayer>(nnet::Dimensions{in, out}, index);
    } else if (kind == "conv") {
      size_t W, H, D, fw, fh, fd, stride, pad, nf;
      ls >> W >> H >> D >> fw >> fh >> fd >> stride >> pad >> nf;
      e.impl = std::make_unique<nnet::ConvolutionLayer>(
          nnet::VolumeDimensions{W, H, D},
          nnet::FilterParams{fw, fh, fd, stride, pad, nf}, index);
    } else if (kind == "pool") {
      size_t W, H, D, ow, oh;
      ls >> W >> H >> D >> ow >> oh;
      e.impl = std::make_unique<nnet::MaxPoolLayer>(
          nnet::VolumeDimensions{W, H, D}, nnet::AreaDimensions{ow, oh}, index);
    } else if (kind == "softmax") {
      size_t n;
      ls >> n;
      e.impl = std::make_unique<nnet::SoftmaxLayer>(n, index);
    } else {
      std::cerr << "ref_gen: unknown directive " << kind << std::endl;
      return 2;
    }
    e.suffix = e.impl->layer_type() + "_" + std::to_string(index);
    layers.push_back(std::move(e));
  }
  if (layers.empty()) {
    std::cerr << "ref_gen: empty spec" << std::endl;
    return 2;
  }

  const std::string eval_tpl = Slurp(root + "/nnet/kernels/evaluate.kernel.cl");
  const std::string bp_tpl = Slurp(root + "/nnet/kernels/back_prop.kernel.cl");
  const std::string err_tpl = Slurp(root + "/nnet/kernels/error.kernel.cl");
  const std::string comb_tpl = Slurp(root + "/nnet/kernels/combine.kernel.cl");

  std::ofstream out(argv[3]);
  out << "// GENERATED by oracle/ref_gen from the reference's kernel generator."
         " Do not commit.\n#include \"ref_shim.h\"\n\n";

  using symbolic::Expression;
  for (const Entry &e : layers) {
    // nnet/layer.cc:140-183
    std::string eval = eval_tpl;
    codegen::CudaGenerator gen;
    e.impl->GenerateOutputCode(Expression::CreateInteger("output_index"), &gen);
    ReplaceOnce(&eval, "EXPRESSION_HERE", gen.code());
    ReplaceAll(&eval, "LAYERID", std::to_string(e.impl->layer_index()));
    out << eval << "\n";

    // nnet/layer.cc:200-238
    std::string bp = bp_tpl;
    codegen::CudaGenerator in_gen;
    e.impl->InputGradientCode(Expression::CreateInteger("input_index"),
                              &in_gen);
    codegen::CudaGenerator w_gen;
    e.impl->WeightGradientCode(Expression::CreateInteger("weight_index"),
                               &w_gen);
    ReplaceOnce(&bp, "INPUT_GRADIENTS_HERE", in_gen.code());
    ReplaceOnce(&bp, "WEIGHT_GRADIENTS_HERE", w_gen.code());
    ReplaceAll(&bp, "LAYERID", e.suffix);
    out << bp << "\n";
  }

  // nnet/error_layer.cc:36-56 (the class is the reference's).
  {
    nnet::ErrorLayer error(loss, layers.back().impl->GetDimensions().num_outputs);
    std::string text = error.GenerateErrorKernels();
    if (text.find("error_value") == std::string::npos) {
      // GenerateErrorKernels reads the template relative to the CWD; when run
      // elsewhere fall back to doing the same substitution on the file we read.
      std::cerr << "ref_gen: run from the reference root (error template)"
                << std::endl;
      return 2;
    }
    out << text << "\n";
  }
  out << comb_tpl << "\n";

  // Dispatch: the part clutil/OpenCL would have played.
  out << "\nextern \"C\" {\n";
  out << "int ref_num_layers(void) { return " << layers.size() << "; }\n";
  out << "int ref_loss(void) { return " << static_cast<int>(loss) << "; }\n";
  out << "void ref_layer_info(int l, long *ni, long *no, long *nw, const char "
         "**name) {\n  switch (l) {\n";
  for (size_t i = 0; i < layers.size(); ++i) {
    const auto d = layers[i].impl->GetDimensions();
    out << "  case " << i << ": *ni=" << d.num_inputs << "; *no="
        << d.num_outputs << "; *nw=" << layers[i].impl->weights().size()
        << "; *name=\"" << layers[i].suffix << "\"; break;\n";
  }
  out << "  default: *ni=*no=*nw=-1; *name=\"\"; }\n}\n";

  auto emit_switch = [&](const std::string &fn, const std::string &params,
                         const std::string &count_field,
                         const std::string &prefix, const std::string &args,
                         bool by_index) {
    out << "void " << fn << "(int l, " << params << ") {\n  switch (l) {\n";
    for (size_t i = 0; i < layers.size(); ++i) {
      const auto d = layers[i].impl->GetDimensions();
      size_t n = count_field == "out"  ? d.num_outputs
                 : count_field == "in" ? d.num_inputs
                                       : layers[i].impl->weights().size();
      std::string name =
          prefix + (by_index ? std::to_string(i) : layers[i].suffix);
      out << "  case " << i << ": REF_NDRANGE(" << n << ", " << name << "("
          << args << ")); break;\n";
    }
    out << "  }\n}\n";
  };
  emit_switch("ref_evaluate", "double *I, double *W, double *O", "out",
              "evaluate_", "I, W, O", true);
  emit_switch("ref_input_delta", "double *I, double *W, double *G, double *dI",
              "in", "input_delta_", "I, W, G, dI", false);
  emit_switch("ref_new_weights",
              "double *I, double *W, double *G, double *Wn, double *lr", "w",
              "new_weights_", "I, W, G, Wn, lr", false);
  emit_switch("ref_weight_deltas",
              "double *I, double *W, double *G, double *dW, double *lr", "w",
              "weight_deltas_", "I, W, G, dW, lr", false);
  const size_t nout = layers.back().impl->GetDimensions().num_outputs;
  out << "void ref_error_value(double *O, double *E, double *c) { REF_NDRANGE("
      << nout << ", error_value(O, E, c)); }\n";
  out << "void ref_error_gradients(double *O, double *E, double *g) { "
         "REF_NDRANGE("
      << nout << ", error_gradients(O, E, g)); }\n";
  out << "void ref_vector_accumulate(double *a, double *b, long n) { "
         "REF_NDRANGE(n, vector_accumulate(a, b)); }\n";
  out << "}  // extern C\n";
  return 0;
}
