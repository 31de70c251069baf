// symbolic::Expression — tag type of the B200 host mirror.
//
// The reference builds every kernel from symbolic expression trees
// (symbolic/expression.h:32-98).  This build has no code generator: an
// Expression is only the *tag* by which the four built-in activation functions
// (symbolic/symbolic_util.h:10-15) are recognised when a caller passes them as
// nnet::Layer::ActivationFunctionType (nnet/layer_impl.h:21-22).  Any other
// function is refused: there is no generic-expression fallback.
#ifndef PLASTICITY_B200_SYMBOLIC_EXPRESSION_H_
#define PLASTICITY_B200_SYMBOLIC_EXPRESSION_H_

namespace symbolic {

class Expression {
 public:
  // kProbe is what the mirror feeds an activation function to identify it.
  enum Tag { kProbe = -1, kIdentity = 0, kRelu = 1, kLeakyRelu = 2, kSigmoid = 3, kUnknown = 99 };
  Expression() : tag_(kUnknown) {}
  explicit Expression(Tag tag) : tag_(tag) {}
  Expression(double) : tag_(kUnknown) {}
  Tag tag() const { return tag_; }

 private:
  Tag tag_;
};

}  // namespace symbolic
#endif
