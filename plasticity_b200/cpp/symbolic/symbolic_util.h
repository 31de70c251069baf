// The four activation functions the reference ships (symbolic/symbolic_util.h:10-15,
// symbolic/symbolic_util.cc:10-28), as tags: Identity x; Relu x>=0?x:0;
// LeakyRelu x>=0?x:x/10; Sigmoid 1/(1+2.71828^-x).  Their arithmetic lives in
// the CUDA kernels (plasticity_b200/csrc/kernels.cuh act_fwd / act_bwd).
#ifndef PLASTICITY_B200_SYMBOLIC_UTIL_H_
#define PLASTICITY_B200_SYMBOLIC_UTIL_H_

#include "symbolic/expression.h"

namespace symbolic {

inline Expression Identity(const Expression &) { return Expression(Expression::kIdentity); }
inline Expression Relu(const Expression &) { return Expression(Expression::kRelu); }
inline Expression LeakyRelu(const Expression &) { return Expression(Expression::kLeakyRelu); }
inline Expression Sigmoid(const Expression &) { return Expression(Expression::kSigmoid); }

}  // namespace symbolic
#endif
