"""Accuracy of split-precision tensor-core products against fp64, on the CPU (numpy):
3xTF32 as the kernels compute it (hi = truncation to 10 mantissa bits, lo = x - hi;
hi*hi + hi*lo + lo*hi, fp32 accumulation) and the bf16x3 variant considered in DESIGN.md
section 9 (x = b0 + b1 + b2, 8 mantissa bits each; the six products b0b0, b0b1, b1b0, b1b1,
b0b2, b2b0), plus the round-to-nearest two-term splits.  K = 400 (the CIFAR conv-4 contraction), operands ~ N(0,1) and a cancelling
case; errors are relative to max |exact|, the tests' bar is 1e-4.
  python tools/split_accuracy.py"""
import numpy as np


def trunc_bits(x, keep):
    """fp32 x with the mantissa truncated to `keep` explicit bits."""
    u = np.asarray(x, np.float32).view(np.uint32)
    mask = np.uint32((0xFFFFFFFF << (23 - keep)) & 0xFFFFFFFF)
    return (u & mask).view(np.float32)


def rn_bits(x, keep):
    """fp32 x rounded to nearest-even at `keep` explicit mantissa bits."""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    sh = 23 - keep
    u = u + ((1 << (sh - 1)) - 1) + ((u >> sh) & 1)
    return (((u >> sh) << sh) & 0xFFFFFFFF).astype(np.uint32).view(np.float32)


def two_term_rn(a, b, keep):
    """x = x0 + x1 with round-to-nearest parts; a0*b0 + a0*b1 + a1*b0, fp32 accumulation."""
    a0, b0 = rn_bits(a, keep), rn_bits(b, keep)
    a1, b1 = rn_bits((a - a0).astype(np.float32), keep), rn_bits((b - b0).astype(np.float32), keep)
    acc = np.zeros((a.shape[0], b.shape[1]), np.float32)
    for x, y in ((a1, b0), (a0, b1), (a0, b0)):
        acc += (x.astype(np.float64) @ y.astype(np.float64)).astype(np.float32)
    return acc


def tf32x3(a, b):
    ah, bh = trunc_bits(a, 10), trunc_bits(b, 10)
    al, bl = (a - ah).astype(np.float32), (b - bh).astype(np.float32)
    # the tensor core reads tf32 operands: the residuals are truncated again on entry
    al, bl = trunc_bits(al, 10), trunc_bits(bl, 10)
    acc = np.zeros((a.shape[0], b.shape[1]), np.float32)
    for x, y in ((al, bh), (ah, bl), (ah, bh)):
        acc += (x.astype(np.float64) @ y.astype(np.float64)).astype(np.float32)
    return acc


def bf16x3(a, b, products=6):
    def split(x):
        p0 = trunc_bits(x, 7)
        r = (x - p0).astype(np.float32)
        p1 = trunc_bits(r, 7)
        p2 = trunc_bits((r - p1).astype(np.float32), 7)
        return p0, p1, p2
    a0, a1, a2 = split(a)
    b0, b1, b2 = split(b)
    terms = [(a2, b0), (a0, b2), (a1, b1), (a1, b0), (a0, b1), (a0, b0)]
    acc = np.zeros((a.shape[0], b.shape[1]), np.float32)
    for x, y in terms[6 - products:]:
        acc += (x.astype(np.float64) @ y.astype(np.float64)).astype(np.float32)
    return acc


def main():
    rng = np.random.default_rng(0)
    M, K, N = 256, 400, 20
    cases = {
        "N(0,1) x N(0,1)": (rng.normal(0, 1, (M, K)), rng.normal(0, 1, (K, N))),
        "positive x positive (no cancellation)": (rng.uniform(0, 1, (M, K)), rng.uniform(0, 1, (K, N))),
        "wide dynamic range": (rng.normal(0, 1, (M, K)) * 10.0 ** rng.uniform(-3, 3, (M, K)),
                               rng.normal(0, 1, (K, N))),
    }
    for name, (a, b) in cases.items():
        a, b = a.astype(np.float32), b.astype(np.float32)
        exact = a.astype(np.float64) @ b.astype(np.float64)
        scale = np.abs(exact).max()
        fp32 = np.zeros_like(exact, dtype=np.float32)
        for k in range(K):   # plain fp32 FMA chain, ascending k (the FFMA kernels)
            fp32 += a[:, k:k + 1] * b[k:k + 1, :]
        print(name)
        for label, got in (("fp32 FFMA chain", fp32), ("1xTF32", None), ("3xTF32", tf32x3(a, b)),
                           ("bf16x3, 6 products", bf16x3(a, b, 6)),
                           ("bf16x3, 5 products (no b1*b1)", None),
                           ("bf16x3, 3 products", bf16x3(a, b, 3)),
                           ("3xTF32, round-to-nearest parts", two_term_rn(a, b, 10)),
                           ("bf16 2-term RN, 3 products", two_term_rn(a, b, 7))):
            if label == "1xTF32":
                got = (trunc_bits(a, 10).astype(np.float64) @ trunc_bits(b, 10).astype(np.float64))
            if label.startswith("bf16x3, 5"):
                def split(x):
                    p0 = trunc_bits(x, 7); r = (x - p0).astype(np.float32)
                    p1 = trunc_bits(r, 7); return p0, p1, trunc_bits((r - p1).astype(np.float32), 7)
                a0, a1, a2 = split(a); b0, b1, b2 = split(b)
                got = sum((x.astype(np.float64) @ y.astype(np.float64)).astype(np.float32)
                          for x, y in ((a2, b0), (a0, b2), (a1, b0), (a0, b1), (a0, b0)))
            print(f"   {label:32s} max |err| / max |exact| = {np.abs(got - exact).max() / scale:.2e}")


if __name__ == "__main__":
    main()
