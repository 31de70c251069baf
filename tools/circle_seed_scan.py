"""CPU model of tests/cpp/_ref/ref_circle_test (nnet/circle_test.cc through the C++ mirror)
for a given PLASTICITY_SEED: the device-side Xavier draw of nnet.cu (splitmix64 + Box-Muller)
re-stated in numpy, glibc rand() for the samples (the test never seeds it), 10 000 per-sample
SGD steps and 1 000 probe points in the fp64 oracle.  Used to pick a seed for
tests/test_cpp_mirror.py::test_reference_circle_test_learns whose result is far from the
900 / 1000 bar (an unseeded run fails the bar now and then).
  python tools/circle_seed_scan.py [first_seed] [count]"""
import ctypes, math, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po

M = (1 << 64) - 1


def xavier(seed, base, n, stddev):
    out = np.zeros(n)
    for i in range(n):
        x = (seed + 0x9E3779B97F4A7C15 * (base + i + 1)) & M
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M
        x = x ^ (x >> 31)
        u1 = np.float32((np.float32(x >> 40) + np.float32(1.0)) * np.float32(1.0 / 16777217.0))
        u2 = np.float32(np.float32((x >> 8) & 0xFFFFFF) * np.float32(1.0 / 16777216.0))
        out[i] = np.float32(stddev) * np.float32(math.sqrt(-2.0 * math.log(u1))) * \
            np.float32(math.cos(2.0 * math.pi * float(u2)))
    return out


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    libc = ctypes.CDLL("libc.so.6")
    RAND_MAX = 2147483647
    S = po.RestatedNet(po.spec_text("circle"))
    for seed in range(first, first + count):
        libc.srand(1)   # the state of a process that never called srand

        def coord():
            return 2.5 * float(libc.rand()) / RAND_MAX - 1.25
        xs = np.zeros((10000, 2)); es = np.zeros((10000, 1))
        for i in range(10000):
            x = coord(); y = coord()
            xs[i] = (x, y); es[i, 0] = 1.0 if x * x + y * y <= 1.0 else 0.0
        w = S.zeros_weights()
        # product layout: layer blocks at 4-float aligned offsets (30 weights at 0, 11 at 32)
        S.layer_weights(w, 1)[:] = xavier(seed, 0, 30, math.sqrt(1.0 / 2))
        S.layer_weights(w, 3)[:] = xavier(seed, 32, 11, math.sqrt(1.0 / 10))
        S.train_loop(xs, es, w, 0.3)
        ok = 0
        for i in range(1000):
            x = coord(); y = coord()
            s = S.evaluate(np.array([x, y]), w)[-1]
            ok += (x * x + y * y <= 1.0) == (s > 0.5)
        print(f"seed {seed}: {ok}/1000", flush=True)


if __name__ == "__main__":
    main()
