"""CPU tests (no GPU) of the host side: the C ABI exports what include/*.h
declares, the Architecture builder and weight layout match the reference, and
the data-parallel plumbing (world_size 2, gloo) reproduces BatchTrain over the
global batch."""
import ctypes
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

import plasticity_b200 as pb
from plasticity_b200 import _lib, parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "plasticity_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) > 50
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/plasticity_b200.h but not exported"
    assert lib.pb_abi_version() == 1
    # every binding the Python mirror uses is declared in the header
    for name in _lib.SIGNATURES:
        assert name in names, f"{name} bound in _lib.py but missing from the header"


def test_no_gpu_means_loud_failure_not_fallback():
    """Without a device the product refuses to create a context."""
    lib = _lib.load()
    n = ctypes.c_int(0)
    rc = lib.pb_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    assert lib.pb_context_create(0, ctypes.byref(h)) != 0
    assert lib.pb_last_error()


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under plasticity_b200/ or
    include/ may import, link or exec it."""
    bad = []
    for base in ("plasticity_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dirpath or "__pycache__" in dirpath:
                continue
            for f in files:
                if not f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp", "Makefile")):
                    continue
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"(import|from)\s+oracle|oracle/|liboracle|libref_|pyoracle", text):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_architecture_builder_matches_reference_layer_list():
    """nnet/architecture.h: implicit identity input layer; Dense => dense +
    activation (default Sigmoid); Conv => conv + activation (default LeakyRelu)."""
    m = pb.Architecture(2)
    m.AddDenseLayer(10).AddDenseLayer(1)  # circle_test.cc:19-22
    assert [l.LayerSuffix() for l in m.layers] == [
        "activation_layer_0", "dense_layer_1", "activation_layer_2", "dense_layer_3",
        "activation_layer_4"]
    assert [l.weights.size for l in m.layers] == [0, 30, 0, 11, 0]
    assert m.layers[2].desc.activation == pb.Sigmoid
    assert m.VerifyArchitecture() and m.input_size() == 2 and m.output_size() == 1

    sys.path.insert(0, ROOT)
    import bench
    c = bench.cifar_model(pb)  # cifar_test.cc:293-350
    assert len(c.layers) == 13
    assert [l.weights.size for l in c.layers if l.weights.size] == [1216, 8020, 10020, 3210]
    assert sum(l.num_outputs for l in c.layers) == 54366
    assert c.layers[2].desc.activation == pb.LeakyRelu
    assert c.VerifyArchitecture()
    with open(os.path.join(ROOT, "oracle", "specs", "cifar.spec")) as f:
        want = "\n".join(l for l in f.read().splitlines() if l and not l.startswith("#")) + "\n"
    assert c.to_spec(pb.CrossEntropy) == want

    bad = pb.Architecture(4)
    bad.AddMaxPoolLayer(pb.VolumeDimensions(4, 4, 3), pb.AreaDimensions(2, 2))
    assert not bad.VerifyArchitecture()  # 4 != 48
    with pytest.raises(_lib.PlasticityError):
        pb.Architecture(9).AddConvolutionLayer(pb.VolumeDimensions(3, 3, 1),
                                               pb.FilterParams(3, 3, 2, 1, 1, 1))
    with pytest.raises(_lib.PlasticityError):
        pb.Architecture(3).AddActivationLayer(17)  # arbitrary std::function: refused


def test_symbol_generator_weight_numbers():
    """nnet_test.cc:486-581: conv weight layout for 3x3x3, 2 filters."""
    s = pb.ConvSymbolGenerator(pb.VolumeDimensions(5, 5, 3), pb.FilterParams(3, 3, 3, 2, 1, 2))
    n = 0
    for f in range(2):
        for z in range(3):
            for row in range(3):
                for col in range(3):
                    assert s.WeightNumber(f, row, col, z) == n
                    n += 1
        assert s.WeightNumber(f) == n  # bias at 27 / 55
        n += 1
    assert s.WeightNumber(0) == 27 and s.WeightNumber(1) == 55
    d = pb.DenseSymbolGenerator(3, 3)  # nnet_test.cc:33-57
    assert [d.WeightNumber(1, e) for e in range(3)] == [4, 5, 6] and d.WeightNumber(1) == 7


def test_shard_range_covers_batch():
    for n in (1, 7, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


DP_WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests'))
from oracle import pyoracle as po
from plasticity_b200 import parallel
import util
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:{port}', rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
ident = parallel.exchange_unique_id(dist, lambda: bytes(range(128)), 128)
assert ident == bytes(range(128))
n = po.RestatedNet(po.spec_text('mixed'))
rng = np.random.default_rng(0)            # same data on both ranks
w0 = util.xavier_like(n, rng)
B = 9
xs = rng.normal(0, 1, (B, n.n_in)); es = rng.uniform(0, 1, (B, n.n_out))
b0, b1 = parallel.shard_range(B, rank, 2)
w = w0.copy()
delta = n.batch_train(xs[b0:b1].copy(), es[b0:b1].copy(), w.copy(), 0.05)   # -lr * sum over the shard
t = torch.from_numpy(delta.copy())
dist.all_reduce(t)                                                          # the one exchange step
w_dp = w0 + t.numpy()
w_full = w0.copy(); n.batch_train(xs, es, w_full, 0.05)
assert np.allclose(w_dp, w_full, rtol=1e-12, atol=1e-15), np.abs(w_dp - w_full).max()
dist.barrier(); dist.destroy_process_group()
print('rank', rank, 'ok')
"""


def test_data_parallel_gradient_sum_world2_gloo(tmp_path):
    """Two ranks each sum -lr*grad over their shard, all-reduce the flat delta
    buffer, apply: identical to BatchTrain over the global batch."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "dp_worker.py"
    script.write_text(DP_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r}:\n{o}"
        assert f"rank {r} ok" in o


def test_yolov1_architecture_is_the_repaired_reference_layer_list():
    """nnet/yolov1.cc:7-231 with the three repairs of SURVEY 8(d): the layer list chains,
    and its totals are the ones the survey pins for configuration 5."""
    from plasticity_b200 import yolov1
    m = yolov1.YoloArchitecture()
    assert m.VerifyArchitecture()
    assert len(m.layers) == 55
    assert m.input_size() == 448 * 448 * 3 and m.output_size() == 7 * 7 * 30
    convs = [l for l in m.layers if l.desc.kind == _lib.PB_CONVOLUTION]
    pools = [l for l in m.layers if l.desc.kind == _lib.PB_MAX_POOL]
    dense = [l for l in m.layers if l.desc.kind == _lib.PB_DENSE]
    assert (len(convs), len(pools), len(dense)) == (23, 4, 2)
    assert sum(l.weights.size for l in m.layers) == 266721790
    macs = sum(l.num_outputs * l.desc.filter_width * l.desc.filter_height * l.desc.filter_depth
               for l in convs) + sum(l.num_inputs * l.num_outputs for l in dense)
    assert macs == 16534396928  # 33.07 GFLOP forward per image
    assert max(l.num_outputs for l in m.layers) == 3211264
    # repair 1: conv 2 keeps 112x112 (padding 1); repair 2: depth chains through the
    # second 14x14 pair; repair 3: the model ends at the 1470-wide dense layer
    assert (convs[1].desc.padding, convs[1].num_outputs) == (1, 112 * 112 * 192)
    assert [c.desc.depth for c in convs[15:19]] == [512, 512, 1024, 512]
    assert dense[0].desc.in_ == 7 * 7 * 1024 and dense[1].desc.out == 1470
    # every convolution is followed by LeakyRelu, every dense layer by Sigmoid (builder defaults)
    for i, l in enumerate(m.layers[:-1]):
        if l.desc.kind == _lib.PB_CONVOLUTION:
            assert m.layers[i + 1].desc.activation == pb.LeakyRelu
        if l.desc.kind == _lib.PB_DENSE:
            assert m.layers[i + 1].desc.activation == pb.Sigmoid
    small = yolov1.YoloArchitecture(64, 128, 30)
    assert small.VerifyArchitecture() and small.output_size() == 30
    with pytest.raises(ValueError):
        yolov1.YoloArchitecture(100)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU path: generated kernels or the
    plain-C port, all host threads) prints exactly ONE JSON line on stdout carrying the
    contract's keys for the same metric / config as the product arm."""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--steps", "1", "--warmup", "0", "--ref-sample", "16"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cifar_cnn_train_samples_per_sec"
    assert d["unit"] == "samples/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "cifar_cnn_batch_train"


def test_cpp_mirror_headers_compile_standalone(tmp_path):
    """Every header of the C++ host mirror (plasticity_b200/cpp: the reference's include
    paths — nnet/*.h, compute/cl_buffer.h, compute/command_queue.h + the CUDA command queue)
    compiles on its own with the host compiler: a reference-side caller needs nothing but
    -I<repo>/plasticity_b200/cpp -I<repo>/include."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cpp = os.path.join(root, "plasticity_b200", "cpp")
    headers = sorted(os.path.relpath(os.path.join(d, f), cpp)
                     for d, _, fs in os.walk(cpp) for f in fs if f.endswith(".h"))
    assert "compute/command_queue.h" in headers and "nnet/nnet.h" in headers
    for h in headers:
        tu = tmp_path / "tu.cc"
        tu.write_text(f'#include "{h}"\nint main() {{ return 0; }}\n')
        r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-w", "-I", cpp,
                            "-I", os.path.join(root, "include"), str(tu)],
                           capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, f"{h} does not compile on its own:\n{r.stderr[-1500:]}"


def test_abi_header_is_plain_c(tmp_path):
    """include/plasticity_b200.h is the drop-in boundary: it must compile as plain C (no C++,
    no CUDA, no torch types) so that cgo / JNI / ctypes-style bindings can consume it."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tu = tmp_path / "abi.c"
    tu.write_text('#include "plasticity_b200.h"\nint main(void) { return pb_abi_version() < 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I",
                        os.path.join(root, "include"), str(tu)],
                       capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr[-2000:]
