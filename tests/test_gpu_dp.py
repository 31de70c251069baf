"""Data-parallel gradient-sum training (the one exchange step of the path, SURVEY 8e;
semantics nnet/nnet.h:624-669: W <- W - lr * sum over the GLOBAL batch) on real GPUs.

* in-process rank groups (pb_comm_create_local): 2 and 3 ranks share cuda:0 and run the same
  fused exchange-and-update kernel over peer memory as the multi-process path;
* with two or more GPUs, additionally the real thing: one process per GPU under
  torch.distributed.run (NCCL handshake, CUDA IPC).
Every variant must leave bit-identical weights on all ranks, equal to the oracle's
batch_train over the global batch within |rel| <= 1e-4."""
import ctypes as C
import os
import subprocess
import sys
import zlib

import numpy as np
import pytest

import plasticity_b200 as pb
from plasticity_b200 import _lib as L, parallel
from oracle import pyoracle as po
import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def device_count():
    n = C.c_int(0)
    L.call("pb_device_count", C.byref(n))
    return n.value


def make_case(name, steps, B, lr, seed):
    """Seeded weights and `steps` global batches with a clear max-pool winner under the
    oracle's weights of that step, plus the oracle's weights after the steps."""
    spec = po.spec_text(name)
    S = po.RestatedNet(spec)
    rng = np.random.default_rng(seed)
    w0 = util.xavier_like(S, rng)
    es = util.f32(rng.uniform(0, 1, (steps, B, S.n_out)))
    w_ref = w0.copy()
    xs = []
    for t in range(steps):
        x = util.draw_with_pool_margin(S, rng, B, w_ref)
        xs.append(x)
        S.batch_train(x.copy(), es[t].copy(), w_ref, lr)
    return spec, S, w0, np.stack(xs), es, w_ref


@pytest.mark.parametrize("name,world,B", [("cifar", 2, 64), ("cifar", 3, 96), ("mixed", 2, 16)])
def test_data_parallel_in_process_ranks(name, world, B):
    steps, lr = 3, 0.002
    spec, S, w0, xs, es, w_ref = make_case(name, steps, B, lr, zlib.crc32(name.encode()) + world)
    model, loss = util.architecture_from_spec(spec)
    # all in-process ranks on cuda:0 (the library keeps ONE current device per process, as in
    # its one-process-per-GPU deployment); several devices are the multi-process test's job
    ndev = 1
    ctxs = [pb.nnet.new_context(r % ndev) for r in range(world)]
    comms = pb.nnet.local_comms(ctxs)
    nets = []
    for r in range(world):
        lo, hi = parallel.shard_range(B, r, world)
        net = pb.Nnet(model, pb.NoWeightInit, loss, max_batch=hi - lo, device=r % ndev, ctx=ctxs[r])
        net.SetLearningParameters(pb.Nnet.LearningParameters(lr))
        util.load_product_weights(net, S, w0)
        nets.append((net, lo, hi))
    # more steps than inbox slots and than the 2 eager runs before a step is replayed as a graph
    # device buffers first: nothing between one rank's step and the next rank's may block
    bufs = {}
    for t in range(steps):
        for r, (net, lo, hi) in enumerate(nets):
            bx = net.MakeBuffer(xs[t, lo:hi].reshape(-1))
            be = net.MakeBuffer(es[t, lo:hi].reshape(-1))
            bx.MoveToGpu()
            be.MoveToGpu()
            bufs[t, r] = (bx, be)
    for rep in range(2):
        for t in range(steps):
            for r, (net, lo, hi) in enumerate(nets):
                # queued rank after rank from this one thread; the ranks meet on the device
                net.BatchTrain(*bufs[t, r], hi - lo, comm=comms[r])
        ws = [util.read_product_weights(net, S) for net, _, _ in nets]
        for r in range(1, world):
            assert np.array_equal(ws[0], ws[r]), f"rank {r} differs from rank 0 (rep {rep})"
        if rep == 0:
            util.assert_close(ws[0], w_ref, f"{name} x{world} data-parallel vs oracle global batch")
            for net, _, _ in nets:   # second round: same buffers again -> recorded graphs replay
                util.load_product_weights(net, S, w0)
    util.assert_close(ws[0], w_ref, f"{name} x{world} data-parallel (graph replay) vs oracle")
    for c in comms:
        L.call("pb_comm_destroy", c)


def test_data_parallel_missing_peer_is_an_error_not_a_corrupt_update():
    """A rank whose peer never arrives must not apply a partial sum: the exchange kernel
    gives up after 5 s, leaves W untouched and every later synchronising call fails."""
    spec, S, w0, xs, es, _ = make_case("mixed", 1, 16, 0.002, 99)
    model, loss = util.architecture_from_spec(spec)
    ctxs = [pb.nnet.new_context(0) for _ in range(2)]
    comms = pb.nnet.local_comms(ctxs)
    net = pb.Nnet(model, pb.NoWeightInit, loss, max_batch=8, device=0, ctx=ctxs[0])
    net.SetLearningParameters(pb.Nnet.LearningParameters(0.002))
    util.load_product_weights(net, S, w0)
    bx, be = net.MakeBuffer(xs[0, :8].reshape(-1)), net.MakeBuffer(es[0, :8].reshape(-1))
    net.BatchTrain(bx, be, 8, comm=comms[0])   # rank 1 never steps
    with pytest.raises(L.PlasticityError, match="peer rank did not arrive"):
        net.Finish()
    dW = C.c_void_p()
    L.call("pb_nnet_weights_device", net.h, C.byref(dW))
    n_flat = sum((n + 3) & ~3 for n in S.nw)
    got = np.zeros(n_flat, dtype=np.float32)
    L.call("pb_buffer_read_f32", ctxs[0], dW, got.ctypes.data_as(C.c_void_p), n_flat)
    flat = got.astype(np.float64)
    # the flat buffer pads every layer's block to 4 floats; compare layer by layer
    off = 0
    for l in range(S.n_layers):
        n = S.nw[l]
        if n:
            assert np.array_equal(flat[off:off + n], S.layer_weights(w0, l)), f"layer {l} changed"
        off += (n + 3) & ~3
    with pytest.raises(L.PlasticityError):
        net.BatchTrain(bx, be, 8, comm=comms[0])


def test_data_parallel_one_process_per_gpu(tmp_path):
    """The multi-process path (one rank per GPU, NCCL + CUDA IPC).  Needs two devices; on a
    single-GPU box the in-process variant above is the coverage and this case only checks
    that a 1-rank communicator is exactly the single-GPU step."""
    ndev = device_count()
    world = min(ndev, 4)
    name, steps, B, lr = "cifar", 3, 48 * max(world, 1), 0.002
    spec, S, w0, xs, es, w_ref = make_case(name, steps, B, lr, 2024 + world)
    np.savez(tmp_path / "inputs.npz", xs=xs, es=es, w0=w0, lr=lr, spec=spec)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(ROOT, "tests", "dp_worker.py"), name, str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    ws = [np.load(tmp_path / f"weights_rank{k}.npy") for k in range(world)]
    for k in range(1, world):
        assert np.array_equal(ws[0], ws[k]), f"rank {k} differs from rank 0"
    util.assert_close(ws[0], w_ref, f"{name}: {world} process(es) vs oracle global batch")
