"""Data-parallel parity on real GPUs (run under torchrun, one rank per GPU):
ranks train 3 gradient-sum steps on shards of a global batch; the result must be
bit-identical across ranks and match a single-GPU run over the whole batch.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py [cifar|mlp]
cifar: 22 466 weights -> the fused NVLink peer-memory exchange; mlp (width 1200, 2.9 M
weights) -> the per-layer NCCL all-reduce overlapped with the backward pass."""
import ctypes as C, os, sys, zlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import plasticity_b200 as pb
from plasticity_b200 import _lib as L, parallel
from oracle import pyoracle as po
import util

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name = sys.argv[1] if len(sys.argv) > 1 else "cifar"
if name == "mlp":
    Wd = 1200
    spec = (f"loss ce\nact {Wd} identity\ndense {Wd} {Wd}\nact {Wd} sigmoid\ndense {Wd} {Wd}\n"
            f"act {Wd} sigmoid\ndense {Wd} 10\nact 10 identity\nsoftmax 10\n")
else:
    spec = po.spec_text(name)
S = po.RestatedNet(spec)
model, loss = util.architecture_from_spec(spec)
rng = np.random.default_rng(123)            # same stream on every rank
w0 = util.xavier_like(S, rng)
# one step for the CNN: later steps amplify ulp-level differences through max-pool routing
B, steps, lr = 64 * world, (1 if name != "mlp" else 3), 0.002
xs = util.f32(rng.normal(0, 1, (steps, B, S.n_in)))
es = util.f32(rng.uniform(0, 1, (steps, B, S.n_out)))
ctx = pb.nnet.context(local)
raw = parallel.exchange_unique_id(dist, lambda: (lambda b: (L.call("pb_comm_unique_id", b), bytes(b.raw))[1])((C.c_char * L.PB_COMM_ID_BYTES)()),
                                  L.PB_COMM_ID_BYTES, device=torch.device("cuda", local))
comm = C.c_void_p()
L.call("pb_comm_create", ctx, raw, rank, world, C.byref(comm))
lo, hi = parallel.shard_range(B, rank, world)
net = pb.Nnet(model, pb.NoWeightInit, loss, max_batch=hi - lo, device=local)
net.SetLearningParameters(pb.Nnet.LearningParameters(lr))
util.load_product_weights(net, S, w0)
for t in range(steps):
    net.BatchTrain(net.MakeBuffer(xs[t, lo:hi].reshape(-1)), net.MakeBuffer(es[t, lo:hi].reshape(-1)), hi - lo, comm=comm)
w = util.read_product_weights(net, S)
allw = [torch.zeros(w.size, dtype=torch.float64, device="cuda") for _ in range(world)]
dist.all_gather(allw, torch.tensor(w, device="cuda"))
same = all(torch.equal(allw[0], a) for a in allw)
msg = f"{name} rank {rank}: ranks bit-identical={same}"
if rank == 0:
    ref = pb.Nnet(model, pb.NoWeightInit, loss, max_batch=B, device=local)
    ref.SetLearningParameters(pb.Nnet.LearningParameters(lr))
    util.load_product_weights(ref, S, w0)
    for t in range(steps):
        ref.BatchTrain(ref.MakeBuffer(xs[t].reshape(-1)), ref.MakeBuffer(es[t].reshape(-1)), B)
    wr = util.read_product_weights(ref, S)
    # weights agree to fp32 rounding (the single-GPU step updates W in place, W - lr*g in
    # one rounding; the data-parallel step adds the summed rank deltas: the two round the
    # WEIGHT differently by an ulp, which is ~1e-5..1e-4 of a small update)
    err_w = np.abs(w - wr).max() / np.abs(wr).max()
    err = np.abs(w - wr).max() / np.abs(wr - w0).max()
    msg += (f"; vs single-GPU global batch: max|w-w_ref|/max|w_ref| = {err_w:.2e}, "
            f"max|dw-dw_ref|/max|dw_ref| = {err:.2e}")
    assert err_w < 2e-6 and err < 1e-3, (err_w, err)
assert same
print(msg, flush=True)
dist.barrier()
L.call("pb_comm_destroy", comm)
dist.destroy_process_group()
