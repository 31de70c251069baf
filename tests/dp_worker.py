"""One rank of the multi-process data-parallel parity check (launched by
tests/test_gpu_dp.py through torch.distributed.run, one rank per GPU): every rank trains
the same seeded global batches on its shard through pb_nnet_batch_train_dp (NCCL handshake,
CUDA-IPC peer memory, the fused exchange-and-update kernel) and writes its final weights."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist

import plasticity_b200 as pb
from plasticity_b200 import _lib as L, parallel
from oracle import pyoracle as po
import util


def main():
    name, out_dir = sys.argv[1], sys.argv[2]
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    data = np.load(os.path.join(out_dir, "inputs.npz"))
    spec = str(data["spec"])
    S = po.RestatedNet(spec)
    model, loss = util.architecture_from_spec(spec)
    xs, es, w0, lr = data["xs"], data["es"], data["w0"], float(data["lr"])
    steps, B = xs.shape[0], xs.shape[1]
    ctx = pb.nnet.context(local)
    raw = parallel.exchange_unique_id(
        dist, lambda: (lambda b: (L.call("pb_comm_unique_id", b), bytes(b.raw))[1])(
            (C.c_char * L.PB_COMM_ID_BYTES)()),
        L.PB_COMM_ID_BYTES, device=torch.device("cuda", local))
    comm = C.c_void_p()
    L.call("pb_comm_create", ctx, raw, rank, world, C.byref(comm))
    lo, hi = parallel.shard_range(B, rank, world)
    net = pb.Nnet(model, pb.NoWeightInit, loss, max_batch=hi - lo, device=local)
    net.SetLearningParameters(pb.Nnet.LearningParameters(lr))
    util.load_product_weights(net, S, w0)
    for t in range(steps):
        net.BatchTrain(net.MakeBuffer(xs[t, lo:hi].reshape(-1)), net.MakeBuffer(es[t, lo:hi].reshape(-1)),
                       hi - lo, comm=comm)
    np.save(os.path.join(out_dir, f"weights_rank{rank}.npy"), util.read_product_weights(net, S))
    dist.barrier()
    L.call("pb_comm_destroy", comm)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
