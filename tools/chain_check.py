"""Race check for the chained (programmatic dependent launch) training step on a GPU:
the CIFAR network trains the same seeded batches for many gradient-sum steps with
PB_STEP_CHAIN=1 and with PB_STEP_CHAIN=0 (each in its own process: the switch is read
once); every kernel of the step is deterministic, so the two weight sets must be
bit-identical.  A kernel that read its inputs before the launch ahead of it had finished
would show up here.
  python tools/chain_check.py [steps] [batch]"""
import hashlib, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def worker(steps, batch):
    import plasticity_b200 as pb
    from oracle import pyoracle as po
    import util
    spec = po.spec_text("cifar")
    S = po.RestatedNet(spec)
    model, loss = util.architecture_from_spec(spec)
    rng = np.random.default_rng(7)
    w0 = util.xavier_like(S, rng)
    nb = 4                                       # resident batches, revisited -> graph replay
    xs = util.f32(rng.normal(0, 1, (nb, batch, S.n_in)))
    es = np.zeros((nb, batch, S.n_out))
    es[np.arange(nb)[:, None], np.arange(batch)[None, :], rng.integers(0, 10, (nb, batch))] = 1.0
    net = pb.Nnet(model, pb.NoWeightInit, loss, max_batch=batch)
    net.SetLearningParameters(pb.Nnet.LearningParameters(1e-4))
    util.load_product_weights(net, S, w0)
    bufs = [(net.MakeBuffer(xs[i].reshape(-1)), net.MakeBuffer(es[i].reshape(-1))) for i in range(nb)]
    for b in bufs:
        b[0].MoveToGpu(); b[1].MoveToGpu()
    for t in range(steps):
        x, e = bufs[t % nb]
        net.BatchTrain(x, e, batch)
    w = util.read_product_weights(net, S)
    assert np.isfinite(w).all()
    print("WEIGHTS", hashlib.sha256(np.ascontiguousarray(w).tobytes()).hexdigest(),
          float(np.abs(w - w0).max()), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--worker":
        worker(int(sys.argv[2]), int(sys.argv[3]))
        sys.exit(0)
    steps = sys.argv[1] if len(sys.argv) > 1 else "200"
    batch = sys.argv[2] if len(sys.argv) > 2 else "1024"
    got = {}
    for chain in ("1", "0", "1"):
        env = dict(os.environ, PB_STEP_CHAIN=chain)
        out = subprocess.run([sys.executable, __file__, "--worker", steps, batch], env=env,
                             capture_output=True, text=True, timeout=600)
        line = [l for l in out.stdout.splitlines() if l.startswith("WEIGHTS")]
        assert out.returncode == 0 and line, (out.returncode, out.stdout[-2000:], out.stderr[-2000:])
        print(f"PB_STEP_CHAIN={chain}: {line[-1]}", flush=True)
        got.setdefault(chain, []).append(line[-1].split()[1])
    same = len({h for v in got.values() for h in v}) == 1
    print("chained == unchained, bit for bit:", same, flush=True)
    sys.exit(0 if same else 1)
