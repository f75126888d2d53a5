"""Multi-GPU parity of the data-parallel path (run under torchrun, one rank per GPU):
N ranks train an MLP on row shards of one global batch with dp.train_step (NCCL all-reduce of parameter gradients,
replicated Adam) and rank 0 compares losses and final weights with the numpy oracle training on the FULL batch in one
process.  Bar: 1e-12 relative (fp64; the all-reduce only re-associates the batch sum)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import autodiff_b200 as ad
from autodiff_b200 import dp
import oracle.np_autodiff as onp

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
ad.init(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sizes, gb, steps = [40, 64, 32, 10], 96, 8
if "big" in sys.argv[1:]:      # a parameter of >= 1 M elements: dp.train_step queues each layer's optimizer update on a side stream
    sizes = [1100, 1024, 10]
rng = np.random.default_rng(7)
xs = [rng.standard_normal((gb, sizes[0])) for _ in range(steps)]
ys = [np.eye(sizes[-1])[rng.integers(0, sizes[-1], gb)] for _ in range(steps)]

np.random.seed(11)
nn = ad.NN(sizes)
opt = ad.Adam(len(nn.w), lr=1e-2)
lo, hi = dp.shard_rows(gb, world, rank)
losses = []
if "deferred" in sys.argv[1:]:  # every loss read one step late (dp.PendingLoss): the host builds step k+1 while the GPU runs step k
    pending = [dp.train_step(nn, opt, xs[s][lo:hi], ys[s][lo:hi], gb, dist if world > 1 else None, read_loss="deferred") for s in range(steps)]
    losses = [p.value() for p in pending]
else:
    for s in range(steps):
        losses.append(dp.train_step(nn, opt, xs[s][lo:hi], ys[s][lo:hi], gb, dist if world > 1 else None, read_loss=True))
w_gpu = [np.asarray(w.value if not isinstance(w.value, ad.DeviceArray) else ad.to_host(w.value)) for w in nn.w]

if rank == 0:
    np.random.seed(11)
    rnn = onp.NN(sizes)
    ropt = onp.Adam(len(rnn.w), lr=1e-2)
    rlosses = []
    for s in range(steps):
        cost = onp.Einsum("i->", onp.SoftmaxCEWithLogits(labels=onp.Variable(ys[s], name="y"), logits=rnn(onp.Variable(xs[s], name="x")))) / gb
        grads = onp.grad(cost, rnn.w)
        rlosses.append(float(cost()))
        new = ropt([w() for w in rnn.w], [g() for g in grads])
        for w, v in zip(rnn.w, new):
            w.value = v
    err_l = max(abs(a - b) / abs(b) for a, b in zip(losses, rlosses))
    err_w = max(float(np.max(np.abs(a - np.asarray(b.value))) / np.max(np.abs(np.asarray(b.value)))) for a, b in zip(w_gpu, rnn.w))
    from autodiff_b200.core import tape
    print({"world": world, "sizes": sizes, "steps": steps, "args": sys.argv[1:], "loss_rel_err": err_l, "weight_rel_err": err_w, "tape": dict(tape.stats), "ok": bool(err_l < 1e-12 and err_w < 1e-12)})
    assert err_l < 1e-12 and err_w < 1e-12
if world > 1:
    dist.destroy_process_group()
