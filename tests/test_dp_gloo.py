"""World-size-2 check (gloo, CPU) of the data-parallel algebra used by autodiff_b200/dp.py: shard rows, divide each
rank's loss by the GLOBAL batch, all-reduce(sum) the parameter gradients => the single-process gradient.
The arithmetic here runs on the numpy oracle (no GPU in this container); the sharding / reduction logic
(dp.shard_rows, the sum all-reduce) is the code under test."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import oracle.np_autodiff as onp
    from autodiff_b200 import dp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gb, sizes = 12, [6, 5, 4]
    rng = np.random.default_rng(0)
    x = rng.standard_normal((gb, sizes[0]))
    y = np.eye(sizes[-1])[rng.integers(0, sizes[-1], gb)]

    def grads_for(xs, ys):
        np.random.seed(3)
        nn = onp.NN(sizes)
        cost = onp.Einsum("i->", onp.SoftmaxCEWithLogits(labels=onp.Variable(ys, name="y"), logits=nn(onp.Variable(xs, name="x")))) / gb
        return [np.array(g()) for g in onp.grad(cost, nn.w)], float(cost())
    lo, hi = dp.shard_rows(gb, world, rank)
    part, loss = grads_for(x[lo:hi], y[lo:hi])
    full, full_loss = grads_for(x, y)
    err = 0.0
    for p, f in zip(part, full):
        t = torch.from_numpy(p.copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        err = max(err, float(np.max(np.abs(t.numpy() - f)) / np.max(np.abs(f))))
    lt = torch.tensor([loss], dtype=torch.float64)
    dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    q.put((rank, lo, hi, err, abs(float(lt.item()) - full_loss)))
    dist.destroy_process_group()


def test_sharded_gradients_sum_to_full_batch_gradient():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 6), (6, 12)]
    for r in res:
        assert r[3] < 1e-13 and r[4] < 1e-13, r


def test_shard_rows_partitions_exactly():
    from autodiff_b200 import dp
    for n in (1, 7, 64, 8192):
        for w in (1, 2, 3, 4, 8):
            spans = [dp.shard_rows(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
