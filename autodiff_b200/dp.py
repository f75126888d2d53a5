"""Batch-sharded data-parallel training over one 8xB200 box (new capability; the reference is single-process).

One process per GPU (torchrun), `torch.distributed` NCCL for the plumbing.  In every training graph the batch
index is a free letter of the forward einsums ('ij,jk->ik') and of all row-wise nodes, so forward and
activation-gradient einsums need no communication; the weight-gradient einsum contracts OVER the batch letter
('ij,ik->jk', reference ops.py:182-213), so each rank holds a partial sum and ONE all-reduce(sum, fp64) per
parameter per step makes every rank's gradient equal to the single-process one (each rank divides its loss by
the GLOBAL batch, as examples/feedforward_mnist.py:25 divides by a constant).  The all-reduce of layer L is
launched as soon as its dW kernel is queued and overlaps the remaining backward GEMMs; the optimizer runs
replicated (identical inputs => identical weights, no broadcast after step 0).
"""
import os
import time

import numpy as np

from .core.engine import check_pending_flags, evaluate_many
from .core.grad import grad
from .core.high_level_ops import NN, Adam
from .core.node import Variable
from .core.ops import Einsum, SoftmaxCEWithLogits
from .device import DeviceArray, pinned_empty, synchronize, to_device, to_host, torch


def world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))


def shard_rows(n_rows, world_size, rank):
    """Contiguous row block [lo, hi) of this rank."""
    per = (n_rows + world_size - 1) // world_size
    lo = min(rank * per, n_rows)
    return lo, min(lo + per, n_rows)


def _flat_view(d):
    """torch view of the storage span a DeviceArray occupies (it must be a whole dense/pitched allocation)."""
    span = 1
    for dim, st in zip(d.shape, d.strides):
        if dim > 1:
            span += (dim - 1) * abs(st)
    return d.base[d.offset:d.offset + span]


def allreduce_async(dev_array, dist):
    """Launches ncclAllReduce(sum) on the array's storage; returns the work handle."""
    return dist.all_reduce(_flat_view(dev_array), op=dist.ReduceOp.SUM, async_op=True)


def train_step(nn, opt, x, y, global_batch, dist=None, read_loss=True):
    """One step of examples/feedforward_mnist.py:17-34 on this rank's shard.  x, y: host ndarrays or DeviceArrays."""
    xv = Variable(x, name="images")
    yv = Variable(y, name="labels")
    logits = nn(xv)
    sce = SoftmaxCEWithLogits(labels=yv, logits=logits)
    cost = Einsum("i->", sce) / global_batch
    grads = grad(cost, nn.w)
    works = []

    def ready(node):                               # fires when a gradient's kernels are queued
        if dist is not None and node is not cost:
            d = node._dev
            if d.size <= 16 or any(st == 0 and dim > 1 for dim, st in zip(d.shape, d.strides)):
                # a gradient that ALIASES a shared constant (broadcast ones, a cached tiny upload) must not be reduced
                # in place: give it storage of its own first
                from .device import copy_device
                own = DeviceArray.empty(d.shape)
                copy_device(d, own)
                node._dev = d = own
            works.append(allreduce_async(d, dist))
    # one fusion plan (and one launch tape) over all outputs; last layer's dW is needed first, so it is evaluated
    # first and its all-reduce overlaps the rest of the backward pass
    evaluate_many(list(reversed(grads)) + ([cost] if read_loss else []), on_ready=ready)
    gdev = [g.eval_device() for g in grads]
    for w in works:
        w.wait()
    new_w = opt(list(nn.w), gdev)
    opt.apply_new_weights(nn.w, new_w)
    loss = None
    if read_loss:
        lv = cost.eval_device()
        if dist is not None:
            dist.all_reduce(_flat_view(lv), op=dist.ReduceOp.SUM)
        loss = float(to_host(lv))
        check_pending_flags()                      # deferred label validation of SoftmaxCEWithLogits (reference ops.py:246-248)
    return loss


def _mlp_flops(sizes, batch):
    """GEMM flops of one step: per layer fwd + dW, + dX for every layer but the first (never evaluated)."""
    f = 0.0
    for i, (a, b) in enumerate(zip(sizes, sizes[1:])):
        f += 2.0 * batch * (a + 1) * b * (3 if i > 0 else 2)
    return f


def bench_mlp(args, dist):
    """MLP training samples/s for BASELINE configs[0] (784-100-10, b=64; rank 0 only, host inputs every step) and
    configs[2] (4096 x 8 layers, global batch 8192 sharded over all ranks, NCCL grad all-reduce)."""
    import torch as t
    ws, rank = world()
    out = {}

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if dist is not None:
            dist.barrier()
        t.cuda.synchronize()
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if dist is not None:
            dist.barrier()
        t.cuda.synchronize()
        ms = max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3) / steps
        if dist is not None:
            tt = t.tensor([ms], dtype=t.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms

    # ---- C1: latency-bound small MLP, through the host API every step (rank 0 only)
    if rank == 0 and ws == 1:
        np.random.seed(1337)
        sizes, batch = [784, 100, 10], 64
        nn = NN(sizes)
        opt = Adam(len(nn.w), lr=1e-3)
        rng = np.random.default_rng(1337)
        xs = pinned_empty((batch, sizes[0])); xs[...] = rng.random((batch, sizes[0]))
        ys = pinned_empty((batch, sizes[-1])); ys[...] = np.eye(sizes[-1])[rng.integers(0, sizes[-1], batch)]
        ms = timed(lambda: train_step(nn, opt, xs, ys, batch, None, read_loss=True), 50, 10)
        out["c1_mlp_784_100_10_b64"] = {"samples_per_s": round(batch / (ms * 1e-3), 1), "ms_per_step": round(ms, 4),
                                        "includes": "graph build + ad.grad + H2D of the batch + kernels + Adam on device + D2H of the loss"}

    # ---- C3: 4096-wide x 8 layers, global batch 8192, strong-scaled over the ranks
    np.random.seed(1337)
    sizes, gbatch = [4096] * 9, 8192
    nn = NN(sizes)
    for w in nn.w:                     # weights resident in HBM from the start
        w.value = to_device(w.value)
    opt = Adam(len(nn.w), lr=1e-3)
    lo, hi = shard_rows(gbatch, ws, rank)
    rng = np.random.default_rng(1337)
    x = rng.standard_normal((gbatch, sizes[0]))[lo:hi]
    lab = rng.integers(0, sizes[-1], gbatch)[lo:hi]
    y = np.zeros((hi - lo, sizes[-1])); y[np.arange(hi - lo), lab] = 1.0
    xd, yd = to_device(x), to_device(y)
    del x, y
    steps = max(2, min(args.steps, 5))
    ms = timed(lambda: train_step(nn, opt, xd, yd, gbatch, dist, read_loss=True), steps, 4)     # (tape: recorded on step 2, first replay on step 3)
    fl = _mlp_flops(sizes, gbatch)
    out["c3_mlp_4096x8_b8192"] = {"samples_per_s": round(gbatch / (ms * 1e-3), 1), "ms_per_step": round(ms, 3), "n_gpus": ws,
                                  "global_batch": gbatch, "scaling": "strong", "gemm_tflops": round(fl / (ms * 1e-3) / 1e12, 2),
                                  "allreduce_bytes_per_step": sum((a + 1) * b * 8 for a, b in zip(sizes, sizes[1:])),
                                  "includes": "graph build + ad.grad + kernels + per-layer NCCL all-reduce overlapped with backward + "
                                              "Adam on device + all-reduced loss read back to the host; inputs resident in HBM"}
    return out
