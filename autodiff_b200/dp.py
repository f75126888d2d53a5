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
import math
import os
import time

import numpy as np

from .core.engine import check_pending_flags, evaluate_many
from .core.grad import grad
from .core.high_level_ops import NN, Adam
from .core.node import Node, Variable
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


_opt_stream = None


def _optimizer_stream():
    global _opt_stream
    if _opt_stream is None:
        _opt_stream = torch().cuda.Stream()
    return _opt_stream


def train_step(nn, opt, x, y, global_batch, dist=None, read_loss=True, overlap_optimizer=True):
    """One step of examples/feedforward_mnist.py:17-34 on this rank's shard.  x, y: host ndarrays or DeviceArrays."""
    def build_cost():
        xv = Variable(x, name="images")
        yv = Variable(y, name="labels")
        sce = SoftmaxCEWithLogits(labels=yv, logits=nn(xv))
        return Einsum("i->", sce) / global_batch
    return step(nn.w, opt, build_cost, dist, read_loss, overlap_optimizer)


def step(params, opt, build_cost, dist=None, read_loss=True, overlap_optimizer=True, broadcast_check=False):
    """One data-parallel training step of ANY model written against the reference API: `build_cost()` returns the loss node
    of this rank's shard, already divided by the GLOBAL batch (the reference divides by a constant, feedforward_mnist.py:25),
    `params` are the Variables to train.  Every parameter gradient is all-reduced (sum) in place as soon as its kernels are
    queued, the optimizer runs replicated.  With `overlap_optimizer` (and an optimizer that has `update_one`) the update of
    each parameter is queued on a side stream as soon as its gradient is reduced, so the HBM-bound update kernels of layer L
    run under the DMMA-bound backward GEMMs of the layers below it (same kernels, same operands: bit-identical results -
    tests/test_gpu_parity.py::test_dp_overlap_schedules_bitwise)."""
    from . import _lib, device as _device
    params = list(params)
    cost = build_cost()
    grads = grad(cost, params)
    works = []
    # (worth two stream switches per parameter only when the update kernels are not tiny: >= 1 M elements somewhere)
    overlap = (overlap_optimizer and opt is not None and hasattr(opt, "update_one") and max(math.prod(w.shape) for w in params) >= (1 << 20)
               and all(isinstance(g, Node) for g in grads))
    new_w = [None] * len(grads)
    if overlap:
        index_of = {id(g): k for k, g in enumerate(grads)}
        t = torch()
        main = t.cuda.current_stream()
        side = _optimizer_stream()

    def ready(node):                               # fires when a gradient's kernels are queued
        if node is cost:
            return
        work = None
        if dist is not None:
            d = node._dev
            if d.size <= 16 or any(st == 0 and dim > 1 for dim, st in zip(d.shape, d.strides)):
                # a gradient that ALIASES a shared constant (broadcast ones, a cached tiny upload) must not be reduced
                # in place: give it storage of its own first
                from .device import copy_device
                own = DeviceArray.empty(d.shape)
                copy_device(d, own)
                node._dev = d = own
            work = allreduce_async(d, dist)
        k = index_of.get(id(node)) if overlap else None
        if k is None or new_w[k] is not None:
            if work is not None:
                works.append(work)
            return
        # the update of parameter k on the side stream: behind the gradient's kernels (and its all-reduce), beside
        # everything the main stream queues from here on.  A recording launch tape must not log these calls.
        done = t.cuda.Event()
        done.record(main)
        saved = (_lib._call_hook, _device._alloc_hook)
        _lib._call_hook = None
        _device._alloc_hook = lambda base: base.record_stream(main)      # the new weights / moments are read on `main` next step
        try:
            with t.cuda.stream(side):
                side.wait_event(done)
                if work is not None:
                    work.wait()
                _device.use_stream(side)
                new_w[k] = opt.update_one(k, params[k], node._dev)
        finally:
            _device.use_stream(main)
            _lib._call_hook, _device._alloc_hook = saved
    # one fusion plan (and one launch tape) over all outputs; last layer's dW is needed first, so it is evaluated
    # first and its all-reduce overlaps the rest of the backward pass
    evaluate_many(list(reversed(grads)) + ([cost] if read_loss else []), on_ready=ready)
    for w in works:
        w.wait()
    if opt is None:
        new_w = None
    elif overlap:
        main.wait_stream(side)
        for k, g in enumerate(grads):              # (a gradient whose callback never fired: update it here)
            if new_w[k] is None:
                new_w[k] = opt.update_one(k, params[k], g.eval_device())
        opt.end_step()
    else:
        new_w = opt(list(params), [g.eval_device() for g in grads])
    if new_w is not None:
        opt.apply_new_weights(params, new_w)
    loss = None
    if read_loss:
        lv = cost.eval_device()
        if dist is not None:
            own = DeviceArray.empty(())            # (never reduce in place into a value that may alias a tape arena / constant)
            from .device import copy_device
            copy_device(lv, own)
            lv = own
            dist.all_reduce(_flat_view(lv), op=dist.ReduceOp.SUM)
        if read_loss == "deferred":
            return PendingLoss(lv)
        loss = float(to_host(lv))
        check_pending_flags()                      # deferred label validation of SoftmaxCEWithLogits (reference ops.py:246-248)
    if opt is None:
        return loss, [g.eval_device() for g in grads]
    return loss


def sharded_eval(tops, dist=None):
    """Evaluates nodes whose value is a SUM over the batch rows of this rank's shard (parameter gradients of any order: the
    second-order `gg` of SURVEY 8e) and all-reduces each in place across the ranks as soon as its kernels are queued.
    Returns the device arrays (identical on every rank)."""
    works = []

    def ready(node):
        if dist is None:
            return
        d = node._dev
        if d.size <= 16 or any(st == 0 and dim > 1 for dim, st in zip(d.shape, d.strides)):
            from .device import copy_device
            own = DeviceArray.empty(d.shape)
            copy_device(d, own)
            node._dev = d = own
        works.append(allreduce_async(d, dist))
    evaluate_many(list(tops), on_ready=ready)
    for w in works:
        w.wait()
    return [t.eval_device() for t in tops]


def broadcast_params(params, dist, src=0):
    """Makes every rank start from rank `src`'s parameter values (ranks otherwise rely on identical seeding)."""
    if dist is None:
        return
    for p in params:
        d = p.eval_device()
        dist.broadcast(_flat_view(d), src=src)


class PendingLoss:
    """The loss of a step whose device->host copy is queued but not waited for (`train_step(..., read_loss="deferred")`).
    A loop that reads step k's loss after queueing step k+1 lets Python build the next graph while the GPU still runs
    (4.7 ms of host work per configs[2] step: graph build, ad.grad, tape replay - measured to be hidden already at >= 1024
    rows per rank, `profiles/r01_shard_step_probe_v2.log`, so bench.py reads synchronously); `value()` blocks until the copy
    is done and raises the deferred label-validation error of SoftmaxCEWithLogits, like the synchronous path."""

    def __init__(self, loss_dev):
        from . import _lib
        from .device import copy_device, stream_ptr
        import ctypes as C
        t = torch()
        self._own = DeviceArray.empty(())             # (not a view of the step's arena: holding one would pin a graph's buffer)
        copy_device(loss_dev, self._own)
        self._host = pinned_empty((1,))
        _lib.check(_lib.lib().adb_memcpy_d2h(C.c_void_p(self._host.ctypes.data), C.c_void_p(self._own.ptr), 8, C.c_void_p(stream_ptr())))
        self._done = t.cuda.Event()
        self._done.record(t.cuda.current_stream())

    def value(self):
        self._done.synchronize()
        check_pending_flags()
        return float(self._host[0])


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
    last = {}

    def c3_step():
        last["loss"] = train_step(nn, opt, xd, yd, gbatch, dist, read_loss=True)
    ms = timed(c3_step, steps, 4)     # (tape: recorded on step 2, first replay on step 3)
    fl = _mlp_flops(sizes, gbatch)
    out["c3_mlp_4096x8_b8192"] = {"samples_per_s": round(gbatch / (ms * 1e-3), 1), "ms_per_step": round(ms, 3), "n_gpus": ws,
                                  "global_batch": gbatch, "scaling": "strong", "gemm_tflops": round(fl / (ms * 1e-3) / 1e12, 2),
                                  "allreduce_bytes_per_step": sum((a + 1) * b * 8 for a, b in zip(sizes, sizes[1:])),
                                  "loss": last.get("loss"),
                                  "includes": "graph build + ad.grad + kernels + per-layer NCCL all-reduce and per-layer Adam (side stream) overlapped "
                                              "with backward + all-reduced loss read back to the host every step; inputs resident in HBM"}
    return out
