"""Measurements of the BASELINE.json training configs (called by bench.py; one process per GPU under torchrun):

  C1  MLP 784-100-10, batch 64 (launch-bound; rank 0 at N = 1 only)
  C3  MLP 4096-wide x 8 layers, global batch 8192, batch-sharded, NCCL all-reduce of the parameter gradients + replicated Adam
  C4  conv net via im2col + einsum, 3x64x64 images, global batch 256, forward + parameter gradients, batch-sharded
  C5  char-rnn hidden 1024, seq 128, global batch 512, second-order gradient gg = grad(grad(loss, w), w), batch-sharded

Every sharded config carries `dp_parity`: the SAME work done single-process on the full batch by rank 0 (plain schedule: no
side-stream optimizer overlap) compared with what the N ranks computed together, at the 1e-12 bar of BASELINE.json's
north_star; bench.py exits non-zero when a parity field is above the bar.  Every config carries a roofline fraction (GEMM
flops of the step / step time against the measured FP64 DMMA peak); bench.py adds the `cpu_baseline` of each config (the
reference's own loop on the host) at N = 1.  Workload definitions: reference examples/feedforward_mnist.py:17-34,
examples/char_rnn.py:50-64, tests/numerical_check.py:43-47; SURVEY.md 8(d)."""
import os
import time

import numpy as np

from . import dp, workloads as wl
from .core.high_level_ops import NN, Adam
from .core.node import Variable
from .device import DeviceArray, pinned_empty, to_device, to_host

PARITY_BAR = 1e-12
DMMA_PEAK_TFLOPS = 37.03      # measured issue-rate peak of DMMA.8x8x4 on this pool's B200 (profiles/r01_probe_fp64_issue_rates.log)


def c3_parity_ok(loss_by_step, grad_err, runs_identical, weight_err, control_weight, control_loss, bar=PARITY_BAR):
    """The data-parallel verdict of configs[2] (see `dp_parity.what` in the bench line): quantities computed from identical inputs -
    the loss at the initial weights, the all-reduced gradients at the initial weights - and run-to-run repeatability are held to the
    1e-12 bar; what follows the first updates is held to the measured noise floor of the trajectory (the single-process run against
    itself from weights one ulp apart): weights within it, losses within max(bar, 4 x it)."""
    return bool(loss_by_step[0] <= bar and grad_err <= bar and runs_identical
                and weight_err <= max(control_weight, bar)
                and max(loss_by_step) <= max(bar, 4.0 * control_loss))


def _rel(a, b):
    """norm-wise relative error max|a-b| / max|b| of two host arrays."""
    a, b = np.asarray(a), np.asarray(b)
    den = float(np.max(np.abs(b))) if b.size else 0.0
    if den == 0.0:
        return float(np.max(np.abs(a - b))) if a.size else 0.0
    return float(np.max(np.abs(a - b))) / den


class _Timer:
    """CUDA-event timing with a barrier on both sides and the max over ranks (bench.py contract)."""

    def __init__(self, dist):
        import torch
        self.t, self.dist = torch, dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.t.cuda.synchronize()

    def run(self, fn, steps, warmup):
        t = self.t
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3) / steps      # (every step ends with a host read of the loss)
        return self.max_over_ranks(ms)

    def max_over_ranks(self, v):
        if self.dist is None:
            return v
        tt = self.t.tensor([v], dtype=self.t.float64, device="cuda")
        self.dist.all_reduce(tt, op=self.dist.ReduceOp.MAX)
        return float(tt.item())


def _roof(tflops_per_gpu, what):
    return {"bound": "tensor", "achieved": round(tflops_per_gpu, 3), "peak": DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
            "frac": round(tflops_per_gpu / DMMA_PEAK_TFLOPS, 4), "what": what}


# ------------------------------------------------------------------------------------------------------------ C1
def bench_c1(timer):
    np.random.seed(1337)
    sizes, batch = [784, 100, 10], 64
    nn = NN(sizes)
    opt = Adam(len(nn.w), lr=1e-3)
    rng = np.random.default_rng(1337)
    xs = pinned_empty((batch, sizes[0])); xs[...] = rng.random((batch, sizes[0]))
    ys = pinned_empty((batch, sizes[-1])); ys[...] = np.eye(sizes[-1])[rng.integers(0, sizes[-1], batch)]
    ms = timer.run(lambda: dp.train_step(nn, opt, xs, ys, batch, None, read_loss=True), 50, 10)
    fl = dp._mlp_flops(sizes, batch)
    out = {"samples_per_s": round(batch / (ms * 1e-3), 1), "ms_per_step": round(ms, 4),
           "includes": "graph build + ad.grad + H2D of the batch + kernels + Adam on device + D2H of the loss",
           "roofline": {"bound": "latency", "what": "%.3g flop per step: 25 kernels of a few microseconds, replayed as ONE CUDA graph; "
                                                     "%.4f TFLOP/s is %.1e of the DMMA peak by construction" % (fl, fl / (ms * 1e-3) / 1e12, fl / (ms * 1e-3) / 1e12 / DMMA_PEAK_TFLOPS)}}
    return out


# ------------------------------------------------------------------------------------------------------------ C3
class _C3:
    sizes, gbatch = [4096] * 9, 8192

    def __init__(self):
        np.random.seed(1337)
        self.w0 = [np.random.randn(a + 1, b) * 0.01 for a, b in zip(self.sizes, self.sizes[1:])]      # (high_level_ops.py:26)
        rng = np.random.default_rng(1337)
        self.x = rng.standard_normal((self.gbatch, self.sizes[0]))
        lab = rng.integers(0, self.sizes[-1], self.gbatch)
        self.y = np.zeros((self.gbatch, self.sizes[-1]))
        self.y[np.arange(self.gbatch), lab] = 1.0

    def model(self):
        nn = NN([2, 2])
        nn.sizes = self.sizes
        nn.w = [Variable(to_device(w), name="w%d" % k) for k, w in enumerate(self.w0)]      # weights resident in HBM from the start
        return nn, Adam(len(nn.w), lr=1e-3)

    def shard(self, world, rank):
        lo, hi = dp.shard_rows(self.gbatch, world, rank)
        return to_device(self.x[lo:hi]), to_device(self.y[lo:hi])


def bench_c3(args, dist, timer, ws, rank):
    import torch as t
    c3 = _C3()
    sizes, gbatch = c3.sizes, c3.gbatch
    xd, yd = c3.shard(ws, rank)
    steps = max(2, min(args.steps, 5))
    # ---- timed run (4 warm-up steps: tape recorded on step 2, first replay on step 3)
    nn, opt = c3.model()
    last = {}

    def c3_step():
        last["loss"] = dp.train_step(nn, opt, xd, yd, gbatch, dist, read_loss=True)
    ms = timer.run(c3_step, steps, 4)
    # the same shard without the exchange, and the host time of a step (queueing only)
    ms_nocomm = None
    if dist is not None:
        ms_nocomm = timer.run(lambda: dp.train_step(nn, opt, xd, yd, gbatch, None, read_loss=True), 3, 1)
    t.cuda.synchronize()
    # host time of ONE step queued into an idle GPU (with several steps in flight the launch queue fills and the host blocks on the
    # GPU: that would measure back-pressure, not host work)
    host_each = []
    pend = []
    for _ in range(3):
        t.cuda.synchronize()
        h0 = time.perf_counter()
        pend.append(dp.train_step(nn, opt, xd, yd, gbatch, dist, read_loss="deferred"))     # (same graph, loss read later)
        host_each.append((time.perf_counter() - h0) * 1e3)
    host_ms = min(host_each)
    for pl in pend:
        pl.value()
    t.cuda.synchronize()
    del nn, opt, pend
    # ---- parity: 9 Adam steps (4 + 5, as the timed run) from the same weights, DP (3 runs) vs single process on the full batch.
    # Run 3 is interleaved with the single-process run on rank 0, so the weights can be compared after EVERY step:
    #   * loss at every step, and the first real Adam update (steps 0-1: the reference's Adam has a = 0 at step 0) from
    #     identical weights, must agree to 1e-12 - that is "the same result on identical inputs";
    #   * after that the two runs are different floating-point trajectories of an expansive map (|g| ~ 1e-7 next to
    #     Adam.eps = 1e-8: the update is lr * m / (sqrt(v) + 1e-8), a 1e7-fold preconditioner), so a 1-ulp difference grows
    #     geometrically; `weight_rel_err_by_step` shows the curve and `control_1ulp_weight_rel_err` what ONE ulp on the
    #     initial weights does to the single-process run itself over the same 9 steps.
    nsteps, runs = 9, 3
    dp_losses, dp_w = [], []

    def wdiff(nn_a, nn_b, ref0=None):
        worst = 0.0
        for k, (a, b) in enumerate(zip(nn_a.w, nn_b.w)):
            fa, fb = dp._flat_view(a.eval_device()), dp._flat_view(b.eval_device())
            if ref0 is None:
                den = float(fb.abs().max())
                num = float((fa - fb).abs().max())
            else:                                            # error of the UPDATE: (a - w0) vs (b - w0)
                f0 = dp._flat_view(ref0[k])
                den = float((fb - f0).abs().max())
                num = float((fa - fb).abs().max())
            worst = max(worst, num / den if den > 0 else num)
        return worst
    # gradients at the initial weights: all-reduced DP gradient vs the single-process gradient of the full batch
    def grads_at_init(xs, ys, d):
        from .core.ops import Einsum, SoftmaxCEWithLogits
        g_nn, _ = c3.model()
        return dp.step(g_nn.w, None, lambda: Einsum("i->", SoftmaxCEWithLogits(labels=Variable(ys, name="labels"),
                                                                                logits=g_nn(Variable(xs, name="images")))) / gbatch, d, read_loss=True)
    _, g_dp = grads_at_init(xd, yd, dist)
    grad_err = None
    if rank == 0:
        _, g_sp = grads_at_init(xd, yd, None) if ws == 1 else grads_at_init(to_device(c3.x), to_device(c3.y), None)
        grad_err = 0.0
        for a, b in zip(g_dp, g_sp):
            fa, fb = dp._flat_view(a), dp._flat_view(b)
            grad_err = max(grad_err, float((fa - fb).abs().max()) / float(fb.abs().max()))
        del g_sp
    del g_dp
    for _ in range(runs - 1):
        nn, opt = c3.model()
        dp_losses.append([dp.train_step(nn, opt, xd, yd, gbatch, dist, read_loss=True) for _ in range(nsteps)])
        dp_w.append([np.array(w.value) for w in nn.w] if rank == 0 else None)
        del nn, opt
    nn, opt = c3.model()
    sp_nn = sp_opt = xf = yf = w_init = None
    if rank == 0:
        sp_nn, sp_opt = c3.model()
        w_init = [to_device(w) for w in c3.w0]
        xf, yf = (xd, yd) if ws == 1 else (to_device(c3.x), to_device(c3.y))
    losses3, sp_losses, by_step, first_update, n1_t = [], [], [], None, 0.0
    for s in range(nsteps):
        losses3.append(dp.train_step(nn, opt, xd, yd, gbatch, dist, read_loss=True))
        if rank == 0:
            t.cuda.synchronize()
            t0 = time.perf_counter()
            sp_losses.append(dp.train_step(sp_nn, sp_opt, xf, yf, gbatch, None, read_loss=True, overlap_optimizer=False))
            t.cuda.synchronize()
            if s >= 4:
                n1_t += time.perf_counter() - t0
            by_step.append(wdiff(nn, sp_nn))
            if s == 1:
                first_update = wdiff(nn, sp_nn, w_init)
    dp_losses.append(losses3)
    dp_w.append([np.array(w.value) for w in nn.w] if rank == 0 else None)
    del nn, opt
    parity, n1_ms = None, None
    if rank == 0:
        n1_ms = n1_t * 1e3 / (nsteps - 4)
        # control: the single-process run again from weights moved by ONE ulp (up, then down; the larger divergence counts)
        control, control_loss = 0.0, 0.0
        for toward in (np.inf, -np.inf):
            c_nn, c_opt = c3.model()
            c_nn.w = [Variable(to_device(np.nextafter(w, toward)), name="w%d" % k) for k, w in enumerate(c3.w0)]
            c_losses = [dp.train_step(c_nn, c_opt, xf, yf, gbatch, None, read_loss=True, overlap_optimizer=False) for s in range(nsteps)]
            control = max(control, wdiff(c_nn, sp_nn))
            control_loss = max(control_loss, max(abs(a - b) / abs(b) for a, b in zip(c_losses, sp_losses)))
            del c_nn, c_opt
        sp_w = [np.array(w.value) for w in sp_nn.w]
        del sp_nn, sp_opt, xf, yf, w_init
        loss_by_step = [max(abs(run[s] - sp_losses[s]) / abs(sp_losses[s]) for run in dp_losses) for s in range(nsteps)]
        parity = {"loss_rel_err_identical_weights": loss_by_step[0], "grad_rel_err_at_init": grad_err,
                  "loss_rel_err": max(loss_by_step), "loss_rel_err_by_step": [float("%.3g" % v) for v in loss_by_step],
                  "first_update_rel_err": first_update,
                  "weight_rel_err": max(_rel(a, b) for run in dp_w for a, b in zip(run, sp_w)),
                  "weight_rel_err_by_step": [float("%.3g" % v) for v in by_step],
                  "control_1ulp_weight_rel_err": float("%.3g" % control), "control_1ulp_loss_rel_err": float("%.3g" % control_loss),
                  "runs": runs, "runs_identical": all(run == dp_losses[0] for run in dp_losses)
                                                   and all(all(np.array_equal(a, b) for a, b in zip(run, dp_w[0])) for run in dp_w), "steps": nsteps, "bar": PARITY_BAR,
                  "reference": "the same %d Adam steps single-process on the full batch (optimizer after backward, no side stream)" % nsteps,
                  "what": "gated at the 1e-12 bar: everything computed from IDENTICAL inputs - the loss at the initial weights and the all-reduced "
                          "gradients at the initial weights - and run-to-run bitwise repeatability of the whole 9-step run (losses and weights).  After "
                          "the first updates the two runs are different fp64 trajectories of an expansive map (8 ReLU layers: a pre-activation "
                          "that rounds to the other side of 0 changes a gradient by O(1) of one unit; weight_rel_err_by_step jumps at such a step), so "
                          "later losses and the final weights are gated against the measured noise floor instead: control_1ulp_* is the "
                          "single-process run against ITSELF from weights one ulp apart over the same 9 steps, and the data-parallel run must stay within it "
                          "(weights) / within max(1e-12, 4 x control) (losses)",
                  "loss_single_process": sp_losses[-1], "loss_dp": dp_losses[0][-1]}
        parity["ok"] = c3_parity_ok(loss_by_step, grad_err, parity["runs_identical"], parity["weight_rel_err"], control, control_loss)
    fl = dp._mlp_flops(sizes, gbatch)
    per_gpu = fl / ws / (ms * 1e-3) / 1e12
    out = {"samples_per_s": round(gbatch / (ms * 1e-3), 1), "ms_per_step": round(ms, 3), "n_gpus": ws, "global_batch": gbatch,
           "scaling": "strong", "gemm_tflops": round(fl / (ms * 1e-3) / 1e12, 2), "gemm_frac_of_peak": round(per_gpu / DMMA_PEAK_TFLOPS, 4),
           "allreduce_bytes_per_step": sum((a + 1) * b * 8 for a, b in zip(sizes, sizes[1:])), "loss": last.get("loss"),
           "host_ms": round(host_ms, 3), "dp_parity": parity,
           "roofline": _roof(per_gpu, "23 GEMMs of 2*rows*4097*4096 flop per step (SURVEY 8d) / whole step time, per GPU"),
           "includes": "graph build + ad.grad + kernels + per-layer NCCL all-reduce and per-layer Adam (side stream) overlapped with "
                       "backward + all-reduced loss read back to the host every step; inputs resident in HBM"}
    if ms_nocomm is not None:
        out["ms_per_step_same_shard_no_exchange"] = round(ms_nocomm, 3)
        out["exposed_comm_ms"] = round(ms - ms_nocomm, 3)
    if n1_ms is not None:
        out["n1_ms_per_step_same_process"] = round(n1_ms, 3)
        out["efficiency_vs_n1"] = round(n1_ms / (ws * ms), 4)
    return out


# ------------------------------------------------------------------------------------------------------------ C4
def bench_c4(args, dist, timer, ws, rank, ad):
    gbatch = 256
    rng = np.random.default_rng(1337)
    x = rng.standard_normal((gbatch, 64, 64, 3))
    y = np.eye(10)[rng.integers(0, 10, gbatch)]
    prm = wl.conv_net_params(rng)
    lo, hi = dp.shard_rows(gbatch, ws, rank)
    xd, yd = to_device(x[lo:hi]), to_device(y[lo:hi])
    pd = [to_device(p) for p in prm]

    def run(xs, ys, d):
        X, Y = Variable(xs, name="x"), Variable(ys, name="y")
        P = [Variable(p, name="p%d" % i) for i, p in enumerate(pd)]
        return dp.step(P, None, lambda: wl.conv_net(ad, X, Y, *P, batch=gbatch), d, read_loss=True)
    last = {}

    def c4_step():
        last["loss"], last["grads"] = run(xd, yd, dist)
    ms = timer.run(c4_step, max(3, min(args.steps, 10)), 6)       # (6 warm-up steps: deferred gradients + tape + CUDA graph are all live)
    parity = None
    if rank == 0:
        got = [to_host(g) for g in last["grads"]]
        loss_sp, grads_sp = run(xd, yd, None) if ws == 1 else run(to_device(x), to_device(y), None)
        want = [to_host(g) for g in grads_sp]
        parity = {"loss_rel_err": abs(last["loss"] - loss_sp) / abs(loss_sp), "grad_rel_err": max(_rel(a, b) for a, b in zip(got, want)),
                  "bar": PARITY_BAR, "reference": "forward + parameter gradients single-process on the full batch"}
        parity["ok"] = bool(parity["loss_rel_err"] <= PARITY_BAR and parity["grad_rel_err"] <= PARITY_BAR)
    last.pop("grads", None)
    fl = wl.conv_net_flops(gbatch)
    per_gpu = fl / ws / (ms * 1e-3) / 1e12
    out = {"samples_per_s": round(gbatch / (ms * 1e-3), 1), "ms_per_step": round(ms, 3), "n_gpus": ws, "global_batch": gbatch, "scaling": "strong",
           "gemm_tflops": round(fl / (ms * 1e-3) / 1e12, 2), "loss": last.get("loss"), "dp_parity": parity,
           "roofline": _roof(per_gpu, "GEMM flops of fwd + both grads (conv1 has no dX) / whole step time, per GPU; the step also moves the im2col matrices "
                                      "through HBM, so the whole-step figure understates the GEMM kernels"),
           "what": "NHWC %dx64x64x3 -> conv5x5x16 -> ReLU -> conv5x5x32 -> ReLU -> dense 10 -> softmax CE; forward + grads of the 3 parameter tensors, "
                   "all-reduced across the ranks; inputs resident, loss read back" % gbatch}
    return out


# ------------------------------------------------------------------------------------------------------------ C5
def char_rnn_flops(hidden, steps, batch, vocab, order=2):
    """GEMM flops of one evaluation: per time step h@w, x@u, h@v forward; the first-order backward doubles every product; the
    second-order graph evaluates roughly the first-order graph's products twice more (counted here as 3x forward per order)."""
    fwd = (steps - 1) * 2.0 * batch * (hidden * hidden + vocab * hidden + hidden * vocab)
    return fwd * (3 if order == 1 else 9)


def bench_c5(args, dist, timer, ws, rank, ad, hidden=1024, steps=128, gbatch=512, vocab=71):
    import torch as t
    rng = np.random.default_rng(1337)
    prm = [to_device(p) for p in wl.char_rnn_params(rng, hidden, vocab)]
    idx = rng.integers(0, vocab, (gbatch, steps))
    lo, hi = dp.shard_rows(gbatch, ws, rank)
    info = {}

    def run(rows, d):
        oh = wl._SlicedDevice(to_device(np.eye(vocab)[idx[rows[0]:rows[1]]]))
        def once():
            info.pop("gg", None)           # (a live value of the previous evaluation would pin the tape's CUDA-graph buffer)
            P = [Variable(p, name="p%d" % i) for i, p in enumerate(prm)]
            t0 = time.perf_counter()
            loss = wl.char_rnn(ad, oh, *P)
            g = ad.grad(loss, [P[0]])[0]
            gg = ad.grad(g, [P[0]])[0]
            info["build_s"] = time.perf_counter() - t0
            vals = dp.sharded_eval([gg, loss], d)
            info["gg"], info["loss"] = vals[0], float(to_host(vals[1]))
        return once
    once = run((lo, hi), dist)
    l0 = ad.launch_count()
    once()
    t.cuda.synchronize()
    kernels = ad.launch_count() - l0
    first_build = info["build_s"]
    ms = timer.run(once, 2, 6)               # steady state of a loop that rebuilds the graph every evaluation (deferred gradients,
                                             # core/grad.py, replay from the 6th evaluation of a structure on)
    s = ms * 1e-3
    parity = None
    if rank == 0:
        got, loss_dp = to_host(info["gg"]), info["loss"]
        build_s = info["build_s"]
        if ws > 1:
            sp = run((0, gbatch), None)
            sp()
        want, loss_sp = to_host(info["gg"]), info["loss"]
        parity = {"gg_rel_err": _rel(got, want), "loss_rel_err": abs(loss_dp - loss_sp) / abs(loss_sp), "bar": PARITY_BAR,
                  "reference": "gg = grad(grad(loss,[w]),[w]) single-process on the full batch" + (" (N = 1: the same evaluation)" if ws == 1 else "")}
        parity["ok"] = bool(parity["gg_rel_err"] <= PARITY_BAR and parity["loss_rel_err"] <= PARITY_BAR)
        info["loss"], info["build_s"] = loss_dp, build_s
    info.pop("gg", None)
    fl = char_rnn_flops(hidden, steps, gbatch, vocab)
    per_gpu = fl / ws / s / 1e12
    out = {"seq_per_s": round(gbatch / s, 1), "s_per_eval": round(s, 4), "graph_build_s": round(info["build_s"], 4), "first_build_s": round(first_build, 3),
           "n_gpus": ws, "hidden": hidden, "seq": steps, "global_batch": gbatch, "vocab": vocab, "scaling": "strong", "loss": info.get("loss"),
           "kernels": kernels, "dp_parity": parity, "gemm_tflops": round(fl / s / 1e12, 2),
           "roofline": _roof(per_gpu, "approximate GEMM flops of the second-order evaluation (9 x the forward products) / whole evaluation time incl. the "
                                      "Python graph build, per GPU"),
           "what": "examples/char_rnn.py graph; loss -> g = grad(loss,[w]) -> gg = grad(g,[w]); evaluate gg on this rank's sequences, all-reduce gg and the loss; "
                   "every repetition rebuilds the graph (graph_build_s is part of s_per_eval); structure-keyed launch tape active"}
    return out


def bench_training(args, dist, ad):
    """All training configs; returns (mlp_train dict, other_configs dict, parity_ok)."""
    import torch as t
    ws, rank = dp.world()
    timer = _Timer(dist)
    mlp, other = {}, {}

    def guarded(name, fn, into):
        try:
            into[name] = fn()
        except Exception as ex:      # reported, never silently dropped
            import traceback
            into[name] = {"error": repr(ex)[:300], "trace": traceback.format_exc()[-600:]}
        t.cuda.synchronize()
    if rank == 0 and ws == 1:
        guarded("c1_mlp_784_100_10_b64", lambda: bench_c1(timer), mlp)
    guarded("c3_mlp_4096x8_b8192", lambda: bench_c3(args, dist, timer, ws, rank), mlp)
    guarded("c4_conv_b256", lambda: bench_c4(args, dist, timer, ws, rank, ad), other)
    guarded("c5_char_rnn_h1024_t128_b512_second_order", lambda: bench_c5(args, dist, timer, ws, rank, ad), other)
    ok = True
    if rank == 0:
        for grp in (mlp, other):
            for v in grp.values():
                p = v.get("dp_parity") if isinstance(v, dict) else None
                if "error" in v or (p is not None and not p.get("ok", False)):
                    ok = False
    return mlp, other, ok
