"""Graph builders for the BASELINE.json configs that are not the headline sweep (used by bench.py and tools/):
C4 conv net via im2col+einsum, C5 char-rnn with a second-order gradient.  `ad` is any module with the reference
API (autodiff_b200 or the numpy oracle), so the same builder feeds the GPU arm and the CPU baseline."""
import numpy as np


def conv_net(ad, x, labels, filters1, filters2, dense, batch=None):
    """NHWC x -> conv5x5 (VALID) -> ReLU -> conv5x5 -> ReLU -> flatten -> dense -> SoftmaxCE mean.
    `batch`: the divisor of the loss (the GLOBAL batch when x is one rank's shard); defaults to the rows of x."""
    n = x.shape[0]
    h1 = ad.ReLU(ad.Conv2D(x, filters1))
    h2 = ad.ReLU(ad.Conv2D(h1, filters2))
    flat = ad.Reshape(h2, (n, -1))
    logits = flat @ dense
    return ad.Einsum("i->", ad.SoftmaxCEWithLogits(labels=labels, logits=logits)) / (n if batch is None else batch)


def conv_net_params(rng, c_in=3, c1=16, c2=32, k=5, hw=64, classes=10):
    o2 = hw - 2 * (k - 1)
    return (rng.standard_normal((k, k, c_in, c1)) * 0.1, rng.standard_normal((k, k, c1, c2)) * 0.05,
            rng.standard_normal((o2 * o2 * c2, classes)) * 0.01)


def conv_net_flops(batch, c_in=3, c1=16, c2=32, k=5, hw=64, classes=10):
    """GEMM flops of fwd + both grads (first layer has no dX)."""
    o1, o2 = hw - k + 1, hw - 2 * (k - 1)
    f1 = 2.0 * batch * o1 * o1 * (k * k * c_in) * c1
    f2 = 2.0 * batch * o2 * o2 * (k * k * c1) * c2
    f3 = 2.0 * batch * (o2 * o2 * c2) * classes
    return 2 * f1 + 3 * f2 + 3 * f3


def char_rnn(ad, onehot, w, u, b_h, v, b_o, chunk=32):
    """examples/char_rnn.py:50-62 on a (B, T, V) one-hot batch; returns the mean loss node."""
    batch, steps, vocab = onehot.shape
    hidden = w.shape[0]
    h = ad.Variable(np.zeros((1, hidden)), name="h")
    costs = []
    for t in range(steps - 1):
        x = ad.Variable(onehot[:, t, :], name="x")
        h = ad.Tanh(h @ w + x @ u + b_h)
        logits = h @ v + b_o
        y = ad.Variable(onehot[:, t + 1, :], name="y")
        costs.append(ad.Einsum("i->", ad.SoftmaxCEWithLogits(labels=y, logits=logits)))
    while len(costs) > chunk:                       # numpy's 64-operand limit in the reference (ops.py:26)
        costs = [ad.Add(*costs[i:i + chunk]) for i in range(0, len(costs), chunk)]
    return ad.Add(*costs) / steps


def char_rnn_params(rng, hidden, vocab):
    return (rng.standard_normal((hidden, hidden)) * 0.01, rng.standard_normal((vocab, hidden)) * 0.01, np.zeros(hidden),
            rng.standard_normal((hidden, vocab)) * 0.01, np.zeros(vocab))


# ------------------------------------------------------------------------------------------------ measurements
def _wall(fn, sync, warmup=1, reps=2):
    import time
    for _ in range(warmup):
        fn()
    sync()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        sync()
        best = min(best, time.perf_counter() - t0)
    return best


def bench_conv(ad, batch=256, reps=3):
    """BASELINE configs[3]: forward + grads of the conv net on one GPU (inputs resident, loss read back)."""
    from .core.engine import evaluate_many
    rng = np.random.default_rng(1337)
    xd = ad.to_device(rng.standard_normal((batch, 64, 64, 3)))
    yd = ad.to_device(np.eye(10)[rng.integers(0, 10, batch)])
    pd = [ad.to_device(p) for p in conv_net_params(rng)]
    last = {}

    def step():
        X, Y = ad.Variable(xd, name="x"), ad.Variable(yd, name="y")
        P = [ad.Variable(p, name="p%d" % i) for i, p in enumerate(pd)]
        loss = conv_net(ad, X, Y, *P)
        evaluate_many(ad.grad(loss, P) + [loss])
        last["loss"] = float(ad.to_host(loss.eval_device()))
    t = _wall(step, ad.synchronize, warmup=2, reps=reps)
    fl = conv_net_flops(batch)
    return {"samples_per_s": round(batch / t, 1), "ms_per_step": round(t * 1e3, 3), "gemm_tflops": round(fl / t / 1e12, 2), "loss": last["loss"],
            "what": "NHWC %dx64x64x3 -> conv5x5x16 -> ReLU -> conv5x5x32 -> ReLU -> dense 10 -> softmax CE; forward + grads of the 3 "
                    "parameter tensors; im2col + DMMA GEMM (split-K for the weight grads); inputs resident, loss read back" % batch}


class _SlicedDevice:
    """Lets char_rnn() slice a device-resident (B,T,V) one-hot array per time step (a view, no copy)."""

    def __init__(self, d):
        self.d, self.shape = d, d.shape

    def __getitem__(self, idx):
        return self.d.getitem(idx)


def bench_char_rnn(ad, hidden=1024, steps=128, batch=512, vocab=71, order=2, reps=1):
    """BASELINE configs[4]: loss -> g = grad(loss,[w]) -> gg = grad(g,[w]); evaluates gg (order=2) or g (order=1)."""
    import time
    rng = np.random.default_rng(1337)
    prm = [ad.to_device(p) for p in char_rnn_params(rng, hidden, vocab)]
    oh = _SlicedDevice(ad.to_device(np.eye(vocab)[rng.integers(0, vocab, (batch, steps))]))
    info = {}

    def run():
        P = [ad.Variable(p, name="p%d" % i) for i, p in enumerate(prm)]
        t0 = time.perf_counter()
        loss = char_rnn(ad, oh, *P)
        top = ad.grad(loss, [P[0]])[0]
        if order == 2:
            top = ad.grad(top, [P[0]])[0]
        info["build_s"] = time.perf_counter() - t0
        top.eval_device()
        info["loss"] = float(ad.to_host(loss.eval_device()))
    l0 = ad.launch_count()
    run()
    ad.synchronize()
    kernels = ad.launch_count() - l0
    first = info["build_s"]
    t = _wall(run, ad.synchronize, warmup=2, reps=max(reps, 2))     # steady state of a loop that rebuilds the graph every step
    return {"seq_per_s": round(batch / t, 1), "s_per_eval": round(t, 3), "graph_build_s": round(info["build_s"], 3), "hidden": hidden,
            "seq": steps, "batch": batch, "vocab": vocab, "loss": info["loss"], "kernels": kernels,
            "what": "examples/char_rnn.py graph; loss -> g = grad(loss,[w])" + (" -> gg = grad(g,[w]); evaluate gg" if order == 2 else "; evaluate g") +
                    "; every repetition rebuilds the graph from scratch (graph_build_s is part of s_per_eval); steady state after 3 repetitions "
                    "(structure-keyed launch tape active)", "first_build_s": round(first, 3)}
