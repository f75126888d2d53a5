"""CPU arm of bench.py: the reference's own implementation of every BASELINE config, timed on the host cores.

`api()` is the UNMODIFIED reference package from baseline/_ref (kind "reference") when it is installed, else the numpy
port oracle/np_autodiff.py (kind "port").  Only bench.py and tests import this module - never the product.
Every function times a BOUNDED sample of its config (sizes stated in the result) so that the default bench run stays
within minutes; the reference path is single-threaded (np.einsum without `optimize` is plain C, ufunc loops), so
`cores` is 1 whatever the box has.  Workload definitions: reference examples/feedforward_mnist.py:17-34 (MLP loop),
examples/char_rnn.py:50-64 (char-rnn graph), tests/numerical_check.py:43-47 (second-order construction)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_API = None


def api():
    """(module with the reference API, "reference" | "port")."""
    global _API
    if _API is None:
        from baseline import ref_shim
        ref = ref_shim.load()
        if ref is not None:
            _API = (ref, "reference")
        else:
            import oracle.np_autodiff as onp
            _API = (onp, "port")
    return _API


def _meta(kind, sample):
    return {"cores": 1, "host_cores_available": os.cpu_count(), "kind": kind, "sample": sample}


def einsum_flops(spec, shapes):
    ext = {}
    for sub, shp in zip(spec.split("->")[0].split(","), shapes):
        for l, d in zip(sub, shp):
            ext[l] = d
    f = 2.0
    for d in ext.values():
        f *= d
    return f


def sweep(cfg, reps=1):
    """The einsum sweep (configs[1]) at the sizes of `cfg`: forward + ad.grad w.r.t. both operands of each member."""
    ad, kind = api()
    rng = np.random.default_rng(1337)
    inputs = {n: [rng.standard_normal(s) for s in shapes] for n, (spec, shapes) in cfg.items()}
    flops = sum(3 * einsum_flops(spec, shapes) for spec, shapes in cfg.values())
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        for name, (spec, _) in cfg.items():
            a = ad.Variable(inputs[name][0], name=name + "_A")
            b = ad.Variable(inputs[name][1], name=name + "_B")
            e = ad.Einsum(spec, a, b)
            da, db = ad.grad(e, [a, b])
            for node in (e, da, db):
                node()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    out = {"value": round(flops / best / 1e12, 6), "unit": "TFLOP/s", "seconds": round(best, 3)}
    out.update(_meta(kind, "same sweep at %s, fwd + 2 grads each (%.3g flop); np.einsum without optimize is single-threaded C "
                           "(reference ops.py:180), so more cores do not help it"
                     % ("; ".join("'%s' %s" % (spec, "x".join(str(d) for d in shapes[0])) for spec, shapes in cfg.values()), flops)))
    return out


def mlp_train(sizes, batch, steps, warmup=1, seed=1337, lr=1e-3):
    """examples/feedforward_mnist.py:17-34 on synthetic data: graph build + ad.grad + evaluation + Adam, every step."""
    ad, kind = api()
    np.random.seed(seed)
    nn = ad.NN(sizes)
    opt = ad.Adam(len(nn.w), lr=lr)
    rng = np.random.default_rng(seed)
    x = rng.random((batch, sizes[0])) if sizes[0] == 784 else rng.standard_normal((batch, sizes[0]))
    y = np.eye(sizes[-1])[rng.integers(0, sizes[-1], batch)]
    loss = None

    def step():
        nonlocal loss
        xv, yv = ad.Variable(x, name="images"), ad.Variable(y, name="labels")
        cost = ad.Einsum("i->", ad.SoftmaxCEWithLogits(labels=yv, logits=nn(xv))) / batch
        grads = ad.grad(cost, nn.w)
        new = opt([w() for w in nn.w], [g() for g in grads])
        opt.apply_new_weights(nn.w, new)
        loss = float(cost())
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    out = {"value": round(batch / dt, 3), "unit": "samples/s", "s_per_step": round(dt, 4), "loss": loss}
    out.update(_meta(kind, "NN(%s), batch %d, %d timed steps after %d warm-up: graph build + ad.grad + eval + Adam (numpy) per step"
                     % ("x".join(str(s) for s in ([sizes[0]] + ["%d*%d" % (sizes[1], len(sizes) - 2)] + [sizes[-1]] if len(sizes) > 4 else sizes)),
                        batch, steps, warmup)))
    return out


def conv_train(batch, steps=1):
    """configs[3] (conv net via im2col + einsum) forward + parameter gradients.  The reference has no runnable convolution
    (reshape.py:137-160 is commented out), so this is always the numpy PORT (oracle Conv2D = as_strided windows + np.einsum)."""
    import oracle.np_autodiff as onp
    from autodiff_b200 import workloads as wl
    rng = np.random.default_rng(1337)
    x = rng.standard_normal((batch, 64, 64, 3))
    y = np.eye(10)[rng.integers(0, 10, batch)]
    prm = wl.conv_net_params(rng)
    best = None
    for _ in range(steps):
        t0 = time.perf_counter()
        X, Y = onp.Variable(x, name="x"), onp.Variable(y, name="y")
        P = [onp.Variable(p, name="p%d" % i) for i, p in enumerate(prm)]
        loss = wl.conv_net(onp, X, Y, *P)
        vals = [g() for g in onp.grad(loss, P)]
        loss()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    out = {"value": round(batch / best, 3), "unit": "samples/s", "s_per_step": round(best, 3)}
    out.update(_meta("port", "same conv net (64x64x3 -> conv5x5x16 -> conv5x5x32 -> dense 10), batch %d instead of 256, forward + 3 parameter gradients" % batch))
    return out


def char_rnn_second_order(hidden, steps, batch, vocab=71):
    """configs[4]: examples/char_rnn.py:50-62 graph, g = grad(loss,[w]), gg = grad(g,[w]); evaluates gg (graph build included)."""
    ad, kind = api()
    from autodiff_b200 import workloads as wl
    rng = np.random.default_rng(1337)
    prm = wl.char_rnn_params(rng, hidden, vocab)
    oh = np.eye(vocab)[rng.integers(0, vocab, (batch, steps))]
    t0 = time.perf_counter()
    P = [ad.Variable(p, name="p%d" % i) for i, p in enumerate(prm)]
    loss = wl.char_rnn(ad, oh, *P)
    g = ad.grad(loss, [P[0]])[0]
    gg = ad.grad(g, [P[0]])[0]
    build = time.perf_counter() - t0
    v = gg()
    dt = time.perf_counter() - t0
    out = {"value": round(batch / dt, 4), "unit": "sequences/s", "s_per_eval": round(dt, 3), "graph_build_s": round(build, 3),
           "checksum": float(np.sum(v))}
    out.update(_meta(kind, "char-rnn hidden %d, seq %d, batch %d (config: 1024 / 128 / 512), loss -> grad -> grad, one evaluation" % (hidden, steps, batch)))
    return out
