"""cProfile of the host side of the launch-bound configs (C1 MLP step, C5 char-rnn second-order evaluation)."""
import cProfile, os, pstats, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autodiff_b200 as ad
from autodiff_b200 import dp, workloads as wl

ad.init(0)
which = sys.argv[1] if len(sys.argv) > 1 else "c1"
if which == "c1":
    np.random.seed(1337)
    sizes, batch = [784, 100, 10], 64
    nn = ad.NN(sizes); opt = ad.Adam(len(nn.w), lr=1e-3)
    rng = np.random.default_rng(1337)
    xs = ad.pinned_empty((batch, 784)); xs[...] = rng.random((batch, 784))
    ys = ad.pinned_empty((batch, 10)); ys[...] = np.eye(10)[rng.integers(0, 10, batch)]
    for _ in range(20):
        dp.train_step(nn, opt, xs, ys, batch, None, True)
    t0 = time.perf_counter()
    for _ in range(200):
        dp.train_step(nn, opt, xs, ys, batch, None, True)
    print("C1: %.3f ms/step" % ((time.perf_counter() - t0) / 200 * 1e3))
    l0 = ad.launch_count(); dp.train_step(nn, opt, xs, ys, batch, None, True); print("launches/step", ad.launch_count() - l0)
    pr = cProfile.Profile(); pr.enable()
    for _ in range(200):
        dp.train_step(nn, opt, xs, ys, batch, None, True)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(35)
else:
    order = 2 if which == "c5" else 1
    r = wl.bench_char_rnn(ad, hidden=1024, steps=64, batch=512, order=order, reps=1)
    print(r)
    rng = np.random.default_rng(1337)
    prm = [ad.to_device(p) for p in wl.char_rnn_params(rng, 1024, 71)]
    oh = wl._SlicedDevice(ad.to_device(np.eye(71)[rng.integers(0, 71, (512, 64))]))
    def run():
        P = [ad.Variable(p, name="p%d" % i) for i, p in enumerate(prm)]
        loss = wl.char_rnn(ad, oh, *P)
        top = ad.grad(loss, [P[0]])[0]
        if order == 2:
            top = ad.grad(top, [P[0]])[0]
        t0 = time.perf_counter()
        top.eval_device()
        t1 = time.perf_counter()
        ad.synchronize()
        print("eval host %.1f ms, +sync %.1f ms" % ((t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3))
    run()
    pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(40)
