"""One GPU stepping the 1024 / 2048 / 4096-row shard of configs[2] (what each rank of an 8 / 4 / 2-GPU job computes) with no
communication: separates the exchange cost from the non-GEMM part of a data-parallel step.  Both optimizer schedules."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autodiff_b200 as ad
from autodiff_b200 import dp
ad.init(0)
np.random.seed(1337)
sizes, gbatch = [4096] * 9, 8192
nn = ad.NN(sizes)
for w in nn.w:
    w.value = ad.to_device(w.value)
opt = ad.Adam(len(nn.w), lr=1e-3)
for local in (1024, 2048, 4096):
    rng = np.random.default_rng(0)
    x = ad.to_device(rng.standard_normal((local, 4096)))
    y = np.zeros((local, 4096)); y[np.arange(local), rng.integers(0, 4096, local)] = 1.0
    y = ad.to_device(y)
    for overlap, deferred in ((True, True), (True, False), (False, False)):
        pend = []

        def step():
            r = dp.train_step(nn, opt, x, y, gbatch, None, "deferred" if deferred else True, overlap_optimizer=overlap)
            if deferred:
                pend.append(r)
                if len(pend) > 1:
                    pend.pop(0).value()
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            step()
        while pend:
            pend.pop(0).value()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        fl = dp._mlp_flops(sizes, local)
        print("local batch %d, optimizer %s, loss read %s: %.2f ms/step, gemm-only floor %.2f ms (35 TF), host+other %.2f ms"
              % (local, "overlapped with backward" if overlap else "after backward", "one step late" if deferred else "synchronously",
                 dt * 1e3, fl / 35e12 * 1e3, dt * 1e3 - fl / 35e12 * 1e3), flush=True)
    # host-only time: how long does the python side take when the GPU is not the bottleneck?
    t0 = time.perf_counter()
    dp.train_step(nn, opt, x, y, gbatch, None, False)
    print("   host time of one step (no loss read): %.2f ms" % ((time.perf_counter() - t0) * 1e3))
    torch.cuda.synchronize()
