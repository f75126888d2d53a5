"""BASELINE configs[3] (conv net via im2col+einsum, 3x64x64, batch 256, fwd+grad) and configs[4] (char-rnn hidden
1024, seq 128, batch 512, grad of grad) on one B200, with the numpy oracle timed on a bounded sample beside them.
    python tools/bench_configs.py [--rnn-steps T] [--skip-cpu]
Writes gpurun_out/configs.json."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autodiff_b200 as ad
import oracle.np_autodiff as onp
from autodiff_b200 import workloads as wl
from autodiff_b200.core.engine import evaluate_many

ap = argparse.ArgumentParser()
ap.add_argument("--rnn-steps", type=int, default=128)
ap.add_argument("--rnn-hidden", type=int, default=1024)
ap.add_argument("--rnn-batch", type=int, default=512)
ap.add_argument("--skip-cpu", action="store_true")
ap.add_argument("--skip-rnn", action="store_true")
args = ap.parse_args()
ad.init(0)
sys.setrecursionlimit(100000)
out = {}


def gpu_time(fn, warmup=2, reps=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


# ---------------------------------------------------------------- C4
rng = np.random.default_rng(1337)
B = 256
xv = rng.standard_normal((B, 64, 64, 3))
yv = np.eye(10)[rng.integers(0, 10, B)]
params = wl.conv_net_params(rng)
xd, yd = ad.to_device(xv), ad.to_device(yv)
pd = [ad.to_device(p) for p in params]


def conv_step(m, x, y, ps):
    X, Y = m.Variable(x, name="x"), m.Variable(y, name="y")
    P = [m.Variable(p, name="p%d" % i) for i, p in enumerate(ps)]
    loss = wl.conv_net(m, X, Y, *P)
    return loss, m.grad(loss, P)


def conv_gpu():
    loss, grads = conv_step(ad, xd, yd, pd)
    evaluate_many(grads + [loss])
    return float(ad.to_host(loss.eval_device()))


l0 = ad.launch_count()
t = gpu_time(conv_gpu)
fl = wl.conv_net_flops(B)
out["c4_conv_b256"] = {"samples_per_s": round(B / t, 1), "ms_per_step": round(t * 1e3, 3), "gemm_tflops": round(fl / t / 1e12, 2),
                       "loss": conv_gpu(), "what": "NHWC 256x64x64x3 -> conv5x5x16 -> ReLU -> conv5x5x32 -> ReLU -> dense 10 -> softmax CE; "
                       "forward + grads of all 3 parameter tensors; im2col + DMMA GEMM; inputs resident, loss read back"}
print(json.dumps(out["c4_conv_b256"]), flush=True)
if not args.skip_cpu:
    bs = 8
    t0 = time.perf_counter()
    loss, grads = conv_step(onp, xv[:bs], yv[:bs], params)
    _ = [g() for g in grads]; lref = float(loss())
    tc = time.perf_counter() - t0
    lg, gg = conv_step(ad, xv[:bs], yv[:bs], params)
    err = max(float(np.max(np.abs(np.asarray(a()) - np.asarray(b()))) / np.max(np.abs(np.asarray(b())))) for a, b in zip(gg, grads))
    out["c4_conv_b256"]["cpu_baseline"] = {"samples_per_s": round(bs / tc, 2), "cores": 1, "kind": "port", "sample": "batch %d of the same net" % bs,
                                           "parity_rel_err_vs_oracle": err}
    print(json.dumps(out["c4_conv_b256"]["cpu_baseline"]), flush=True)
del xd, yd, pd

# ---------------------------------------------------------------- C5
if not args.skip_rnn:
    H, T, Bn, V = args.rnn_hidden, args.rnn_steps, args.rnn_batch, 71
    rng = np.random.default_rng(1337)
    prm = wl.char_rnn_params(rng, H, V)
    onehot = np.eye(V)[rng.integers(0, V, (Bn, T))]
    prm_d = [ad.to_device(p) for p in prm]
    oh_d = ad.to_device(onehot)

    class _Sliced:                      # lets the shared builder slice a device-resident (B,T,V) array per time step
        def __init__(self, d):
            self.d, self.shape = d, d.shape

        def __getitem__(self, idx):
            return self.d.getitem(idx)

    def rnn_gpu(order):
        P = [ad.Variable(p, name="p%d" % i) for i, p in enumerate(prm_d)]
        t0 = time.perf_counter()
        loss = wl.char_rnn(ad, _Sliced(oh_d), *P)
        g = ad.grad(loss, [P[0]])[0]
        top = g
        if order == 2:
            top = ad.grad(g, [P[0]])[0]
        tb = time.perf_counter() - t0
        top.eval_device()
        torch.cuda.synchronize()
        return tb, time.perf_counter() - t0, float(ad.to_host(loss.eval_device()))

    for order in (1, 2):
        torch.cuda.empty_cache()
        l0 = ad.launch_count()
        tb, tt, lossv = rnn_gpu(order)          # warm-up (pinned pools, module load)
        l1 = ad.launch_count()
        tb, tt, lossv = rnn_gpu(order)
        free, total = torch.cuda.mem_get_info()
        out["c5_char_rnn_order%d" % order] = {"seq_per_s": round(Bn / tt, 1), "s_per_eval": round(tt, 3), "graph_build_s": round(tb, 3),
                                              "hidden": H, "seq": T, "batch": Bn, "vocab": V, "loss": lossv, "kernels": l1 - l0,
                                              "hbm_in_use_gb": round((total - free) / 2 ** 30, 1),
                                              "what": "loss -> g = grad(loss,[w])" + (" -> gg = grad(g,[w]), evaluate gg" if order == 2 else ", evaluate g")}
        print(json.dumps(out["c5_char_rnn_order%d" % order]), flush=True)
    if not args.skip_cpu:
        h, tt_, b = 64, 8, 16
        prm_s = wl.char_rnn_params(np.random.default_rng(1), h, V)
        oh_s = np.eye(V)[np.random.default_rng(2).integers(0, V, (b, tt_))]

        def build(m):
            P = [m.Variable(p, name="p%d" % i) for i, p in enumerate(prm_s)]
            loss = wl.char_rnn(m, oh_s, *P)
            g = m.grad(loss, [P[0]])[0]
            return m.grad(g, [P[0]])[0]
        t0 = time.perf_counter()
        ref = np.asarray(build(onp)())
        tc = time.perf_counter() - t0
        got = np.asarray(build(ad)())
        out["c5_char_rnn_order2"]["cpu_baseline"] = {"seq_per_s": round(b / tc, 2), "cores": 1, "kind": "port",
                                                     "sample": "hidden %d, seq %d, batch %d, same second-order graph" % (h, tt_, b),
                                                     "parity_rel_err_vs_oracle": float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))}
        print(json.dumps(out["c5_char_rnn_order2"]["cpu_baseline"]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/configs.json", "w"), indent=1)
