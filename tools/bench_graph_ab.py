"""A/B of the launch-tape replay modes on the launch-bound configs: call-by-call replay vs ONE CUDA graph per step.
C1 = BASELINE configs[0] (MLP 784-100-10, batch 64, host inputs every step), C4 = conv net, C5 = char-rnn second order.
    python tools/bench_graph_ab.py [--rnn-steps T]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autodiff_b200 as ad
from autodiff_b200 import dp, workloads as wl
from autodiff_b200.core import tape
from autodiff_b200.core.high_level_ops import NN, Adam

ap = argparse.ArgumentParser()
ap.add_argument("--rnn-steps", type=int, default=128)
ap.add_argument("--skip-rnn", action="store_true")
args = ap.parse_args()
ad.init(0)
out = {}


def c1():
    np.random.seed(1337)
    sizes, batch = [784, 100, 10], 64
    nn = NN(sizes)
    opt = Adam(len(nn.w), lr=1e-3)
    rng = np.random.default_rng(1337)
    xs = ad.pinned_empty((batch, sizes[0])); xs[...] = rng.random((batch, sizes[0]))
    ys = ad.pinned_empty((batch, sizes[-1])); ys[...] = np.eye(sizes[-1])[rng.integers(0, sizes[-1], batch)]
    losses = []
    for _ in range(20):
        losses.append(dp.train_step(nn, opt, xs, ys, batch, None, read_loss=True))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        losses.append(dp.train_step(nn, opt, xs, ys, batch, None, read_loss=True))
    torch.cuda.synchronize()
    return {"ms_per_step": round((time.perf_counter() - t0) / 200 * 1e3, 4), "loss_after_220_steps": losses[-1]}


for mode in ("graph", "calls"):
    tape.GRAPH = mode == "graph"
    tape.clear()
    s0 = dict(tape.stats)
    res = {"c1": c1(), "c4": wl.bench_conv(ad)}
    if not args.skip_rnn:
        res["c5"] = wl.bench_char_rnn(ad, steps=args.rnn_steps, reps=3)
    res["tape_stats"] = {k: tape.stats[k] - s0.get(k, 0) for k in tape.stats}
    out[mode] = res
    print(mode, json.dumps(res), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/graph_ab.json", "w"), indent=1)
