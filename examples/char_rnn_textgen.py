"""Character-level RNN (tanh cell, unrolled, softmax cross-entropy per step) with text sampling - the workload of the
reference's examples/char_rnn.py, on a small built-in corpus.  The graph is rebuilt every step, as in the reference;
from the third step on the structure-keyed launch tape replays it.

    python examples/char_rnn_textgen.py [--steps 200] [--hidden 100] [--unroll 20] [--corpus some_text.txt]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import autodiff_b200 as ad

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--hidden", type=int, default=100)
ap.add_argument("--unroll", type=int, default=20)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--corpus", default=None, help="a text file to train on (e.g. the reference's examples/sample_text_for_generation.txt); default: a built-in pangram corpus")
args = ap.parse_args()

corpus = open(args.corpus, encoding="utf-8", errors="ignore").read() if args.corpus else ("the quick brown fox jumps over the lazy dog. pack my box with five dozen liquor jugs. "
          "how vexingly quick daft zebras jump. sphinx of black quartz, judge my vow. ") * 40
chars = sorted(set(corpus))
index = {c: i for i, c in enumerate(chars)}
V, H = len(chars), args.hidden
data = np.array([index[c] for c in corpus])
rng = np.random.default_rng(0)

w = ad.Variable(rng.standard_normal((H, H)) * 0.01, name="w")
u = ad.Variable(rng.standard_normal((V, H)) * 0.01, name="u")
b_h = ad.Variable(np.zeros(H), name="bias_hidden")
v = ad.Variable(rng.standard_normal((H, V)) * 0.01, name="v")
b_o = ad.Variable(np.zeros(V), name="bias_out")
params = [w, u, b_h, v, b_o]
optimizer = ad.Adam(len(params), lr=0.01)


def one_hot(idx):
    return np.eye(V)[idx]


def cell(h, x):
    h = ad.Tanh(h @ w + x @ u + b_h)
    return h, h @ v + b_o


def sample(seed="the ", n=60, temperature=0.7):
    h = ad.Variable(np.zeros((1, H)), name="h")
    for c in seed:
        h, logits = cell(h, ad.Variable(one_hot([index[c]]), name="x"))
    out = seed
    for _ in range(n):
        p = ad.Softmax(logits / temperature)()[0]
        k = rng.choice(V, p=p / p.sum())
        out += chars[k]
        h, logits = cell(h, ad.Variable(one_hot([k]), name="x"))
    return out


for step in range(args.steps):
    starts = rng.integers(0, len(data) - args.unroll - 1, args.batch)
    x_batch = one_hot(np.stack([data[s:s + args.unroll] for s in starts]))          # (batch, unroll, V)
    h = ad.Variable(np.zeros((1, H)), name="h")
    costs = []
    for t in range(args.unroll - 1):
        h, logits = cell(h, ad.Variable(x_batch[:, t, :], name="x"))
        costs.append(ad.Einsum("i->", ad.SoftmaxCEWithLogits(labels=ad.Variable(x_batch[:, t + 1, :]), logits=logits)))
    total_cost = ad.Add(*costs) / args.unroll
    grads = ad.grad(total_cost, params)
    new_params = optimizer([p() for p in params], [g() for g in grads])
    optimizer.apply_new_weights(params, new_params)
    if step % 40 == 0:
        print("step %4d  cost %.2f\n    %s" % (step, float(total_cost()), sample()))
