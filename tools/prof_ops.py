"""Runs a handful of HBM-bound ops a few times each (for `ncu -k regex:...` captures; not a benchmark).
usage: python tools/prof_ops.py [op ...]   ops: add relu copy exp tanh sumrows sumcols sumall c2c_fwd c2c_dB c2c_dA sce adam"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autodiff_b200 as ad

ad.init(0)
rng = np.random.default_rng(0)
ops = sys.argv[1:] or ["add", "relu", "copy"]
R, Cc = 8192, 4096
x = ad.to_device(rng.standard_normal((R, Cc)))
y = ad.to_device(rng.random((R, Cc)) + 0.5)
V = lambda d: ad.Variable(d)
S = 128
big = None
if any(o.startswith("c2c") for o in ops):
    big = ad.to_device(rng.standard_normal((S, S, S, S)))
    b2 = ad.to_device(rng.standard_normal((S, S)))
    g2 = ad.to_device(rng.standard_normal((S, S)))
lab = None
if "sce" in ops:
    l = np.zeros((R, Cc)); l[np.arange(R), rng.integers(0, Cc, R)] = 1.0
    lab = ad.to_device(l)
builders = {
    "add": lambda: V(x) + V(y),
    "relu": lambda: ad.ReLU(V(x)),
    "copy": lambda: ad.Concat(V(x), ad.ConstFill((R, 1), 1.0), axis=1),
    "exp": lambda: ad.Exp(V(x)),
    "tanh": lambda: ad.Tanh(V(x)),
    "sumrows": lambda: ad.Einsum("ab->a", V(x)),
    "sumcols": lambda: ad.Einsum("ab->b", V(x)),
    "sumall": lambda: ad.Einsum("ab->", V(x)),
    "c2c_fwd": lambda: ad.Einsum("ijkl,jl->ik", V(big), V(b2)),
    "c2c_dB": lambda: ad.Einsum("ijkl,ik->jl", V(big), V(g2)),
    "c2c_dA": lambda: ad.Einsum("ik,jl->ijkl", V(g2), V(b2)),
    "sce": lambda: ad.SoftmaxCEWithLogits(labels=V(lab), logits=V(x)),
}
for name in ops:
    for _ in range(3):
        builders[name]().eval_device()
    torch.cuda.synchronize()
print("ok", ops)
