"""The small GEMMs of the char-rnn config (C5) through the public API: time per einsum (16 queued back to back, so the
tiny ones are bounded by the ~20 us host cost of a node evaluation) and kernels launched (2 = split-K + reduction).
    python tools/bench_gemm_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autodiff_b200 as ad

ad.init(0)
rng = np.random.default_rng(0)
SHAPES = [  # (spec, A shape, B shape): forward, dX and dW forms of examples/char_rnn.py at H=1024, B=512, V=71
    ("ij,jk->ik", (512, 1024), (1024, 1024)), ("ik,jk->ij", (512, 1024), (1024, 1024)), ("ij,ik->jk", (512, 1024), (512, 1024)),
    ("ij,jk->ik", (512, 1024), (1024, 71)), ("ik,jk->ij", (512, 71), (1024, 71)), ("ij,ik->jk", (512, 1024), (512, 71)),
    ("ij,jk->ik", (512, 71), (71, 1024)), ("ij,jk->ik", (256, 1024), (1024, 1024)), ("ij,jk->ik", (1024, 1024), (1024, 1024)),
    ("ij,jk->ik", (2048, 1024), (1024, 1024)), ("ij,jk->ik", (64, 785), (785, 100)),
]


def run(spec, a, b, reps=16):
    nodes = [ad.Einsum(spec, ad.Variable(a, name="a"), ad.Variable(b, name="b")) for _ in range(reps)]
    l0 = ad.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for n in nodes:
        n.eval_device()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (ad.launch_count() - l0) // reps


for spec, sa, sb in SHAPES:
    a, b = ad.to_device(rng.standard_normal(sa)), ad.to_device(rng.standard_normal(sb))
    ins, out = spec.split("->")
    la, lb = ins.split(",")
    ext = dict(zip(la, sa)); ext.update(zip(lb, sb))
    flops = 2.0 * np.prod([ext[c] for c in ext])
    row = {"spec": spec, "A": sa, "B": sb}
    for _ in range(2):
        run(spec, a, b, 4)
    ms, launches = min(run(spec, a, b) for _ in range(4))
    row.update({"us": round(ms * 1e3, 1), "TFLOP/s": round(float(flops) / ms / 1e9, 2), "launches": launches})
    print(row, flush=True)
