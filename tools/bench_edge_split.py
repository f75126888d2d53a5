"""A/B of the ragged-edge GEMM split (csrc/einsum.cu: ragged_edge) on the configs[2] layer shapes: dW and dX of a 4096-wide
layer with the bias row (4097), at 8192 rows (1 GPU) and 1024 rows (one rank of 8)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autodiff_b200 as ad

ad.init(0)
rng = np.random.default_rng(0)


def run(spec, a, b, reps=6):
    nodes = [ad.Einsum(spec, ad.Variable(a, name="a"), ad.Variable(b, name="b")) for _ in range(reps)]
    l0 = ad.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for n in nodes:
        n.eval_device()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (ad.launch_count() - l0) // reps


for rows in (8192, 1024):
    x = ad.to_device(rng.standard_normal((rows, 4097))); dy = ad.to_device(rng.standard_normal((rows, 4096)))
    w = ad.to_device(rng.standard_normal((4097, 4096)))
    for name, spec, a, b, flops in (("dW", "ij,ik->jk", x, dy, 2.0 * rows * 4097 * 4096), ("dX", "ik,jk->ij", dy, w, 2.0 * rows * 4097 * 4096),
                                    ("fwd", "ij,jk->ik", x, w, 2.0 * rows * 4097 * 4096)):
        row = {"rows": rows, "gemm": name}
        for mode in ("1", "0"):
            os.environ["ADB_EDGE_SPLIT"] = mode
            run(spec, a, b, 2)
            ms, launches = min(run(spec, a, b) for _ in range(3))
            row["split" if mode == "1" else "whole"] = {"ms": round(ms, 3), "TFLOP/s": round(flops / ms / 1e9, 2), "launches": launches}
        print(row, flush=True)
