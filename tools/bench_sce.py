"""A/B of the two SoftmaxCEWithLogits kernels (persistent bulk-copy ring vs register-resident) per row width.
Usage: python tools/bench_sce.py [cols ...]   (256 MiB of logits + 256 MiB of labels per case; CUDA events, best of 5).
ADB_SCE_RING=1 is the default dispatch (ring only where softmax_ce_ring_try accepts the shape), 0 pins the register-resident kernels."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autodiff_b200 as ad

ad.init(0)
PEAK = 6555.8
N = 8192 * 4096
rng = np.random.default_rng(0)
zf = ad.to_device(rng.standard_normal(N))
lab = np.zeros(N)
widths = [int(a) for a in sys.argv[1:]] or [512, 1024, 1536, 2048, 3072, 4096, 8192]


def run(z, l, reps=8):
    nodes = [ad.SoftmaxCEWithLogits(labels=ad.Variable(l, name="l"), logits=ad.Variable(z, name="z")) for _ in range(reps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for n in nodes:
        n.eval_device()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for cols in widths:
    rows = N // cols
    l = lab[:rows * cols].reshape(rows, cols).copy()
    l[np.arange(rows), rng.integers(0, cols, rows)] = 1.0
    ld = ad.to_device(l)
    z = zf.getitem(slice(0, rows * cols)).reshape_view((rows, cols))
    if os.environ.get("BENCH_SCE_SIMPLE"):        # the default dispatch only (run once per ADB_SCE_WARP setting: read once per process)
        for _ in range(3):
            run(z, ld, 2)
        ms = min(run(z, ld) for _ in range(5))
        nbytes = 8 * (2 * rows * cols + rows)
        print({"rows": rows, "cols": cols, "ms": round(ms, 4), "frac": round(nbytes / ms / 1e6 / PEAK, 3), "warp_kernel": os.environ.get("ADB_SCE_WARP", "1")}, flush=True)
        del ld
        continue
    out = {}
    for ring in ("1", "0"):
        os.environ["ADB_SCE_RING"] = ring
        for _ in range(3):
            run(z, ld, 2)
        ms = min(run(z, ld) for _ in range(5))
        out[ring] = ms
    nbytes = 8 * (2 * rows * cols + rows)
    print({"rows": rows, "cols": cols, "ring_ms": round(out["1"], 4), "ring_frac": round(nbytes / out["1"] / 1e6 / PEAK, 3),
           "regs_ms": round(out["0"], 4), "regs_frac": round(nbytes / out["0"] / 1e6 / PEAK, 3)}, flush=True)
    del ld
