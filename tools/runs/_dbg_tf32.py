import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import autodiff_b200 as ad
from autodiff_b200 import dp
ad.init(0)
for (m, k, n) in [(1024, 4097, 1536), (1024, 4096, 1536), (1024, 6144, 1536), (4096, 6144, 4096), (256, 6144, 256)]:
    rng = np.random.default_rng(m + k + n)
    a = ad.to_device(rng.standard_normal((m, k))); b = ad.to_device(rng.standard_normal((k, n)))
    w = ad.to_device(rng.standard_normal((m, n)))
    def run():
        A, B, W = ad.Variable(a, name="a"), ad.Variable(b, name="b"), ad.Variable(w, name="w")
        e = ad.Einsum("ij,jk->ik", A, B)
        loss = ad.Einsum("ik,ik->", e, W)
        return [x.eval_device() for x in [e] + ad.grad(loss, [A, B])]
    ref = run()
    with ad.precision("tf32x3"):
        got = run()
    for name, g, r in zip(("e", "dA", "dB"), got, ref):
        tg, tr = dp._flat_view(g), dp._flat_view(r)
        d = (tg - tr).abs()
        print((m, k, n), name, tuple(g.shape), g.strides, "err %.3e" % float(d.max() / tr.abs().max()), "bad elems", int((d > 1e-3 * tr.abs().max()).sum()), "first bad", int(torch.nonzero(d > 1e-3 * tr.abs().max())[0]) if (d > 1e-3 * tr.abs().max()).any() else -1, flush=True)
