import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import autodiff_b200 as ad
ad.init(0)
rng = np.random.default_rng(0)
A = ad.to_device(rng.standard_normal((8192, 8192))); B = ad.to_device(rng.standard_normal((8192, 8192)))
with ad.precision(sys.argv[1] if len(sys.argv) > 1 else "tf32x3"):
    for rep in range(4):
        x = ad.Variable(A, name="A"); y = ad.Variable(B, name="B")
        e = ad.Einsum("ij,jk->ik", x, y)
        da, db = ad.grad(e, [x, y])
        for node, name in ((e, "e"), (da, "dA"), (db, "dB")):
            torch.cuda.synchronize(); t0 = time.perf_counter(); l0 = ad.launch_count()
            if rep == 3 and name == "dA":
                pr = cProfile.Profile(); pr.enable()
            node.eval_device()
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            if rep == 3 and name == "dA":
                pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
            print("rep", rep, name, "host %.2f ms, total %.2f ms" % ((t1 - t0) * 1e3, (time.perf_counter() - t0) * 1e3), "launches", ad.launch_count() - l0, flush=True)
