"""The reference's own CPU path (numpy oracle = np.einsum without BLAS, one thread) at growing sizes up to the FULL C2a
forward einsum (8192^2, ~2 x 8192^3 flop ~ 4-5 minutes of one host core).  Writes gpurun_out/cpu_fullsize.json."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle.np_autodiff as onp

out = {"host_cores": os.cpu_count(), "what": "oracle Einsum('ij,jk->ik') forward, float64, np.einsum(optimize=False) as reference ops.py:180", "runs": []}
rng = np.random.default_rng(0)
sizes = [int(s) for s in sys.argv[1:]] or [1024, 2048, 4096, 8192]
for n in sizes:
    a, b = rng.standard_normal((n, n)), rng.standard_normal((n, n))
    t0 = time.perf_counter()
    c = onp.Einsum("ij,jk->ik", onp.Variable(a, name="a"), onp.Variable(b, name="b"))()
    dt = time.perf_counter() - t0
    out["runs"].append({"n": n, "seconds": round(dt, 3), "gflops": round(2.0 * n ** 3 / dt / 1e9, 3), "checksum": float(c[0, 0])})
    print(out["runs"][-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/cpu_fullsize.json", "w"), indent=1)
