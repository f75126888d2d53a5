"""Where does an end-to-end sweep step spend its time?  (host wall clock at a few checkpoints + CUDA events)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autodiff_b200 as ad
import bench
from autodiff_b200.core import engine

ad.init(0)
cfg = bench.FULL
rng = np.random.default_rng(0)
host_inputs = {}
for name, (spec, shapes) in cfg.items():
    host_inputs[name] = []
    for shp in shapes:
        buf = ad.pinned_empty(shp)
        rng.standard_normal(out=buf.reshape(-1))
        host_inputs[name].append(buf)

def step():
    t0 = time.perf_counter()
    nodes = [node for name, e, da, db in bench.build_sweep(ad, cfg, host_inputs) for node in (e, da, db)]
    t1 = time.perf_counter()
    hosts = ad.eval_many(nodes)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, hosts

for i in range(4):
    b, e, hosts = step()
    del hosts
    print("step %d: build %.1f ms, eval_many %.1f ms" % (i, b * 1e3, e * 1e3), flush=True)

engine.pipeline_trace = []
step()
torch.cuda.synchronize()
t0 = engine.pipeline_trace[0][1]
for label, ev in engine.pipeline_trace[1:]:
    print("  %7.1f ms  %s" % (t0.elapsed_time(ev), label))
engine.pipeline_trace = None

# raw copy speeds
x = host_inputs["C2c"][0]
d = ad.to_device(x)
for i in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d = ad.to_device(x); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("H2D 2 GiB pinned: %.1f ms = %.1f GB/s" % ((t1 - t0) * 1e3, x.nbytes / (t1 - t0) / 1e9))
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); h = ad.to_host(d); t1 = time.perf_counter()
    print("D2H 2 GiB -> fresh pinned: %.1f ms = %.1f GB/s" % ((t1 - t0) * 1e3, x.nbytes / (t1 - t0) / 1e9))
    del h
t0 = time.perf_counter(); p = ad.pinned_empty((1 << 28,)); t1 = time.perf_counter()
print("pinned_empty 2 GiB (cached?): %.1f ms" % ((t1 - t0) * 1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); step(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
