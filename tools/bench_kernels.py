"""Per-kernel achieved bandwidth / throughput on one B200 (CUDA events, inputs > L2, 3 warm-ups, best of 5).
Algorithmic bytes = 8 * (distinct input elements + output elements), SURVEY.md 8d.  Writes gpurun_out/kernels.json."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autodiff_b200 as ad

ad.init(0)
PEAK = 6555.8
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
rng = np.random.default_rng(0)
results = []
only = sys.argv[1] if len(sys.argv) > 1 else ""


def dev(shape, positive=False):
    a = rng.random(shape) + 0.5 if positive else rng.standard_normal(shape)
    return ad.to_device(a)


def cpu_time(fn, nbytes_sample):
    """The reference's numpy expression for the same op on a bounded host sample (1 core: ufuncs and np.einsum are
    single-threaded): returns GB/s of algorithmic bytes."""
    import time
    fn()
    t0 = time.perf_counter()
    fn()
    return nbytes_sample / (time.perf_counter() - t0) / 1e9


SAMPLE = (2048, 2048)
hx = rng.standard_normal(SAMPLE); hy = rng.random(SAMPLE) + 0.5; hb = rng.standard_normal(SAMPLE[1])
hn = hx.size
CPU = {   # name prefix -> (numpy restatement of the reference's _eval body, algorithmic bytes of the sample)
    "copy": (lambda: np.concatenate((hx, np.ones((SAMPLE[0], 1))), axis=1), 16 * hn),
    "add x+y": (lambda: np.array(sum([hx, hy])), 24 * hn),
    "mul x*y*b": (lambda: (hx * hy) * hb, 24 * hn),
    "relu": (lambda: hx * (hx > 0), 16 * hn),
    "recipr": (lambda: 1 / (hy + 1e-12), 16 * hn),
    "exp": (lambda: np.exp(hx), 16 * hn),
    "sigmoid": (lambda: 1 / (1 + np.exp(-hx)), 16 * hn),
    "log": (lambda: np.log(hy + 1e-12), 16 * hn),
    "tanh macro": (lambda: (lambda v: (1 + -v) * (1 / (1 + v + 1e-12)))(np.exp(-2 * hx)), 16 * hn),
    "relu grad": (lambda: (lambda r: hy * r * (1 / (r + 1e-12)))(hx * (hx > 0)), 24 * hn),
    "sum rows": (lambda: np.einsum("ab->a", hx), 8 * hn),
    "sum cols": (lambda: np.einsum("ab->b", hx), 8 * hn),
    "sum all": (lambda: np.einsum("ab->", hx), 8 * hn),
    "softmax denominator": (lambda: np.einsum("bi,o->bo", hx, np.array([1])), 8 * hn),
    "frobenius": (lambda: np.sqrt(np.sum(np.square(hx))), 8 * hn),
    "softmax_ce": (lambda: (lambda z, l: -np.sum(l * z - l * (np.max(z, -1, keepdims=True) + np.log(np.sum(np.exp(z - np.max(z, -1, keepdims=True)), -1, keepdims=True))), -1))(hx, hy), 16 * hn),
    "adam": (lambda: (lambda m, v: hx - 1e-3 * m / (np.sqrt(v) + 1e-8))(0.9 * hx + 0.1 * hy, 0.999 * hy + 0.001 * hy * hy), 56 * hn),
}
h4 = rng.standard_normal((48, 48, 48, 48)); h2 = rng.standard_normal((48, 48))
CPU["C2c fwd"] = (lambda: np.einsum("ijkl,jl->ik", h4, h2), 8 * h4.size)
CPU["C2c dB"] = (lambda: np.einsum("ijkl,ik->jl", h4, h2), 8 * h4.size)
CPU["C2c dA"] = (lambda: np.einsum("ik,jl->ijkl", h2, h2), 8 * h4.size)
hg = rng.standard_normal((768, 768))
CPU["gemm"] = (lambda: np.einsum("ij,jk->ik", hg, hg), None)


def timeit(name, build, nbytes, flops=None, reps=8):
    """GPU time per op with `reps` ops queued back to back (graphs are built before the clock starts, so host
    graph-construction time is excluded and launch overhead overlaps with the previous kernel)."""
    if only and only not in name:
        return
    import time
    for _ in range(3):
        build().eval_device()
    torch.cuda.synchronize()
    best, host = 1e9, 1e9
    for _ in range(3):
        nodes = [build() for _ in range(reps)]
        l0 = ad.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for node in nodes:
            node.eval_device()
        e1.record()
        host = min(host, (time.perf_counter() - t0) / reps * 1e3)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
        launches = (ad.launch_count() - l0) // reps
        del nodes
    gbs = nbytes / (best * 1e-3) / 1e9
    row = {"name": name, "ms": round(best, 4), "GB/s": round(gbs, 1), "frac_hbm": round(gbs / PEAK, 3), "launches": launches,
           "host_ms": round(host, 4)}
    if flops:
        row["TFLOP/s"] = round(flops / (best * 1e-3) / 1e12, 2)
    for prefix, (fn, nb) in CPU.items():
        if name.startswith(prefix):
            if nb is None:
                import time
                fn(); t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
                row["numpy_1core_TFLOP/s"] = round(2.0 * 768 ** 3 / dt / 1e12, 5)
            else:
                row["numpy_1core_GB/s"] = round(cpu_time(fn, nb), 2)
            break
    results.append(row)
    print(row, flush=True)


R, Cc = 8192, 4096
n = R * Cc
x = dev((R, Cc)); y = dev((R, Cc), positive=True); b = dev((Cc,)); xo = dev((R, Cc + 1))
V = lambda d, name="v": ad.Variable(d, name=name)
timeit("copy (Concat column block) 8192x4096", lambda: ad.Concat(V(x), ad.ConstFill((R, 1), 1.0), axis=1), 8 * (2 * n + R))
timeit("add x+y", lambda: V(x) + V(y), 8 * 3 * n)
timeit("mul x*y*b", lambda: V(x) * V(y) * V(b), 8 * (3 * n + Cc))
timeit("relu", lambda: ad.ReLU(V(x)), 8 * 2 * n)
timeit("recipr", lambda: ad.Recipr(V(y)), 8 * 2 * n)
timeit("exp", lambda: ad.Exp(V(x)), 8 * 2 * n)
timeit("sigmoid", lambda: ad.Sigmoid(V(x)), 8 * 2 * n)
timeit("log", lambda: ad.Log(V(y)), 8 * 2 * n)
timeit("tanh macro (7 nodes)", lambda: ad.Tanh(V(x)), 8 * 2 * n)
timeit("relu grad prev*self*Recipr(self)", lambda: (lambda r: V(y) * r * ad.Recipr(r))(V(x)), 8 * 3 * n)
timeit("pitched relu 8192x4097", lambda: ad.ReLU(V(xo)), 8 * 2 * R * (Cc + 1))
timeit("sum rows ab->a", lambda: ad.Einsum("ab->a", V(x)), 8 * (n + R))
timeit("sum cols ab->b", lambda: ad.Einsum("ab->b", V(x)), 8 * (n + Cc))
timeit("sum all ab->", lambda: ad.Einsum("ab->", V(x)), 8 * n)
one = ad.to_device(np.array([1.0]))     # (the reference passes the literal np.array([1]); resident here so the clock sees the kernel, not its upload)
timeit("softmax denominator bi,o->bo", lambda: ad.Einsum("bi,o->bo", V(x), V(one)), 8 * (n + R))
timeit("frobenius", lambda: ad.FrobeniusNorm(V(x)), 8 * n)
lab = np.zeros((R, Cc)); lab[np.arange(R), rng.integers(0, Cc, R)] = 1.0
labd = ad.to_device(lab)
timeit("softmax_ce 8192x4096", lambda: ad.SoftmaxCEWithLogits(labels=V(labd), logits=V(x)), 8 * (2 * n + R))
os.environ["ADB_SCE_RING"] = "0"
timeit("softmax_ce 8192x4096 (register-resident kernel, ADB_SCE_RING=0)", lambda: ad.SoftmaxCEWithLogits(labels=V(labd), logits=V(x)), 8 * (2 * n + R))
del os.environ["ADB_SCE_RING"]
for cc in (512, 1024, 2048):
    rr = n // cc
    zz, ll = x.reshape_view((rr, cc)), labd.reshape_view((rr, cc))
    timeit("softmax_ce %dx%d" % (rr, cc), lambda: ad.SoftmaxCEWithLogits(labels=V(ll), logits=V(zz)), 8 * (2 * n + rr))
# the same reductions on a 1 GiB operand: separates the fixed ramp / tail cost of a 40 us kernel from its streaming rate
xl = dev((4 * R, Cc))
timeit("sum rows ab->a 1 GiB", lambda: ad.Einsum("ab->a", V(xl)), 8 * (4 * n + 4 * R))
timeit("sum cols ab->b 1 GiB", lambda: ad.Einsum("ab->b", V(xl)), 8 * (4 * n + Cc))
timeit("sum all ab-> 1 GiB", lambda: ad.Einsum("ab->", V(xl)), 8 * 4 * n)
timeit("frobenius 1 GiB", lambda: ad.FrobeniusNorm(V(xl)), 8 * 4 * n)
del xl
p, g, m, v = dev((4097, 4096)), dev((4097, 4096)), dev((4097, 4096)), dev((4097, 4096), positive=True)


class _AdamNode:
    def __init__(self):
        self.opt = ad.Adam(1, lr=1e-3); self.opt.m[0] = m; self.opt.v[0] = v; self.opt.step = 3

    def eval_device(self):
        return self.opt([p], [g])[0]


timeit("adam update 4097x4096 (4 in, 3 out)", lambda: _AdamNode(), 8 * 7 * 4097 * 4096)
del x, y, xo, labd, p, g, m, v
S = 128
A4 = dev((S, S, S, S)); B2 = dev((S, S)); G2 = dev((S, S))
nb = 8 * (S ** 4 + 2 * S * S)
timeit("C2c fwd ijkl,jl->ik", lambda: ad.Einsum("ijkl,jl->ik", V(A4), V(B2)), nb)
timeit("C2c dB ijkl,ik->jl", lambda: ad.Einsum("ijkl,ik->jl", V(A4), V(G2)), nb)
timeit("C2c dA ik,jl->ijkl", lambda: ad.Einsum("ik,jl->ijkl", V(G2), V(B2)), nb)
del A4
S = 64                                        # SURVEY 8d asks for the 64^4 member too: 128 MiB, 20 us of traffic - ramp/tail-bound
A4 = dev((S, S, S, S)); B2 = dev((S, S)); G2 = dev((S, S))
nb = 8 * (S ** 4 + 2 * S * S)
timeit("C2c 64^4 fwd ijkl,jl->ik", lambda: ad.Einsum("ijkl,jl->ik", V(A4), V(B2)), nb)
timeit("C2c 64^4 dB ijkl,ik->jl", lambda: ad.Einsum("ijkl,ik->jl", V(A4), V(G2)), nb)
timeit("C2c 64^4 dA ik,jl->ijkl", lambda: ad.Einsum("ik,jl->ijkl", V(G2), V(B2)), nb)
del A4
for (M, N, K, tag) in ((8192, 8192, 8192, "C2a"), (8192, 4096, 4097, "C3 fwd"), (1024, 4096, 4097, "C3 fwd b/8"), (512, 1024, 1024, "C5 h@w")):
    a, bb = dev((M, K)), dev((K, N))
    timeit("gemm %s %dx%dx%d" % (tag, M, N, K), lambda: V(a) @ V(bb), 8 * (M * K + K * N + M * N), flops=2.0 * M * N * K, reps=4)
    del a, bb
os.makedirs("gpurun_out", exist_ok=True)
json.dump(results, open("gpurun_out/kernels.json", "w"), indent=1)
open("gpurun_out/ewise_stats_kernels.txt", "w").write(ad.ewise_stats())
