#!/usr/bin/env python
"""bench.py — headline benchmark of the autodiff_b200 hot path (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--skip-e2e] [--skip-mlp]

Metric (BASELINE.json): "einsum fwd+grad TFLOP/s (fp64)" on the einsum sweep (configs[1]):
    C2a 'ij,jk->ik' 8192^2     C2b 'bij,bjk->bik' b=64 1024^2     C2c 'ijkl,jl->ik' A=(128,)*4
each evaluated forward + ad.grad w.r.t. both operands (9 einsums per step, 3.712e12 flop).  One "step" is one
pass over the sweep on synthetic N(0,1) fp64 inputs.  With --gpus N every rank runs the same sweep on its own
GPU (the sweep has no exchange step: "replicas only") and `value` is the aggregate; the batch-
sharded MLP training step WITH its NCCL gradient all-reduce (configs[0] and configs[2]) is reported in the same
JSON line under "mlp_train".

value  : inputs resident in HBM, results stay in HBM, CUDA-event timed.
e2e    : the same sweep through the public API with host buffers: ad.Variable(pinned ndarray) ... node() ->
         ndarray; host->device and device->host copies inside the timed region.
--impl reference : the UNMODIFIED reference (baseline/_ref, installed by baseline/install_ref.sh; np.einsum without BLAS,
         single thread) timed on the host cores on a bounded sample of the same sweep; the numpy port
         oracle/np_autodiff.py stands in only when baseline/_ref is absent.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "einsum fwd+grad TFLOP/s (fp64)"
FULL = {"C2a": ("ij,jk->ik", [(8192, 8192), (8192, 8192)]),
        "C2b": ("bij,bjk->bik", [(64, 1024, 1024), (64, 1024, 1024)]),
        "C2c": ("ijkl,jl->ik", [(128, 128, 128, 128), (128, 128)])}
# bounded sample of the same sweep for the CPU arm: ~8 s of single-threaded np.einsum per pass on the GPU box's host
SAMPLE = {"C2a": ("ij,jk->ik", [(1536, 1536), (1536, 1536)]),
          "C2b": ("bij,bjk->bik", [(16, 384, 384), (16, 384, 384)]),
          "C2c": ("ijkl,jl->ik", [(64, 64, 64, 64), (64, 64)])}


def einsum_flops(spec, shapes):
    ext = {}
    for sub, shp in zip(spec.split("->")[0].split(","), shapes):
        for l, d in zip(sub, shp):
            ext[l] = d
    f = 2.0
    for d in ext.values():
        f *= d
    return f


def sweep_flops(cfg):
    return sum(3 * einsum_flops(spec, shapes) for spec, shapes in cfg.values())   # fwd + 2 grads each


def sweep_bytes(cfg):
    """Algorithmic bytes of the HBM-bound member (C2c): 8*(|A|+|B|+|out|) per einsum, SURVEY.md 8d."""
    spec, shapes = cfg["C2c"]
    a = int(np.prod(shapes[0])); b = int(np.prod(shapes[1])); o = shapes[0][0] * shapes[0][2]
    return 3 * 8 * (a + b + o)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (recipe in B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.rows, self.stop = index, [], False
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop:
            try:
                r = subprocess.run(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=5)
                parts = [p.strip() for p in r.stdout.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(r[3 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "power_w_max": max(float(r[2]) for r in self.rows),
                "samples": len(self.rows), "reasons": reasons}


# ------------------------------------------------------------------------------------------------- GPU arm
def build_sweep(ad, cfg, inputs):
    """One step's graphs: E = Einsum(spec, A, B); dA, dB = ad.grad(E, [A, B]) for each member of the sweep."""
    graphs = []
    for name, (spec, _) in cfg.items():
        a = ad.Variable(inputs[name][0], name=name + "_A")
        b = ad.Variable(inputs[name][1], name=name + "_B")
        e = ad.Einsum(spec, a, b)
        da, db = ad.grad(e, [a, b])
        graphs.append((name, e, da, db))
    return graphs


def run_gpu(args):
    import torch
    import autodiff_b200 as ad
    from autodiff_b200.core import engine
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ad.init(local)
    dist = None
    if world > 1:
        # one slice of the allowed host cores per rank: the upload / download staging of N ranks shares one NUMA node on this box
        # class (topology: all GPUs report CPU affinity 0-31, node 0), so the ranks are at least kept off each other's cores
        try:
            if os.environ.get("ADB_BENCH_BIND", "1") == "0":
                raise OSError
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            mine = cores[(local * per) % len(cores):][:per]
            if mine:
                os.sched_setaffinity(0, mine)
        except (AttributeError, OSError):
            pass
        import torch.distributed as dist
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=600))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cfg = FULL
    rng = np.random.default_rng(1337 + rank)
    host_inputs = {}
    for name, (spec, shapes) in cfg.items():
        host_inputs[name] = []
        for shp in shapes:
            buf = ad.pinned_empty(shp)
            rng.standard_normal(out=buf.reshape(-1))
            host_inputs[name].append(buf)
    dev_inputs = {n: [ad.to_device(v) for v in vs] for n, vs in host_inputs.items()}
    flops = sweep_flops(cfg)

    def step_resident(timers=None):
        for name, e, da, db in build_sweep(ad, cfg, dev_inputs):
            for node in (e, da, db):
                node.eval_device()

    def step_e2e():
        nodes = [node for name, e, da, db in build_sweep(ad, cfg, host_inputs) for node in (e, da, db)]
        hosts = ad.eval_many(nodes)          # == [n() for n in nodes]: host ndarrays, uploads/kernels/downloads pipelined
        return sum(float(h.reshape(-1)[0]) for h in hosts)

    # ---- resident (kernel-path) number
    for _ in range(args.warmup):
        step_resident()
    barrier()
    engine.kernel_timings = []           # per-einsum CUDA-event pairs recorded by Einsum._eval_device
    launches0 = ad.launch_count()
    with ClockSampler(local) as clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step_resident()
        e1.record()
        barrier()
    launches = ad.launch_count() - launches0
    timings = engine.kernel_timings
    engine.kernel_timings = None
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    value = world * flops / (ms * 1e-3) / 1e12

    # ---- roofline of the dominant kernel: the 8192^3 DMMA GEMM launches of C2a (89% of the step's flops)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json")))
    except Exception:
        pass
    gemm_ms = [a.elapsed_time(b) for (desc, a, b) in timings if "C2a" in desc]
    gemm_flops = einsum_flops(*cfg["C2a"])
    roof = None
    if gemm_ms:
        avg = sum(gemm_ms) / len(gemm_ms)
        achieved = gemm_flops / (avg * 1e-3) / 1e12
        peak = float(peaks.get("dmma_issue_peak_tflops", 37.03))
        roof = {"bound": "tensor", "kernel": "gemm_sk_kernel (stream-K, FP64 DMMA.8x8x4), C2a launches 2*8192^3 flop each",
                "achieved": round(achieved, 3), "peak": peak, "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
                "avg_launch_ms": round(avg, 4), "launches_timed": len(gemm_ms), "traffic": peaks.get("c2a_gemm_dram_bytes_per_launch"),
                "peak_source": "measured DMMA.8x8x4 issue-rate peak on this pool's B200 (profiles/r01_probe_fp64_issue_rates.log); "
                               "MEASURED_PEAKS.json has no fp64 entry; cuBLAS DGEMM sustained measured 35.39"}
    c2c_ms = [a.elapsed_time(b) for (desc, a, b) in timings if "C2c" in desc]
    hbm = None
    if c2c_ms:
        mp = {}
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        pk = float(mp.get("hbm_gbs", 6650.0))
        per_einsum = sweep_bytes(cfg) / 3
        avg = sum(c2c_ms) / len(c2c_ms)
        hbm = {"bound": "hbm", "kernel": "C2c einsums (reduce_kernel / ewise_kernel)", "achieved": round(per_einsum / (avg * 1e-3) / 1e9, 1),
               "peak": pk, "unit": "GB/s", "frac": round(per_einsum / (avg * 1e-3) / 1e9 / pk, 4),
               "peak_source": "MEASURED_PEAKS.json hbm_gbs" if mp else "fallback 6650 GB/s"}

    # ---- end-to-end number (host buffers in, host arrays out)
    e2e = None
    if not args.skip_e2e:
        for _ in range(max(1, min(args.warmup, 2))):
            step_e2e()
        barrier()
        k = max(1, min(args.steps, 5))
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        t0.record()
        for _ in range(k):
            step_e2e()
        t1.record()
        barrier()
        wall = (time.perf_counter() - w0) * 1e3
        ems = max_over_ranks(max(t0.elapsed_time(t1), wall)) / k
        h2d = sum(int(np.prod(s)) * 8 for _, shapes in cfg.values() for s in shapes)
        d2h = 0
        for spec, shapes in cfg.values():
            ext = {}
            for sub, shp in zip(spec.split("->")[0].split(","), shapes):
                ext.update(dict(zip(sub, shp)))
            out_elems = int(np.prod([ext[l] for l in spec.split("->")[1]]))
            d2h += 8 * (out_elems + int(np.prod(shapes[0])) + int(np.prod(shapes[1])))     # E, dA, dB
        e2e = {"value": round(world * flops / (ems * 1e-3) / 1e12, 4), "unit": "TFLOP/s", "ms_per_step": round(ems, 3), "steps": k,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "h2d_gbs_per_rank": round(h2d / (ems * 1e-3) / 1e9, 1), "d2h_gbs_per_rank": round(d2h / (ems * 1e-3) / 1e9, 1),
               "host_link_gbs_all_ranks": round(world * (h2d + d2h) / (ems * 1e-3) / 1e9, 1),
               "host_link_ceiling": "tools/probe_host_link.py, profiles/r02_host_link.md: aggregate pinned-copy GB/s per direction at N = 1, 2, 4, 8 on this box class",
               "api": "ad.Variable(pinned ndarray) -> ad.Einsum / ad.grad -> ad.eval_many(nodes) returns host ndarrays "
                      "(== [n() for n in nodes]; H2D, kernels and D2H pipelined on three streams)"}

    # ---- the training configs (C1, C3, C4, C5): batch-sharded over all ranks, each with dp_parity + roofline
    mlp, extra, parity_ok = None, None, True
    if not args.skip_mlp:
        from autodiff_b200 import training_bench
        mlp, extra, parity_ok = training_bench.bench_training(args, dist, ad)
        if rank == 0 and world == 1 and not args.skip_cpu:
            attach_cpu_baselines(mlp, extra)

    # ---- the TF32 variant of the contraction path (north_star: 1e-4 bar): C2a forward + both grads as 3xTF32 on tcgen05
    tf32 = None
    if rank == 0 and world == 1 and not args.skip_mlp:
        try:
            def tf_graph():
                x = ad.Variable(dev_inputs["C2a"][0], name="A"); y = ad.Variable(dev_inputs["C2a"][1], name="B")
                e = ad.Einsum("ij,jk->ik", x, y)
                return [o.eval_device() for o in [e] + ad.grad(e, [x, y])]
            ref = tf_graph()                                                    # the FP64 DMMA path: forward + both grads

            def tf_step():
                with ad.precision("tf32x3"):
                    return tf_graph()
            for _ in range(2):
                got = tf_step()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(3):
                got = tf_step()
            t1.record()
            torch.cuda.synchronize()

            def view(d):
                return torch.as_strided(d.base, tuple(d.shape), tuple(d.strides), d.offset)
            rels = []
            for g, r in zip(got, ref):                                          # over the FULL matrices, on the device
                tg, tr = view(g), view(r)
                rels.append(float((tg - tr).abs().max() / tr.abs().max()))
            rel = max(rels)
            tf_ms = t0.elapsed_time(t1) / 3
            eff = 3 * einsum_flops(*cfg["C2a"]) / (tf_ms * 1e-3) / 1e12
            tf_peak, tf_src = 753.4, "cuBLAS TF32 8192^3 burst on this pool's B200 (tools/probe_tf32_peak.py, profiles/tf32_peak.json; sustained 628.9)"
            tf32 = {"tflops_effective": round(eff, 2), "ms_per_step": round(tf_ms, 3),
                    "rel_err_vs_fp64": rel, "rel_err_fwd_dA_dB": rels, "bar": 1e-4,
                    "roofline": {"bound": "tensor", "achieved": round(3 * eff, 1), "peak": tf_peak, "unit": "TFLOP/s (raw tf32 MMA: 3 products per fp64 product; "
                                 "split passes and the fp64 epilogue inside the timed region)", "frac": round(3 * eff / tf_peak, 4), "peak_source": tf_src,
                                 "traffic": 8247000000, "traffic_note": "ncu --set full, gemm_tf32x3_2cta_kernel at 8192^3 (profiles/r02_ncu_summary.md): 6.43 GB read + 1.82 GB "
                                 "written per launch vs 1.07 GB of fp32 hi/lo operands + 0.54 GB of fp64 result; kernel alone 4.23 ms = 780 TFLOP/s raw; tensor pipe 78 % active"},
                    "what": "C2a 'ij,jk->ik' 8192^2 forward + both grads with ad.precision('tf32x3'): fp64 operands split into hi+lo tf32, "
                            "three tcgen05.mma.cta_group::2 (kind::tf32) products per 256x256 CTA-pair tile accumulated in TMEM in chains of <= 2048, chains added in fp64 by the epilogue warps; "
                            "fp64 in / fp64 out; includes the split passes; error measured over the full matrices against the FP64 path"}
            del got, ref
        except Exception as ex:
            tf32 = {"error": repr(ex)[:300]}

    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cpu = run_reference_sample(reps=2, courtesy=True)

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 3), "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "impl": "autodiff_b200",
                "config": {"workload": "einsum sweep fp64 (BASELINE configs[1]): 'ij,jk->ik' 8192^2 + 'bij,bjk->bik' b=64 1024^2 + "
                                       "'ijkl,jl->ik' 128^4, each with ad.grad wrt both operands",
                           "flop_per_step": flops, "l2": "inputs (512 MiB - 2 GiB each) exceed the 126 MB L2",
                           "parallelism": "replicas x%d (the sweep has no exchange step)" % world},
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches) * world,
                "roofline": roof, "roofline_hbm": hbm, "cpu_baseline": cpu, "mlp_train": mlp, "other_configs": extra, "tf32x3_variant": tf32}
        print(json.dumps(line))
        if os.environ.get("ADB_EW_STATS"):        # which elementwise programs ran (input of tools/gen_ewise_catalog.py)
            open(os.environ["ADB_EW_STATS"], "w").write(ad.ewise_stats())
    if dist is not None:
        dist.destroy_process_group()
    if rank == 0 and not parity_ok:
        sys.stderr.write("bench.py: a data-parallel parity field is above the 1e-12 bar (or a training config failed); see dp_parity in the line above\n")
        sys.exit(3)


# ------------------------------------------------------------------------------------------------- CPU arm
def run_reference_sample(reps=1, courtesy=False):
    """The reference's own implementation (baseline/_ref through baseline/ref_shim.py; the numpy port when that is not
    installed) of the sweep on a bounded sample: np.einsum without BLAS, single thread (reference ops.py:180)."""
    from baseline import cpu_arm
    out = cpu_arm.sweep(SAMPLE, reps=reps)
    if courtesy:
        # NOT the reference path (it never calls BLAS): what all host cores reach on the same matmul through numpy's BLAS
        rng = np.random.default_rng(1337)
        a, b = rng.standard_normal(SAMPLE["C2a"][1][0]), rng.standard_normal(SAMPLE["C2a"][1][1])
        t0 = time.perf_counter()
        _ = a @ b
        dt = time.perf_counter() - t0
        out["blas_all_cores_courtesy"] = {"tflops": round(2.0 * a.shape[0] ** 3 / dt / 1e12, 4), "what": "numpy a @ b (OpenBLAS, all cores) on the 1536^2 member; "
                                          "not what the reference executes"}
    return out


def attach_cpu_baselines(mlp, extra):
    """`cpu_baseline` beside every training number: the reference's own loop timed on this box's host cores (bounded samples)."""
    from baseline import cpu_arm
    jobs = [(mlp, "c1_mlp_784_100_10_b64", lambda: cpu_arm.mlp_train([784, 100, 10], 64, steps=30, warmup=3)),
            (mlp, "c3_mlp_4096x8_b8192", lambda: cpu_arm.mlp_train([4096] * 9, 16, steps=1, warmup=0)),
            (extra, "c4_conv_b256", lambda: cpu_arm.conv_train(4)),
            (extra, "c5_char_rnn_h1024_t128_b512_second_order", lambda: cpu_arm.char_rnn_second_order(1024, 16, 64, 71))]
    for grp, key, fn in jobs:
        if isinstance(grp, dict) and isinstance(grp.get(key), dict):
            try:
                grp[key]["cpu_baseline"] = fn()
            except Exception as ex:
                grp[key]["cpu_baseline"] = {"error": repr(ex)[:200]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(args.warmup):
        run_reference_sample()
    t0 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        last = run_reference_sample()
    dt = (time.perf_counter() - t0) / args.steps
    flops = sweep_flops(SAMPLE)
    v = round(flops / dt / 1e12, 6)
    last["value"] = v
    line = {"metric": METRIC, "value": v, "unit": "TFLOP/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": "einsum sweep fp64 (BASELINE configs[1]), bounded sample per step: " + last["sample"],
                       "reference": "unmodified bgavran/autodiff from baseline/_ref (graphviz stub + tuple-ised slice lists only)"
                                    if last["kind"] == "reference" else "numpy port oracle/np_autodiff.py (baseline/_ref not installed)"},
            "cpu_baseline": last, "e2e": {"value": v, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="autodiff_b200")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-mlp", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
