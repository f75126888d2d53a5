"""bench.py --impl reference (the CPU arm the driver runs beside the GPU arm): one JSON line with the contract's keys on rank 0,
nothing from the other ranks.  Runs the unmodified reference from baseline/_ref (the numpy port when it is not installed) on the
bounded sample once (~5-10 s of one core)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                          cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)


def test_reference_arm_prints_one_contract_line_on_rank_0():
    r = _run({"RANK": "0", "WORLD_SIZE": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "TFLOP/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["metric"].startswith("einsum fwd+grad TFLOP/s") and d["steps"] == 1 and d["n_gpus"] == 1
    assert 0 < d["value"] < 1.0 and d["ms_per_step"] > 0
    have_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "autodiff"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert "1536x1536" in d["cpu_baseline"]["sample"] and "workload" in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    r = _run({"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""
