"""Peak probe (NOT product path): cuBLAS DGEMM through torch.matmul, fp64, burst and sustained.
Writes gpurun_out/fp64_peak.json; the committed copy lives in profiles/fp64_peak.json."""
import json, time, subprocess, threading, torch

def run(n, trans, iters):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    fa = (lambda: a.t()) if trans[0] == "T" else (lambda: a)
    fb = (lambda: b.t()) if trans[1] == "T" else (lambda: b)
    for _ in range(2):
        torch.matmul(fa(), fb())
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(fa(), fb()); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / best * 1e-9, best

out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "how": "torch.matmul fp64 (cuBLAS DGEMM), CUDA events"}
for n in (4096, 8192):
    for tr in ("NN", "NT", "TN"):
        tf, ms = run(n, tr, 5)
        out[f"dgemm_{tr}_{n}_tflops_burst"] = round(tf, 2)
        print(n, tr, f"{tf:.2f} TFLOP/s {ms:.2f} ms", flush=True)
# sustained: back-to-back for ~4 s
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
clk = []
stop = False
def sample():
    while not stop:
        try:
            r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True, timeout=5)
            clk.append(r.stdout.strip())
        except Exception as ex:
            clk.append(str(ex))
        time.sleep(0.2)
th = threading.Thread(target=sample); th.start()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
reps = 120
e0.record()
for _ in range(reps):
    torch.matmul(a, b)
e1.record(); torch.cuda.synchronize()
stop = True; th.join()
ms = e0.elapsed_time(e1)
out["dgemm_NN_8192_tflops_sustained"] = round(2.0 * n ** 3 * reps / ms * 1e-9, 2)
out["sustained_seconds"] = round(ms / 1e3, 2)
out["clock_samples"] = clk[:40]
print(json.dumps(out, indent=1))
import os
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fp64_peak.json", "w"), indent=1)
