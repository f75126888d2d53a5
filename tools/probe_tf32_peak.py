"""Peak probe (NOT product path): cuBLAS TF32 GEMM through torch.matmul (fp32 operands, allow_tf32), burst and sustained -
the roof the 3xTF32 tcgen05 variant (csrc/gemm_tf32x3.cu) is quoted against.  Writes gpurun_out/tf32_peak.json."""
import json, time, torch
torch.backends.cuda.matmul.allow_tf32 = True
out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "how": "torch.matmul fp32 with allow_tf32 (cuBLAS), CUDA events"}
for n in (4096, 8192):
    a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out["tf32_%d_tflops_burst" % n] = round(2.0 * n ** 3 / best * 1e-9, 1)
n = 8192
a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); it = 0
while time.perf_counter() - t0 < 4.0:
    for _ in range(20):
        a @ b
    it += 20
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
out["tf32_8192_tflops_sustained"] = round(2.0 * n ** 3 * it / e0.elapsed_time(e1) * 1e-9, 1)
torch.backends.cuda.matmul.allow_tf32 = False
a @ b; torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
out["fp32_8192_tflops_burst"] = round(2.0 * n ** 3 / e0.elapsed_time(e1) * 1e-9, 1)
print(json.dumps(out))
open("gpurun_out/tf32_peak.json", "w").write(json.dumps(out, indent=1))
