ADB_TF32_2CTA=2 timeout 120 tools/gemm_selftest tf32 > gpurun_out/r2x_tf32_pair.log 2>&1; echo rc=$?; grep -E "FAIL|correctness|time|error" gpurun_out/r2x_tf32_pair.log | head -40
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
