python -m pytest tests -m gpu -x -q -k "tf32" 2>&1 | tail -2
ADB_TF32_2CTA=2 python -m pytest tests -m gpu -x -q -k "tf32" 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -c 1 -f -o gpurun_out/r2y_tf32x3_pair tools/gemm_selftest tf32one 8192 > gpurun_out/r2y_ncu.log 2>&1; echo ncu rc=$?
timeout 200 tools/gemm_selftest tf32 2>&1 | grep time
