timeout 300 tools/gemm_selftest tf32 > gpurun_out/r2q_tf32_selftest.log 2>&1; echo rc=$?; grep -E "FAIL|correctness|time" gpurun_out/r2q_tf32_selftest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -c 1 -f -o gpurun_out/r2q_tf32x3 tools/gemm_selftest tf32one 8192 > gpurun_out/r2q_ncu.log 2>&1; echo ncu rc=$?
