timeout 300 tools/gemm_selftest sk > gpurun_out/r2i_sk.log 2>&1; echo rc=$?
grep -c "^ok" gpurun_out/r2i_sk.log; grep -E "FAIL|correctness|error" gpurun_out/r2i_sk.log | head -20
ADB_GEMM_SK=0 timeout 300 tools/gemm_selftest sk 2>&1 | grep "time.*path=0" > gpurun_out/r2i_sk_off.log
grep time gpurun_out/r2i_sk.log | tail -80
