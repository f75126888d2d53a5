timeout 300 tools/gemm_selftest sk > gpurun_out/r2l_sk.log 2>&1; echo rc=$?
grep -c "^ok" gpurun_out/r2l_sk.log; grep -E "FAIL|correctness|error" gpurun_out/r2l_sk.log | head -20
grep "time" gpurun_out/r2l_sk.log | grep -E "path=(0|7)" 
timeout 120 tools/gemm_selftest skab 2>&1 | grep time
