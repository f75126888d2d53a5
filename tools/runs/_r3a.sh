timeout 300 tools/gemm_selftest quick > gpurun_out/r3a_selftest_quick.log 2>&1; echo rc=$?; grep -E "FAIL|correctness" gpurun_out/r3a_selftest_quick.log | head
timeout 200 tools/gemm_selftest sk check 2>&1 | grep -E "FAIL|correctness"
timeout 200 tools/gemm_selftest sanitize 2>&1 | grep -E "FAIL|sanitize set"
ADB_GEMM_SK=0 timeout 200 tools/gemm_selftest tail 2>&1 | grep -E "FAIL|time" | head -12
python -m pytest tests -m gpu -x -q -k "gemm or einsum or dp or conv" 2>&1 | tail -2
