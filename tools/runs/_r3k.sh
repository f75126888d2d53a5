timeout 200 tools/gemm_selftest sk check 2>&1 | grep -E "FAIL|correctness"
python -m pytest tests -m gpu -x -q -k "stream_k or baseline_size or host_logic or c_abi" 2>&1 | tail -2
tools/gemm_selftest one 8192 8192 8192 1 1 0 0 2>&1 | grep time
