timeout 120 tools/gemm_selftest skab 2>&1 | grep time
python -m pytest tests -m gpu -x -q > gpurun_out/r2m_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r2m_tests.log
python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo bench rc=$?
cut -c1-200 gpurun_out/r2m_bench.json
