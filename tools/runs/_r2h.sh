python -m pytest tests -m gpu -x -q > gpurun_out/r2h_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r2h_tests.log
python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo bench rc=$?
cut -c1-300 gpurun_out/r2h_bench.json
