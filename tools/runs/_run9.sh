timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; tail -15 gpurun_out/r2_pytest1.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench1.err; cut -c1-3000 gpurun_out/r2_bench1.json
