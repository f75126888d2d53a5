python -m pytest tests -m gpu -x -q > gpurun_out/r2t_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r2t_tests.log
python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo bench rc=$?
cut -c1-200 gpurun_out/r2t_bench.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_sk -c 1 -f -o gpurun_out/r2t_gemm_sk_8192 tools/gemm_selftest one 8192 8192 8192 1 1 0 0 > gpurun_out/r2t_ncu_sk.log 2>&1; echo ncu rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2t_launches_sweep.csv python bench.py --steps 2 --warmup 1 --skip-mlp --skip-cpu --skip-e2e > gpurun_out/r2t_ncu_bench.log 2>&1; echo launches rc=$?
