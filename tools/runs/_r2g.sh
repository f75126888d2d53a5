timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches_c4.csv python tools/bench_c4.py > gpurun_out/r2g_ncu_c4.log 2>&1; echo rc=$?
tail -2 gpurun_out/r2g_ncu_c4.log | cut -c1-300
