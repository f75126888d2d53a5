BISECT_REPS=2 python tools/dp_bisect.py d > gpurun_out/r2_bisect5_d.log 2>&1
BISECT_REPS=2 ADB_GEMM_TILE=big python tools/dp_bisect.py d > gpurun_out/r2_bisect5_d_big.log 2>&1
BISECT_REPS=2 ADB_GEMM_BLOCKED=0 python tools/dp_bisect.py d > gpurun_out/r2_bisect5_d_unblocked.log 2>&1
grep -h "^D " gpurun_out/r2_bisect5_d.log | cut -c1-330
echo ==== big; grep -h "^D " gpurun_out/r2_bisect5_d_big.log | grep -v Variable | cut -c1-330 | tail -30
echo ==== unblocked; grep -h "^D " gpurun_out/r2_bisect5_d_unblocked.log | grep -v Variable | cut -c1-330 | tail -30
