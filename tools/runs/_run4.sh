python tools/dp_bisect.py d > gpurun_out/r2_bisect4_d.log 2>&1
export BISECT_QUICK=1 BISECT_REPS=3
ADB_DP_DEBUG=allocmain python tools/dp_bisect.py b > gpurun_out/r2_bisect4_allocmain.log 2>&1
grep -h "^B:\|^D \|Error\|error" gpurun_out/r2_bisect4_*.log | cut -c1-400 | tail -60
