python tools/dp_bisect.py c > gpurun_out/r2_bisect3_c.log 2>&1
export BISECT_QUICK=1 BISECT_REPS=2
for m in inline syncboth syncbefore syncafter; do ADB_DP_DEBUG=$m python tools/dp_bisect.py b > gpurun_out/r2_bisect3_$m.log 2>&1; done
ADB_EDGE_SPLIT=0 python tools/dp_bisect.py b > gpurun_out/r2_bisect3_noedge.log 2>&1
PYTORCH_NO_CUDA_MEMORY_CACHING=1 python tools/dp_bisect.py b > gpurun_out/r2_bisect3_nocache.log 2>&1
grep -h "^B:\|^C \|Error\|error" gpurun_out/r2_bisect3_*.log | tail -60
