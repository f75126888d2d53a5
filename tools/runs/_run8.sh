cd tools
for h in 1 0; do for p in 3 4; do for l in 0 1; do ./gemm_stress $p 2048 4096 4096 20 $h $l | tail -2; done; done; done
./gemm_stress 3 1024 4096 4096 30 1 0 | tail -2
./gemm_stress 3 512 1024 1024 50 1 0 | tail -2
./gemm_stress 0 8192 8192 8192 3 1 0 | tail -2
./gemm_stress 0 8192 8192 8192 3 0 0 | tail -2
cd ..
BISECT_REPS=3 python tools/dp_bisect.py b 2>&1 | grep "^B:" | cut -c1-150
