cd tools
for p in 3 7 8 9 10; do ./gemm_stress $p 2048 4096 4096 20 1 0 | tail -1; done
for h in 2 3 4; do ./gemm_stress 3 2048 4096 4096 20 $h 0 | tail -1; done
