cd tools
for p in 3 4; do for h in 1 0; do ./gemm_stress $p 2048 4096 4096 30 $h 0; done; done
./gemm_stress 3 2048 4096 4096 30 1 1
./gemm_stress 3 1024 4096 4096 30 1 0
./gemm_stress 3 512 1024 1024 50 1 0
