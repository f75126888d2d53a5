for mu in 8 16 21 32 64; do for path in 7 8; do echo -n "min_units=$mu "; ADB_GEMM_SK_MIN_UNITS=$mu tools/gemm_selftest one 512 1024 1024 1 1 0 $path 2>&1 | grep time; done; done
for path in 3 4; do tools/gemm_selftest one 512 1024 1024 1 1 0 $path 2>&1 | grep time; done
for mu in 8 16 32; do for path in 7 8; do echo -n "min_units=$mu "; ADB_GEMM_SK_MIN_UNITS=$mu tools/gemm_selftest one 1024 1024 512 1 0 0 $path 2>&1 | grep time; done; done
tools/gemm_selftest one 1024 1024 512 1 0 0 3 2>&1 | grep time
