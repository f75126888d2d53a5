cd tools
./gemm_selftest quick > ../gpurun_out/r2_selftest_quick.log 2>&1; tail -3 ../gpurun_out/r2_selftest_quick.log
ADB_GEMM_TAIL=0 ./gemm_selftest tail 2>&1 | grep -v "^ok\|correctness" > ../gpurun_out/r2_tail_off.log
ADB_GEMM_TAIL=1 ./gemm_selftest tail 2>&1 | grep -v "^ok\|correctness" > ../gpurun_out/r2_tail_on.log
echo "--- tail off"; cat ../gpurun_out/r2_tail_off.log; echo "--- tail on"; cat ../gpurun_out/r2_tail_on.log
for p in 0 12 13 14; do ./gemm_stress $p 8192 8192 8192 4 0 0 | tail -2; done
for p in 3 11; do ./gemm_stress $p 1024 4096 4096 10 0 0 | tail -2; done
cd ..
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest2.log 2>&1; tail -15 gpurun_out/r2_pytest2.log
