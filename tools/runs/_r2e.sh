cd tools && timeout 120 ./conv_selftest time > ../gpurun_out/r2e_conv_selftest2.log 2>&1; echo rc=$?; cat ../gpurun_out/r2e_conv_selftest2.log | grep -v "^ok   forward\|^ok   data"
cd .. && timeout 200 python tools/bench_c4.py 2>&1 | tail -2
