cd tools
for v in s6p4 s7p4 s7p5 s6p3; do
  echo "=== variant $v"
  ./gemm_selftest_$v 2>&1 | grep "time\|FAIL\|failing" > ../gpurun_out/r2b_selftest_$v.log
  ./gemm_selftest_$v tail 2>&1 | grep "time" >> ../gpurun_out/r2b_selftest_$v.log
  cat ../gpurun_out/r2b_selftest_$v.log
  for p in 3 4; do ./gemm_stress_$v $p 2048 4096 4096 20 1 0 | tail -1; done
  ./gemm_stress_$v 0 8192 8192 8192 3 1 0 | tail -1
done > ../gpurun_out/r2b_variants.log 2>&1
cd ..
tail -100 gpurun_out/r2b_variants.log
timeout 900 python -m pytest tests -m gpu -x -q -k "pad or softmax_ce_nd or reference_classes" 2>&1 | tail -5
