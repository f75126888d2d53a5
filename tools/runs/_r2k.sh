for path in 4 7; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_ -c 1 -f -o gpurun_out/r2k_path$path tools/gemm_selftest one 4736 4096 4096 1 1 0 $path > gpurun_out/r2k_ncu_$path.log 2>&1; echo rc=$?
done
