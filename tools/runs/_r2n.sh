python -m pytest tests -m gpu -x -q -k "softmax or sce or Softmax" 2>&1 | tail -3
for w in 1 0; do ADB_SCE_WARP=$w BENCH_SCE_SIMPLE=1 python tools/bench_sce.py 64 128 200 256 384 512 768 1024 2>&1 | grep rows; done | tee gpurun_out/r2n_sce_warp.log
