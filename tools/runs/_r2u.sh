python -m pytest tests -m gpu -x -q -k "softmax or sce or Softmax" 2>&1 | tail -4
for w in 129 0; do ADB_SCE_NARROW_RING=$w BENCH_SCE_SIMPLE=1 python tools/bench_sce.py 128 132 200 256 384 512 2>&1 | grep rows; done | tee gpurun_out/r2u_sce_narrow_ring.log
