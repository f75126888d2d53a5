python -m pytest tests -m gpu -x -q -k "softmax or sce or Softmax or mlp or golden" 2>&1 | tail -2
for m in tab taylor; do ADB_SCE_EXP=$m BENCH_SCE_SIMPLE=1 python tools/bench_sce.py 128 256 512 1024 2048 4096 8192 2>&1 | grep rows | sed "s/^/$m /"; done | tee gpurun_out/r3o_sce_exp.log
