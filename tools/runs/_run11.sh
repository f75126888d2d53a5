cd tools
ADB_GEMM_TAIL=0 ./gemm_selftest tail 2>&1 | grep "time\|FAIL" > ../gpurun_out/r2_tail_off2.log
ADB_GEMM_TAIL=1 ./gemm_selftest tail 2>&1 | grep "time\|FAIL" > ../gpurun_out/r2_tail_on2.log
echo "--- tail off"; cat ../gpurun_out/r2_tail_off2.log; echo "--- tail on"; cat ../gpurun_out/r2_tail_on2.log
cd ..
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench2.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench2.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
for g in ('mlp_train','other_configs'):
    for k,v in d[g].items():
        print(k, {a:b for a,b in v.items() if a in ('samples_per_s','ms_per_step','seq_per_s','s_per_eval','graph_build_s','host_ms','gemm_tflops','error','trace')}, v.get('dp_parity',{}) and v['dp_parity'].get('ok'))
P
