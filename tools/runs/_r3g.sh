python -m pytest tests -m gpu -x -q > gpurun_out/r3g_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r3g_tests.log
python bench.py > gpurun_out/r3g_bench.json 2> gpurun_out/r3g_bench.err; echo bench rc=$?
python -c "
import json
d=json.load(open('gpurun_out/r3g_bench.json'))
print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline_hbm']['frac'], d['gpu_launches'])
for g in ('mlp_train','other_configs'):
    for k,v in d[g].items():
        print(k, {a:b for a,b in v.items() if a in ('samples_per_s','ms_per_step','seq_per_s','s_per_eval','host_ms','gemm_tflops')}, v['dp_parity']['ok'] if v.get('dp_parity') else None)
t=d['tf32x3_variant']; print({k:v for k,v in t.items() if k not in ('what','roofline')}, t['roofline']['frac'])
"
python -c "import __graft_entry__ as g; g.smoke()"
