timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r3i_bench_n8.json 2> gpurun_out/r3i_bench_n8.err; echo "bench n8 rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/r3i_bench_n8.json") if l.startswith("{")][0])
print(d['value'], d['ms_per_step'], {k:v for k,v in d['e2e'].items() if k not in ('api','host_link_ceiling')})
for g in ('mlp_train','other_configs'):
    for k,v in (d.get(g) or {}).items():
        print('  ',k, {a:b for a,b in v.items() if a in ('samples_per_s','ms_per_step','seq_per_s','s_per_eval','efficiency_vs_n1','exposed_comm_ms','host_ms','gemm_frac_of_peak','loss','error')}, (v.get('dp_parity') or {}).get('ok'))
p=d['mlp_train']['c3_mlp_4096x8_b8192']['dp_parity']
print({k:v for k,v in p.items() if k not in ('what','reference')})
P
