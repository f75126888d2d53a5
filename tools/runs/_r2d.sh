timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2d_bench_n4.json 2> gpurun_out/r2d_bench_n4.err; echo "bench rc=$?"; tail -5 gpurun_out/r2d_bench_n4.err
python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/r2d_bench_n4.json") if l.startswith("{")][0])
print(d['value'], d['ms_per_step'], d['e2e'])
for g in ('mlp_train','other_configs'):
    for k,v in d[g].items():
        print(k, {a:b for a,b in v.items() if a not in ('what','includes','dp_parity','cpu_baseline','roofline')})
        print('   parity', v.get('dp_parity'))
P
