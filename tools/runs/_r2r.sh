# 8-GPU box: host-link ceiling at N = 1, 2, 4, 8, then the bench at N = 8 (bound / unbound ranks) and N = 2
python tools/probe_host_link.py 2>&1 | tail -1
for n in 2 4 8; do python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n tools/probe_host_link.py 2>&1 | grep '^{' ; done
nvidia-smi topo -m > gpurun_out/r2r_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2r_bench_n8.json 2> gpurun_out/r2r_bench_n8.err; echo "bench n8 rc=$?"
ADB_BENCH_BIND=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 5 --warmup 3 --skip-mlp > gpurun_out/r2r_bench_n8_unbound.json 2> gpurun_out/r2r_bench_n8_unbound.err; echo "bench n8 unbound rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29529 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2r_bench_n2.json 2> gpurun_out/r2r_bench_n2.err; echo "bench n2 rc=$?"
python - <<'P'
import json
for f in ("r2r_bench_n8","r2r_bench_n8_unbound","r2r_bench_n2"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][0])
    except Exception as ex:
        print(f, "no line", ex); continue
    print(f, d['value'], d['ms_per_step'], {k:v for k,v in d['e2e'].items() if k!='api'})
    for g in ('mlp_train','other_configs'):
        for k,v in (d.get(g) or {}).items():
            print('  ',k, {a:b for a,b in v.items() if a in ('samples_per_s','ms_per_step','seq_per_s','s_per_eval','efficiency_vs_n1','exposed_comm_ms','host_ms','gemm_frac_of_peak','loss','error')}, (v.get('dp_parity') or {}).get('ok'))
P
