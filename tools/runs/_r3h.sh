timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29529 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r3h_bench_n2.json 2> gpurun_out/r3h_bench_n2.err; echo "bench n2 rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/r3h_bench_n2.json") if l.startswith("{")][0])
print(d['value'], d['ms_per_step'], {k:v for k,v in d['e2e'].items() if k not in ('api','host_link_ceiling')})
p=d['mlp_train']['c3_mlp_4096x8_b8192']['dp_parity']
print({k:v for k,v in p.items() if k not in ('what','reference')})
P
