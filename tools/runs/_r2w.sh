python -m pytest tests -m gpu -x -q -k "eval_many or pipelin" 2>&1 | tail -2
python tools/e2e_trace.py 2>&1 | head -26
python bench.py --skip-mlp --skip-cpu > gpurun_out/r2w_bench.json 2>gpurun_out/r2w_bench.err; echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r2w_bench.json')); print(d['value'], {k:v for k,v in d['e2e'].items() if k not in ('api','host_link_ceiling')})"
