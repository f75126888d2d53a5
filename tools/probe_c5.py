"""C5 (char-rnn second order) alone: s per evaluation call by call vs as one CUDA graph, tape statistics, kernel time share.
    python tools/probe_c5.py            (ADB_TAPE_GRAPH=0 for call-by-call replay)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import autodiff_b200 as ad
from autodiff_b200 import training_bench as tb
from autodiff_b200.core import tape
ad.init(0)


class A:
    steps = 5


out = tb.bench_c5(A, None, tb._Timer(None), 1, 0, ad)
print(json.dumps({k: v for k, v in out.items() if k not in ("what", "roofline", "dp_parity")}))
print("tape stats", dict(tape.stats))
