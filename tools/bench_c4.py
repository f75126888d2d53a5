"""C4 alone (conv net, batch 256, forward + parameter gradients) through training_bench.bench_c4; ADB_IMPLICIT_CONV=0 for the explicit path."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import autodiff_b200 as ad
from autodiff_b200 import training_bench as tb
ad.init(0)


class A:
    steps = 10


out = tb.bench_c4(A, None, tb._Timer(None), 1, 0, ad)
print(json.dumps({k: v for k, v in out.items() if k not in ("what", "roofline")}))
