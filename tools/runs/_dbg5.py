import sys, faulthandler, numpy as np
faulthandler.dump_traceback_later(50, exit=True)
sys.path.insert(0, '.')
import autodiff_b200 as ad
import oracle.np_autodiff as onp
ad.init(0)
rng = np.random.default_rng(31)
spec = sys.argv[1]; shapes = eval(sys.argv[2])
vals = [rng.standard_normal(s) for s in shapes]
vs = [ad.Variable(v, name="v%d" % i) for i, v in enumerate(vals)]
e = ad.Einsum(spec, *vs)
print("fwd", float(np.abs(e() - np.einsum(spec, *vals)).max()), flush=True)
gs = ad.grad(e, vs)
ovs = [onp.Variable(v, name="v%d" % i) for i, v in enumerate(vals)]
og = onp.grad(onp.Einsum(spec, *ovs), ovs)
for k, g in enumerate(gs):
    print("grad", k, g.op_str if hasattr(g, "op_str") else type(g).__name__, flush=True)
    print("   err", float(np.abs(g() - og[k]()).max()), flush=True)
