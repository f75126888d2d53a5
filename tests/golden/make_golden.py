"""Generate the golden vectors by running the REAL reference (/root/reference) in the build container.

    python tests/golden/make_golden.py        # writes tests/golden/*.npz

The reference is imported unmodified, with the two host-side shims of SURVEY.md Appendix D (a stub
``graphviz`` module, and tuple-ising the list-of-slices index that numpy >= 1.23 rejects); no arithmetic is
touched.  /root/reference does not exist on the GPU box, so only the committed .npz files travel.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.environ.get("AUTODIFF_REFERENCE", "/root/reference")


def load_reference():
    g = types.ModuleType("graphviz")
    g.Digraph = type("Digraph", (), {"__init__": lambda s, *a, **k: None})
    sys.modules["graphviz"] = g
    sys.path.insert(0, REF_DIR)
    sys.dont_write_bytecode = True
    sys.setrecursionlimit(20000)
    import autodiff as ref
    from autodiff.core import reshape as _r
    _orig = _r.Slice.__init__

    def _init(self, node, slice_val, name="Slice"):
        _orig(self, node, tuple(slice_val) if isinstance(slice_val, list) else slice_val, name)
    _r.Slice.__init__ = _init
    from autodiff.core.grad import grad
    ref.grad = grad
    sys.path.pop(0)
    return ref


def main():
    ref = load_reference()
    sys.path.insert(0, os.path.dirname(HERE))
    import cases
    out = {}
    for name in cases.CASES:
        nodes = cases.derive(ref, name)
        for key, node in nodes.items():
            out[name + "/" + key] = np.array(node())
    np.savez_compressed(os.path.join(HERE, "cases.npz"), **out)
    print("cases.npz:", len(out), "arrays")

    tanh = {"d%d" % i: np.array(n()) for i, n in enumerate(cases.tanh_derivatives(ref))}
    np.savez_compressed(os.path.join(HERE, "tanh_derivs.npz"), **tanh)

    hist, weights = cases.mlp_train_steps(ref)
    mlp = {}
    for s, h in enumerate(hist):
        mlp["step%d/cost" % s] = h["cost"]
        mlp["step%d/gnorm" % s] = h["gnorm"]
        for i, g in enumerate(h["grads"]):
            mlp["step%d/grad%d" % (s, i)] = g
    for i, w in enumerate(weights):
        mlp["final/w%d" % i] = w
    np.savez_compressed(os.path.join(HERE, "mlp_steps.npz"), **mlp)

    total, params = cases.char_rnn_graph(ref)
    g1 = ref.grad(total, params)
    rnn = {"loss": np.array(total())}
    for i, g in enumerate(g1):
        rnn["g%d" % i] = np.array(g())
    gg = ref.grad(g1[0], [params[0]])[0]
    rnn["gg_w"] = np.array(gg())
    np.savez_compressed(os.path.join(HERE, "char_rnn.npz"), **rnn)

    # structural golden: the einsum strings ad.grad produces (host logic, checked without a GPU)
    lines = []
    for name in sorted(cases.CASES):
        nodes = cases.derive(ref, name)
        for key in sorted(nodes):
            lines.append(name + "/" + key + "\t" + " ".join(sorted(_einsum_strings(nodes[key]))))
    with open(os.path.join(HERE, "grad_einsum_strings.tsv"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("done")


def _einsum_strings(top):
    seen, stack, out = set(), [top], []
    while stack:
        n = stack.pop()
        if id(n) in seen:
            continue
        seen.add(id(n))
        if hasattr(n, "op_str"):
            out.append(n.op_str)
        stack.extend(n.children)
    return out


if __name__ == "__main__":
    main()
