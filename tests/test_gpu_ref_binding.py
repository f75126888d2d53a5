"""INTEGRATION.md section B, executed (pytest -m gpu): libadb200.so bound INSIDE the unmodified reference's own classes
(baseline/ref_binding.py patches only the numpy line of `_eval` in Einsum, Add, Mul, Negate, Recipr, ReLU, Exp, Log, Sigmoid,
ReduceSumKeepDims).  The reference's own graph code - constructors, `grad`, memoisation - then drives the C ABI, and every
value must match the stock numpy evaluation of the same graph within 1e-12."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _graphs(ref, seed):
    """Graphs built with the REFERENCE's classes: forward values, first- and second-order gradients."""
    rng = np.random.default_rng(seed)
    out = {}
    x = ref.Variable(rng.standard_normal((37, 20)), name="x")
    w = ref.Variable(rng.standard_normal((20, 11)) * 0.3, name="w")
    b = ref.Variable(rng.standard_normal((11,)), name="b")
    h = ref.Sigmoid(x @ w + b) * ref.ReLU(x @ w)                       # Einsum 'ij,jk->ik', Add (broadcast), Mul, Sigmoid, ReLU
    loss = ref.Einsum("ab->", h / (ref.Exp(h) + 2.0))                  # Recipr (every `/`), Exp, full reduction
    g = ref.grad(loss, [w, b])                                         # gradient graph: swapped einsums, ones operands, 'ab->b'
    out["mlp"] = [h, loss] + g + ref.grad(g[0], [w])                   # ... and a second-order gradient
    a = ref.Variable(rng.random((6, 5, 4)) + 0.5, name="a")
    c = ref.Variable(rng.random((5, 4)) + 0.5, name="c")
    e = ref.Einsum("bij,ij->bi", ref.Log(a), c)                        # Log with the reference's epsilon, 3-D einsum
    r = ref.ReduceSumKeepDims(a * c, axes=[0, 2])                      # np.sum(keepdims) replaced by adb_reduce_sum
    out["misc"] = [e, r, -e + 3.0] + ref.grad(ref.Einsum("bi->", e), [a, c])
    t = ref.Tanh(ref.Variable(rng.standard_normal((9, 7)), name="t"))  # macro: Exp, Negate, Add, Mul, Recipr nodes
    out["tanh"] = [t, ref.Einsum("...i,...i->...", t, t)]              # ellipsis subscripts expanded by the binding
    return out


def test_reference_classes_bound_to_the_c_abi():
    from baseline import ref_binding, ref_shim
    ref = ref_shim.load()
    if ref is None:
        pytest.skip("baseline/_ref not installed (baseline/install_ref.sh)")
    stock = {k: [np.asarray(n()) for n in nodes] for k, nodes in _graphs(ref, 3).items()}
    bound = ref_binding.bind(ref)
    assert {"Einsum", "Add", "Recipr", "ReduceSumKeepDims"} <= set(bound)
    try:
        got = {k: [np.asarray(n()) for n in nodes] for k, nodes in _graphs(ref, 3).items()}
    finally:
        ref_binding.unbind(ref)
    for k in stock:
        for i, (a, b) in enumerate(zip(got[k], stock[k])):
            assert a.shape == b.shape, (k, i, a.shape, b.shape)
            scale = float(np.max(np.abs(b))) or 1.0
            assert float(np.max(np.abs(a - b))) <= 1e-12 * scale, (k, i)
    # the patch is gone: the reference evaluates with numpy again
    again = [np.asarray(n()) for n in _graphs(ref, 3)["mlp"]]
    assert all(np.array_equal(a, b) for a, b in zip(again, stock["mlp"]))
