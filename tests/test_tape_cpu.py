"""Host-side logic of the launch tapes (autodiff_b200/core/tape.py) that needs no GPU: the structural signature must be
equal for isomorphic graphs and differ whenever the kernels or their wiring would differ."""
import numpy as np

import pytest

import autodiff_b200 as ad
from autodiff_b200.core import engine, grad as grad_mod, tape


@pytest.fixture(autouse=True)
def _plain_grad():
    """These tests compare signatures of EXPANDED gradient graphs: keep `ad.grad` on the plain path (tests/test_lazy_grad_cpu.py
    covers the deferred one)."""
    saved = grad_mod.LAZY
    grad_mod.LAZY = False
    yield
    grad_mod.LAZY = saved


def _sig(tops):
    r = tape.scan(list(tops), engine.scalar_of, engine._has_value, engine._is_device_node)
    return None if r is None else r[2]


def _mlp(seed, scale=2.0, share=True, width=5):
    rng = np.random.default_rng(seed)
    x = ad.Variable(rng.standard_normal((4, width)), name="x")
    w = ad.Variable(rng.standard_normal((width, 3)), name="w")
    h = ad.ReLU(x @ w)
    h2 = h if share else ad.ReLU(x @ w)
    loss = ad.Einsum("ab->", h * h2) * scale
    return ad.grad(loss, [w]) + [loss]


def test_isomorphic_graphs_share_a_signature():
    a, b = _sig(_mlp(0)), _sig(_mlp(1))
    assert a is not None and a == b and hash(a) == hash(b)


def test_signature_sees_scalars_shapes_sharing_and_attributes():
    base = _sig(_mlp(0))
    assert _sig(_mlp(0, scale=3.0)) != base          # a folded scalar is part of the program
    assert _sig(_mlp(0, width=6)) != base            # leaf shapes
    assert _sig(_mlp(0, share=False)) != base        # the same tree with a different sharing pattern is a different DAG
    x = ad.Variable(np.ones((3, 4)))
    assert _sig([ad.Einsum("ab->a", x)]) != _sig([ad.Einsum("ab->b", x)])
    assert _sig([x[0:2]]) != _sig([x[1:3]])
    assert _sig([ad.ReduceSumKeepDims(x, axes=[0])]) != _sig([ad.ReduceSumKeepDims(x, axes=[1])])
    assert _sig([ad.Concat(x, x, axis=0)]) != _sig([ad.Concat(x, x, axis=1)])
    assert _sig([ad.Reshape(x, (4, 3))]) != _sig([ad.Reshape(x, (2, 6))])
    assert _sig([ad.Pad(x, ((1, 1), (0, 0)), 0.0)]) != _sig([ad.Pad(x, ((1, 1), (0, 0)), 1.0)])


def test_order_and_multiplicity_of_outputs_matter():
    x = ad.Variable(np.ones((3, 4)))
    a, b = ad.Exp(x), ad.Log(x)
    assert _sig([a, b]) != _sig([b, a])
    assert _sig([a, a]) != _sig([a])


def test_graphs_that_cannot_be_taped_are_rejected():
    class Mine(ad.Node):                             # a user-defined numpy op (README: "easy to add your own operations")
        def __init__(self, n):
            super().__init__([n], "Mine")
            self.shape = self.children[0].shape

        def _eval(self):
            return np.sin(self.children[0]())

    class MyRelu(ad.ReLU):                           # a subclass of a known op may compute anything
        pass

    x = ad.Variable(np.ones((3, 4)))
    assert _sig([Mine(x) + 1.0]) is None
    assert _sig([MyRelu(x)]) is None
    assert _sig([ad.ReLU(x)]) is not None
    assert _sig([ad.checkpoint(lambda v: ad.Exp(v))(x)]) is None


def test_lookup_records_on_second_sighting_and_replays_after_store():
    tape.clear()
    sig = ("unit-test", 1)
    assert tape.lookup(sig) == (None, False)
    assert tape.lookup(sig) == (None, True)
    tape.store(sig, None)                            # recording failed: never try again
    assert tape.lookup(sig) == (None, False)
    sig2 = ("unit-test", 2)
    tape.lookup(sig2); tape.lookup(sig2)
    marker = object()
    tape.store(sig2, marker)
    assert tape.lookup(sig2) == (marker, False)
    tape.clear()


def test_graph_replay_gating_without_a_gpu():
    """When a tape may be replayed as one CUDA graph (tape._graph_for): only small-leaf tapes with enough calls and after
    the first call-by-call replay; never while a value of the previous graph replay is still alive in the graph's buffer."""
    import sys
    import torch
    from autodiff_b200.device import DeviceArray

    class FakeBase:
        def __init__(self, n):
            self.n = n

        def numel(self):
            return self.n

    t = tape.Tape()
    t.ext_size, t.calls, t.replays, t.graph = [80], [None] * 10, 1, None
    bases = [FakeBase(10)]
    if not tape.GRAPH:
        return
    assert tape._graph_for(t, bases) is None and t.graph is None          # first replay: call by call
    t.replays, t.calls = 5, [None] * (tape.GRAPH_MIN_CALLS - 1)
    assert tape._graph_for(t, bases) is None and t.graph is None          # too few kernels to be worth a graph
    t.calls, t.ext_size = [None] * 10, [tape.GRAPH_MAX_STAGE_BYTES + 8]
    assert tape._graph_for(t, [FakeBase(tape.GRAPH_MAX_STAGE_BYTES // 8 + 1)]) is None and t.graph is None   # leaves too big to stage
    t.ext_size = [80]
    g = tape._Graph()
    g.exec, g.buf = None, torch.zeros(8, dtype=torch.float64)
    tape.stats["graph_bytes"] += 64                                     # (what _capture books for the buffer; released by __del__)
    g.refs = sys.getrefcount(g.buf)
    t.graph = g
    assert tape._graph_for(t, bases) is g
    assert tape._graph_for(t, [FakeBase(11)]) is None                     # a leaf is now a view of a differently sized array
    busy = tape.stats["graph_busy"]
    alive = DeviceArray._make(g.buf, 0, (2,), (1,))                        # what a node of the previous replay holds
    assert tape._graph_for(t, bases) is None and tape.stats["graph_busy"] == busy + 1
    del alive
    assert tape._graph_for(t, bases) is g
    t.graph = False                                                       # capture failed once: never tried again
    assert tape._graph_for(t, bases) is None


def test_scan_never_materialises_a_constfill():
    """The signature walk must not touch ConstFill._value: that property builds the host array (512 MiB for the ones seed of
    an 8192^2 gradient - 70 ms per evaluation when a first version of the inlined scan did)."""
    x = ad.Variable(np.ones((6, 5)), name="x")
    ones = ad.ConstFill((6, 5), 1.0)
    e = ad.Einsum("ij,ij->ij", ones, x)
    g = ad.grad(ad.Einsum("ij,jk->ik", x, ad.Variable(np.ones((5, 4)))), [x])[0]
    fills = [ones]
    stack = [g]
    seen = set()
    while stack:
        n = stack.pop()
        if id(n) in seen:
            continue
        seen.add(id(n))
        if isinstance(n, ad.ConstFill):
            fills.append(n)
        stack.extend(n.children)
    assert len(fills) >= 2                           # the gradient seed is a ConstFill too
    assert _sig([e]) is not None and _sig([g]) is not None
    assert all(f._materialised is None for f in fills)


def test_plan_streams_orders_every_dependent_pair():
    """core/tape.py: plan_streams puts independent calls of a tape on different capture streams (parallel branches of the CUDA
    graph).  For random read / write sets: every RAW / WAW / WAR pair must be ordered by stream order + event waits (checked through
    the transitive happens-before closure), unknown calls disable the plan, a pure chain stays on one stream."""
    import random

    class Call:
        def __init__(self, reads, writes):
            self.reads, self.writes = reads, writes

    rnd = random.Random(5)
    for trial in range(60):
        n, nbuf, k = rnd.randint(8, 70), rnd.randint(3, 30), rnd.randint(2, 5)
        calls = []
        for i in range(n):
            reads = set(rnd.sample(range(nbuf), rnd.randint(0, min(3, nbuf))))
            writes = {rnd.randrange(nbuf)} if rnd.random() < 0.85 else {rnd.randrange(nbuf), rnd.randrange(nbuf)}
            if rnd.random() < 0.1:
                reads |= writes                                  # in place
            calls.append(Call(reads, writes))
        plan = tape.plan_streams(calls, k)
        assert plan is not None
        assign, waits, need = plan
        assert len(assign) == n and all(0 <= a < k for a in assign)
        assert need == {d for w in waits for d in w}
        before = [set() for _ in range(n)]                       # happens-before closure
        tail = {}
        for i in range(n):
            preds = list(waits[i])
            if assign[i] in tail:
                preds.append(tail[assign[i]])
            for d in preds:
                assert d < i and (d == tail.get(assign[d]) or assign[d] != assign[i] or True)
                before[i].add(d)
                before[i] |= before[d]
            tail[assign[i]] = i
        for i in range(n):
            for j in range(i):
                a, b = calls[j], calls[i]
                conflict = (a.writes & (b.reads | b.writes)) or (a.reads & b.writes)
                if conflict:
                    assert j in before[i], (trial, j, i)
    chain = [Call({i - 1} if i else set(), {i}) for i in range(20)]
    assign, waits, need = tape.plan_streams(chain, 4)
    assert len(set(assign)) == 1 and not need
    fan = [Call(set(), {0})] + [Call({0}, {i}) for i in range(1, 9)]
    assign, waits, need = tape.plan_streams(fan, 4)
    assert len(set(assign)) == 4                                 # eight independent consumers spread over all streams
    assert tape.plan_streams([Call(None, None)] * 10, 4) is None and tape.plan_streams(chain, 1) is None


def test_nd_softmax_ce_is_never_taped():
    """N-D SoftmaxCEWithLogits decides its label check on the host (core/ops.py): the scan must refuse such graphs."""
    z = ad.Variable(np.zeros((3, 4, 5)), name="z")
    lab = ad.Variable(np.full((3, 4, 5), 0.25), name="l")
    n = ad.SoftmaxCEWithLogits(labels=lab, logits=z)
    assert tape.scan([n], engine.scalar_of, engine._has_value, engine._is_device_node) is None
    n2 = ad.SoftmaxCEWithLogits(labels=ad.Variable(np.full((3, 5), 0.2), name="l"), logits=ad.Variable(np.zeros((3, 5)), name="z"))
    assert tape.scan([n2], engine.scalar_of, engine._has_value, engine._is_device_node) is not None
