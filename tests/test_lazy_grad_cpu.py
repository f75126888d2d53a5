"""Host-side logic of the deferred gradients (autodiff_b200/core/grad.py) that needs no GPU: when `ad.grad` defers, what a
deferred node looks like, that walking into it yields exactly the plain gradient graph, and how evaluations are keyed."""
import numpy as np
import pytest

import autodiff_b200 as ad
from autodiff_b200.core import engine, grad as G, tape


@pytest.fixture(autouse=True)
def _fresh():
    tape.clear()
    saved = G.LAZY
    G.LAZY = True
    yield
    G.LAZY = saved
    tape.clear()


def _sig(tops):
    r = tape.scan(list(tops), engine.scalar_of, engine._has_value, engine._is_device_node)
    return None if r is None else r[2]


def _net(seed, width=5, scale=2.0):
    rng = np.random.default_rng(seed)
    x = ad.Variable(rng.standard_normal((4, width)), name="x")
    w = ad.Variable(rng.standard_normal((width, 3)), name="w")
    b = ad.Variable(rng.standard_normal((3,)), name="b")
    loss = ad.Einsum("ab->", ad.Tanh(x @ w + b)) * scale
    return loss, [w, b]


def test_first_sighting_is_plain_then_deferred_with_same_shapes_and_names():
    loss, wrt = _net(0)
    first = ad.grad(loss, wrt)
    assert all(type(g) is not G.DeferredGrad for g in first)
    loss2, wrt2 = _net(1)
    second = ad.grad(loss2, wrt2)
    assert all(type(g) is G.DeferredGrad and g._real is None for g in second)
    assert [g.shape for g in second] == [g.shape for g in first] and [g.name for g in second] == [g.name for g in first]
    # a different structure (other width / folded scalar) is a first sighting again
    loss3, wrt3 = _net(0, width=6)
    assert type(ad.grad(loss3, wrt3)[0]) is not G.DeferredGrad
    loss4, wrt4 = _net(0, scale=3.0)
    assert type(ad.grad(loss4, wrt4)[0]) is not G.DeferredGrad


def test_walking_into_a_deferred_gradient_yields_the_plain_graph():
    G.LAZY = False
    loss, wrt = _net(0)
    plain = _sig(ad.grad(loss, wrt))
    G.LAZY = True
    ad.grad(*_net(1))
    loss, wrt = _net(2)
    deferred = ad.grad(loss, wrt)
    assert all(g._real is None for g in deferred)
    reals = [g.children[0] for g in deferred]          # expansion on demand
    assert all(g._real is not None for g in deferred) and deferred[0]._req is deferred[1]._req
    assert _sig(reals) == plain                          # the very graph the plain path builds
    assert G.stats["expanded"] >= 1


def test_nested_requests_stay_deferred_and_key_on_the_inner_request():
    def second_order(seed):
        loss, wrt = _net(seed)
        g = ad.grad(loss, [wrt[0]])[0]
        return g, ad.grad(g, [wrt[0]])[0]
    for s in range(3):
        g, gg = second_order(s)
    assert type(g) is G.DeferredGrad and type(gg) is G.DeferredGrad
    assert g._real is None and gg._real is None and gg._req.top is g
    assert gg.shape == (5, 3)
    # evaluation key: equal for isomorphic requests, different when an output is added or the order changes
    g2, gg2 = second_order(7)
    k1 = G.lazy_context([gg], ("f",)).key
    assert k1 == G.lazy_context([gg2], ("f",)).key
    loss, wrt = _net(9)
    assert G.lazy_context([gg2, g2], ("f",)).key != G.lazy_context([g2, gg2], ("f",)).key != k1
    assert G.lazy_context([loss], ("f",)) is None        # nothing deferred among the outputs


def test_fallbacks_take_the_plain_path():
    for s in range(2):
        loss, wrt = _net(s)
        out = ad.grad(loss, wrt, previous_grad=ad.Variable(np.array(2.0)))
        assert type(out[0]) is not G.DeferredGrad       # an explicit incoming gradient is never deferred
    for s in range(2):
        rng = np.random.default_rng(s)
        x = ad.Variable(rng.standard_normal((4, 5)))
        h = ad.Exp(x)
        top = ad.Einsum("ab->", h * h)
        out = ad.grad(top, [h])                          # wrt is an interior node
        assert type(out[0]) is not G.DeferredGrad
    G.LAZY = False
    for s in range(3):
        assert type(ad.grad(*_net(s))[0]) is not G.DeferredGrad


def test_deferred_gradient_differentiates_like_identity_once_expanded():
    for s in range(2):
        loss, wrt = _net(s)
        g = ad.grad(loss, [wrt[0]])[0]
    assert type(g) is G.DeferredGrad
    outer = ad.Einsum("ab->", g * g)                     # a regular node above a deferred one: walking expands it
    gg = ad.grad(outer, [wrt[0]])[0]
    assert g._real is not None and gg.shape == (5, 3)
