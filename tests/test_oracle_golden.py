"""Pins the numpy oracle (oracle/np_autodiff.py) against vectors produced by the REAL reference
(tests/golden/make_golden.py).  Same numpy calls in the same order => bit-for-bit equality."""
import os

import numpy as np
import pytest

import cases
import oracle.np_autodiff as onp

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(a, b, equal_nan=True), float(np.max(np.abs(a - b)))


@pytest.fixture(scope="module")
def gold_cases():
    return np.load(os.path.join(GOLD, "cases.npz"))


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_case_bit_exact(name, gold_cases):
    nodes = cases.derive(onp, name)
    for key, node in nodes.items():
        _same(node(), gold_cases[name + "/" + key])


def test_readme_numbers():
    # README.md:49-55 literal values
    g, wrt = cases.CASES["readme_scalar"](onp)
    assert float(g()) == 415.4287934927351
    assert float(onp.grad(g, [wrt[0]])[0]()) == 407.4287934927351


def test_tanh_derivatives():
    gold = np.load(os.path.join(GOLD, "tanh_derivs.npz"))
    for i, n in enumerate(cases.tanh_derivatives(onp)):
        _same(n(), gold["d%d" % i])


def test_mlp_steps():
    gold = np.load(os.path.join(GOLD, "mlp_steps.npz"))
    hist, weights = cases.mlp_train_steps(onp)
    for s, h in enumerate(hist):
        _same(h["cost"], gold["step%d/cost" % s])
        _same(h["gnorm"], gold["step%d/gnorm" % s])
        for i, g in enumerate(h["grads"]):
            _same(g, gold["step%d/grad%d" % (s, i)])
    for i, w in enumerate(weights):
        _same(w, gold["final/w%d" % i])


def test_char_rnn_second_order():
    gold = np.load(os.path.join(GOLD, "char_rnn.npz"))
    total, params = cases.char_rnn_graph(onp)
    g1 = onp.grad(total, params)
    _same(total(), gold["loss"])
    for i, g in enumerate(g1):
        _same(g(), gold["g%d" % i])
    _same(onp.grad(g1[0], [params[0]])[0](), gold["gg_w"])


def test_adam_step0_is_noop():
    # step starts at 0 => first update multiplies by sqrt(1-b2**0) = 0
    opt = onp.Adam(1, lr=0.1)
    p = [np.array([1.0])]
    p = opt(p, [np.array([1.0])])
    assert p[0][0] == 1.0
    p = opt(p, [np.array([1.0])])
    assert abs(p[0][0] - (1.0 - 0.13438)) < 1e-4


def test_conv_oracle_matches_direct_loop():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 6, 5, 3)); w = rng.standard_normal((3, 2, 3, 4))
    y = onp.conv_valid_nhwc(x, w)
    ref = np.zeros((2, 4, 4, 4))
    for a in range(3):
        for b in range(2):
            ref += np.einsum("nhwc,co->nhwo", x[:, a:a + 4, b:b + 4, :], w[a, b])
    np.testing.assert_allclose(y, ref, rtol=1e-13, atol=1e-13)
