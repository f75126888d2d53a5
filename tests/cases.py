"""Graph catalogue shared by the golden generator, the oracle tests and the GPU parity tests.

Every case is ``builder(ad) -> (graph, wrt_vars)`` where ``ad`` is ANY module exposing the
reference API (the real reference, ``oracle.np_autodiff`` or ``autodiff_b200``).  Seeds, shapes and
graphs are the reference's own test fixtures:
  tests/test_einsum.py:12-28,41-111      tests/test_matrix_ops.py:16-42,54-236
  tests/test_complete_graphs.py:16-38,49-77,96-117      tests/test_scalar_ops.py:12-51
  README.md:44-55      examples/higher_order_deriv_visualization.py:5-20
plus scaled-down versions of the BASELINE.json configs (MLP step, char-rnn, einsum sweep).
"""
import numpy as np


# ----------------------------------------------------------------------------- fixtures
def _einsum_fixture(ad):
    np.random.seed(1337)                                    # tests/test_einsum.py:12-17
    vals = [np.random.randn(2, 3), np.random.randn(3, 5), np.random.randn(5, 5),
            np.random.randn(2, 3, 5, 7), np.random.randn(5)]
    return [ad.Variable(v, name="w%d_val" % i) for i, v in enumerate(vals)]


def _matrix_fixture(ad):
    np.random.seed(1337)                                    # tests/test_matrix_ops.py:16-27
    w0 = np.random.rand(2, 3); w1 = np.random.rand(3, 5); w2 = np.random.rand(10, 3)
    w3 = np.random.rand(10, 5); w4 = np.random.rand(10, 5); w5 = np.linspace(-7, 7, 200)
    w6 = np.random.rand(2, 3, 5, 7)
    w3 /= np.expand_dims(np.sum(w3, axis=1), axis=1)
    return [ad.Variable(v, name="w%d_val" % i) for i, v in enumerate([w0, w1, w2, w3, w4, w5, w6])]


def _einsum_case(op_str, idx):
    def build(ad):
        w = _einsum_fixture(ad)
        args = [w[i] for i in idx]
        return ad.Einsum(op_str, *args), args
    return build


def _matrix_case(fn, wrt_idx):
    def build(ad):
        w = _matrix_fixture(ad)
        return fn(ad, w), [w[i] for i in wrt_idx]
    return build


def _mlp_sigmoid(ad):                                       # tests/test_complete_graphs.py:16-38
    np.random.seed(1337)
    x = ad.Variable(np.random.randn(16, 20), name="x_val")
    w1 = ad.Variable(np.random.randn(20, 40), name="w1_val")
    w2 = ad.Variable(np.random.randn(40, 5), name="w2_val")
    h = ad.Sigmoid(x @ w1)
    return ad.Sigmoid(h @ w2), [x, w1, w2]


def _bcast_fixture(ad):                                     # tests/test_complete_graphs.py:49-58
    np.random.seed(1337)
    h = ad.Variable(np.random.randn(2, 5), name="h")
    b0 = ad.Variable(np.random.randn(5), name="b0")
    b1 = ad.Variable(np.random.randn(1, 5), name="b1")
    b2 = ad.Variable(7, name="b2")
    return h, b0, b1, b2


def _bcast_sum(ad):
    h, b0, b1, b2 = _bcast_fixture(ad)
    return h + b0 + b1 + b2, [b0, b1, b2]


def _bcast_mul(ad):
    h, b0, b1, b2 = _bcast_fixture(ad)
    return h * b0 * b1 + b2, [b0, b1, b2]


def _matmul_bias_sigmoid(ad):                               # tests/test_complete_graphs.py:96-117
    np.random.seed(1337)
    x = ad.Variable(np.random.randn(2, 3), name="x_val")
    w = ad.Variable(np.random.randn(3, 5), name="w_val")
    b = ad.Variable(np.random.randn(5), name="b_val")
    y = x @ w + b
    return ad.Sigmoid(y), [x, w, b, y]


def _scalar_fixture(ad):                                    # tests/test_scalar_ops.py:12-20
    np.random.seed(1336)
    return ad.Variable(np.random.rand() * 2 + 1, name="x"), ad.Variable(np.random.rand() * 2 + 1, name="y")


def _scalar_case(fn):
    def build(ad):
        x, y = _scalar_fixture(ad)
        return fn(ad, x, y), [x, y]
    return build


def _readme(ad):                                            # README.md:44-55
    x = ad.Variable(3, name="x"); y = ad.Variable(4, name="y")
    return x * y + ad.Exp(x + 3), [x, y]


def _sweep_case(op_str, shapes, seed=1337):
    def build(ad):
        rng = np.random.default_rng(seed)
        args = [ad.Variable(rng.standard_normal(s), name="op%d" % i) for i, s in enumerate(shapes)]
        return ad.Einsum(op_str, *args), args
    return build


CASES = {
    "readme_scalar": _readme,
    # --- tests/test_einsum.py
    "reducesumkeepdims": lambda ad: (lambda w: (ad.ReduceSumKeepDims(w[3], axes=[0, 2]), [w[3]]))(_einsum_fixture(ad)),
    "einsum_identity": _einsum_case("ijkl->ijkl", [3]),
    "einsum_sum1": _einsum_case("ijkl->ijk", [3]),
    "einsum_sum2": _einsum_case("ijkl->ij", [3]),
    "einsum_sum3": _einsum_case("ijkl->i", [3]),
    "einsum_sum5": _einsum_case("ijkl->", [3]),
    "einsum_twovars": _einsum_case("dt,tp->dp", [0, 1]),
    "einsum_threevars": _einsum_case("dt,tp,pr->dtp", [0, 1, 2]),
    "einsum_summation1": _einsum_case("dt,tp->d", [0, 1]),
    "einsum_summation2": _einsum_case("dt,tp->p", [0, 1]),
    "einsum_summation3": _einsum_case("dt,tp->", [0, 1]),
    # --- tests/test_matrix_ops.py
    "normal_distribution": _matrix_case(lambda ad, w: ad.NormalDistribution(w[5], 0, 1), [5]),
    "sigmoid_ce": _matrix_case(lambda ad, w: ad.SigmoidCEWithLogits(labels=w[3], logits=w[4]), [4]),
    "softmax_ce": _matrix_case(lambda ad, w: ad.SoftmaxCEWithLogits(labels=w[3], logits=w[4]), [4]),
    "softmax_2d": _matrix_case(lambda ad, w: ad.Softmax(w[4]), [4]),
    "softmax_1d": _matrix_case(lambda ad, w: ad.Softmax(w[5]), [5]),
    "softmax_4d": _matrix_case(lambda ad, w: ad.Softmax(w[6]), [6]),
    "reshape": _matrix_case(lambda ad, w: ad.Reshape(w[1], (1, 15)), [1]),
    "pad": _matrix_case(lambda ad, w: ad.Pad(w[0], [[1, 0], [2, 2]], constant_values=[0, 0]), [0]),
    "slice1": _matrix_case(lambda ad, w: w[1][:, :-1], [1]),
    "slice2": _matrix_case(lambda ad, w: w[1][0, 0], [1]),
    "slice3": _matrix_case(lambda ad, w: w[1][2, 2:4], [1]),
    "slice4": _matrix_case(lambda ad, w: w[1][1], [1]),
    "slice5": _matrix_case(lambda ad, w: w[1][::-2], [1]),
    "sigmoid": _matrix_case(lambda ad, w: ad.Sigmoid(w[0]), [0]),
    "relu": _matrix_case(lambda ad, w: ad.ReLU(w[0]), [0]),
    "transpose": _matrix_case(lambda ad, w: ad.Transpose(w[0]), [0]),
    "recipr": _matrix_case(lambda ad, w: ad.Recipr(w[0]), [0]),
    "negate": _matrix_case(lambda ad, w: ad.Negate(w[0]), [0]),
    "log": _matrix_case(lambda ad, w: ad.Log(w[0]), [0]),
    "exp": _matrix_case(lambda ad, w: ad.Exp(w[0]), [0]),
    "identity": _matrix_case(lambda ad, w: ad.Identity(w[0]), [0]),
    "tanh": _matrix_case(lambda ad, w: ad.Tanh(w[0]), [0]),
    "matmul": _matrix_case(lambda ad, w: ad.MatMul(w[0], w[1]), [0, 1]),
    "concat1": _matrix_case(lambda ad, w: ad.Concat(w[0], w[2], axis=0), [0, 2]),
    "concat2": _matrix_case(lambda ad, w: ad.Concat(w[2], w[3], axis=1), [2, 3]),
    "frobenius": _matrix_case(lambda ad, w: ad.FrobeniusNorm(w[3]), [3]),
    # --- tests/test_complete_graphs.py
    "mlp_sigmoid": _mlp_sigmoid,
    "bcast_sum": _bcast_sum,
    "bcast_mul": _bcast_mul,
    "matmul_bias_sigmoid": _matmul_bias_sigmoid,
    # --- tests/test_scalar_ops.py
    "scalar_add": _scalar_case(lambda ad, x, y: x + y),
    "scalar_mul": _scalar_case(lambda ad, x, y: x * y),
    "scalar_pow": _scalar_case(lambda ad, x, y: x ** y),
    "scalar_sqdiff": _scalar_case(lambda ad, x, y: ad.SquaredDifference(x, y)),
    # --- BASELINE config 2 (einsum sweep), scaled down; odd sizes exercise ragged tiles
    "sweep_ij_jk": _sweep_case("ij,jk->ik", [(37, 53), (53, 29)]),
    "sweep_ij_jk_tile": _sweep_case("ij,jk->ik", [(256, 192), (192, 160)]),
    "sweep_bij_bjk": _sweep_case("bij,bjk->bik", [(3, 33, 47), (3, 47, 21)]),
    "sweep_ijkl_jl": _sweep_case("ijkl,jl->ik", [(6, 5, 7, 9), (5, 9)]),
}

# second-order through these evaluates eagerly in the reference (reshape.py:96-102) but is well defined
ORDER = {name: 2 for name in CASES}
ORDER.update({"mlp_sigmoid": 2, "sweep_ij_jk_tile": 1})


def derive(ad, name):
    """Outputs checked for a case: f, d1 (grad wrt every var), d2 (grad of that grad wrt the same var,
    as tests/numerical_check.py:36-47 builds it), and d1 of Exp(graph) (tests/numerical_check.py:74-76)."""
    graph, wrt = CASES[name](ad)
    outs = {"f": graph}
    d1 = ad.grad(graph, wrt)
    for i, g in enumerate(d1):
        outs["d1_%d" % i] = g
    if ORDER[name] >= 2:
        for i, (g, v) in enumerate(zip(d1, wrt)):
            outs["d2_%d" % i] = ad.grad(g, [v])[0]
    e = ad.Exp(graph, name="extra_exp_op")
    for i, g in enumerate(ad.grad(e, wrt)):
        outs["e_d1_%d" % i] = g
    return outs


# ----------------------------------------------------------------------------- bigger workloads
def tanh_derivatives(ad, n=200, order=4):
    """examples/higher_order_deriv_visualization.py:5-20."""
    x = ad.Variable(np.linspace(-7, 7, n), name="x")
    y = ad.Tanh(x)
    outs = [y]
    for _ in range(order):
        y = ad.grad(y, [x])[0]
        outs.append(y)
    return outs


def mlp_train_steps(ad, sizes=(20, 16, 5), batch=8, steps=3, lr=1e-3, seed=1337):
    """BASELINE config 1 shape: NN + SoftmaxCE + Einsum('i->')/batch + Adam
    (examples/feedforward_mnist.py:17-34), synthetic data."""
    np.random.seed(seed)
    nn = ad.NN(list(sizes))
    opt = ad.Adam(len(nn.w), lr=lr)
    rng = np.random.default_rng(seed)
    history = []
    for _ in range(steps):
        xv = rng.random((batch, sizes[0]))
        yv = np.eye(sizes[-1])[rng.integers(0, sizes[-1], batch)]
        x = ad.Variable(xv, name="images"); y = ad.Variable(yv, name="labels")
        logits = nn(x)
        cost = ad.Einsum("i->", ad.SoftmaxCEWithLogits(labels=y, logits=logits)) / batch
        grads = ad.grad(cost, nn.w)
        gvals = [g() for g in grads]
        history.append({"cost": np.asarray(cost()), "grads": [np.array(g) for g in gvals],
                        "gnorm": np.asarray(ad.FrobeniusNorm(*grads)())})
        new_w = opt([w() for w in nn.w], gvals)
        opt.apply_new_weights(nn.w, new_w)
    return history, [np.array(w()) for w in nn.w]


def char_rnn_graph(ad, hidden=16, vocab=7, batch=4, steps=5, seed=1337, chunk=32):
    """BASELINE config 5 shape (examples/char_rnn.py:15-22,50-62): returns loss, g=dL/dw, gg=d(g)/dw."""
    rng = np.random.default_rng(seed)
    w = ad.Variable(rng.standard_normal((hidden, hidden)) * 0.01, name="w")
    u = ad.Variable(rng.standard_normal((vocab, hidden)) * 0.01, name="u")
    b_h = ad.Variable(np.zeros(hidden), name="bias_hidden")
    v = ad.Variable(rng.standard_normal((hidden, vocab)) * 0.01, name="v")
    b_o = ad.Variable(np.zeros(vocab), name="bias_hidden")
    onehot = np.eye(vocab)[rng.integers(0, vocab, (batch, steps))]
    h = ad.Variable(np.zeros((1, hidden)), name="h")
    costs = []
    for t in range(steps - 1):
        x = ad.Variable(onehot[:, t, :], name="x")
        h = ad.Tanh(h @ w + x @ u + b_h)
        logits = h @ v + b_o
        y = ad.Variable(onehot[:, t + 1, :], name="y")
        costs.append(ad.Einsum("i->", ad.SoftmaxCEWithLogits(labels=y, logits=logits)))
    while len(costs) > chunk:                       # numpy's 64-array limit (reference ops.py:26)
        costs = [ad.Add(*costs[i:i + chunk]) for i in range(0, len(costs), chunk)]
    total = ad.Add(*costs) / steps
    params = [w, u, b_h, v, b_o]
    return total, params
