"""GPU parity tests (pytest -m gpu): the CUDA path vs (a) golden vectors produced by the real reference and
(b) the numpy oracle on fresh seeded inputs.  Tolerance: norm-wise relative error <= 1e-12 in fp64
(BASELINE.json north_star); integer/index-like results (shapes, flags) exact."""
import ctypes as C
import os

import numpy as np
import pytest

import cases
import oracle.np_autodiff as onp

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-12


@pytest.fixture(scope="module")
def ad():
    import autodiff_b200 as ad
    ad.init()
    return ad


def rel_err(x, ref):
    x, ref = np.asarray(x, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    x = np.broadcast_to(x, ref.shape) if x.shape != ref.shape and x.size == 1 else x
    assert x.shape == ref.shape, (x.shape, ref.shape)
    if ref.size == 0:
        return 0.0
    scale = np.max(np.abs(ref))
    return float(np.max(np.abs(x - ref)) / (scale if scale > 0 else 1.0))


def check(x, ref, tol=TOL, what=""):
    e = rel_err(x, ref)
    assert e <= tol, "%s rel err %.3e > %.1e" % (what, e, tol)


# Outputs that are ZERO BY MATHEMATICS (the gradient of sum(softmax(x)) == const, and derivatives of it): the
# reference's value there is pure rounding noise (~1e-14), so "relative to max|ref|" has no meaning.  They are
# held to the same 1e-12 but relative to the magnitude of the case's forward value instead.
ZERO_BY_MATH = {"softmax_2d/d1_0", "softmax_2d/d2_0", "softmax_1d/d1_0", "softmax_1d/d2_0", "softmax_4d/d1_0", "softmax_4d/d2_0",
                "softmax_ce/d2_0"}


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_case_against_reference_golden(name, ad):
    gold = np.load(os.path.join(GOLD, "cases.npz"))
    nodes = cases.derive(ad, name)
    for key, node in nodes.items():
        ref = gold[name + "/" + key]
        if name + "/" + key in ZERO_BY_MATH:
            fscale = float(np.max(np.abs(gold[name + "/f"])))
            assert np.max(np.abs(ref)) < 1e-12 * fscale                      # the reference itself is at noise level
            got = np.broadcast_to(np.asarray(node()), ref.shape)
            assert np.max(np.abs(got - ref)) <= TOL * fscale, name + "/" + key
        else:
            check(node(), ref, what=name + "/" + key)


def test_readme_numbers(ad):
    g, wrt = cases.CASES["readme_scalar"](ad)
    assert abs(float(g()) - 415.4287934927351) <= 1e-12 * 415.4287934927351
    assert abs(float(ad.grad(g, [wrt[0]])[0]()) - 407.4287934927351) <= 1e-12 * 407.4287934927351


def test_tanh_derivatives(ad):
    gold = np.load(os.path.join(GOLD, "tanh_derivs.npz"))
    for i, n in enumerate(cases.tanh_derivatives(ad)):
        check(n(), gold["d%d" % i], what="tanh d%d" % i)


def test_mlp_training_steps(ad):
    gold = np.load(os.path.join(GOLD, "mlp_steps.npz"))
    hist, weights = cases.mlp_train_steps(ad)
    for s, h in enumerate(hist):
        check(h["cost"], gold["step%d/cost" % s], what="cost")
        check(h["gnorm"], gold["step%d/gnorm" % s], what="gnorm")
        for i, g in enumerate(h["grads"]):
            check(g, gold["step%d/grad%d" % (s, i)], what="grad")
    for i, w in enumerate(weights):
        check(w, gold["final/w%d" % i], what="weights")


def test_char_rnn_second_order(ad):
    gold = np.load(os.path.join(GOLD, "char_rnn.npz"))
    total, params = cases.char_rnn_graph(ad)
    g1 = ad.grad(total, params)
    check(total(), gold["loss"], what="loss")
    for i, g in enumerate(g1):
        check(g(), gold["g%d" % i], what="g%d" % i)
    check(ad.grad(g1[0], [params[0]])[0](), gold["gg_w"], what="gg")


# ----------------------------------------------------------------------------- vs oracle on fresh inputs
SWEEP = [
    ("ij,jk->ik", [(300, 260), (260, 140)]),
    ("ij,jk->ik", [(129, 131), (131, 127)]),                 # odd everything: pitched rows, ragged tiles
    ("bij,bjk->bik", [(5, 130, 70), (5, 70, 90)]),
    ("ijkl,jl->ik", [(12, 10, 14, 16), (10, 16)]),
    ("ijkl,jl->ik", [(20, 8, 24, 70), (8, 70)]),
]


@pytest.mark.parametrize("spec,shapes", SWEEP)
def test_einsum_forward_and_grads_vs_oracle(spec, shapes, ad):
    rng = np.random.default_rng(7)
    vals = [rng.standard_normal(s) for s in shapes]
    outs = {}
    for mod in (onp, ad):
        vs = [mod.Variable(v.copy(), name="v%d" % i) for i, v in enumerate(vals)]
        e = mod.Einsum(spec, *vs)
        g = mod.grad(e, vs)
        outs[mod] = [np.array(e())] + [np.array(x()) for x in g]
    for a, b in zip(outs[ad], outs[onp]):
        check(a, b, what=spec)


EINSUM_FORMS = [
    ("ab->b", [(70, 300)]), ("ab->a", [(70, 300)]), ("ab->", [(70, 300)]), ("i->", [(100003,)]), ("ab->ba", [(33, 65)]),
    ("ijkl->ijk", [(3, 4, 5, 70)]), ("ijkl->ij", [(3, 4, 5, 70)]), ("ijkl->l", [(3, 4, 5, 70)]), ("ijkl->jl", [(3, 4, 5, 70)]),
    ("bi,o->bo", [(64, 10), (1,)]), ("i,bo,o->bi", [(10,), (64, 1), (1,)]), ("ik,jl->ijkl", [(6, 7), (5, 9)]),
    ("ijkl,ik->jl", [(6, 5, 7, 9), (6, 7)]), ("bi,bij->bij", [(9, 5), (9, 5, 4)]), ("ij,ij->ij", [(9, 5), (9, 5)]),
    ("ij,ij->", [(9, 5), (9, 5)]), ("dt,tp,pr->dtp", [(2, 3), (3, 5), (5, 5)]), ("dt,tp->d", [(20, 30), (30, 50)]),
    ("ii->i", [(7, 7)]), ("ii->", [(7, 7)]), ("i,i->", [(50000,), (50000,)]), ("ij,j->i", [(300, 257), (257,)]),
    ("i,->i", [(11,), ()]), ("ij,jk->ki", [(40, 50), (50, 60)]), ("bij,bjk->ikb", [(3, 40, 50), (3, 50, 60)]),
]


@pytest.mark.parametrize("spec,shapes", EINSUM_FORMS)
def test_einsum_forms_vs_numpy(spec, shapes, ad):
    rng = np.random.default_rng(11)
    vals = [rng.standard_normal(s) for s in shapes]
    ref = np.einsum(spec, *vals)
    got = ad.Einsum(spec, *[ad.Variable(v, name="v") for v in vals])()
    check(got, ref, what=spec)


def test_einsum_on_noncontiguous_operands(ad):
    rng = np.random.default_rng(3)
    big = rng.standard_normal((40, 6, 50))
    a = big[:, 2, :]                         # strided view, like Variable(x[:, t, :]) in examples/char_rnn.py:55
    b = rng.standard_normal((60, 50)).T      # transposed view
    check(ad.Einsum("ij,jk->ik", ad.Variable(a, name="a"), ad.Variable(b, name="b"))(), a @ b)
    t = ad.Transpose(ad.Variable(a, name="a"))
    check((t @ ad.Variable(a, name="a2"))(), a.T @ a)
    check(ad.Variable(big, name="big")[::-2, 1:5:2, ::3](), big[::-2, 1:5:2, ::3])


def test_elementwise_ops_vs_oracle(ad):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((37, 53)); y = rng.random((37, 53)) + 0.5; b = rng.standard_normal(53); c = rng.standard_normal((37, 1))

    def graphs(m):
        X, Y, B, Cc = (m.Variable(v.copy(), name=n) for v, n in ((x, "x"), (y, "y"), (b, "b"), (c, "c")))
        return [X + Y + B + Cc + 3, X * Y * B * Cc * 0.5, X - Y, X / Y, 2 / Y, -X, m.Recipr(Y), m.Exp(X), m.Log(Y), m.ReLU(X),
                m.Sigmoid(X), m.Tanh(X), m.Pow(Y, X), Y ** 2.5, m.Pow(2.0, X), m.SquaredDifference(X, Y), m.Identity(X),
                m.NormalDistribution(X, 0.3, 1.7), m.SigmoidCEWithLogits(labels=Y, logits=X), m.Softmax(X), m.FrobeniusNorm(X, Y, B),
                m.ReduceSumKeepDims(X, axes=[0]), m.ReduceSumKeepDims(X, axes=[1]), m.ReduceSumKeepDims(X, axes=[0, 1]),
                m.Concat(X, Y, axis=0), m.Concat(X, Cc, axis=1), m.Reshape(X, (53, 37)), m.Reshape(X, (-1,)), X[3:30:2, ::-1], X[5],
                m.Pad(X, [[1, 2], [3, 0]], constant_values=0), m.Transpose(X) @ Y, m.Add(X, X, X), m.Mul(X, X)]
    for i, (g, r) in enumerate(zip(graphs(ad), graphs(onp))):
        check(g(), r(), what="graph %d" % i)


def test_large_elementwise_and_reduction_shapes(ad):
    rng = np.random.default_rng(9)
    x = rng.standard_normal((1031, 1027)); y = rng.standard_normal((1031, 1027)); b = rng.standard_normal(1027)
    X, Y, B = ad.Variable(x, name="x"), ad.Variable(y, name="y"), ad.Variable(b, name="b")
    check((X * Y + B)(), x * y + b)
    check(ad.Einsum("ab->b", X)(), x.sum(0))
    check(ad.Einsum("ab->a", X)(), x.sum(1))
    check(ad.Einsum("ab->", X)(), x.sum())
    check(ad.Tanh(X)(), onp.Tanh(onp.Variable(x, name="x"))())


def test_softmax_ce_matches_and_validates_labels(ad):
    rng = np.random.default_rng(2)
    z = rng.standard_normal((33, 17)) * 5
    lab = rng.random((33, 17)); lab /= lab.sum(1, keepdims=True)
    got = ad.SoftmaxCEWithLogits(labels=ad.Variable(lab, name="l"), logits=ad.Variable(z, name="z"))()
    ref = onp.SoftmaxCEWithLogits(labels=onp.Variable(lab, name="l"), logits=onp.Variable(z, name="z"))()
    check(got, ref)
    bad = lab.copy(); bad[4] *= 1.5
    with pytest.raises(ValueError, match="Labels must be a valid probability distribution!"):
        ad.SoftmaxCEWithLogits(labels=ad.Variable(bad, name="l"), logits=ad.Variable(z, name="z"))()


def test_cached_protocol_finite_difference_style(ad):
    """The reference's checker edits `var.cached` in place and resets caches (tests/numerical_check.py:5-27)."""
    rng = np.random.default_rng(4)
    wv = rng.standard_normal((4, 3))
    w = ad.Variable(wv.copy(), name="w")
    graph = ad.Sigmoid(w @ ad.Variable(rng.standard_normal((3, 2)), name="u"))
    base = graph().copy()

    def uncache(n):
        n.cached = None
        for c in n.children:
            uncache(c)
    old = w().copy()
    uncache(graph)
    w.cached = old
    w.cached[1, 2] += 0.5
    moved = graph().copy()
    assert np.max(np.abs(moved - base)) > 1e-3
    uncache(graph)
    check(graph(), base)            # Variable falls back to its own value again
    # numerical gradient agrees with backprop at the reference's own tolerance
    gb = ad.grad(graph, [w])[0]()
    eps, num = 1e-6, np.zeros_like(wv)
    for idx in np.ndindex(*wv.shape):
        for sgn in (1, -1):
            uncache(graph)
            w.cached = wv.copy()
            w.cached[idx] += sgn * eps
            num[idx] += sgn * np.sum(graph()) / (2 * eps)
    np.testing.assert_allclose(gb, num, rtol=1e-2, atol=1e-5)


def test_user_defined_numpy_node_and_checkpoint(ad):
    class Cube(ad.Node):                       # README: "easy to add your own operations"
        def __init__(self, node):
            super().__init__([node], "Cube")
            self.node = self.children[0]
            self.shape = self.node.shape

        def _eval(self):
            return self.node() ** 3

        def _partial_derivative(self, wrt, previous_grad):
            return previous_grad * 3 * self.node * self.node if wrt == self.node else 0
    xv = np.linspace(-2, 2, 11)
    x = ad.Variable(xv, name="x")
    y = ad.Exp(Cube(x * 2))
    check(y(), np.exp((xv * 2) ** 3))
    check(ad.grad(y, [x])[0](), np.exp((xv * 2) ** 3) * 3 * (2 * xv) ** 2 * 2, tol=1e-11)

    def block(a):
        return ad.Tanh(a) * a
    c = ad.checkpoint(block)(x)
    check(c(), block(onp.Variable(xv, name="x")) () if False else onp.Tanh(onp.Variable(xv, name="x"))() * xv)
    ox = onp.Variable(xv, name="x")
    check(ad.grad(c, [x])[0](), onp.grad(onp.Tanh(ox) * ox, [ox])[0]())


def test_device_resident_variables_and_optimizers(ad):
    rng = np.random.default_rng(8)
    p0 = rng.standard_normal((33, 9)); g0 = rng.standard_normal((33, 9))
    for name, mk in (("sgd", lambda m: m.SGDOptimizer(lr=0.1)), ("mom", lambda m: m.Momentum(1, lr=0.1)),
                     ("adagrad", lambda m: m.Adagrad(1, lr=0.1)), ("adam", lambda m: m.Adam(1, lr=0.1))):
        o_ref, o_dev = mk(onp), mk(ad)
        pr, pd = [p0.copy()], [ad.to_device(p0)]
        for step in range(3):
            g = g0 * (step + 1)
            pr = o_ref([pr[0]], [g])
            pd = o_dev([pd[0]], [ad.to_device(g)])
            assert isinstance(pd[0], ad.DeviceArray)
            check(ad.to_host(pd[0]), pr[0], what="%s step %d" % (name, step))
    v = ad.Variable(ad.to_device(p0), name="dev")
    check((v * 2)(), p0 * 2)
    check(v.value, p0)


def test_gemm_c_abi_all_layouts(ad):
    """Calls adb_gemm_f64 directly through ctypes with plain device pointers (the drop-in boundary)."""
    import torch
    from autodiff_b200 import _lib
    rng = np.random.default_rng(1)
    for (M, N, K, nb) in ((128, 128, 64, 1), (200, 136, 100, 1), (257, 129, 50, 3), (512, 384, 1030, 2)):
        for a_kc in (True, False):
            for b_kc in (True, False):
                A = rng.standard_normal((nb, M, K)); B = rng.standard_normal((nb, K, N))
                Ah = A if a_kc else np.ascontiguousarray(A.transpose(0, 2, 1))
                Bh = np.ascontiguousarray(B.transpose(0, 2, 1)) if b_kc else B
                At, Bt = torch.from_numpy(Ah).cuda(), torch.from_numpy(Bh).cuda()
                Ct = torch.full((nb, M, N), float("nan"), dtype=torch.float64, device="cuda")
                d = _lib.GemmDesc()
                d.M, d.N, d.K, d.batch = M, N, K, nb
                d.A, d.B, d.C = At.data_ptr(), Bt.data_ptr(), Ct.data_ptr()
                d.a_sm, d.a_sk, d.a_sb = (K, 1, M * K) if a_kc else (1, M, M * K)
                d.b_sk, d.b_sn, d.b_sb = (1, K, N * K) if b_kc else (N, 1, N * K)
                d.c_sm, d.c_sn, d.c_sb = N, 1, M * N
                aligned = (K % 2 == 0 if a_kc else M % 2 == 0) and (K % 2 == 0 if b_kc else N % 2 == 0)
                for path in ((1, 2, 3, 4) if aligned else (0, 2)):
                    Ct.fill_(float("nan"))
                    _lib.check(_lib.lib().adb_gemm_f64_path(C.byref(d), C.c_void_p(torch.cuda.current_stream().cuda_stream), path))
                    torch.cuda.synchronize()
                    check(Ct.cpu().numpy(), A @ B, what="gemm %s path %d" % ((M, N, K, nb, a_kc, b_kc), path))


def test_gemm_stream_k(ad):
    """The stream-K kernel (csrc/gemm_f64.cu: gemm_sk_kernel; forced paths 7 = 128x128 tiles, 8 = 64x64 tiles, and the automatic
    choice): whole-tile CTAs plus CTAs that share the tiles of the ragged last wave along K.  Values against numpy's matmul, and
    bitwise run-to-run repeatability - the parts of a split tile are re-added in k order whichever CTA arrives last."""
    import torch
    from autodiff_b200 import _lib
    rng = np.random.default_rng(7)
    cases = [(384, 256, 1024, 1), (1000, 520, 333, 2), (2176, 2176, 528, 1), (1281, 2047, 1000, 1), (640, 640, 4098, 1), (256, 256, 272, 37)]
    for (M, N, K, nb) in cases:
        for a_kc, b_kc in ((True, False), (False, True), (True, True), (False, False)):
            A = rng.standard_normal((nb, M, K)); B = rng.standard_normal((nb, K, N))
            Ah = A if a_kc else np.ascontiguousarray(A.transpose(0, 2, 1))
            Bh = np.ascontiguousarray(B.transpose(0, 2, 1)) if b_kc else B
            At, Bt = torch.from_numpy(Ah).cuda(), torch.from_numpy(Bh).cuda()
            Ct = torch.empty((nb, M, N), dtype=torch.float64, device="cuda")
            d = _lib.GemmDesc()
            d.M, d.N, d.K, d.batch = M, N, K, nb
            d.A, d.B, d.C = At.data_ptr(), Bt.data_ptr(), Ct.data_ptr()
            d.a_sm, d.a_sk, d.a_sb = (K, 1, M * K) if a_kc else (1, M, M * K)
            d.b_sk, d.b_sn, d.b_sb = (1, K, N * K) if b_kc else (N, 1, N * K)
            d.c_sm, d.c_sn, d.c_sb = N, 1, M * N
            aligned = (K % 2 == 0 if a_kc else M % 2 == 0) and (K % 2 == 0 if b_kc else N % 2 == 0)
            want = A @ B
            for path in ((7, 8, 0) if aligned else (0,)):
                runs = []
                for _ in range(3):
                    Ct.fill_(float("nan"))
                    _lib.check(_lib.lib().adb_gemm_f64_path(C.byref(d), C.c_void_p(torch.cuda.current_stream().cuda_stream), path))
                    torch.cuda.synchronize()
                    runs.append(Ct.cpu().numpy())
                check(runs[0], want, what="stream-K gemm %s path %d" % ((M, N, K, nb, a_kc, b_kc), path))
                assert all(np.array_equal(runs[0], r) for r in runs[1:]), "stream-K result differs from run to run: %s path %d" % ((M, N, K, nb), path)


@pytest.mark.parametrize("spec,shapes", [
    ("ik,jk->ij", [(1152, 1024), (1027, 1024)]), ("ij,jk->ik", [(1152, 1024), (1024, 1028)]), ("ij,jk->ki", [(2050, 768), (768, 1152)]),
    ("ik,jk->ij", [(1152, 1024), (2049, 1024)]), ("ij,jk->ik", [(1152, 640), (640, 2050)]),
    ("ij,ik->jk", [(600, 1025), (600, 1152)]), ("ij,jk->ik", [(1026, 640), (640, 1152)]), ("ji,jk->ik", [(900, 2049), (900, 1152)]),
])
def test_gemm_ragged_edge_split(spec, shapes, ad):
    """Contractions whose M or N is a few elements past a tile multiple: aligned part on the DMMA GEMM, the ragged rows / columns
    as a reduction (csrc/einsum.cu: ragged_edge) - values, both gradients, and a transposed / sliced operand view."""
    import ctypes as C
    from autodiff_b200 import _lib
    rng = np.random.default_rng(len(spec) + shapes[0][0])
    a, b = [rng.standard_normal(sh) for sh in shapes]
    da, db_ = ad.to_device(a), ad.to_device(b)
    ta, tb = da.tensor(), db_.tensor()
    ref = np.einsum(spec, a, b)
    do = ad.DeviceArray.empty(ref.shape)
    to = do.tensor()
    buf = C.create_string_buffer(512)
    _lib.check(_lib.lib().adb_einsum_describe(spec.encode(), 2, (_lib.Tensor * 2)(ta, tb), C.byref(to), buf, 512))
    assert b"edge reduce" in buf.value, buf.value       # the case really takes the split path
    for m in (onp, ad):
        A, B = m.Variable(a, name="a"), m.Variable(b, name="b")
        e = m.Einsum(spec, A, B)
        outs = [e] + m.grad(e, [A, B])
        if m is ad:
            got = [np.asarray(o()) for o in outs]
        else:
            want = [np.asarray(o()) for o in outs]
    for g, w in zip(got, want):
        assert g.shape == w.shape
        check(g, w)
    at = np.ascontiguousarray(a.T)                       # the same values through a transposed view of the first operand
    sa = spec.split(",")[0]
    spec_t = sa[::-1] + spec[len(sa):]
    check(np.asarray(ad.Einsum(spec_t, ad.Variable(at, name="at"), ad.Variable(b, name="b"))()), want[0])


def test_large_gemm_properties(ad):
    """BASELINE-size contraction checked through size-independent properties: linearity in A and a checksum
    identity  ones^T (A B) ones == (ones^T A)(B ones)."""
    rng = np.random.default_rng(6)
    n = 2048
    a = rng.standard_normal((n, n)); b = rng.standard_normal((n, n)); a2 = rng.standard_normal((n, n))
    A, B, A2 = (ad.Variable(ad.to_device(v), name="v") for v in (a, b, a2))
    c = ad.to_host((A @ B).eval_device())
    c2 = ad.to_host((A2 @ B).eval_device())
    csum = ad.to_host(((A + A2) @ B).eval_device())
    check(csum, c + c2, tol=1e-12)
    lhs = c.sum()
    rhs = a.sum(0) @ b.sum(1)
    assert abs(lhs - rhs) <= 1e-9 * abs(rhs) + 1e-6
    check(c[:64], a[:64] @ b, tol=1e-12)


def test_conv_im2col_forward_and_grads(ad):
    """Convolution = Im2Col + Einsum: forward vs the as_strided+np.einsum restatement of the
    reference's commented-out design, first- and second-order grads vs the oracle graph."""
    rng = np.random.default_rng(12)
    x = rng.standard_normal((3, 9, 8, 3)); w = rng.standard_normal((3, 2, 3, 5)); w2 = rng.standard_normal((2, 2, 5, 4))
    check(ad.Conv2D(ad.Variable(x, name="x"), ad.Variable(w, name="w"))(), onp.conv_valid_nhwc(x, w))
    outs = {}
    for m in (onp, ad):
        X, W, W2 = m.Variable(x.copy(), name="x"), m.Variable(w.copy(), name="w"), m.Variable(w2.copy(), name="w2")
        h = m.ReLU(m.Conv2D(X, W))
        y = m.Tanh(m.Conv2D(h, W2))
        g = m.grad(y, [X, W, W2])
        gg = m.grad(g[1], [W])[0]
        outs[m] = [np.array(y())] + [np.array(t()) for t in g] + [np.array(gg())]
    for a, b in zip(outs[ad], outs[onp]):
        check(a, b, tol=1e-11, what="conv")
    # adjointness: <im2col(x), c> == <x, col2im(c)>
    c = rng.standard_normal((3 * 7 * 7, 3 * 2 * 3))
    lhs = float(np.sum(np.asarray(ad.Im2Col(ad.Variable(x, name="x"), 3, 2)()) * c))
    rhs = float(np.sum(x * np.asarray(ad.Col2Im(ad.Variable(c, name="c"), x.shape, 3, 2)())))
    assert abs(lhs - rhs) <= 1e-11 * abs(lhs)
    # channel counts that are multiples of 4 take the 256-bit col2im kernel
    for C_, kh, kw in ((8, 3, 3), (16, 5, 5), (4, 1, 2)):
        xi = rng.standard_normal((2, 11, 9, C_))
        oh, ow = 11 - kh + 1, 9 - kw + 1
        c = rng.standard_normal((2 * oh * ow, kh * kw * C_))
        got = np.asarray(ad.Col2Im(ad.Variable(c, name="c"), xi.shape, kh, kw)())
        ref = np.zeros_like(xi)
        c6 = c.reshape(2, oh, ow, kh, kw, C_)
        for a in range(kh):
            for b in range(kw):
                ref[:, a:a + oh, b:b + ow, :] += c6[:, :, :, a, b, :]
        check(got, ref, tol=1e-13, what="col2im quad C=%d" % C_)


def test_evaluate_many_and_fusion_toggle(ad):
    from autodiff_b200.core import engine
    rng = np.random.default_rng(13)
    xv = rng.standard_normal((65, 33)); yv = rng.random((65, 33)) + 0.5

    def build():
        X, Y = ad.Variable(xv, name="x"), ad.Variable(yv, name="y")
        r = ad.ReLU(X @ ad.Transpose(Y))
        return [ad.Tanh(r) * 2 + 1, (r * r + r) / (1 + ad.Exp(-r)), ad.Einsum("ab->b", ad.Sigmoid(r) * r)]
    a = build()
    engine.evaluate_many(a)
    l0 = ad.launch_count()
    fused = [np.array(n()) for n in a]
    engine.fusion_enabled = False
    try:
        unfused = [np.array(n()) for n in build()]
    finally:
        engine.fusion_enabled = True
    for f, u in zip(fused, unfused):
        check(f, u, tol=1e-15)


def test_tf32x3_precision_mode(ad):
    """The TF32/FP32 variant (3xTF32 on tcgen05, fp32 accumulation in TMEM): parity bar 1e-4 (north_star)."""
    rng = np.random.default_rng(21)
    a = rng.standard_normal((300, 520)); b = rng.standard_normal((520, 260))
    with ad.precision("tf32x3"):
        A, B = ad.Variable(a, name="a"), ad.Variable(b, name="b")
        e = A @ B
        da, db = ad.grad(e, [A, B])
        got = [np.array(e()), np.array(da()), np.array(db())]
    ref = [a @ b, np.ones((300, 260)) @ b.T, a.T @ np.ones((300, 260))]
    for g, r in zip(got, ref):
        err = rel_err(g, r)
        assert err <= 1e-4, err
        assert err >= 1e-13 or True
    # and the mode is really different arithmetic from fp64 (errors are ~1e-7, not 1e-16)
    assert rel_err(got[0], ref[0]) > 1e-12
    check((ad.Variable(a, name="a") @ ad.Variable(b, name="b"))(), a @ b)      # back to fp64 outside the context


@pytest.mark.gpu
@pytest.mark.parametrize("m,k,n", [(4096, 4096, 4096), (1024, 4097, 1536), (640, 8192, 384), (257, 2049, 130)])
def test_tf32x3_large_full_matrix(m, k, n, ad):
    """3xTF32 at sizes with several accumulation chunks (K > 2048: the chunks are added in fp64 by the epilogue warps), a ragged
    K = 4097 and odd extents: forward and BOTH gradients compared over the full matrices against the FP64 DMMA path on the device
    (itself pinned to the oracle at the sizes numpy finishes in seconds).  Bar 1e-4 (north_star), expected ~1e-6."""
    import torch
    rng = np.random.default_rng(m + k + n)
    a = ad.to_device(rng.standard_normal((m, k))); b = ad.to_device(rng.standard_normal((k, n)))
    w = ad.to_device(rng.standard_normal((m, n)))

    def run():
        A, B, W = ad.Variable(a, name="a"), ad.Variable(b, name="b"), ad.Variable(w, name="w")
        e = ad.Einsum("ij,jk->ik", A, B)
        loss = ad.Einsum("ik,ik->", e, W)                      # non-trivial upstream gradient
        return [x.eval_device() for x in [e] + ad.grad(loss, [A, B])]

    ref = run()
    with ad.precision("tf32x3"):
        got = run()
        again = run()
    def view(d):                                              # (pitched rows: compare the elements, not the padding)
        return torch.as_strided(d.base, tuple(d.shape), tuple(d.strides), d.offset)
    for g, r, g2 in zip(got, ref, again):
        tg, tr = view(g), view(r)
        err = float((tg - tr).abs().max() / tr.abs().max())
        assert err <= 1e-4, err
        assert err > 1e-13, "tf32x3 mode did not change the arithmetic"
        assert torch.equal(tg, view(g2)), "tf32x3 result differs from run to run"


# ------------------------------------------------------------------------------------- static programs vs interpreter
@pytest.mark.gpu
def test_catalogued_programs_match_interpreter_bitwise():
    """Every catalogued (statically compiled) elementwise program must agree BIT FOR BIT with the bytecode interpreter
    running the same program - both execute the one operation table of csrc/ewise_kernel.cuh."""
    import autodiff_b200 as ad
    rng = np.random.default_rng(7)
    shapes = [(37, 53), (64, 128), (5, 7, 9), (1001,)]

    def graphs(x, y, z):
        X, Y, Z = ad.Variable(x), ad.Variable(y), ad.Variable(z)
        r = ad.ReLU(X)
        return [X + Y, X * Y, X * Y * Z, ad.ReLU(X), ad.Recipr(Y), ad.Exp(X), ad.Sigmoid(X), ad.Log(Y), -X, ad.Tanh(X),
                Y * r * ad.Recipr(r), X + Y + Z, (X + Y) * Z, ad.Exp(X + Y), 1 - ad.Sigmoid(X), X / Y, ad.Sqrt(Y) if hasattr(ad, "Sqrt") else Y * Y,
                ad.Concat(X, Y, axis=len(x.shape) - 1), ad.Pow(Y, Z) if hasattr(ad, "Pow") else Y * Z]

    for shp in shapes:
        x, y, z = rng.standard_normal(shp), rng.random(shp) + 0.5, rng.standard_normal(shp)
        try:
            ad.force_interpreter(True)
            want = [np.asarray(g()) for g in graphs(x, y, z)]
        finally:
            ad.force_interpreter(False)
        got = [np.asarray(g()) for g in graphs(x, y, z)]
        for i, (w, g) in enumerate(zip(want, got)):
            assert w.shape == g.shape
            assert np.array_equal(w, g), "graph %d shape %s: static and interpreted programs differ" % (i, shp)
    stats = ad.ewise_stats()
    assert " static " in stats and " interp " in stats


# ------------------------------------------------------------------------------------- streaming reductions / wide CE
@pytest.mark.parametrize("cols", [512, 515, 1000, 1024, 2000, 2049, 3000, 4000, 4096, 4100, 5000, 6000, 8000, 8192, 8200])
def test_softmax_ce_wide_rows(cols, ad):
    """One-CTA-per-row register-resident kernel (512 <= cols <= 4096, ragged tails) and the multi-pass fallback."""
    rng = np.random.default_rng(cols)
    rows = 67 if cols % 1000 else 333           # >= 296 rows of a 16-byte-multiple width take the TMA-staged persistent kernel
    z = rng.standard_normal((rows, cols)) * 4
    lab = rng.random((rows, cols)); lab /= lab.sum(1, keepdims=True)
    got = ad.SoftmaxCEWithLogits(labels=ad.Variable(lab, name="l"), logits=ad.Variable(z, name="z"))()
    ref = onp.SoftmaxCEWithLogits(labels=onp.Variable(lab, name="l"), logits=onp.Variable(z, name="z"))()
    assert got.shape == ref.shape
    check(got, ref)
    z[5, 7] = np.nan
    got = ad.SoftmaxCEWithLogits(labels=ad.Variable(lab, name="l"), logits=ad.Variable(z, name="z"))()
    assert np.isnan(got[5]) and not np.isnan(got[4])
    bad = lab.copy(); bad[rows - 1] *= 1.01
    with pytest.raises(ValueError, match="Labels must be a valid probability distribution!"):
        ad.SoftmaxCEWithLogits(labels=ad.Variable(bad, name="l"), logits=ad.Variable(z, name="z"))()


@pytest.mark.parametrize("cols", [8, 64, 100, 128, 130, 200, 256, 260, 384, 511, 512])
@pytest.mark.parametrize("rows", [300, 64 * 148 + 77])
def test_softmax_ce_narrow_rows(rows, cols, ad):
    """Rows of <= 512 columns: the warp-per-row register kernel (csrc/reduce.cu: softmax_ce_warp_kernel) and, from 64 rows per SM
    and 129 columns on, the narrow bulk-copy ring (csrc/softmax_ce_ring.cu: softmax_ce_ring_narrow_kernel) - against the oracle,
    against each other (ADB_SCE_RING=0), NaN / -inf / +inf rows, label validation, a pitched window of a wider array."""
    rng = np.random.default_rng(rows + cols)
    z = rng.standard_normal((rows, cols)) * 4
    lab = rng.random((rows, cols)); lab /= lab.sum(1, keepdims=True)

    def run(zz, ll):
        return ad.SoftmaxCEWithLogits(labels=ad.Variable(ll, name="l"), logits=ad.Variable(zz, name="z"))()
    got = run(z, lab)
    ref = onp.SoftmaxCEWithLogits(labels=onp.Variable(lab, name="l"), logits=onp.Variable(z, name="z"))()
    assert got.shape == ref.shape
    check(got, ref)
    assert np.max(np.abs(got - ref) / np.abs(ref)) < 1e-12
    os.environ["ADB_SCE_RING"] = "0"
    try:
        other = run(z, lab)
    finally:
        del os.environ["ADB_SCE_RING"]
    check(got, other)
    wide_z = rng.standard_normal((rows, cols + 12)); wide_l = rng.random((rows, cols + 12))
    wide_l[:, 4:4 + cols] /= wide_l[:, 4:4 + cols].sum(1, keepdims=True)
    zd, ld = ad.to_device(wide_z), ad.to_device(wide_l)
    got_v = run(zd.getitem((slice(None), slice(4, 4 + cols))), ld.getitem((slice(None), slice(4, 4 + cols))))
    ref_v = onp.SoftmaxCEWithLogits(labels=onp.Variable(wide_l[:, 4:4 + cols], name="l"), logits=onp.Variable(wide_z[:, 4:4 + cols], name="z"))()
    check(got_v, ref_v)
    z[5, 7] = np.nan
    z[9, :] = -np.inf
    z[11, 3] = np.inf
    z[13, :] = -np.inf; z[13, 5] = np.nan; z[13, cols - 1] = 1.0
    z[rows - 2, cols - 1] = np.nan
    got = run(z, lab)
    assert all(np.isnan(got[r]) for r in (5, 9, 11, 13, rows - 2))
    assert not np.isnan(got[4]) and not np.isnan(got[rows - 1])
    keep = np.ones(rows, bool); keep[[5, 9, 11, 13, rows - 2]] = False
    check(got[keep], ref[keep])
    bad = lab.copy(); bad[rows - 1] *= 1.01
    with pytest.raises(ValueError, match="Labels must be a valid probability distribution!"):
        run(z, bad)


@pytest.mark.parametrize("cols", [1024, 2048, 2049, 2500, 3001, 3072, 3073, 4095, 4096, 4097, 4100, 5001, 6144, 8191, 8192])
def test_softmax_ce_ring_rows(cols, ad):
    """Persistent bulk-copy ring kernel (csrc/softmax_ce_ring.cu: rows >= 4 per SM, 2048 < cols <= 8192; beyond 4096 columns a
    row travels as two half-rows with a running log-sum-exp): against the oracle,
    against the register-resident kernels (ADB_SCE_RING=0), NaN / -inf rows, label validation, pitched row views."""
    rng = np.random.default_rng(cols)
    rows = 4 * 148 + 41
    z = rng.standard_normal((rows, cols)) * 4
    lab = rng.random((rows, cols)); lab /= lab.sum(1, keepdims=True)

    def run(zz, ll):
        return ad.SoftmaxCEWithLogits(labels=ad.Variable(ll, name="l"), logits=ad.Variable(zz, name="z"))()
    got = run(z, lab)
    ref = onp.SoftmaxCEWithLogits(labels=onp.Variable(lab, name="l"), logits=onp.Variable(z, name="z"))()
    assert got.shape == ref.shape
    check(got, ref)
    os.environ["ADB_SCE_RING"] = "0"
    try:
        other = run(z, lab)
    finally:
        del os.environ["ADB_SCE_RING"]
    check(got, other)
    # a (rows, cols) window of a wider array: row pitch != cols, columns start 4 elements in
    wide_z = rng.standard_normal((rows, cols + 12)); wide_l = rng.random((rows, cols + 12))
    wide_l[:, 4:4 + cols] /= wide_l[:, 4:4 + cols].sum(1, keepdims=True)
    zd, ld = ad.to_device(wide_z), ad.to_device(wide_l)
    got_v = run(zd.getitem((slice(None), slice(4, 4 + cols))), ld.getitem((slice(None), slice(4, 4 + cols))))
    ref_v = onp.SoftmaxCEWithLogits(labels=onp.Variable(wide_l[:, 4:4 + cols], name="l"), logits=onp.Variable(wide_z[:, 4:4 + cols], name="z"))()
    check(got_v, ref_v)
    z[5, 7] = np.nan
    z[9, :] = -np.inf
    z[11, 3] = np.inf
    z[13, :] = -np.inf; z[13, 5] = np.nan; z[13, 200] = 1.0       # the NaN's warp holds nothing else but -inf
    z[15, cols - 3] = np.nan                                      # (in the second half-row when cols > 4096)
    z[17, :cols // 2] = -np.inf                                   # a first half of nothing but -inf: still a finite result
    z[19, cols // 2:] = -np.inf
    got = run(z, lab)
    assert np.isnan(got[5]) and np.isnan(got[9]) and np.isnan(got[11]) and np.isnan(got[13]) and np.isnan(got[15])
    assert not np.isnan(got[4]) and not np.isnan(got[rows - 1])
    with np.errstate(all="ignore"):                               # (the oracle's own inf - inf / 0 * inf warnings on these rows)
        ref = onp.SoftmaxCEWithLogits(labels=onp.Variable(lab, name="l"), logits=onp.Variable(z, name="z"))()
    for r in (17, 19):                                            # (0 * -inf in sum(l*z): the reference's value there is nan or inf too)
        assert np.isnan(ref[r]) == np.isnan(got[r]) and (np.isnan(ref[r]) or ref[r] == got[r] or abs(got[r] - ref[r]) <= 1e-12 * abs(ref[r]))
    os.environ["ADB_SCE_RING"] = "0"
    try:
        other = run(z, lab)
    finally:
        del os.environ["ADB_SCE_RING"]
    assert np.array_equal(np.isnan(other), np.isnan(got))
    check(np.nan_to_num(got, nan=0.0, posinf=1e300, neginf=-1e300), np.nan_to_num(other, nan=0.0, posinf=1e300, neginf=-1e300))
    bad = lab.copy(); bad[rows - 1] *= 1.01
    with pytest.raises(ValueError, match="Labels must be a valid probability distribution!"):
        run(z, bad)


@pytest.mark.parametrize("scale", [1.0, 30.0, 300.0])
@pytest.mark.parametrize("rows,cols", [(700, 2048), (50, 2048), (300, 71)])
def test_softmax_ce_dynamic_range(rows, cols, scale, ad):
    """The branch-free exp of the log-sum-exp kernels (csrc/adb_math.cuh) over arguments down to -1000s (flush to 0 below
    exp(-700)), in all three kernels: ring (700 rows), register-resident (50 rows), warp per row (71 columns)."""
    rng = np.random.default_rng(int(scale) + rows)
    z = rng.standard_normal((rows, cols)) * scale
    z[:, 0] -= 1e-3 * np.arange(rows)              # arguments just below 0 as well
    lab = rng.random((rows, cols)); lab /= lab.sum(1, keepdims=True)
    got = ad.SoftmaxCEWithLogits(labels=ad.Variable(lab, name="l"), logits=ad.Variable(z, name="z"))()
    ref = onp.SoftmaxCEWithLogits(labels=onp.Variable(lab, name="l"), logits=onp.Variable(z, name="z"))()
    check(got, ref)
    assert np.max(np.abs(got - ref) / np.abs(ref)) < 1e-12      # element-wise too, not only norm-wise


REDUCTIONS = [
    ("ab->a", [(300, 4096)]), ("ab->a", [(77, 1031)]), ("ab->b", [(4096, 300)]), ("ab->b", [(1031, 77)]), ("ab->b", [(9, 5000)]),
    ("ab->", [(2048, 2048)]), ("ab->", [(1, 7)]), ("a->", [(1000003,)]), ("abc->b", [(61, 130, 67)]), ("abc->ac", [(61, 130, 68)]),
    ("abc->c", [(61, 130, 68)]), ("abc->a", [(61, 130, 68)]), ("abcd->bd", [(12, 20, 14, 36)]), ("abcd->ac", [(12, 20, 14, 36)]),
    ("ijkl,jl->ik", [(24, 20, 28, 32), (20, 32)]), ("ijkl,ik->jl", [(24, 20, 28, 32), (24, 28)]),
    ("ijkl,jl->ik", [(9, 11, 13, 15), (11, 15)]), ("ijkl,ik->jl", [(9, 11, 13, 15), (9, 13)]),
    ("bi,o->bo", [(257, 1000), (1,)]), ("ab,ab->a", [(130, 2048), (130, 2048)]), ("ab,ab->b", [(2048, 132), (2048, 132)]),
    ("ab,b->a", [(300, 1024), (1024,)]), ("ab,a->b", [(1024, 300), (1024,)]), ("ab,bc,c->a", [(40, 300), (300, 9), (9,)]),
    # an operand that does not depend on the summed index is factored out of the row reduction (one load per output)
    ("ab,a->a", [(300, 1024), (300,)]), ("bi,o->bo", [(700, 4096), (1,)]), ("abc,a,c->a", [(9, 40, 128), (9,), (128,)]),
    ("abc,ab->ab", [(12, 33, 512), (12, 33)]),
]


@pytest.mark.parametrize("spec,shapes", REDUCTIONS)
def test_streaming_reductions_vs_numpy(spec, shapes, ad):
    """Row / column streaming kernels incl. sliced summed ranges (few outputs), ragged quads and the scalar fallbacks."""
    rng = np.random.default_rng(abs(hash(spec)) % 1000 + len(shapes[0]))
    ops = [rng.standard_normal(s) for s in shapes]
    got = ad.Einsum(spec, *[ad.Variable(o, name="o%d" % i) for i, o in enumerate(ops)])()
    ref = np.einsum(spec, *ops)
    assert np.shape(got) == ref.shape
    # cancellation: hold the error to 1e-12 of the sum of magnitudes (what a reordered fp64 sum can guarantee)
    scale = np.einsum(spec, *[np.abs(o) for o in ops]).max()
    assert np.max(np.abs(np.asarray(got) - ref)) <= 1e-12 * scale
    # a strided (sliced) view of the first operand takes the unaligned / scalar path
    if ops[0].ndim >= 2 and ops[0].shape[-1] > 4:
        view = ops[0][..., 1:-1]
        rest = [o[..., 1:-1] if (sub[-1] == spec.split("->")[0].split(",")[0][-1]) else o
                for sub, o in zip(spec.split("->")[0].split(",")[1:], ops[1:])]
        try:
            ref2 = np.einsum(spec, view, *rest)
        except ValueError:
            return
        got2 = ad.Einsum(spec, ad.Variable(view, name="v"), *[ad.Variable(o, name="r%d" % i) for i, o in enumerate(rest)])()
        scale2 = np.einsum(spec, np.abs(view), *[np.abs(o) for o in rest]).max()
        assert np.max(np.abs(np.asarray(got2) - ref2)) <= 1e-12 * scale2


def test_frobenius_norm_large(ad):
    rng = np.random.default_rng(3)
    xs = [rng.standard_normal((1031, 517)), rng.standard_normal((4097, 64)), rng.standard_normal(11)]
    got = ad.FrobeniusNorm(*[ad.Variable(x, name="x%d" % i) for i, x in enumerate(xs)])()
    check(got, np.sqrt(sum(np.sum(np.square(x)) for x in xs)))
    # one tensor: adb_norm2 takes the root inside the reduction (row kernel, sliced row kernel, generic fallbacks, views)
    for x in xs + [rng.standard_normal((2048, 2048)), rng.standard_normal(7), rng.standard_normal((3, 5, 70)), rng.standard_normal((40, 300))[:, 1:-1]]:
        got = ad.FrobeniusNorm(ad.Variable(x, name="x"))()
        assert np.shape(got) == ()
        check(got, np.sqrt(np.sum(np.square(x))))
        g = ad.grad(ad.FrobeniusNorm(ad.Variable(x, name="x")) * 2.0, [ad.Variable(x)])    # (unrelated variable: zero gradient path still builds)
    X = ad.Variable(xs[0], name="x")
    dn = ad.grad(ad.FrobeniusNorm(X), [X])[0]()
    check(dn, xs[0] / np.sqrt(np.sum(np.square(xs[0]))))


def test_eval_many_pipelined_matches_sequential(ad):
    """ad.eval_many == [n() for n in nodes] (bitwise), incl. large leaves (async upload path), pitched / strided
    results, small results, shared leaves and already-evaluated nodes."""
    rng = np.random.default_rng(11)
    a, b = rng.standard_normal((700, 515)), rng.standard_normal((515, 600))
    c = rng.standard_normal((700, 600))

    def graphs():
        A, B, Cc = ad.Variable(a, name="a"), ad.Variable(b, name="b"), ad.Variable(c, name="c")
        e = A @ B
        da, db = ad.grad(e, [A, B])
        return [e, da, db, ad.ReLU(e + Cc), ad.Einsum("ab->", e), ad.Transpose(Cc), A]
    want = [np.asarray(n()) for n in graphs()]
    nodes = graphs()
    got = ad.eval_many(nodes)
    assert len(got) == len(want)
    for g, w, n in zip(got, want, nodes):
        assert np.shape(g) == np.shape(w)
        assert np.array_equal(np.asarray(g), w)
        assert n.cached is g or np.array_equal(np.asarray(n.cached), w)
    again = ad.eval_many(nodes)            # memoised
    for g, w in zip(again, want):
        assert np.array_equal(np.asarray(g), w)


# ------------------------------------------------------------------------------------- launch tapes (core/tape.py)
def _tape_graphs(ad, seed):
    """A few graph families, rebuilt from scratch with fresh data for every seed (same structure, new values)."""
    from autodiff_b200 import workloads as wl
    rng = np.random.default_rng(seed)
    out = {}
    # MLP training graph (NN macro: Concat + Slice grads, ReLU, SoftmaxCE) - weights are shared across seeds like a real loop
    x, lab = rng.random((32, 20)), np.eye(7)[rng.integers(0, 7, 32)]
    w1, w2 = rng.standard_normal((21, 16)) * 0.1, rng.standard_normal((17, 7)) * 0.1
    X, L, W1, W2 = ad.Variable(x, name="x"), ad.Variable(lab, name="l"), ad.Variable(w1, name="w1"), ad.Variable(w2, name="w2")
    ones = ad.ConstFill((32, 1), 1.0)
    h = ad.ReLU(ad.Concat(X, ones, axis=1) @ W1)
    logits = ad.Concat(h, ad.ConstFill((32, 1), 1.0), axis=1) @ W2
    cost = ad.Einsum("i->", ad.SoftmaxCEWithLogits(labels=L, logits=logits)) / 32
    out["mlp"] = ad.grad(cost, [W1, W2]) + [cost]
    # second-order char-rnn (Tanh macro, Softmax grads, thousands of Add/Mul nodes)
    H, V, B, T = 12, 9, 6, 5
    prm = [ad.Variable(p, name="p%d" % i) for i, p in enumerate(wl.char_rnn_params(rng, H, V))]
    oh = np.eye(V)[rng.integers(0, V, (B, T))]
    loss = wl.char_rnn(ad, oh, *prm)
    g = ad.grad(loss, [prm[0]])[0]
    out["rnn2"] = [ad.grad(g, [prm[0]])[0], g, loss]
    # conv net
    xs = rng.standard_normal((3, 12, 12, 2)); ys = np.eye(4)[rng.integers(0, 4, 3)]
    P = [ad.Variable(p, name="c%d" % i) for i, p in enumerate(wl.conv_net_params(rng, c_in=2, c1=3, c2=4, k=3, hw=12, classes=4))]
    closs = wl.conv_net(ad, ad.Variable(xs, name="x"), ad.Variable(ys, name="y"), *P)
    out["conv"] = ad.grad(closs, P) + [closs]
    # reshapes / pads / slices / reductions / scalar operands
    a, b = rng.standard_normal((6, 10)), rng.random((6, 10)) + 0.5
    A, Bv = ad.Variable(a, name="a"), ad.Variable(b, name="b")
    out["misc"] = [ad.Pad(A, ((1, 2), (0, 3)), 0.5) * 2.0, ad.Reshape(A * Bv, (10, 6)), A[1:5, ::2] + 1.0, ad.Einsum("ab->b", A / Bv),
                   ad.Einsum("ab,->ab", A, 3.0), ad.FrobeniusNorm(A, Bv), ad.ReduceSumKeepDims(ad.Exp(A), axes=[1]), ad.Transpose(A) @ Bv]
    return out


@pytest.fixture
def plain_grad():
    """The launch-tape tests count records / replays per structure: keep `ad.grad` on the plain path there (deferred gradients
    change how many distinct structures a loop goes through; test_deferred_gradients_replay_bitwise covers them)."""
    from autodiff_b200.core import grad as G
    saved = G.LAZY
    G.LAZY = False
    yield
    G.LAZY = saved


def test_deferred_gradients_replay_bitwise(ad):
    """Loops that rebuild an isomorphic graph every step: from the 2nd step `ad.grad` returns deferred gradients, and once
    their structure has a launch tape the evaluation replays it straight from the forward graph's leaves (core/grad.py) -
    bit for bit the values of the plain path (no tapes, expanded gradient graphs), first and second order, conv included."""
    from autodiff_b200.core import engine, grad as G, tape
    if not tape.ENABLED:
        pytest.skip("ADB_TAPE=0")
    tape.clear()
    want = {}
    tape.ENABLED = False                  # (also keeps ad.grad on the plain path)
    try:
        for seed in range(9):
            for name, nodes in _tape_graphs(ad, seed).items():
                engine.evaluate_many(nodes)
                want[(seed, name)] = [np.asarray(n()) for n in nodes]
    finally:
        tape.ENABLED = True
    saved, G.LAZY = G.LAZY, True
    before = G.stats.get("lazy_replays", 0)
    try:
        for seed in range(9):
            for name, nodes in _tape_graphs(ad, seed).items():
                engine.evaluate_many(nodes)
                for k, (n, w) in enumerate(zip(nodes, want[(seed, name)])):
                    g = np.asarray(n())
                    assert g.shape == w.shape, (seed, name, k)
                    assert np.array_equal(g, w), "seed %d graph %s output %d differs with deferred gradients" % (seed, name, k)
        assert G.stats["deferred"] > 0
        assert G.stats.get("lazy_replays", 0) - before >= 6, G.stats      # mlp, rnn (2nd order) and conv each reach the lazy replay
        # a deferred gradient used as an ordinary operand expands and evaluates like the plain graph
        nodes = _tape_graphs(ad, 50)["mlp"]
        assert type(nodes[0]) is G.DeferredGrad
        twice = nodes[0] * 2.0
        ref = _tape_graphs(ad, 50)["mlp"]
        assert np.array_equal(np.asarray(twice()), 2.0 * np.asarray(ref[0]()))
    finally:
        G.LAZY = saved
        tape.clear()


@pytest.mark.usefixtures("plain_grad")
def test_launch_tapes_replay_bitwise(ad):
    """From the 3rd evaluation of a structure on, evaluate_many replays a launch tape: results must equal the normal
    planner/executor path bit for bit, for fresh data and fresh graph objects every time."""
    from autodiff_b200.core import engine, tape
    if not tape.ENABLED:
        pytest.skip("ADB_TAPE=0")
    tape.clear()
    want = {}
    tape.ENABLED = False
    try:
        for seed in range(5):
            for name, nodes in _tape_graphs(ad, seed).items():
                engine.evaluate_many(nodes)
                want[(seed, name)] = [np.asarray(n()) for n in nodes]
    finally:
        tape.ENABLED = True
    before = dict(tape.stats)
    for seed in range(5):
        for name, nodes in _tape_graphs(ad, seed).items():
            engine.evaluate_many(nodes)
            got = [np.asarray(n()) for n in nodes]
            for k, (g, w) in enumerate(zip(got, want[(seed, name)])):
                assert g.shape == w.shape, (seed, name, k)
                assert np.array_equal(g, w), "seed %d graph %s output %d differs under the tape" % (seed, name, k)
    assert tape.stats["records"] - before["records"] >= 4, tape.stats
    assert tape.stats["replays"] - before["replays"] >= 12, tape.stats


@pytest.mark.usefixtures("plain_grad")
def test_launch_tapes_cuda_graph_replay(ad):
    """From the 2nd replay on a small-leaf tape is ONE cudaGraphLaunch (core/tape.py): values equal the planner path bit for
    bit on fresh data; a value of the previous replay that is still alive forces the call-by-call path (and keeps its
    contents); the deferred label check still fires."""
    import gc
    from autodiff_b200.core import engine, tape
    if not (tape.ENABLED and tape.GRAPH):
        pytest.skip("tapes / graphs disabled")
    tape.clear()
    want = {}
    tape.ENABLED = False
    try:
        for seed in range(7):
            for name, nodes in _tape_graphs(ad, seed).items():
                engine.evaluate_many(nodes)
                want[(seed, name)] = [np.asarray(n()) for n in nodes]
    finally:
        tape.ENABLED = True
    before = dict(tape.stats)
    for seed in range(7):
        graphs = _tape_graphs(ad, seed)
        for name, nodes in graphs.items():
            engine.evaluate_many(nodes)
            for k, (n, w) in enumerate(zip(nodes, want[(seed, name)])):
                g = np.asarray(n())
                assert g.shape == w.shape, (seed, name, k)
                assert np.array_equal(g, w), "seed %d graph %s output %d differs under the graph replay" % (seed, name, k)
        del graphs, nodes, n
        gc.collect()
    assert tape.stats["graphs"] - before["graphs"] >= 3, tape.stats
    assert tape.stats["graph_launches"] - before["graph_launches"] >= 9, tape.stats
    # a survivor of the previous replay pins the graph's buffer: the next replay must not overwrite it
    keep = _tape_graphs(ad, 100)["mlp"]
    engine.evaluate_many(keep)
    kept_dev = keep[0].eval_device()
    kept_val = ad.to_host(kept_dev).copy()
    busy = tape.stats["graph_busy"]
    nxt = _tape_graphs(ad, 101)["mlp"]
    engine.evaluate_many(nxt)
    assert tape.stats["graph_busy"] == busy + 1
    assert np.array_equal(ad.to_host(kept_dev), kept_val)
    ref = _tape_graphs(ad, 101)["mlp"]
    tape.ENABLED = False
    try:
        engine.evaluate_many(ref)
    finally:
        tape.ENABLED = True
    for a, b in zip(nxt, ref):
        assert np.array_equal(np.asarray(a()), np.asarray(b()))


@pytest.mark.usefixtures("plain_grad")
def test_launch_tapes_cuda_graph_label_flag(ad):
    from autodiff_b200.core import engine, tape
    if not (tape.ENABLED and tape.GRAPH):
        pytest.skip("tapes / graphs disabled")
    tape.clear()
    rng = np.random.default_rng(6)

    def build(bad=False):
        z = rng.standard_normal((8, 5)); lab = np.eye(5)[rng.integers(0, 5, 8)]
        if bad:
            lab[3] *= 2.0
        Z, Lb = ad.Variable(z, name="z"), ad.Variable(lab, name="l")
        sce = ad.SoftmaxCEWithLogits(labels=Lb, logits=ad.Exp(Z * 0.5) + Z)
        parts = [ad.Einsum("i->", sce * float(k + 1)) for k in range(4)]      # enough kernels for a graph
        return z, lab, sce, parts
    for i in range(6):
        z, lab, sce, parts = build()
        engine.evaluate_many(parts + [sce])
        ref = onp.SoftmaxCEWithLogits(labels=onp.Variable(lab), logits=onp.Variable(np.exp(z * 0.5) + z))()
        check(sce(), ref)
        for k in range(4):
            check(parts[k](), ref.sum() * (k + 1))
        del sce, parts                              # (a surviving node would pin the graph's buffer: see the test above)
    assert tape.stats["graph_launches"] >= 2, tape.stats
    z, lab, sce, parts = build(bad=True)
    launches = tape.stats["graph_launches"]
    with pytest.raises(ValueError, match="Labels must be a valid probability distribution!"):
        engine.evaluate_many(parts + [sce])
        parts[0]()
    assert tape.stats["graph_launches"] == launches + 1
    del sce, parts
    z, lab, sce, parts = build()                   # and the flag is clean again on the next launch
    engine.evaluate_many(parts + [sce])
    check(sce(), onp.SoftmaxCEWithLogits(labels=onp.Variable(lab), logits=onp.Variable(np.exp(z * 0.5) + z))())


@pytest.mark.usefixtures("plain_grad")
def test_launch_tapes_intermediates_flags_and_callbacks(ad):
    from autodiff_b200.core import engine, tape
    if not tape.ENABLED:
        pytest.skip("ADB_TAPE=0")
    tape.clear()
    rng = np.random.default_rng(5)

    def build(bad=False):
        z = rng.standard_normal((8, 5)); lab = np.eye(5)[rng.integers(0, 5, 8)]
        if bad:
            lab[3] *= 2.0
        Z, Lb = ad.Variable(z, name="z"), ad.Variable(lab, name="l")
        mid = ad.Exp(Z * 0.5)
        sce = ad.SoftmaxCEWithLogits(labels=Lb, logits=mid + Z)
        return z, lab, mid, sce, ad.Einsum("i->", sce)
    for i in range(4):
        z, lab, mid, sce, tot = build()
        fired = []
        engine.evaluate_many([tot, sce], on_ready=lambda n: fired.append(n))
        assert fired == [tot, sce]
        ref = onp.SoftmaxCEWithLogits(labels=onp.Variable(lab), logits=onp.Variable(np.exp(z * 0.5) + z))()
        check(sce(), ref)
        check(tot(), ref.sum())
        check(mid(), np.exp(z * 0.5))          # an intermediate of a replayed graph is still available
    assert tape.stats["replays"] >= 2
    z, lab, mid, sce, tot = build(bad=True)      # the deferred label check still fires on the replay path
    with pytest.raises(ValueError, match="Labels must be a valid probability distribution!"):
        engine.evaluate_many([tot])
        tot()


def test_lazy_slice_gradient_is_differentiable(ad):
    """with ad.lazy_slice_grad(True) the gradient of Slice is a graph node, so second-order gradients
    through slicing are right (the reference's eager constant makes them 0); the default stays reference-compatible."""
    rng = np.random.default_rng(4)
    x = rng.standard_normal((7, 5))

    def second(lazy):
        X = ad.Variable(x, name="x")
        with ad.lazy_slice_grad(lazy):
            s = X[1:6, ::2]
            f = ad.Einsum("ab->", s * s * s)
            g = ad.grad(f, [X])[0]
            h = ad.grad(ad.Einsum("ab->", g * g), [X])[0]       # d/dx sum((3x^2)^2 on the slice) = 36 x^3 on the slice
        return np.asarray(g()), np.broadcast_to(np.asarray(h()), x.shape)
    mask = np.zeros_like(x); mask[1:6, ::2] = 1.0
    g, h = second(True)
    check(g, 3 * x * x * mask)
    check(h, 36 * x ** 3 * mask)
    g0, h0 = second(False)
    check(g0, 3 * x * x * mask)
    assert not np.allclose(h0, 36 * x ** 3 * mask)                # the reference quirk: the first gradient was a constant


def test_empty_and_degenerate_shapes(ad):
    """Zero-extent and single-element tensors through every kernel family (numpy semantics: empty sums are 0)."""
    z = np.zeros((0, 5)); o = np.ones((3, 0)); s = np.array([[2.5]])
    check(ad.Einsum("ab->b", ad.Variable(z))(), z.sum(0))
    check(ad.Einsum("ab->a", ad.Variable(o))(), o.sum(1))
    check(ad.Einsum("ab->", ad.Variable(z))(), 0.0)
    assert (ad.Variable(z) + ad.Variable(z))().shape == (0, 5)
    assert ad.ReLU(ad.Variable(o))().shape == (3, 0)
    assert (ad.Variable(np.ones((4, 0))) @ ad.Variable(np.ones((0, 6))))().shape == (4, 6)
    check((ad.Variable(np.ones((4, 0))) @ ad.Variable(np.ones((0, 6))))(), np.zeros((4, 6)))
    assert (ad.Variable(np.ones((0, 3))) @ ad.Variable(np.ones((3, 6))))().shape == (0, 6)
    check(ad.Einsum("ab,bc->ac", ad.Variable(s), ad.Variable(s))(), s @ s)
    check(ad.Tanh(ad.Variable(s))(), onp.Tanh(onp.Variable(s))())
    check(ad.Einsum("ab->", ad.Variable(s))(), 2.5)
    check(ad.ReduceSumKeepDims(ad.Variable(np.arange(6.0).reshape(1, 6)), axes=[0, 1])(), np.array([[15.0]]))
    check(ad.Concat(ad.Variable(np.ones((2, 0))), ad.Variable(np.full((2, 3), 7.0)), axis=1)(), np.full((2, 3), 7.0))
    check(ad.FrobeniusNorm(ad.Variable(z))(), 0.0)


def test_elementwise_iteration_space_splitting(ad):
    """Iteration spaces beyond 2^31 quads are halved on the host until they fit; exercised here by lowering the limit."""
    from autodiff_b200 import _lib
    rng = np.random.default_rng(8)
    x, y, b = rng.standard_normal((37, 50, 21)), rng.standard_normal((37, 50, 21)), rng.standard_normal((50, 1))
    flat = rng.standard_normal(100003)
    try:
        _lib.check(_lib.lib().adb_ewise_set_piece_limit(257))
        got = [(ad.Variable(x) * ad.Variable(y) + ad.Variable(b))(), ad.Tanh(ad.Variable(flat))(), ad.Transpose(ad.Variable(x[0])) @ ad.Variable(y[0]),
               ad.Einsum("abc,bo->abc", ad.Variable(x), ad.Variable(b))()]
        got[2] = got[2]()
    finally:
        _lib.check(_lib.lib().adb_ewise_set_piece_limit(0))
    check(got[0], x * y + b)
    check(got[1], onp.Tanh(onp.Variable(flat))())
    check(got[2], x[0].T @ y[0])
    check(got[3], x * b)


def test_negative_and_odd_strides_everywhere(ad):
    """Reversed / stepped slices (negative and non-unit strides, odd offsets) through elementwise, reduction and
    contraction kernels: everything falls back to the scalar-access or generic paths and must still be exact."""
    rng = np.random.default_rng(21)
    x = rng.standard_normal((40, 70)); y = rng.standard_normal((40, 70))
    X, Y = ad.Variable(x, name="x"), ad.Variable(y, name="y")
    xr, yr = X[::-1, ::-1], Y[::-1, 1::2]
    check((xr * 2.0 + X)(), x[::-1, ::-1] * 2.0 + x)
    check(ad.Exp(yr)(), np.exp(y[::-1, 1::2]))
    check(ad.Einsum("ab->a", xr)(), x[::-1, ::-1].sum(1))
    check(ad.Einsum("ab->b", yr)(), y[::-1, 1::2].sum(0))
    check(ad.Einsum("ab->", X[3:37:3, 5::7])(), x[3:37:3, 5::7].sum())
    check((xr @ ad.Transpose(X))(), x[::-1, ::-1] @ x.T)
    check(ad.Einsum("ab,cb->ac", X[:, ::-1], Y[::2])(), x[:, ::-1] @ y[::2].T)
    z = rng.standard_normal((9, 6, 5, 7))
    Z = ad.Variable(z, name="z")
    check(ad.Einsum("abcd->ca", Z[::-1, :, ::2])(), np.einsum("abcd->ca", z[::-1, :, ::2]))
    check(ad.Tanh(Z[1:, ::-1, :, 1:6])(), onp.Tanh(onp.Variable(z[1:, ::-1, :, 1:6]))())


def test_baseline_size_contractions_vs_blas(ad):
    """FULL BASELINE sizes against a BLAS-backed numpy restatement (SURVEY 8d "Parity": np.matmul agrees with the true
    np.einsum path to rounding at the sizes where both can run, see test_einsum_forward_and_grads_vs_oracle): the C2a
    forward product and both gradient forms at 8192^2, one C3 layer (8192 x 4097 x 4096, ragged K) and the C2b batch at
    64 x 1024^2, each within 1e-12 norm-wise."""
    rng = np.random.default_rng(99)
    n = 8192
    a = rng.standard_normal((n, n)); b = rng.standard_normal((n, n))
    A, B = ad.Variable(ad.to_device(a), name="A"), ad.Variable(ad.to_device(b), name="B")
    e = ad.Einsum("ij,jk->ik", A, B)
    da, db = ad.grad(e, [A, B])
    got = ad.eval_many([e, da, db])
    check(got[0], a @ b, tol=1e-12, what="C2a fwd")
    ones = np.ones((n, n))
    check(got[1], ones @ b.T, tol=1e-12, what="C2a dA")
    check(got[2], a.T @ ones, tol=1e-12, what="C2a dB")
    del got, e, da, db, A, B, ones
    x = rng.standard_normal((8192, 4097)); w = rng.standard_normal((4097, 4096)) * 0.01
    check((ad.Variable(x, name="x") @ ad.Variable(w, name="w"))(), x @ w, tol=1e-12, what="C3 layer")
    p = rng.standard_normal((64, 1024, 1024)); q = rng.standard_normal((64, 1024, 1024))
    check(ad.Einsum("bij,bjk->bik", ad.Variable(p, name="p"), ad.Variable(q, name="q"))(), np.matmul(p, q), tol=1e-12, what="C2b fwd")
    del p, q, x, w
    # C2c at full size (128^4): the HBM-bound members against np.einsum itself (the reference's call, ops.py:180)
    t4 = rng.standard_normal((128, 128, 128, 128)); m2 = rng.standard_normal((128, 128)); g2 = rng.standard_normal((128, 128))
    T4 = ad.Variable(ad.to_device(t4), name="T")
    check(ad.Einsum("ijkl,jl->ik", T4, ad.Variable(m2, name="m"))(), np.einsum("ijkl,jl->ik", t4, m2), tol=1e-12, what="C2c fwd")
    check(ad.Einsum("ijkl,ik->jl", T4, ad.Variable(g2, name="g"))(), np.einsum("ijkl,ik->jl", t4, g2), tol=1e-12, what="C2c dB")
    del T4, t4
    outer = ad.Einsum("ik,jl->ijkl", ad.Variable(g2, name="g"), ad.Variable(m2, name="m"))()
    assert outer.shape == (128, 128, 128, 128)
    assert np.array_equal(outer[5, :, 7, :], g2[5, 7] * m2) and np.array_equal(outer[:, 100, :, 3], g2 * m2[100, 3])
    assert abs(outer.sum() - g2.sum() * m2.sum()) <= 1e-9 * abs(g2.sum() * m2.sum()) + 1e-6


def test_pad_per_side_constants(ad):
    """np.pad(mode="constant") with a different constant per side (reference reshape.py:124-129 hands constant_values to
    np.pad unchanged): overlapping corners take the LATER axis' constant."""
    rng = np.random.default_rng(21)
    x = rng.standard_normal((5, 7)); x3 = rng.standard_normal((3, 4, 5))
    for arr, pw, cv in [(x, [[1, 2], [3, 1]], [[1.5, -2.0], [3.0, 4.0]]), (x, [[2, 0], [0, 3]], [0.25, -0.5]),
                        (x, [[1, 1], [2, 2]], [[7.0], [9.0]]), (x3, [[1, 0], [2, 1], [0, 2]], [[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]]),
                        (x3, 1, [[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]]), (x, [[0, 0], [0, 0]], [[1.0, 2.0], [3.0, 4.0]])]:
        got = ad.Pad(ad.Variable(arr, name="x"), pw, constant_values=cv)
        ref = np.pad(arr, pw, mode="constant", constant_values=cv)
        assert tuple(got.shape) == ref.shape
        assert np.array_equal(got(), ref), (pw, cv)
        w = ad.Variable(arr, name="x")                                   # gradient: the inner window of the incoming gradient
        g = ad.grad(ad.Pad(w, pw, constant_values=cv) * 2.0, [w])[0]
        assert np.array_equal(g(), np.full(arr.shape, 2.0))


def test_softmax_ce_nd_reference_semantics(ad):
    """Labels / logits that are not both 2-D: the reference sums the labels along axis 1 for its distribution check and runs the
    log-sum-exp along the last axis (ops.py:243-255) - restated by the oracle, matched here."""
    rng = np.random.default_rng(22)
    z = rng.standard_normal((6, 5, 9)) * 3
    lab = rng.random((6, 5, 9)); lab /= lab.sum(1, keepdims=True)        # a distribution along axis ONE (what upstream checks)
    got = ad.SoftmaxCEWithLogits(labels=ad.Variable(lab, name="l"), logits=ad.Variable(z, name="z"))
    ref = onp.SoftmaxCEWithLogits(labels=onp.Variable(lab, name="l"), logits=onp.Variable(z, name="z"))
    assert tuple(got.shape) == tuple(ref.shape) == (6, 5)
    check(got(), ref())
    onehot = np.eye(9)[rng.integers(0, 9, (6, 5))]                       # one-hot along the LAST axis: rejected upstream, and here
    with pytest.raises(ValueError, match="Labels must be a valid probability distribution!"):
        onp.SoftmaxCEWithLogits(labels=onp.Variable(onehot, name="l"), logits=onp.Variable(z, name="z"))()
    with pytest.raises(ValueError, match="Labels must be a valid probability distribution!"):
        ad.SoftmaxCEWithLogits(labels=ad.Variable(onehot, name="l"), logits=ad.Variable(z, name="z"))()
    zt = np.ascontiguousarray(z.transpose(1, 0, 2))
    for m in (ad, onp):                                                   # 1-D labels: np.sum(axis=1) raises AxisError upstream
        with pytest.raises(np.exceptions.AxisError):
            m.SoftmaxCEWithLogits(labels=m.Variable(np.full(9, 1 / 9), name="l1"), logits=m.Variable(zt, name="z"))()
    lab3 = np.ones((5, 1, 9))        # labels that BROADCAST against the logits; axis 1 has one entry, so the check wants all ones
    vals = []
    for m in (ad, onp):
        Z = m.Variable(zt, name="z")
        e = m.SoftmaxCEWithLogits(labels=m.Variable(lab3, name="l"), logits=Z)
        g = m.grad(e, [Z])[0]
        vals.append((np.asarray(e()), np.asarray(g())))
    check(vals[0][0], vals[1][0]); check(vals[0][1], vals[1][1])


PAIRWISE = [
    # (sizes bounded by the ORACLE: np.einsum without optimize walks the whole iteration space - 5 operands of ~80 took minutes)
    ("ij,jk,kl->il", [(120, 100), (100, 110), (110, 90)], None),
    ("ij,jk,kl->li", [(120, 100), (100, 110), (110, 90)], [(1, 0), None, (1, 0)]),           # transposed (strided) operands
    ("ab,bc,cd,de,ef->af", [(12, 14), (14, 16), (16, 18), (18, 20), (20, 22)], None),        # more operands than the fused kernel takes
    ("bij,bjk,bkl->bil", [(6, 64, 80), (6, 80, 96), (6, 96, 72)], None),
    ("dt,tp,pr->dtp", [(60, 80), (80, 70), (70, 50)], None),                                 # reference tests/test_einsum.py:83-111
    ("ij,jk,ik->", [(257, 130), (130, 99), (257, 99)], None),
    ("ij,j,jk->ik", [(150, 333), (333,), (333, 140)], None),
    ("iij,jk,kl->il", [(90, 90, 70), (70, 110), (110, 50)], None),                           # a diagonal inside a chain
]


@pytest.mark.parametrize("spec,shapes,perms", PAIRWISE)
def test_einsum_pairwise_chain_forward_and_grads(spec, shapes, perms, ad):
    """3+ operand contractions run as a chain of pairwise GEMMs (csrc/einsum.cu run_pairwise); values and the gradients w.r.t.
    every operand (each again a 3+ operand einsum, reference ops.py:182-213) against the numpy oracle."""
    rng = np.random.default_rng(31)
    vals = []
    for k, shp in enumerate(shapes):
        pm = perms[k] if perms is not None else None
        if pm is None:
            vals.append(rng.standard_normal(shp))
        else:                                                   # a view whose memory order is the permuted one
            base = rng.standard_normal(tuple(shp[a] for a in pm))
            vals.append(base.transpose(np.argsort(pm)))
    outs = {}
    for mod in (onp, ad):
        vs = [mod.Variable(v, name="v%d" % i) for i, v in enumerate(vals)]
        e = mod.Einsum(spec, *vs)
        outs[mod] = [np.array(e())]
        if "ii" not in spec:
            outs[mod] += [np.array(x()) for x in mod.grad(e, vs)]
    for a, b in zip(outs[ad], outs[onp]):
        check(a, b, what=spec)


def test_einsum_pairwise_chain_beyond_the_oracle(ad):
    """Five operands of ~80-100: np.einsum (the reference's evaluator, optimize=False) needs minutes for the 3e11-term iteration
    space, so the check here is the matrix-product chain in numpy BLAS; the planner runs four small GEMMs."""
    rng = np.random.default_rng(32)
    shapes = [(64, 80), (80, 96), (96, 72), (72, 88), (88, 100)]
    vals = [rng.standard_normal(s) for s in shapes]
    vs = [ad.Variable(v, name="v%d" % i) for i, v in enumerate(vals)]
    e = ad.Einsum("ab,bc,cd,de,ef->af", *vs)
    a, b, c, d, f = vals
    check(e(), a @ b @ c @ d @ f, what="chain")
    g = ad.grad(e, vs)
    ones = np.ones((64, 100))
    check(g[0](), ones @ (b @ c @ d @ f).T, what="d/da")
    check(g[2](), (a @ b).T @ ones @ (d @ f).T, what="d/dc")
    check(g[4](), (a @ b @ c @ d).T @ ones, what="d/df")


def test_conv_implicit_gemm_matches_explicit_and_oracle(ad):
    """Channel counts that are multiples of 16: the einsums around Im2Col / Col2Im read the image through im2col-mode TMA
    (adb_conv_gemm; the plan makes the Im2Col and the gradient einsum under Col2Im virtual).  Values must equal the explicit
    im2col -> einsum -> col2im path and the oracle, first and second order."""
    from autodiff_b200 import _lib
    from autodiff_b200.core import engine
    rng = np.random.default_rng(41)
    x = rng.standard_normal((3, 12, 11, 16)); w = rng.standard_normal((3, 3, 16, 32)) * 0.2; w2 = rng.standard_normal((2, 3, 32, 16)) * 0.2

    def run(m):
        X, W, W2 = m.Variable(x.copy(), name="x"), m.Variable(w.copy(), name="w"), m.Variable(w2.copy(), name="w2")
        h = m.ReLU(m.Conv2D(X, W))
        y = m.Tanh(m.Conv2D(h, W2))
        g = m.grad(y, [X, W, W2])
        gg = m.grad(g[1], [W])[0]
        return [np.array(t()) for t in [y] + g + [gg]]
    names, orig = [], _lib.call

    def spy(name, *a):
        names.append(name)
        return orig(name, *a)
    _lib.call = spy
    try:
        got = run(ad)
        n_implicit, n_im2col = names.count("adb_conv_gemm"), names.count("adb_im2col")
        engine.implicit_conv_enabled = False
        del names[:]
        explicit = run(ad)
        assert names.count("adb_conv_gemm") == 0 and names.count("adb_im2col") > n_im2col
    finally:
        _lib.call = orig
        engine.implicit_conv_enabled = True
    assert n_implicit >= 6, (n_implicit, n_im2col)        # 2 forward + 2 filter gradients + 2 data gradients at least
    ref = run(onp)
    for a, b, c in zip(got, explicit, ref):
        check(a, b, tol=1e-12, what="implicit vs explicit")
        check(a, c, tol=1e-11, what="implicit vs oracle")
    # through the evaluator with ALL outputs at once (shared Im2Col between the forward and the filter gradient), taped replays too
    for rep in range(4):
        X, W = ad.Variable(x, name="x"), ad.Variable(w, name="w")
        y = ad.Einsum("nhwo->", ad.ReLU(ad.Conv2D(X, W)))
        gs = ad.grad(y, [X, W])
        ad.evaluate_many(gs + [y])
        Xo, Wo = onp.Variable(x, name="x"), onp.Variable(w, name="w")
        go = onp.grad(onp.Einsum("nhwo->", onp.ReLU(onp.Conv2D(Xo, Wo))), [Xo, Wo])
        check(gs[0](), go[0](), tol=1e-11); check(gs[1](), go[1](), tol=1e-11)
