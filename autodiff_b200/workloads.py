"""Graph builders for the BASELINE.json configs that are not the headline sweep (used by bench.py and tools/):
C4 conv net via im2col+einsum, C5 char-rnn with a second-order gradient.  `ad` is any module with the reference
API (autodiff_b200 or the numpy oracle), so the same builder feeds the GPU arm and the CPU baseline."""
import numpy as np


def conv_net(ad, x, labels, filters1, filters2, dense):
    """NHWC x -> conv5x5 (VALID) -> ReLU -> conv5x5 -> ReLU -> flatten -> dense -> SoftmaxCE mean."""
    n = x.shape[0]
    h1 = ad.ReLU(ad.Conv2D(x, filters1))
    h2 = ad.ReLU(ad.Conv2D(h1, filters2))
    flat = ad.Reshape(h2, (n, -1))
    logits = flat @ dense
    return ad.Einsum("i->", ad.SoftmaxCEWithLogits(labels=labels, logits=logits)) / n


def conv_net_params(rng, c_in=3, c1=16, c2=32, k=5, hw=64, classes=10):
    o2 = hw - 2 * (k - 1)
    return (rng.standard_normal((k, k, c_in, c1)) * 0.1, rng.standard_normal((k, k, c1, c2)) * 0.05,
            rng.standard_normal((o2 * o2 * c2, classes)) * 0.01)


def conv_net_flops(batch, c_in=3, c1=16, c2=32, k=5, hw=64, classes=10):
    """GEMM flops of fwd + both grads (first layer has no dX)."""
    o1, o2 = hw - k + 1, hw - 2 * (k - 1)
    f1 = 2.0 * batch * o1 * o1 * (k * k * c_in) * c1
    f2 = 2.0 * batch * o2 * o2 * (k * k * c1) * c2
    f3 = 2.0 * batch * (o2 * o2 * c2) * classes
    return 2 * f1 + 3 * f2 + 3 * f3


def char_rnn(ad, onehot, w, u, b_h, v, b_o, chunk=32):
    """examples/char_rnn.py:50-62 on a (B, T, V) one-hot batch; returns the mean loss node."""
    batch, steps, vocab = onehot.shape
    hidden = w.shape[0]
    h = ad.Variable(np.zeros((1, hidden)), name="h")
    costs = []
    for t in range(steps - 1):
        x = ad.Variable(onehot[:, t, :], name="x")
        h = ad.Tanh(h @ w + x @ u + b_h)
        logits = h @ v + b_o
        y = ad.Variable(onehot[:, t + 1, :], name="y")
        costs.append(ad.Einsum("i->", ad.SoftmaxCEWithLogits(labels=y, logits=logits)))
    while len(costs) > chunk:                       # numpy's 64-operand limit in the reference (ops.py:26)
        costs = [ad.Add(*costs[i:i + chunk]) for i in range(0, len(costs), chunk)]
    return ad.Add(*costs) / steps


def char_rnn_params(rng, hidden, vocab):
    return (rng.standard_normal((hidden, hidden)) * 0.01, rng.standard_normal((vocab, hidden)) * 0.01, np.zeros(hidden),
            rng.standard_normal((hidden, vocab)) * 0.01, np.zeros(vocab))
