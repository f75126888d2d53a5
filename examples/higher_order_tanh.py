"""Derivatives of tanh up to 4th order by differentiating the gradient graph again and again (the reference's
examples/higher_order_deriv_visualization.py without matplotlib: prints a small table and checks the closed forms).

    python examples/higher_order_tanh.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import autodiff_b200 as ad

xs = np.linspace(-3, 3, 13)
x = ad.Variable(xs, name="x")
curves = [ad.Tanh(x)]
for order in range(4):
    curves.append(ad.grad(curves[-1], [x])[0])
values = [np.broadcast_to(np.asarray(c()), xs.shape) for c in curves]

t = np.tanh(xs)
closed = [t, 1 - t ** 2, -2 * t * (1 - t ** 2), (1 - t ** 2) * (6 * t ** 2 - 2), (1 - t ** 2) * (16 * t - 24 * t ** 3)]
print("    x    " + "".join("   d^%d tanh " % k for k in range(5)))
for i, xv in enumerate(xs):
    print("%7.2f  " % xv + "".join("%11.6f " % values[k][i] for k in range(5)))
for k in range(5):
    err = np.max(np.abs(values[k] - closed[k]))
    print("order %d: max |graph - closed form| = %.2e" % (k, err))
    assert err < 1e-9          # the Tanh macro carries the reference's 1e-12 epsilon in its denominator
