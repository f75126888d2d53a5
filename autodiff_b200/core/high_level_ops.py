"""Modules and optimizers — API of reference autodiff/core/high_level_ops.py.

`NN` builds the same graph as the reference (bias folded into the weights through a concatenated ones column).
The optimizers keep the reference's call convention (`new = opt(param_values, grad_values)`; `apply_new_weights`)
and its arithmetic, quirks included (Adam's step counter starts at 0, so the first update is a no-op,
high_level_ops.py:118-131), but the update itself is ONE fused multi-output device program per parameter;
moments never leave HBM.  Values may be passed as host ndarrays (as the reference's examples do), as
`DeviceArray`s, or directly as graph nodes (then nothing crosses PCIe).
"""
import math

import numpy as np

from .. import _lib
from ..device import DeviceArray, to_device, to_host
from .engine import run_program
from .node import ConstFill, Node, Variable, add_context
from .ops import *      # noqa: F401,F403  (the reference re-exports everything from here)
from .reshape import *  # noqa: F401,F403

IN, IMM, TMP = _lib.SRC_IN, _lib.SRC_IMM, _lib.SRC_TMP


class Module:
    def forward(self, *args, **kwargs):
        with add_context(type(self).__name__):
            return self._forward(*args, **kwargs)

    def _forward(self, *args, **kwargs):
        raise NotImplementedError()

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)


class NN(Module):
    def __init__(self, sizes):
        if len(sizes) < 2:
            raise ValueError("Sizes represent the layer sizes and needs to have at least two values for one weight"
                             " matrix to exist!")
        self.sizes = sizes
        self.w = []
        for i, (inp, out) in enumerate(zip(self.sizes, self.sizes[1:])):
            w_initial = np.random.randn(inp + 1, out) * 0.01
            self.w.append(Variable(w_initial, name="w" + str(i)))

    def _forward(self, x):
        first_dim = x.shape[0]
        bias = ConstFill((first_dim, 1), 1.0, name="bias")     # reference: Variable(np.ones((first_dim, 1)))
        for i, w in enumerate(self.w):
            x = Concat(x, bias, axis=1)
            x = x @ w
            if i < len(self.w) - 1:
                x = ReLU(x)
        return x


# ------------------------------------------------------------------------------------------------ optimizers
def _as_device(v):
    """(DeviceArray, kind) where kind remembers how to hand the result back."""
    if isinstance(v, Node):
        return v.eval_device(), "device"
    if isinstance(v, DeviceArray):
        return v, "device"
    return to_device(np.asarray(v)), "host"


def _give_back(d, kind):
    return d if kind == "device" else to_host(d)


class Optimizer(Module):
    @staticmethod
    def apply_new_weights(weight_list, new_weight_values):
        for w, new_w_value in zip(weight_list, new_weight_values):
            w.value = new_w_value

    def _forward(self, *args, **kwargs):
        raise NotImplementedError()


class SGDOptimizer(Optimizer):
    def __init__(self, lr=0.001):
        self.lr = lr

    def _forward(self, params_list, grads_list):
        return [self.fn(param, grad) for param, grad in zip(params_list, grads_list)]

    def fn(self, param, grad):
        # param - lr * grad
        p, kind = _as_device(param)
        g, _ = _as_device(grad)
        code = [(_lib.EW_LOAD | IN, 1, 0.0), (_lib.EW_MUL | IMM, 0, float(self.lr)), (_lib.EW_NEG, 0, 0.0),
                (_lib.EW_ADD | IN, 0, 0.0), (_lib.EW_STORE_OUT, 0, 0.0)]
        return _give_back(run_program(code, [p, g], np.broadcast_shapes(p.shape, g.shape)), kind)


class Momentum(Optimizer):
    """Accumulates gradient change through time steps (reference high_level_ops.py:64-82)."""

    def __init__(self, num_params, lr=0.001, gamma=0.8):
        self.lr = lr
        self.gamma = gamma
        self.v = [0 for _ in range(num_params)]

    def _forward(self, params_list, grads_list):
        for i, (param, grad) in enumerate(zip(params_list, grads_list)):
            params_list[i], self.v[i] = self.fn(self.v[i], param, grad)
        return params_list

    def fn(self, prev_m, param, grad):
        # new_m = gamma * prev_m + grad ; param - lr * new_m
        p, kind = _as_device(param)
        g, _ = _as_device(grad)
        m = prev_m if isinstance(prev_m, DeviceArray) else DeviceArray.full(g.shape, float(prev_m))
        code = [(_lib.EW_LOAD | IN, 2, 0.0), (_lib.EW_MUL | IMM, 0, float(self.gamma)), (_lib.EW_ADD | IN, 1, 0.0),
                (_lib.EW_STORE_OUT, 1, 0.0),
                (_lib.EW_MUL | IMM, 0, float(self.lr)), (_lib.EW_NEG, 0, 0.0), (_lib.EW_ADD | IN, 0, 0.0), (_lib.EW_STORE_OUT, 0, 0.0)]
        new_p, new_m = run_program(code, [p, g, m], np.broadcast_shapes(p.shape, g.shape), n_out=2)
        return _give_back(new_p, kind), new_m


class Adagrad(Optimizer):
    """Per-parameter learning rates from the running sum of squared gradients (reference high_level_ops.py:85-105)."""

    def __init__(self, num_params, lr=0.01):
        self.lr = lr
        self.run_avg = [0 for _ in range(num_params)]

    def _forward(self, params_list, grads_list):
        for i, (param, grad) in enumerate(zip(params_list, grads_list)):
            params_list[i], self.run_avg[i] = self.fn(self.run_avg[i], param, grad)
        return params_list

    def fn(self, prev_avg, param, grad):
        # new_avg = prev_avg + grad**2 ; param - lr / sqrt(new_avg + 1e-8) * grad
        p, kind = _as_device(param)
        g, _ = _as_device(grad)
        a = prev_avg if isinstance(prev_avg, DeviceArray) else DeviceArray.full(g.shape, float(prev_avg))
        code = [(_lib.EW_LOAD | IN, 1, 0.0), (_lib.EW_SQUARE, 0, 0.0), (_lib.EW_STORE_T, 0, 0.0),
                (_lib.EW_LOAD | IN, 2, 0.0), (_lib.EW_ADD | TMP, 0, 0.0), (_lib.EW_STORE_OUT, 1, 0.0),
                (_lib.EW_ADD | IMM, 0, 1e-8), (_lib.EW_SQRT, 0, 0.0), (_lib.EW_STORE_T, 0, 0.0),
                (_lib.EW_LOAD | IMM, 0, float(self.lr)), (_lib.EW_DIV | TMP, 0, 0.0), (_lib.EW_MUL | IN, 1, 0.0),
                (_lib.EW_NEG, 0, 0.0), (_lib.EW_ADD | IN, 0, 0.0), (_lib.EW_STORE_OUT, 0, 0.0)]
        new_p, new_avg = run_program(code, [p, g, a], np.broadcast_shapes(p.shape, g.shape), n_out=2)
        return _give_back(new_p, kind), new_avg


class Adam(Optimizer):
    eps = 1e-8

    def __init__(self, num_params, lr=0.001, b1=0.9, b2=0.999):
        self.lr = lr
        self.b1 = b1
        self.b2 = b2
        self.num_params = num_params
        self.m = [0 for _ in range(num_params)]
        self.v = [0 for _ in range(num_params)]
        self.step = 0

    def _forward(self, params_list, grads_list):
        params_list = list(params_list)
        for i, (param, grad) in enumerate(zip(params_list, grads_list)):
            params_list[i], self.m[i], self.v[i] = self.fn(self.step, self.m[i], self.v[i], param, grad)
        self.step += 1
        return params_list

    def fn(self, step, prev_m, prev_v, param, grad):
        # new_m = b1*prev_m + (1-b1)*grad ; new_v = b2*prev_v + (1-b2)*grad**2
        # a = lr*sqrt(1-b2**step)/(1-b1**step+eps) ; param - a*new_m/(sqrt(new_v)+eps)     (high_level_ops.py:126-131)
        p, kind = _as_device(param)
        g, _ = _as_device(grad)
        m = prev_m if isinstance(prev_m, DeviceArray) else DeviceArray.full(g.shape, float(prev_m))
        v = prev_v if isinstance(prev_v, DeviceArray) else DeviceArray.full(g.shape, float(prev_v))
        a = self.lr * math.sqrt(1 - self.b2 ** step) / (1 - self.b1 ** step + Adam.eps)
        code = [
            (_lib.EW_LOAD | IN, 1, 0.0), (_lib.EW_MUL | IMM, 0, float(1 - self.b1)), (_lib.EW_STORE_T, 0, 0.0),
            (_lib.EW_LOAD | IN, 2, 0.0), (_lib.EW_MUL | IMM, 0, float(self.b1)), (_lib.EW_ADD | TMP, 0, 0.0),
            (_lib.EW_STORE_T, 1, 0.0), (_lib.EW_STORE_OUT, 1, 0.0),                                          # new_m
            (_lib.EW_LOAD | IN, 1, 0.0), (_lib.EW_SQUARE, 0, 0.0), (_lib.EW_MUL | IMM, 0, float(1 - self.b2)), (_lib.EW_STORE_T, 0, 0.0),
            (_lib.EW_LOAD | IN, 3, 0.0), (_lib.EW_MUL | IMM, 0, float(self.b2)), (_lib.EW_ADD | TMP, 0, 0.0),
            (_lib.EW_STORE_OUT, 2, 0.0),                                                                      # new_v
            (_lib.EW_SQRT, 0, 0.0), (_lib.EW_ADD | IMM, 0, float(Adam.eps)), (_lib.EW_STORE_T, 0, 0.0),
            (_lib.EW_LOAD | TMP, 1, 0.0), (_lib.EW_MUL | IMM, 0, float(a)), (_lib.EW_DIV | TMP, 0, 0.0),
            (_lib.EW_NEG, 0, 0.0), (_lib.EW_ADD | IN, 0, 0.0), (_lib.EW_STORE_OUT, 0, 0.0),                    # new_param
        ]
        new_p, new_m, new_v = run_program(code, [p, g, m, v], np.broadcast_shapes(p.shape, g.shape), n_out=3)
        return _give_back(new_p, kind), new_m, new_v


class NesterovMomentum(Optimizer):
    """Unfinished in the reference too (no `fn`, high_level_ops.py:134-150): calling it raises AttributeError."""

    def __init__(self, num_params, lr=0.001, gamma=0.8):
        self.lr = lr
        self.gamma = gamma
        self.v = [0 for _ in range(num_params)]

    def _forward(self, params_list, grads_list):
        for i, (param, grad) in enumerate(zip(params_list, grads_list)):
            params_list[i], self.v[i] = self.fn(self.v[i], param, grad)
        return params_list
