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
    """Callable graph builder: `module(...)` runs `_forward` inside a name scope carrying the class name
    (reference high_level_ops.py:6-15; `forward` is the same entry point)."""

    def __call__(self, *args, **kwargs):
        scope = add_context(type(self).__name__)
        scope.__enter__()
        try:
            return self._forward(*args, **kwargs)
        finally:
            scope.__exit__(None, None, None)

    forward = __call__

    def _forward(self, *args, **kwargs):
        raise NotImplementedError("%s does not define _forward" % type(self).__name__)


_NN_SIZES_ERROR = ("Sizes represent the layer sizes and needs to have at least two values for one weight"
                   " matrix to exist!")


class NN(Module):
    """MLP of reference high_level_ops.py:18-38: layer i owns ONE (in_i + 1, out_i) matrix drawn as randn * 0.01 whose
    last row is the bias - the input gets a column of ones concatenated before every matmul - and every layer but the
    last is followed by ReLU.  The ones column is a `ConstFill` (never materialised) instead of np.ones((batch, 1))."""

    def __init__(self, sizes):
        if len(sizes) < 2:
            raise ValueError(_NN_SIZES_ERROR)
        self.sizes = sizes
        fan = list(zip(sizes[:-1], sizes[1:]))
        self.w = [Variable(np.random.randn(n_in + 1, n_out) * 0.01, name="w%d" % k) for k, (n_in, n_out) in enumerate(fan)]

    def _forward(self, x):
        ones_col = ConstFill((x.shape[0], 1), 1.0, name="bias")
        last = len(self.w) - 1
        for k, weights in enumerate(self.w):
            x = Concat(x, ones_col, axis=1) @ weights
            if k != last:
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
    """Call convention of reference high_level_ops.py:41-131: `new_params = opt(param_values, grad_values)` followed by
    `opt.apply_new_weights(variables, new_params)`.  Stateful optimizers keep one entry per parameter in the lists named
    by `STATE`; `fn` takes (leading args..., state..., param, grad) and returns (new_param, new_state...).  As in the
    reference the stateful ones write the new values into the list they were given and return that same list."""
    STATE = ()

    @staticmethod
    def apply_new_weights(weight_list, new_weight_values):
        for variable, fresh in zip(weight_list, new_weight_values):
            variable.value = fresh

    def _reset_state(self, num_params):
        for name in self.STATE:
            setattr(self, name, [0] * num_params)

    def _leading(self):
        return ()

    def _stepped(self):
        pass

    def update_one(self, k, param, grad):
        """The update of parameter k alone (same arithmetic as one iteration of `_forward`): lets a training loop queue
        each parameter's update as soon as its gradient exists (autodiff_b200.dp overlaps it with the rest of the backward
        pass).  Call `end_step()` once after the last parameter of a step."""
        slots = [getattr(self, name) for name in self.STATE]
        out = self.fn(*self._leading(), *[slot[k] for slot in slots], param, grad)
        if not self.STATE:
            return out
        for slot, value in zip(slots, out[1:]):
            slot[k] = value
        return out[0]

    def end_step(self):
        self._stepped()

    def _forward(self, params_list, grads_list):
        if not isinstance(params_list, list):
            params_list = list(params_list)
        slots = [getattr(self, name) for name in self.STATE]
        for k in range(min(len(params_list), len(grads_list))):
            out = self.fn(*self._leading(), *[slot[k] for slot in slots], params_list[k], grads_list[k])
            params_list[k] = out[0]
            for slot, value in zip(slots, out[1:]):
                slot[k] = value
        self._stepped()
        return params_list


def _moment(prev, like):
    """Optimizer state starts as the python number 0 (reference) and lives in HBM from the first update on."""
    return prev if isinstance(prev, DeviceArray) else DeviceArray.full(like.shape, float(prev))


class SGDOptimizer(Optimizer):
    def __init__(self, lr=0.001):
        self.lr = lr

    def _forward(self, params_list, grads_list):
        return [self.fn(p, g) for p, g in zip(params_list, grads_list)]          # (a new list, as in the reference)

    def fn(self, param, grad):
        # param - lr * grad
        p, kind = _as_device(param)
        g, _ = _as_device(grad)
        code = [(_lib.EW_LOAD | IN, 1, 0.0), (_lib.EW_MUL | IMM, 0, float(self.lr)), (_lib.EW_NEG, 0, 0.0),
                (_lib.EW_ADD | IN, 0, 0.0), (_lib.EW_STORE_OUT, 0, 0.0)]
        return _give_back(run_program(code, [p, g], np.broadcast_shapes(p.shape, g.shape)), kind)


class Momentum(Optimizer):
    """Accumulates gradient change through time steps (reference high_level_ops.py:64-82)."""
    STATE = ("v",)

    def __init__(self, num_params, lr=0.001, gamma=0.8):
        self.lr, self.gamma = lr, gamma
        self._reset_state(num_params)

    def fn(self, prev_m, param, grad):
        # new_m = gamma * prev_m + grad ; param - lr * new_m
        p, kind = _as_device(param)
        g, _ = _as_device(grad)
        m = _moment(prev_m, g)
        code = [(_lib.EW_LOAD | IN, 2, 0.0), (_lib.EW_MUL | IMM, 0, float(self.gamma)), (_lib.EW_ADD | IN, 1, 0.0),
                (_lib.EW_STORE_OUT, 1, 0.0),
                (_lib.EW_MUL | IMM, 0, float(self.lr)), (_lib.EW_NEG, 0, 0.0), (_lib.EW_ADD | IN, 0, 0.0), (_lib.EW_STORE_OUT, 0, 0.0)]
        new_p, new_m = run_program(code, [p, g, m], np.broadcast_shapes(p.shape, g.shape), n_out=2)
        return _give_back(new_p, kind), new_m


class Adagrad(Optimizer):
    """Per-parameter learning rates from the running sum of squared gradients (reference high_level_ops.py:85-105)."""
    STATE = ("run_avg",)

    def __init__(self, num_params, lr=0.01):
        self.lr = lr
        self._reset_state(num_params)

    def fn(self, prev_avg, param, grad):
        # new_avg = prev_avg + grad**2 ; param - lr / sqrt(new_avg + 1e-8) * grad
        p, kind = _as_device(param)
        g, _ = _as_device(grad)
        a = _moment(prev_avg, g)
        code = [(_lib.EW_LOAD | IN, 1, 0.0), (_lib.EW_SQUARE, 0, 0.0), (_lib.EW_STORE_T, 0, 0.0),
                (_lib.EW_LOAD | IN, 2, 0.0), (_lib.EW_ADD | TMP, 0, 0.0), (_lib.EW_STORE_OUT, 1, 0.0),
                (_lib.EW_ADD | IMM, 0, 1e-8), (_lib.EW_SQRT, 0, 0.0), (_lib.EW_STORE_T, 0, 0.0),
                (_lib.EW_LOAD | IMM, 0, float(self.lr)), (_lib.EW_DIV | TMP, 0, 0.0), (_lib.EW_MUL | IN, 1, 0.0),
                (_lib.EW_NEG, 0, 0.0), (_lib.EW_ADD | IN, 0, 0.0), (_lib.EW_STORE_OUT, 0, 0.0)]
        new_p, new_avg = run_program(code, [p, g, a], np.broadcast_shapes(p.shape, g.shape), n_out=2)
        return _give_back(new_p, kind), new_avg


class Adam(Optimizer):
    """Reference high_level_ops.py:108-131, quirk included: `step` starts at 0, so the first update is a no-op."""
    eps = 1e-8
    STATE = ("m", "v")

    def __init__(self, num_params, lr=0.001, b1=0.9, b2=0.999):
        self.lr, self.b1, self.b2 = lr, b1, b2
        self.num_params = num_params
        self.step = 0
        self._reset_state(num_params)

    def _leading(self):
        return (self.step,)

    def _stepped(self):
        self.step += 1

    def fn(self, step, prev_m, prev_v, param, grad):
        # new_m = b1*prev_m + (1-b1)*grad ; new_v = b2*prev_v + (1-b2)*grad**2
        # a = lr*sqrt(1-b2**step)/(1-b1**step+eps) ; param - a*new_m/(sqrt(new_v)+eps)     (high_level_ops.py:126-131)
        p, kind = _as_device(param)
        g, _ = _as_device(grad)
        m, v = _moment(prev_m, g), _moment(prev_v, g)
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
    """Unfinished in the reference too (it defines no `fn`, high_level_ops.py:134-150): calling it raises AttributeError."""
    STATE = ("v",)

    def __init__(self, num_params, lr=0.001, gamma=0.8):
        self.lr, self.gamma = lr, gamma
        self._reset_state(num_params)
