"""TEST INFRASTRUCTURE ONLY — numpy restatement of bgavran/autodiff's graph + eval path.

This is the ORACLE the CUDA path is checked against.  It restates, in table-driven
form, what the reference does for every node on the hot path: the memoised ``eval``
(reference autodiff/core/node.py:42-46), the numpy bodies of every ``_eval``
(autodiff/core/ops.py, autodiff/core/reshape.py), the graph->graph ``grad`` transform
(autodiff/core/grad.py:9-38) and the MLP builder / optimizers
(autodiff/core/high_level_ops.py).  All arithmetic is the reference's own numpy call
(``np.einsum`` without ``optimize``, ufuncs, ``np.sum``), so results are bit-identical
to the reference on the same numpy build.

Parity pinning: ``tests/golden/make_golden.py`` imports the REAL reference from
/root/reference (in the build container), runs the reference's own test fixtures
(seeds/shapes of tests/test_einsum.py, test_matrix_ops.py, test_complete_graphs.py,
test_scalar_ops.py, the README scalars and the tanh 0..4th-derivative sweep) and
stores the outputs under tests/golden/*.npz; ``tests/test_oracle_golden.py`` checks
this file against those vectors bit-for-bit.  The reference's own tests pin only
finite-difference consistency at rtol 1e-2 (tests/numerical_check.py:66-67), so the
1e-12 bar is pinned by those generated vectors.

Deliberate, documented deviations (host-side only, no arithmetic touched):
  * list-of-slices indices are tuple-ised (reference reshape.py:41-48,133-134 index
    with a *list*, an error on numpy >= 1.23; old numpy treated it as a tuple);
  * ``reverse_topo_sort`` / ``eval`` are iterative so BASELINE-size graphs (44k nodes)
    do not hit the recursion limit; visiting order is identical to the recursive DFS.
"""
import collections
import functools
import numbers
import re
import time
from contextlib import contextmanager
from string import ascii_lowercase

import numpy as np

EPS = 1e-12  # reference node.py:8


# --------------------------------------------------------------------------- graph core
class Node:
    """Restates reference autodiff/core/node.py:7-108."""
    epsilon = EPS
    id = 0
    context_list = []

    def __init__(self, children, name="Node"):
        self.children = [c if isinstance(c, Node) else Variable(c) for c in children]  # node.py:14
        self.name = name
        self.cached = None
        self.shape = None
        self.context_list = Node.context_list.copy()                                   # node.py:19
        self.id = Node.id
        Node.id += 1

    def _eval(self):
        raise NotImplementedError()

    def _partial_derivative(self, wrt, previous_grad):
        raise NotImplementedError()

    def eval(self):
        # node.py:42-46 — memoised; iterative post-order so deep graphs do not recurse.
        if self.cached is not None:
            return self.cached
        stack = [(self, False)]
        while stack:
            node, ready = stack.pop()
            if node.cached is not None:
                continue
            if ready or not _is_builtin(node):
                node.cached = node._eval()
                continue
            stack.append((node, True))
            for ch in reversed(node.children):
                if ch.cached is None:
                    stack.append((ch, False))
        return self.cached

    def partial_derivative(self, wrt, previous_grad):                                  # node.py:48-50
        with add_context(self.name + " PD" + " wrt " + str(wrt)):
            return self._partial_derivative(wrt, previous_grad)

    def __call__(self, *a, **k):
        return self.eval()

    def __str__(self):
        return self.name

    # operator sugar — node.py:62-108
    def __add__(self, o): return Add(self, o)
    def __neg__(self): return Negate(self)
    def __sub__(self, o): return self.__add__(o.__neg__())
    def __rsub__(self, o): return self.__neg__().__add__(o)
    def __mul__(self, o): return Mul(self, o)
    def __matmul__(self, o): return MatMul(self, o)
    def __rmatmul__(self, o): return MatMul(o, self)
    def __imatmul__(self, o): return self.__matmul__(o)
    def __truediv__(self, o): return self.__mul__(Recipr(o))
    def __rtruediv__(self, o): return Recipr(self).__mul__(o)
    def __pow__(self, p, modulo=None): return Pow(self, p)
    __rmul__ = __mul__
    __radd__ = __add__
    def __getitem__(self, item): return Slice(self, item)


def _is_builtin(node):
    """True when the node's _eval only calls child() on its children (safe to pre-evaluate them)."""
    return type(node).__module__ == __name__ and "_eval" not in node.__dict__


class Variable(Node):
    """node.py:111-137."""
    def __init__(self, value, name=None):
        if name is None:
            name = str(value)
        super().__init__([], name)
        if isinstance(value, numbers.Number):
            self._value = np.array(value, dtype=np.float64)                            # node.py:117-118
        else:
            self._value = value
        self.shape = self._value.shape

    @property
    def value(self):
        return self._value

    @value.setter
    def value(self, val):
        self.cached = self._value = val                                                # node.py:127-129

    def _eval(self):
        return self._value

    def _partial_derivative(self, wrt, previous_grad):
        return previous_grad if self == wrt else 0


@contextmanager
def add_context(ctx):                                                                  # node.py:140-146
    Node.context_list.append(ctx + "_" + str(time.time()))
    try:
        yield
    finally:
        del Node.context_list[-1]


def module_wrapper(fn):                                                                # ops.py:11-16
    @functools.wraps(fn)
    def wrapped(*a, **k):
        with add_context(fn.__name__):
            return fn(*a, **k)
    return wrapped


def letters_from_tuple(tpl):                                                           # ops.py:19-20
    return ascii_lowercase[:len(tpl)]


def shape_from_elems(*elems):                                                          # ops.py:23-26
    if len(elems) == 0:
        return 1,
    return np.broadcast_shapes(*[tuple(e.shape) for e in elems])


# --------------------------------------------------------------------------- reshape.py
class ReduceSumKeepDims(Node):                                                         # reshape.py:6-19
    def __init__(self, node, axes):
        super().__init__([node], name="ReduceSumKeepDims")
        self.node = self.children[0]
        self.axes = tuple(axes)
        self.shape = [1 if i in self.axes else s for i, s in enumerate(self.node.shape)]

    def _eval(self):
        return np.sum(self.node(), axis=self.axes, keepdims=True)

    def _partial_derivative(self, wrt, previous_grad):
        return previous_grad * np.ones(self.node.shape) if self.node == wrt else 0


class Concat(Node):                                                                    # reshape.py:22-49
    def __init__(self, a, b, axis=0):
        assert axis >= 0
        super().__init__([a, b], name="Concat")
        self.a, self.b = self.children
        self.axis = axis
        self.shape = list(self.a.shape)
        self.shape[axis] += self.b.shape[axis]

    def _eval(self):
        return np.concatenate((self.a(), self.b()), axis=self.axis)

    def _partial_derivative(self, wrt, previous_grad):
        previous_grad = Reshape(previous_grad, self.shape)
        split = self.a.shape[self.axis]
        sl = [slice(None, None, None) for _ in range(self.axis + 1)]
        if wrt == self.a:
            sl[self.axis] = slice(None, split, None)
            return previous_grad[sl]
        elif wrt == self.b:
            sl[self.axis] = slice(split, None, None)
            return previous_grad[sl]
        return 0


class Reshape(Node):                                                                   # reshape.py:52-77
    def __init__(self, node, shape, name="Reshape"):
        super().__init__([node], name)
        self.node = self.children[0]
        self.shape = self.infer_shape(shape)

    def _eval(self):
        v = self.node()
        if isinstance(v, numbers.Number):
            return np.broadcast_to(v, self.shape)
        return np.reshape(v, self.shape)

    def _partial_derivative(self, wrt, previous_grad):
        return Reshape(previous_grad, self.node.shape) if self.node == wrt else 0

    def infer_shape(self, shape):
        if isinstance(shape, numbers.Number):
            return shape
        if -1 in shape:
            shape = list(shape)
            for i in range(len(shape)):
                if shape[i] == -1:
                    shape[i] = int(-np.prod(self.node.shape) / np.prod(shape))
        return shape


class Slice(Node):                                                                     # reshape.py:82-102
    def __init__(self, node, slice_val, name="Slice"):
        if name is None:
            name = str(slice_val)
        super().__init__([node], name)
        self.node = self.children[0]
        # documented deviation: tuple-ise list indices (old-numpy behaviour)
        self.slice_val = tuple(slice_val) if isinstance(slice_val, list) else slice_val
        self.shape = np.zeros(self.node.shape)[self.slice_val].shape

    def _eval(self):
        return self.node()[self.slice_val]

    def _partial_derivative(self, wrt, previous_grad):
        if self.node == wrt:                       # eager evaluation, reshape.py:96-102
            g = np.zeros(wrt.shape)
            g[self.slice_val] = previous_grad()
            return g
        return 0


class Pad(Node):                                                                       # reshape.py:105-135
    def __init__(self, node, pad_width, constant_values, name="Slice"):
        super().__init__([node], name)
        self.node = self.children[0]
        self.pad_width = pad_width
        self.constant_values = constant_values
        self.shape = np.pad(np.ones(self.node.shape), self.pad_width, mode="constant",
                            constant_values=self.constant_values).shape

    def _eval(self):
        return np.pad(self.node(), self.pad_width, mode="constant", constant_values=self.constant_values)

    def _partial_derivative(self, wrt, previous_grad):
        if self.node == wrt:
            sl = [slice(p[0], s - p[1]) for p, s in zip(self.pad_width, self.shape)]
            return previous_grad[sl]
        return 0


# --------------------------------------------------------------------------- ops.py
@module_wrapper
def ReduceSumToShape(tensor, to_shape):                                                # ops.py:29-41
    if tensor.shape == to_shape:       # list-vs-tuple compare is part of the behaviour
        return tensor
    g_letters = letters_from_tuple(tensor.shape)
    w_letters = "" if len(to_shape) == 0 else g_letters[-len(to_shape):]
    summed = Einsum(str(g_letters) + "->" + str(w_letters), tensor)
    return ReduceSumKeepDims(summed, axes=[i for i, v in enumerate(to_shape) if v == 1])


class Add(Node):                                                                       # ops.py:44-61
    def __init__(self, *elems, name="Add"):
        if not elems:
            name = "0-" + name
        super().__init__(list(elems), name)
        self.shape = shape_from_elems(*self.children)

    def _eval(self):
        return np.array(sum([e() for e in self.children]))

    def _partial_derivative(self, wrt, previous_grad):
        g = previous_grad * Variable(self.children.count(wrt))
        return ReduceSumToShape(g, wrt.shape)


class Mul(Node):                                                                       # ops.py:64-86
    def __init__(self, *elems, name="Mul"):
        if not elems:
            name = "1-" + name
        super().__init__(list(elems), name)
        self.shape = shape_from_elems(*self.children)

    def _eval(self):
        return functools.reduce(lambda x, y: x * y, [c() for c in self.children], 1)

    def _partial_derivative(self, wrt, previous_grad):
        terms = []
        for loc, child in enumerate(self.children):
            if child == wrt:
                terms.append(Mul(*[ch for i, ch in enumerate(self.children) if loc != i]))
        return ReduceSumToShape(previous_grad * Add(*terms), wrt.shape)


class _Unary(Node):
    """Shared shape/child plumbing of the one-input ops (ops.py:89-121,220-233,302-359)."""
    default_name = "Unary"

    def __init__(self, node, name=None):
        super().__init__([node], self.default_name if name is None else name)
        self.node = self.children[0]
        self.shape = self.node.shape

    def _partial_derivative(self, wrt, previous_grad):
        return self._pd(previous_grad) if self.node == wrt else 0


class Negate(_Unary):                                                                  # ops.py:89-102
    default_name = "Negate"
    def _eval(self): return -self.node()
    def _pd(self, g): return -g


class Recipr(_Unary):                                                                  # ops.py:105-121
    default_name = "Reciprocal"
    def _eval(self): return 1 / (self.node() + Node.epsilon)
    def _pd(self, g): return - g * self * self


class ReLU(_Unary):                                                                    # ops.py:220-233
    default_name = "ReLU"
    def _eval(self):
        v = self.node()
        return v * (v > 0)
    def _pd(self, g): return g * self * Recipr(self)


class Log(_Unary):                                                                     # ops.py:302-314
    default_name = "Log"
    def _eval(self): return np.log(self.node() + Node.epsilon)
    def _pd(self, g): return g * Recipr(self.node)


class Identity(_Unary):                                                                # ops.py:317-329
    default_name = "Identity"
    def _eval(self): return self.node()
    def _pd(self, g): return g


class Exp(_Unary):                                                                     # ops.py:332-344
    default_name = "Exp"
    def _eval(self): return np.exp(self.node())
    def _pd(self, g): return g * self


class Sigmoid(_Unary):                                                                 # ops.py:347-359
    default_name = "Sigmoid"
    def _eval(self): return 1 / (1 + np.exp(-self.node()))
    def _pd(self, g): return g * self * (1 - self)


class NormalDistribution(Node):                                                        # ops.py:379-396
    def __init__(self, node, mean=0, variance=1, name="Normal Distribution"):
        super().__init__([node], name=name)
        self.node = self.children[0]
        self.mean = mean
        self.variance = variance
        self.shape = self.node.shape

    def _eval(self):
        x, m, v = self.node(), self.mean, self.variance
        return 1 / np.sqrt(2 * np.pi * (v ** 2)) * np.exp(-(x - m) ** 2 / (2 * v ** 2))

    def _partial_derivative(self, wrt, previous_grad):
        return -previous_grad * self.node * self if self.node == wrt else 0


class Einsum(Node):                                                                    # ops.py:124-217
    def __init__(self, op_str, *operands, name="EinSum"):
        super().__init__(list(operands), name + " " + op_str)
        self.op_str = op_str
        self.operands = self.children
        self.opnames = re.split(",|->", self.op_str)
        self.all_letters = "".join(set("".join(self.opnames[:-1])))
        self.letter_to_dim = {}
        if len(self.operands) + 1 != len(self.opnames):
            raise ValueError("Number of operands doesn't match the einsum string!")
        for op, letters in zip(self.operands, self.opnames[:-1]):
            if len(op.shape) != 0 and len(op.shape) != len(letters) and "..." not in letters and letters != "":
                raise ValueError("Dimension of operand " + str(op) + " doesn't match the string! " +
                                 "Shape: " + str(op.shape) + " , string: '" + letters + "'")
            shp = op.shape
            if letters[:3] == "...":                                                   # ops.py:146-148
                letters = letters[::-1]
                shp = op.shape[::-1]
            for i, lett in enumerate(Einsum.split_dots(letters)):
                try:
                    dim = [shp[i]] if len(lett) == 1 else shp[i:]
                    if self.letter_to_dim.get(lett, dim) != dim:
                        raise ValueError("Inconsistent dimension names!")
                    self.letter_to_dim[lett] = dim
                except IndexError:
                    pass
        shape = []
        for let in Einsum.split_dots(self.opnames[-1]):
            for l in self.letter_to_dim.get(let, [1]):
                shape.append(l)
        self.shape = tuple(shape)

    @staticmethod
    def split_dots(op_str):                                                            # ops.py:167-170
        return re.findall(r"\.{3}|\S", op_str)

    def _eval(self):                                                                   # ops.py:172-180
        arr = [op() for op in self.operands]
        for i, val in enumerate(arr):
            if isinstance(val, numbers.Number):
                shp = [l for let in Einsum.split_dots(self.opnames[i]) for l in self.letter_to_dim.get(let, [1])]
                arr[i] = np.broadcast_to(val, shp)
        return np.einsum(self.op_str, *arr)

    def _partial_derivative(self, wrt, previous_grad):                                 # ops.py:182-213
        order = list(range(len(self.opnames)))
        try:
            loc = self.operands.index(wrt)
        except ValueError:
            return 0
        order[loc], order[-1] = order[-1], order[loc]
        pool = self.operands + [previous_grad]
        operands = [pool[i] for i in order]
        opnames = [self.opnames[i] for i in order]
        for i, letter in enumerate(Einsum.split_dots(self.opnames[loc])):
            if letter not in Einsum.split_dots("".join(opnames[:-1])):
                opnames.insert(0, letter)
                dim = wrt.shape[i]
                operands.insert(0, Variable(np.ones(dim), name="np.ones(" + str(dim) + ")"))
        return Einsum(Einsum.to_einsum_string(opnames), *operands[:-1])

    @staticmethod
    def to_einsum_string(ops):                                                         # ops.py:215-217
        return ",".join(ops[:-1]) + "->" + ops[-1]


class SoftmaxCEWithLogits(Node):                                                       # ops.py:236-262
    def __init__(self, labels, logits, name="SoftmaxCEWithLogits"):
        super().__init__([labels, logits], name=name)
        self.labels, self.logits = self.children
        self.shape = self.logits.shape[:-1]

    def _eval(self):
        lab, z = self.labels(), self.logits()
        lab_sum = np.sum(lab, axis=1)
        if not np.allclose(lab_sum, np.ones_like(lab_sum)):
            raise ValueError("Labels must be a valid probability distribution!")
        mx = np.expand_dims(np.max(z, axis=-1), axis=-1)
        lse = mx + np.expand_dims(np.log(np.sum(np.exp(z - mx), axis=-1)), axis=-1)
        return -np.sum(lab * z - lab * lse, axis=-1)

    def _partial_derivative(self, wrt, previous_grad):
        if wrt == self.logits:
            return Einsum("...i,...ij->...ij", previous_grad, Softmax(self.logits) - self.labels)
        elif wrt == self.labels:
            return Variable(0)
        return 0


class SigmoidCEWithLogits(Node):                                                       # ops.py:265-279
    def __init__(self, labels, logits, name="SigmoidCEWithLogits"):
        super().__init__([labels, logits], name)
        self.labels, self.logits = self.children
        self.shape = self.logits.shape

    def _eval(self):
        z, x = self.labels(), self.logits()
        return np.maximum(x, 0) - x * z + np.log(1 + np.exp(-abs(x)))

    def _partial_derivative(self, wrt, previous_grad):
        if wrt == self.logits:
            return Einsum("...ij,...ij->...ij", previous_grad, Sigmoid(self.logits) - self.labels)
        return 0


class Pow(Node):                                                                       # ops.py:282-299
    def __init__(self, first, second, name="Pow"):
        super().__init__([first, second], name)
        self.first, self.second = self.children
        self.shape = shape_from_elems(*self.children)

    def _eval(self):
        return np.power(self.first(), self.second())

    def _partial_derivative(self, wrt, previous_grad):
        if self.first == self.second == wrt:
            return previous_grad * self * (Log(self.first) + 1)
        elif self.first == wrt:
            return previous_grad * self.second * Pow(self.first, self.second - 1)
        elif self.second == wrt:
            return previous_grad * Log(self.first) * self
        return 0


class FrobeniusNorm(Node):                                                             # ops.py:362-376
    def __init__(self, *nodes, name="Frobenius Norm"):
        super().__init__(list(nodes), name=name)
        self.nodes = self.children
        self.shape = 1,

    def _eval(self):
        return np.sqrt(sum([np.sum(np.square(n())) for n in self.nodes]))

    def _partial_derivative(self, wrt, previous_grad):
        try:
            loc = self.nodes.index(wrt)
        except ValueError:
            return 0
        return previous_grad * self.nodes[loc] / self


@module_wrapper
def Tanh(x):                                                                           # ops.py:399-402
    val = Exp(-2 * x)
    return (1 - val) / (1 + val)


@module_wrapper
def SquaredDifference(x, y):                                                           # ops.py:405-408
    diff = x - y
    return diff * diff


@module_wrapper
def MatMul(x, y):                                                                      # ops.py:411-413
    return Einsum("ij,jk->ik", x, y)


@module_wrapper
def Transpose(x):                                                                      # ops.py:416-418
    return Einsum("ij->ji", x)


@module_wrapper
def Softmax(x):                                                                        # ops.py:421-434
    exp = Exp(x)
    if len(x.shape) == 1:
        return exp / Einsum("i->", exp)
    elif len(x.shape) == 2:
        return exp / Einsum("bi,o->bo", exp, np.array([1]))
    elif len(x.shape) == 3:
        return exp / Einsum("abi,o->abo", exp, np.array([1]))
    elif len(x.shape) == 4:
        return exp / Einsum("abci,o->abco", exp, np.array([1]))
    else:
        raise ValueError("5D tensors not yet supported")


# --------------------------------------------------------------------------- grad.py / utils.py / wrappers.py
def reverse_topo_sort(top_node):
    """utils.py:1-17 — same order as the recursive DFS post-order, reversed; iterative."""
    order, visited = [], set()
    stack = [(top_node, iter(top_node.children))]
    visited.add(top_node)
    while stack:
        node, it = stack[-1]
        for ch in it:
            if ch not in visited:
                visited.add(ch)
                stack.append((ch, iter(ch.children)))
                break
        else:
            order.append(node)
            stack.pop()
    return reversed(order)


def grad(top_node, wrt_list, previous_grad=None):                                      # grad.py:9-38
    assert isinstance(wrt_list, list) or isinstance(wrt_list, tuple)
    if previous_grad is None:
        previous_grad = Variable(np.ones(top_node.shape), name="'" + top_node.name + "' grad_sum")
    dct = collections.defaultdict(lambda: Variable(0))
    dct[top_node] += previous_grad
    for node in reverse_topo_sort(top_node):
        for child in set(node.children):
            dct[child] += node.partial_derivative(wrt=child, previous_grad=dct[node])
    return [dct[w] for w in wrt_list]


def checkpoint(fn):                                                                    # wrappers.py:5-15
    def wrap_in_primitive(*fn_args):
        op = Node(children=fn_args, name=fn.__name__)
        op._eval = lambda: fn(*fn_args)()
        op._partial_derivative = lambda wrt, previous_grad: grad(fn(*fn_args), [wrt], previous_grad=previous_grad)[0]
        return op
    return wrap_in_primitive


# --------------------------------------------------------------------------- high_level_ops.py
class Module:                                                                          # high_level_ops.py:6-15
    def forward(self, *a, **k):
        with add_context(type(self).__name__):
            return self._forward(*a, **k)

    def _forward(self, *a, **k):
        raise NotImplementedError()

    def __call__(self, *a, **k):
        return self.forward(*a, **k)


class NN(Module):                                                                      # high_level_ops.py:18-38
    def __init__(self, sizes):
        if len(sizes) < 2:
            raise ValueError("Sizes represent the layer sizes and needs to have at least two values for one weight"
                             " matrix to exist!")
        self.sizes = sizes
        self.w = []
        for i, (inp, out) in enumerate(zip(self.sizes, self.sizes[1:])):
            self.w.append(Variable(np.random.randn(inp + 1, out) * 0.01, name="w" + str(i)))

    def _forward(self, x):
        bias = Variable(np.ones((x.shape[0], 1)), name="bias")
        for i, w in enumerate(self.w):
            x = Concat(x, bias, axis=1)
            x = x @ w
            if i < len(self.w) - 1:
                x = ReLU(x)
        return x


class Optimizer(Module):                                                               # high_level_ops.py:41-48
    @staticmethod
    def apply_new_weights(weight_list, new_weight_values):
        for w, v in zip(weight_list, new_weight_values):
            w.value = v


class SGDOptimizer(Optimizer):                                                         # high_level_ops.py:51-60
    def __init__(self, lr=0.001):
        self.lr = lr

    def _forward(self, params_list, grads_list):
        return [p - self.lr * g for p, g in zip(params_list, grads_list)]


class Momentum(Optimizer):                                                             # high_level_ops.py:64-82
    def __init__(self, num_params, lr=0.001, gamma=0.8):
        self.lr, self.gamma = lr, gamma
        self.v = [0 for _ in range(num_params)]

    def _forward(self, params_list, grads_list):
        for i, (p, g) in enumerate(zip(params_list, grads_list)):
            new_m = self.gamma * self.v[i] + g
            params_list[i], self.v[i] = p - self.lr * new_m, new_m
        return params_list


class Adagrad(Optimizer):                                                              # high_level_ops.py:85-105
    def __init__(self, num_params, lr=0.01):
        self.lr = lr
        self.run_avg = [0 for _ in range(num_params)]

    def _forward(self, params_list, grads_list):
        for i, (p, g) in enumerate(zip(params_list, grads_list)):
            new_avg = self.run_avg[i] + g ** 2
            params_list[i], self.run_avg[i] = p - self.lr / np.sqrt(new_avg + 1e-8) * g, new_avg
        return params_list


class Adam(Optimizer):                                                                 # high_level_ops.py:108-131
    eps = 1e-8

    def __init__(self, num_params, lr=0.001, b1=0.9, b2=0.999):
        self.lr, self.b1, self.b2 = lr, b1, b2
        self.num_params = num_params
        self.m = [0 for _ in range(num_params)]
        self.v = [0 for _ in range(num_params)]
        self.step = 0

    def _forward(self, params_list, grads_list):
        for i, (p, g) in enumerate(zip(params_list, grads_list)):
            new_m = self.b1 * self.m[i] + (1 - self.b1) * g
            new_v = self.b2 * self.v[i] + (1 - self.b2) * g ** 2
            a = self.lr * np.sqrt(1 - self.b2 ** self.step) / (1 - self.b1 ** self.step + Adam.eps)
            params_list[i], self.m[i], self.v[i] = p - a * new_m / (np.sqrt(new_v) + Adam.eps), new_m, new_v
        self.step += 1
        return params_list


# --------------------------------------------------------------------------- convolution
def conv_valid_nhwc(x, w):
    """VALID, stride-1 cross-correlation, NHWC input and HWIO filter, as decoded from the
    reference's commented-out design (reshape.py:137-160, high_level_ops.py:152-165,
    tests/test_convolution.py:17-33): windows via as_strided, contraction via np.einsum."""
    from numpy.lib.stride_tricks import as_strided
    n, h, wd, c = x.shape
    kh, kw, ci, co = w.shape
    oh, ow = h - kh + 1, wd - kw + 1
    s = x.strides
    win = as_strided(x, shape=(n, oh, ow, kh, kw, c), strides=(s[0], s[1], s[2], s[1], s[2], s[3]))
    return np.einsum("nhwabc,abco->nhwo", win, w)


class Im2Col(Node):
    """Window gather as the reference's AsStrided design would do it (reshape.py:137-160), then flattened."""
    def __init__(self, node, kh, kw, name="Im2Col"):
        super().__init__([node], name)
        self.node = self.children[0]
        self.kh, self.kw = kh, kw
        n, h, w, c = self.node.shape
        self.shape = (n * (h - kh + 1) * (w - kw + 1), kh * kw * c)

    def _eval(self):
        from numpy.lib.stride_tricks import as_strided
        x = np.ascontiguousarray(self.node())
        n, h, w, c = x.shape
        oh, ow = h - self.kh + 1, w - self.kw + 1
        s = x.strides
        win = as_strided(x, shape=(n, oh, ow, self.kh, self.kw, c), strides=(s[0], s[1], s[2], s[1], s[2], s[3]))
        return win.reshape(self.shape)

    def _partial_derivative(self, wrt, previous_grad):
        return Col2Im(previous_grad, tuple(self.node.shape), self.kh, self.kw) if self.node == wrt else 0


class Col2Im(Node):
    def __init__(self, node, image_shape, kh, kw, name="Col2Im"):
        super().__init__([node], name)
        self.node = self.children[0]
        self.kh, self.kw = kh, kw
        self.shape = tuple(image_shape)

    def _eval(self):
        n, h, w, c = self.shape
        oh, ow = h - self.kh + 1, w - self.kw + 1
        g = np.broadcast_to(self.node(), (n * oh * ow, self.kh * self.kw * c)).reshape(n, oh, ow, self.kh, self.kw, c)
        out = np.zeros(self.shape)
        for a in range(self.kh):
            for b in range(self.kw):
                out[:, a:a + oh, b:b + ow, :] += g[:, :, :, a, b, :]
        return out

    def _partial_derivative(self, wrt, previous_grad):
        return Im2Col(previous_grad, self.kh, self.kw) if self.node == wrt else 0


@module_wrapper
def Conv2D(x, w):
    kh, kw, c, o = w.shape
    n, h, wd, _ = x.shape
    out = Einsum("pk,ko->po", Im2Col(x, kh, kw), Reshape(w, (kh * kw * c, o)))
    return Reshape(out, (n, h - kh + 1, wd - kw + 1, o))
