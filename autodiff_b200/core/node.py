"""Graph core: `Node`, `Variable`, `add_context` — the API of reference autodiff/core/node.py:7-146, unchanged,
with node evaluation moved to the GPU.

What stays exactly as in the reference: children wrapping (node.py:14), the global id counter and name-scope
snapshot (node.py:9-10,19-21), memoised `eval()`/`__call__` returning a HOST numpy value (node.py:42-46,56-57),
operator overloads (node.py:62-108), `Variable.value` semantics (node.py:123-129) and the fact that tests may
write `node.cached` directly (reference tests/numerical_check.py:5-8,19-27).

What changes: a node's value lives in `_dev` (a `DeviceArray` in HBM); `cached` is the lazily downloaded host
copy.  Assigning `cached` (or `Variable.value`) drops the device copy so the next evaluation re-uploads, which
keeps external in-place edits of a host array coherent.  Evaluation is an iterative post-order walk (no
recursion limit, reference recursion at node.py:42-46 fails beyond ~1000 deep graphs).
"""
import numbers
import time
from contextlib import contextmanager

import numpy as np

from ..device import DeviceArray, to_device, to_host


_ops = None        # autodiff_b200.core.ops / .reshape / .engine, bound on first use (they import this module, so not at import time)
_reshape = None
_engine = None


def _mods():
    global _ops, _reshape, _engine
    from . import engine, ops, reshape
    _ops, _reshape, _engine = ops, reshape, engine


class Node:
    epsilon = 1e-12          # reference node.py:8
    id = 0
    context_list = []
    _scalar_as_array = False  # 0-d results come back as np.float64 (numpy scalar) unless the op says ndarray

    def __init__(self, children, name="Node"):
        self.children = [child if isinstance(child, Node) else Variable(child) for child in children]
        self.name = name
        self._host = None
        self._dev = None
        self._dev_override = False
        self.shape = None
        self.context_list = Node.context_list.copy()
        self.id = Node.id
        Node.id += 1

    # ---- value slots ------------------------------------------------------------------------------
    @property
    def cached(self):
        if self._host is None and self._dev is not None:
            self._host = self._download(self._dev)
        return self._host

    @cached.setter
    def cached(self, value):
        if isinstance(value, DeviceArray):
            self._dev, self._host = value, None
            self._dev_override = True       # the device value no longer mirrors a host-known leaf value
        else:
            self._host, self._dev = value, None
            self._dev_override = False

    def _download(self, dev):
        host = to_host(dev)
        if _engine is None:
            _mods()
        _engine.check_pending_flags()
        if host.ndim == 0 and not self._scalar_as_array:
            return host[()]          # numpy scalar, like the reference's ufunc results on 0-d input
        return host

    # ---- protocol ---------------------------------------------------------------------------------
    def _eval(self):
        """Built-in ops evaluate on the device; user subclasses may override this with numpy code
        (reference README: "easy to add your own operations") and their result is uploaded on demand."""
        dev = self._eval_device()
        return self._download(dev)

    def _eval_device(self):
        raise NotImplementedError()

    def _partial_derivative(self, wrt, previous_grad):
        raise NotImplementedError()

    def eval(self):
        if self._host is None:
            if self._dev is None:
                if _engine is None:
                    _mods()
                _engine.evaluate(self)
            if self._host is None:
                self._host = self._download(self._dev)
        return self._host

    def eval_device(self):
        """Evaluates (memoised) and returns the device-resident value without any device->host copy."""
        if self._dev is None:
            if _engine is None:
                _mods()
            _engine.evaluate(self)
            if self._dev is None:           # foreign node that produced a host value
                self._dev = to_device(np.asarray(self._host))
        return self._dev

    def partial_derivative(self, wrt, previous_grad):
        with add_context(self.name + " PD" + " wrt " + str(wrt)):
            return self._partial_derivative(wrt, previous_grad)

    def plot_comp_graph(self, view=True, name="comp_graph"):
        raise NotImplementedError("graph drawing (reference visualization/graph_visualization.py) is out of scope "
                                  "of the B200 backend; node.context_list is kept so an external plotter can cluster by scope")

    def __call__(self, *args, **kwargs):
        return self.eval()

    def __str__(self):
        return self.name

    def __add__(self, other):
        if _ops is None:
            _mods()
        return _ops.Add(self, other)

    def __neg__(self):
        if _ops is None:
            _mods()
        return _ops.Negate(self)

    def __sub__(self, other):
        return self.__add__(other.__neg__())

    def __rsub__(self, other):
        return self.__neg__().__add__(other)

    def __mul__(self, other):
        if _ops is None:
            _mods()
        return _ops.Mul(self, other)

    def __matmul__(self, other):
        if _ops is None:
            _mods()
        return _ops.MatMul(self, other)

    def __rmatmul__(self, other):
        if _ops is None:
            _mods()
        return _ops.MatMul(other, self)

    def __imatmul__(self, other):
        return self.__matmul__(other)

    def __truediv__(self, other):
        if _ops is None:
            _mods()
        return self.__mul__(_ops.Recipr(other))

    def __rtruediv__(self, other):
        if _ops is None:
            _mods()
        return _ops.Recipr(self).__mul__(other)

    def __pow__(self, power, modulo=None):
        if _ops is None:
            _mods()
        return _ops.Pow(self, power)

    __rmul__ = __mul__
    __radd__ = __add__

    def __getitem__(self, item):
        if _reshape is None:
            _mods()
        return _reshape.Slice(self, item)


class Variable(Node):
    """Leaf holding a value (reference node.py:111-137).  The value may be a host ndarray (uploaded on first
    use, re-uploaded whenever `cached`/`value` is reassigned) or an already device-resident `DeviceArray`."""
    _scalar_as_array = True

    def __init__(self, value, name=None):
        # The name defaults to str(value) (reference node.py:113-114); it is computed on demand - str() of an array is
        # slow and a second-order char-rnn graph creates thousands of leaves.  Node.__init__ is inlined (no children).
        self._lazy_name = value if name is None else None
        self._name = name
        self.children = []
        self._host = None
        self._dev = None
        self._dev_override = False
        self.context_list = Node.context_list.copy()
        self.id = Node.id
        Node.id += 1
        tv = type(value)
        if tv is float or tv is int or (tv is not np.ndarray and tv is not DeviceArray and isinstance(value, numbers.Number)):
            self._value = np.array(value, dtype=np.float64)
            self.shape = ()
        else:
            self._value = value
            self.shape = tuple(value.shape)

    @property
    def name(self):
        if self._name is None:
            v = self._lazy_name
            self._name = "DeviceArray%s" % (v.shape,) if isinstance(v, DeviceArray) else str(v)
        return self._name

    @name.setter
    def name(self, n):
        self._name = n

    @property
    def value(self):
        if isinstance(self._value, DeviceArray):
            return self._download(self._value)
        return self._value

    @value.setter
    def value(self, val):
        self._value = val
        self.cached = val

    def host_scalar(self):
        """Python float when this leaf is a host-known 0-d constant (folded into kernels as an immediate).
        An uploaded copy in `_dev` does not hide the host value; only an explicit `cached = DeviceArray` does."""
        if self._host is not None:
            v = self._host
        elif self._dev_override:
            return None
        else:
            v = self._value
        tv = type(v)
        if tv is np.ndarray:
            return float(v) if v.ndim == 0 else None
        if tv is DeviceArray:
            return None
        if tv is float or tv is int or isinstance(v, numbers.Number):
            return float(v)
        return None

    def eval(self):
        # a leaf's host value needs no device round trip (reference node.py:131-132 returns _value)
        if self._host is None:
            if self._dev is None and not isinstance(self._value, DeviceArray):
                self._host = self._value
            else:
                if self._dev is None:
                    self._dev = self._value
                self._host = self._download(self._dev)
        return self._host

    def _eval(self):
        v = self._value
        if isinstance(v, DeviceArray):
            return self._download(v)
        return v

    def _eval_device(self):
        v = self._host if self._host is not None else self._value
        if isinstance(v, DeviceArray):
            return v
        return to_device(v)

    def _partial_derivative(self, wrt, previous_grad):
        if self == wrt:
            return previous_grad
        return 0


class ConstFill(Variable):
    """`Variable(np.ones(shape))` / `np.zeros` that is never materialised: a one-element device buffer viewed with
    zero strides.  Stands in for the ones-operands the reference creates at grad.py:26, ops.py:209,
    reshape.py:18 and high_level_ops.py:32 (256 MiB each at BASELINE config 3).  `fill` is the known constant
    (None once somebody has overwritten the value)."""

    def __init__(self, shape, fill=1.0, name=None):
        self.fill = float(fill)
        self._orig_fill = float(fill)
        self._fill_shape = tuple(int(s) for s in shape)
        self._materialised = None
        Node.__init__(self, [], name if name is not None else "np.full(%s, %g)" % (self._fill_shape, fill))
        self._lazy_name = None
        self.shape = self._fill_shape

    @property
    def _value(self):
        if self._materialised is None:
            self._materialised = np.full(self._fill_shape, self._orig_fill, dtype=np.float64)
        return self._materialised

    @_value.setter
    def _value(self, v):
        self._materialised = v
        self.fill = None

    @property
    def cached(self):
        return Node.cached.fget(self)

    @cached.setter
    def cached(self, value):
        if value is not None:
            self.fill = None
        Node.cached.fset(self, value)

    def host_scalar(self):
        if self.fill is not None and self._host is None:
            return self.fill if self._fill_shape == () else None     # never materialise just to look
        return Variable.host_scalar(self)

    def _eval(self):
        return self._value

    def _eval_device(self):
        if self.fill is not None and self._host is None:
            return DeviceArray.full(self._fill_shape, self.fill)
        return Variable._eval_device(self)


class add_context:
    """Name scope (reference node.py:140-146): pushes `ctx_<timestamp>` on Node.context_list for the `with` body."""
    __slots__ = ("ctx",)

    def __init__(self, ctx):
        self.ctx = ctx

    def __enter__(self):
        Node.context_list.append(self.ctx + "_" + str(time.time()))

    def __exit__(self, *exc):
        del Node.context_list[-1]
        return False
