"""Shape / data-movement nodes — API of reference autodiff/core/reshape.py.  Views alias device memory (offset +
strides, like numpy); real copies (Concat, Pad, the scatter in Slice's gradient, Reshape of a non-dense view) run
the strided-copy program of the elementwise kernel."""
import ctypes as C
import numbers

import numpy as np

from .. import _lib
from ..device import DeviceArray, copy_device, fill_device, stream_ptr
from .engine import dev
from .node import ConstFill, Node, Variable


class ReduceSumKeepDims(Node):
    def __init__(self, node, axes):
        super().__init__([node], name="ReduceSumKeepDims")
        self.node = self.children[0]
        self.axes = tuple(axes)
        self.shape = [1 if i in self.axes else shp for i, shp in enumerate(self.node.shape)]

    def _ew_spec(self):
        shp = self.node.shape
        if all(shp[a] == 1 for a in self.axes):
            return ("alias",)          # axes=() (inserted by ReduceSumToShape, ops.py:40) sums nothing
        return None

    def _eval_device(self):
        x = dev(self.node)
        axes = set(a % max(x.ndim, 1) for a in self.axes)
        if all(x.shape[a] == 1 for a in axes):
            return x                      # axes=() (reference ops.py:40 inserts these) or size-1 axes: nothing to add
        out = DeviceArray.empty([1 if i in axes else d for i, d in enumerate(x.shape)])
        tx, to = x.tensor(), out.tensor()
        _lib.call("adb_reduce_sum", C.byref(tx), C.byref(to), C.c_void_p(stream_ptr()))
        return out

    def _partial_derivative(self, wrt, previous_grad):
        if self.node == wrt:
            return previous_grad * ConstFill(tuple(self.node.shape), 1.0)    # reference: prev * np.ones(shape), reshape.py:18
        return 0


class Concat(Node):
    def __init__(self, a, b, axis=0):
        assert axis >= 0  # reference reshape.py:24
        super().__init__([a, b], name="Concat")
        self.a, self.b = self.children
        self.axis = axis
        self.shape = list(self.a.shape)
        self.shape[axis] += self.b.shape[axis]

    def _eval_device(self):
        a, b = dev(self.a), dev(self.b)
        shape = list(a.shape)
        shape[self.axis] += b.shape[self.axis]
        out = DeviceArray.empty(shape)
        split = a.shape[self.axis]
        lo = [slice(None)] * len(shape)
        hi = [slice(None)] * len(shape)
        lo[self.axis] = slice(0, split)
        hi[self.axis] = slice(split, None)
        copy_device(a, out.getitem(tuple(lo)))
        copy_device(b, out.getitem(tuple(hi)))
        return out

    def _partial_derivative(self, wrt, previous_grad):
        previous_grad = Reshape(previous_grad, self.shape)  # in case it's just a scalar Variable
        split = self.a.shape[self.axis]
        slice_val = [slice(None, None, None) for _ in range(self.axis + 1)]
        if wrt == self.a:
            slice_val[self.axis] = slice(None, split, None)
            return previous_grad[slice_val]
        elif wrt == self.b:
            slice_val[self.axis] = slice(split, None, None)
            return previous_grad[slice_val]
        return 0


class Reshape(Node):
    def __init__(self, node, shape, name="Reshape"):
        super().__init__([node], name)
        self.node = self.children[0]
        self.shape = self.infer_shape(shape)

    def _eval_device(self):
        x = dev(self.node)
        shape = (self.shape,) if isinstance(self.shape, numbers.Number) else tuple(self.shape)
        target = int(np.prod(shape, dtype=np.int64))
        if x.ndim == 0 and target != 1:
            return x.broadcast_to(shape)       # a number is broadcast (reference reshape.py:60-61)
        if x.size != target:
            raise ValueError("cannot reshape array of size %d into shape %s" % (x.size, shape))
        v = x.reshape_view(shape)
        if v is not None:
            return v
        dense = DeviceArray.empty((x.size,))
        copy_device(x, dense.view(x.shape, _dense(x.shape)))
        return dense.view(shape, _dense(shape))

    def _partial_derivative(self, wrt, previous_grad):
        if self.node == wrt:
            return Reshape(previous_grad, self.node.shape)
        return 0

    def infer_shape(self, shape):
        if isinstance(shape, numbers.Number):
            return shape
        if -1 in shape:
            shape = list(shape)
            for i in range(len(shape)):
                if shape[i] == -1:
                    shape[i] = int(-np.prod(self.node.shape) / np.prod(shape))
        return shape


def _dense(shape):
    st, acc = [0] * len(shape), 1
    for i in range(len(shape) - 1, -1, -1):
        st[i] = acc
        acc *= max(int(shape[i]), 1)
    return st


def _basic_index_shape(shape, idx):
    """Result shape of numpy basic indexing without allocating np.zeros(shape) (reference reshape.py:90 does)."""
    probe = np.lib.stride_tricks.as_strided(np.zeros(1), shape=tuple(shape), strides=(0,) * len(shape), writeable=False)
    return probe[idx].shape


class Slice(Node):
    def __init__(self, node, slice_val, name="Slice"):
        if name is None:
            name = str(slice_val)
        super().__init__([node], name)
        self.node = self.children[0]
        # numpy >= 1.23 rejects a *list* of slices; older numpy (the reference's era) read it as a tuple
        self.slice_val = tuple(slice_val) if isinstance(slice_val, list) else slice_val
        self.shape = _basic_index_shape(self.node.shape, self.slice_val)

    def _eval_device(self):
        return dev(self.node).getitem(self.slice_val)

    def _partial_derivative(self, wrt, previous_grad):
        # The reference evaluates previous_grad HERE, at graph-build time, and returns a constant ndarray
        # (reshape.py:96-102).  Same eagerness, but the scatter runs on the device and the constant stays in HBM.
        if self.node == wrt and lazy_slice_grad.enabled:
            return SliceScatter(previous_grad, tuple(wrt.shape), self.slice_val)     # a real graph node
        if self.node == wrt:
            g = DeviceArray.empty(tuple(wrt.shape))
            fill_device(g, 0.0)
            pg = previous_grad.eval_device()
            copy_device(pg, g.getitem(self.slice_val))
            return Variable(g, name="Slice grad")
        return 0


class SliceScatter(Node):
    """zeros(shape) with [slice_val] = node: the gradient of `Slice` as a GRAPH NODE.  The
    reference evaluates that gradient eagerly into a constant (reshape.py:96-102), which makes second-order gradients
    through Slice / Concat / NN silently constant; with `ad.lazy_slice_grad(True)` this node is used instead and
    gradients of any order through slicing are correct.  Its own derivative is the Slice of the incoming gradient."""

    def __init__(self, node, shape, slice_val, name="SliceScatter"):
        super().__init__([node], name)
        self.node = self.children[0]
        self.shape = tuple(int(s) for s in shape)
        self.slice_val = tuple(slice_val) if isinstance(slice_val, list) else slice_val

    def _eval_device(self):
        g = DeviceArray.empty(self.shape)
        fill_device(g, 0.0)
        copy_device(dev(self.node), g.getitem(self.slice_val))
        return g

    def _partial_derivative(self, wrt, previous_grad):
        if self.node == wrt:
            return Reshape(previous_grad, self.shape)[self.slice_val]
        return 0


class lazy_slice_grad:
    """`ad.lazy_slice_grad(True)` (or `with ad.lazy_slice_grad(True): ...`) makes Slice's gradient a graph node instead
    of the reference's eagerly evaluated constant.  Default off: bug-for-bug parity with the reference."""
    enabled = False

    def __init__(self, on=True):
        self.prev = lazy_slice_grad.enabled
        lazy_slice_grad.enabled = bool(on)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        lazy_slice_grad.enabled = self.prev
        return False


class Pad(Node):
    def __init__(self, node, pad_width, constant_values, name="Slice"):
        """pad_width as in np.pad (reference reshape.py:106-122; the node really is named "Slice" there)."""
        super().__init__([node], name)
        self.node = self.children[0]
        self.pad_width = pad_width
        self.constant_values = constant_values
        self._pairs = _as_pairs(pad_width, len(self.node.shape))
        self.shape = tuple(d + lo + hi for d, (lo, hi) in zip(self.node.shape, self._pairs))

    def _eval_device(self):
        """np.pad(mode="constant") pads axis by axis, first axis first, each time over the full extent of the array padded so
        far (numpy lib/arraypad.py `_set_pad_area` loop): where the pad slabs of two axes overlap, the LATER axis' constant
        wins.  One constant for all sides is a single fill; per-side constants are one fill per non-empty slab in that order."""
        x = dev(self.node)
        nd = len(self.shape)
        cv = np.broadcast_to(np.asarray(self.constant_values, dtype=np.float64), (nd, 2)) if nd else np.zeros((0, 2))
        out = DeviceArray.empty(self.shape)
        if np.unique(cv).size <= 1:
            fill_device(out, float(cv.flat[0]) if cv.size else 0.0)
        else:
            for ax, (lo, hi) in enumerate(self._pairs):
                full = [slice(None)] * nd
                if lo:
                    fill_device(out.getitem(tuple(full[:ax] + [slice(0, lo)] + full[ax + 1:])), float(cv[ax, 0]))
                if hi:
                    fill_device(out.getitem(tuple(full[:ax] + [slice(self.shape[ax] - hi, None)] + full[ax + 1:])), float(cv[ax, 1]))
        inner = tuple(slice(lo, lo + d) for d, (lo, hi) in zip(x.shape, self._pairs))
        copy_device(x, out.getitem(inner))
        return out

    def _partial_derivative(self, wrt, previous_grad):
        if self.node == wrt:
            slice_val = [slice(pad[0], shp - pad[1]) for pad, shp in zip(self._pairs, self.shape)]
            return previous_grad[slice_val]
        return 0


def _as_pairs(pad_width, ndim):
    pw = np.asarray(pad_width)
    if pw.ndim == 0:
        return [(int(pw), int(pw))] * ndim
    if pw.ndim == 1 and pw.size == 2:
        return [(int(pw[0]), int(pw[1]))] * ndim
    if pw.shape == (1, 2):
        return [(int(pw[0, 0]), int(pw[0, 1]))] * ndim
    if pw.shape != (ndim, 2):
        raise ValueError("Unable to create correctly shaped tuple from %s" % (pad_width,))
    return [(int(a), int(b)) for a, b in pw]


# ------------------------------------------------------------------------------------------------ launch tapes
from . import tape as _tape  # noqa: E402
_tape.register(Concat, "axis")
_tape.register(Reshape, "shape")
_tape.register(Slice, "slice_val")
_tape.register(SliceScatter, "shape", "slice_val")
_tape.register(Pad, "_pairs", "constant_values")
