"""Convolution as im2col + Einsum (SURVEY.md 8 row a23 — the reference only sketches it in comments:
reshape.py:137-160 `AsStrided`, high_level_ops.py:152-165 `Convolution`).  NHWC input, HWIO filter, VALID,
stride 1, cross-correlation — the semantics of the reference's dead test (tests/test_convolution.py:17-33).

`Im2Col` and `Col2Im` are each other's adjoint, so gradients of any order are ordinary graph nodes."""
import ctypes as C

from .. import _lib
from ..device import DeviceArray, stream_ptr
from .engine import dev
from .node import Node
from .ops import Einsum, module_wrapper
from .reshape import Reshape


class Im2Col(Node):
    """(N,H,W,C) -> (N*OH*OW, kh*kw*C): row p = (n,oh,ow) holds the window x[n, oh:oh+kh, ow:ow+kw, :] flattened."""

    def __init__(self, node, kh, kw, name="Im2Col"):
        super().__init__([node], name)
        self.node = self.children[0]
        self.kh, self.kw = int(kh), int(kw)
        n, h, w, c = self.node.shape
        if self.kh > h or self.kw > w:
            raise ValueError("Im2Col: window %dx%d larger than the image %dx%d" % (kh, kw, h, w))
        self.shape = (n * (h - self.kh + 1) * (w - self.kw + 1), self.kh * self.kw * c)

    def _eval_device(self):
        x = dev(self.node)
        out = DeviceArray.empty(self.shape)
        tx, to = x.tensor(), out.tensor()
        _lib.call("adb_im2col", C.byref(tx), self.kh, self.kw, C.byref(to), C.c_void_p(stream_ptr()))
        return out

    def _partial_derivative(self, wrt, previous_grad):
        if self.node == wrt:
            return Col2Im(previous_grad, tuple(self.node.shape), self.kh, self.kw)
        return 0


class Col2Im(Node):
    """Adjoint of Im2Col: scatters-adds window rows back onto the (N,H,W,C) image."""

    def __init__(self, node, image_shape, kh, kw, name="Col2Im"):
        super().__init__([node], name)
        self.node = self.children[0]
        self.kh, self.kw = int(kh), int(kw)
        self.shape = tuple(int(s) for s in image_shape)

    def _eval_device(self):
        cols = dev(self.node)
        n, h, w, c = self.shape
        want = (n * (h - self.kh + 1) * (w - self.kw + 1), self.kh * self.kw * c)
        cols = cols.broadcast_to(want) if cols.shape != want else cols
        if cols.strides[1] != 1:                      # broadcast / transposed gradient: make the rows contiguous
            from ..device import copy_device
            dense = DeviceArray.empty(want)
            copy_device(cols, dense)
            cols = dense
        out = DeviceArray.empty(self.shape)
        tc, to = cols.tensor(), out.tensor()
        _lib.call("adb_col2im", C.byref(tc), self.kh, self.kw, C.byref(to), C.c_void_p(stream_ptr()))
        return out

    def _partial_derivative(self, wrt, previous_grad):
        if self.node == wrt:
            return Im2Col(previous_grad, self.kh, self.kw)
        return 0


@module_wrapper
def Conv2D(x, w):
    """x: (N,H,W,C), w: (kh,kw,C,O)  ->  (N,OH,OW,O), VALID cross-correlation = im2col @ reshape(w)."""
    kh, kw, c, o = w.shape
    n, h, wd, _ = x.shape
    cols = Im2Col(x, kh, kw)
    out = Einsum("pk,ko->po", cols, Reshape(w, (kh * kw * c, o)))
    return Reshape(out, (n, h - kh + 1, wd - kw + 1, o))


# ------------------------------------------------------------------------------------------------ launch tapes
from . import tape as _tape  # noqa: E402
_tape.register(Im2Col, "kh", "kw")
_tape.register(Col2Im, "shape", "kh", "kw")
