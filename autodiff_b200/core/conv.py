"""Convolution as im2col + Einsum (SURVEY.md 8 row a23 — the reference only sketches it in comments:
reshape.py:137-160 `AsStrided`, high_level_ops.py:152-165 `Convolution`).  NHWC input, HWIO filter, VALID,
stride 1, cross-correlation — the semantics of the reference's dead test (tests/test_convolution.py:17-33).

`Im2Col` and `Col2Im` are each other's adjoint, so gradients of any order are ordinary graph nodes."""
import ctypes as C

from .. import _lib
from ..device import DeviceArray, stream_ptr
from .engine import dev, scalar_of
from .node import Node
from .ops import Add, Einsum, module_wrapper
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
        got = _implicit_data_gradient(self)
        if got is not None:
            return got
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


# ------------------------------------------------------------------------------------------------ implicit-GEMM convolution
# The einsums on either side of Im2Col / Col2Im never need the (N*OH*OW, kh*kw*C) matrix itself: adb_conv_gemm reads the
# windows straight from the NHWC image with im2col-mode TMA.  The evaluator's plan marks such an Im2Col (every consumer is one of
# the two einsum forms below) and the `Variable(0) + Einsum('po,ko->pk')` chain under a Col2Im as VIRTUAL: they get no value,
# their consumers look through them.  At configs[3] that removes a 2.57 GB matrix written once and read twice, and a second one
# (its gradient) written by a GEMM and read by col2im.  The graph the user sees - and its values - are unchanged.
def _two_letter(s):
    return len(s) == 2 and s[0] != s[1] and s.isalpha()


def einsum_role(e, cols):
    """0: e = Einsum('pk,ko->po', cols, w) (forward);  1: e = Einsum('pk,po->ko', cols, d) (filter gradient);  None: neither.
    Operand order is free; `cols` must appear exactly once."""
    if type(e) is not Einsum or len(e.operands) != 2 or type(cols) is not Im2Col:
        return None
    a, b = e.operands
    if (a is cols) == (b is cols):
        return None
    i = 0 if a is cols else 1
    sc, so, sout = e.opnames[i], e.opnames[1 - i], e.opnames[2]
    if not (_two_letter(sc) and _two_letter(so) and _two_letter(sout)) or len(e.operands[1 - i].shape) != 2:
        return None
    if cols.node.shape[3] % 16:
        return None                                   # the TMA rows are 16 channels wide
    p, k = sc
    if so[0] == k and so[1] not in sc and sout == p + so[1]:
        return 0
    if so[0] == p and so[1] not in sc and sout == k + so[1]:
        return 1
    return None


def _zero_add_inner(node):
    """`Variable(0) + x` - how grad accumulates a node's first partial (reference grad.py:28-31) - gives x."""
    if type(node) is Add and len(node.children) == 2:
        a, b = node.children
        if scalar_of(a) == 0.0 and tuple(b.shape) == tuple(node.shape):
            return b
        if scalar_of(b) == 0.0 and tuple(a.shape) == tuple(node.shape):
            return a
    return None


def data_gradient_form(e, col2im):
    """e = Einsum('po,ko->pk', d, w) under `col2im`, with the image's channel count and d's column count multiples of 16/any:
    returns (d node, w node) or None."""
    if type(e) is not Einsum or len(e.operands) != 2:
        return None
    s0, s1, sout = e.opnames
    if not (_two_letter(s0) and _two_letter(s1) and _two_letter(sout)):
        return None
    for i in (0, 1):
        sd, sw = (s0, s1) if i == 0 else (s1, s0)
        p, o = sd
        if sw[1] == o and sw[0] not in sd and sout == p + sw[0]:
            d, w = e.operands[i], e.operands[1 - i]
            if len(d.shape) == 2 and len(w.shape) == 2 and d.shape[1] % 16 == 0 and d.shape[1] >= 16 \
                    and tuple(e.shape) == (d.shape[0], w.shape[0]) and w.shape[0] == col2im.kh * col2im.kw * col2im.shape[3]:
                return d, w
    return None


def mark_virtual(order, consumers, top_ids, virtual):
    """Called by the evaluator's planner: which nodes of `order` need no value of their own."""
    for node in order:
        k = id(node)
        tn = type(node)
        if tn is Im2Col:
            cons = consumers.get(k)
            if cons and k not in top_ids and all(einsum_role(c, node) is not None for c in cons):
                virtual[k] = node
        elif tn is Col2Im:
            chain, x = [], node.node
            inner = _zero_add_inner(x)
            if inner is not None:
                chain.append(x)
                x = inner
            if data_gradient_form(x, node) is None:
                continue
            chain.append(x)
            # every link must be consumed by the next link only, and be neither an output nor already evaluated
            ok, nxt = True, node
            for link in chain:
                cons = consumers.get(id(link))
                if id(link) in top_ids or cons is None or len(cons) != 1 or cons[0] is not nxt:
                    ok = False
                    break
                nxt = link
            if ok:
                for link in chain:
                    virtual[id(link)] = link


def _has_value(node):
    return node._dev is not None or node._host is not None


def _conv_call(mode, x, kh, kw, w, out):
    tx, tw, to = x.tensor(), w.tensor(), out.tensor()
    if not _lib.lib().adb_conv_gemm_ok(mode, C.byref(tx), kh, kw, C.byref(tw), C.byref(to)):
        return False
    _lib.call("adb_conv_gemm", mode, C.byref(tx), kh, kw, C.byref(tw), C.byref(to), C.c_void_p(stream_ptr()))
    return True


def materialise(cols_node):
    """The plan made `cols_node` virtual but a consumer cannot read through it after all (alignment, strides): evaluate it."""
    if not _has_value(cols_node):
        cols_node._dev = cols_node._eval_device()
    return cols_node._dev


def implicit_einsum(e):
    """Einsum._eval_device hook: the value of `e` when one operand is a virtual Im2Col, else None."""
    a, b = e.operands
    cols = a if (type(a) is Im2Col and not _has_value(a)) else (b if (type(b) is Im2Col and not _has_value(b)) else None)
    if cols is None:
        return None
    role = einsum_role(e, cols)
    if role is not None:
        other = dev(b if cols is a else a)
        x = dev(cols.node)
        out = DeviceArray.empty(e.shape)
        if _conv_call(role, x, cols.kh, cols.kw, other, out):
            return out
    materialise(cols)
    return None


def _implicit_data_gradient(col2im):
    x = col2im.node
    if _has_value(x):
        return None
    chain = [x]
    inner = _zero_add_inner(x)
    if inner is not None and not _has_value(inner):
        chain.append(inner)
        x = inner
    form = data_gradient_form(x, col2im)
    if form is not None:
        d, w = dev(form[0]), dev(form[1])
        n, h, wd, c = col2im.shape
        oh, ow = h - col2im.kh + 1, wd - col2im.kw + 1
        d4 = d.reshape_view((n, oh, ow, d.shape[1])) if d.shape[0] == n * oh * ow else None
        if d4 is not None:
            out = DeviceArray.empty(col2im.shape)
            if _conv_call(2, d4, col2im.kh, col2im.kw, w, out):
                return out
    # cannot read through after all: evaluate the chain bottom-up the ordinary way
    for link in reversed(chain):
        if not _has_value(link):
            link._dev = link._eval_device()
    return None


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
