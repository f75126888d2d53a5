"""Op library — same names, constructor arguments, shapes, error messages and gradient GRAPHS as reference
autodiff/core/ops.py; each op's numpy `_eval` body is replaced by `_eval_device`, which launches sm_100a kernels
through libadb200.so on device-resident operands.

Bit-level rules kept from the reference: Recipr = 1/(x+1e-12) (ops.py:116), Log =
log(x+1e-12) (ops.py:309), ReLU = x*(x>0) with gradient prev*relu*Recipr(relu) (ops.py:228,232), n-ary Add/Mul
folded left to right (ops.py:53,75), Tanh/Softmax as graph macros (ops.py:399-402,421-434), Einsum gradients as
swapped-subscript Einsums with inserted ones-operands (ops.py:182-213).  Exact-identity work is elided on the
device: x+0, x*1, 'ab->ab', axes=() and pure permutations alias their input instead of copying.
"""
import ctypes as C
import math
import numbers
import re
from string import ascii_lowercase, ascii_uppercase

import numpy as np

from .. import _lib
from ..device import DeviceArray, copy_device, fill_device, stream_ptr, to_host, torch
from . import engine
from .engine import dev, run_program, scalar_of
from .fusion import broadcast_shapes
from .node import ConstFill, Node, Variable, add_context
from .reshape import ReduceSumKeepDims

IN, IMM, TMP = _lib.SRC_IN, _lib.SRC_IMM, _lib.SRC_TMP


def module_wrapper(fn):
    def wrap_in_context(*args, **kwargs):
        with add_context(fn.__name__):
            return fn(*args, **kwargs)

    wrap_in_context.__name__ = fn.__name__
    return wrap_in_context


def letters_from_tuple(tpl):
    return ascii_lowercase[:len(tpl)]


def shape_from_elems(*elems):
    """Broadcast shape of the children (reference ops.py:23-26, which allocates np.ones per child to find it)."""
    if len(elems) == 0:
        return 1,
    return broadcast_shapes(*[elem.shape for elem in elems])


@module_wrapper
def ReduceSumToShape(tensor, to_shape):
    # `==` between a list shape and a tuple shape is False on purpose: the reference inserts the identity
    # einsum in that case (ops.py:31) and the device aliases it away.
    if tensor.shape == to_shape:
        return tensor
    previous_grad_letters = letters_from_tuple(tensor.shape)
    if len(to_shape) == 0:
        wrt_letters = ""
    else:
        wrt_letters = previous_grad_letters[-len(to_shape):]
    new_curr_grad = Einsum(str(previous_grad_letters) + "->" + str(wrt_letters), tensor)
    return ReduceSumKeepDims(new_curr_grad, axes=[i for i, val in enumerate(to_shape) if val == 1])


# ------------------------------------------------------------------------------------------------ n-ary
def _single(node):
    """Evaluates one fusable node on its own through the same emitter the fusion groups use."""
    from . import fusion
    try:
        return fusion.run_group(node, {id(node): node}, dev, scalar_of, run_program)
    except fusion.FusionOverflow:
        spec = node._ew_spec()
        if spec[0] != "nary":
            raise
        return _nary_chunked(node, spec[1], spec[2])


def _nary_chunked(node, kind, identity):
    """n-ary fold with more operands than one program can take: left-to-right in chunks (same association)."""
    terms, shapes = [], []
    for ch in node.children:
        s = scalar_of(ch)
        if s is not None:
            if s != identity:
                terms.append((IMM, s))
        else:
            a = dev(ch)
            terms.append((IN, a))
            shapes.append(a.shape)
    out_shape = tuple(np.broadcast_shapes(*shapes)) if shapes else ()
    acc, i = None, 0
    while i < len(terms):
        code, inputs = [], []
        if acc is not None:
            code.append((_lib.EW_LOAD | IN, 0, 0.0))
            inputs.append(acc)
        while i < len(terms) and len(inputs) < _lib.EW_MAX_IN and len(code) < _lib.EW_MAX_INSTR - 1:
            src, val = terms[i]
            op = _lib.EW_LOAD if not code else kind
            if src == IN:
                code.append((op | IN, len(inputs), 0.0))
                inputs.append(val)
            else:
                code.append((op | IMM, 0, val))
            i += 1
        code.append((_lib.EW_STORE_OUT, 0, 0.0))
        shape = out_shape if i == len(terms) else tuple(np.broadcast_shapes(*[t.shape for t in inputs]))
        acc = run_program(code, inputs, shape)
    return acc


class Add(Node):
    _scalar_as_array = True      # reference returns np.array(sum(...)) (ops.py:53)

    def __init__(self, *elems, name="Add"):
        if not elems:
            name = "0-" + name
        super().__init__(list(elems), name)
        self.shape = shape_from_elems(*self.children)

    def _ew_spec(self):
        return ("nary", _lib.EW_ADD, 0.0)

    def _eval_device(self):
        return _single(self)

    def _partial_derivative(self, wrt, previous_grad):
        wrt_count = self.children.count(wrt)
        grad = previous_grad * Variable(wrt_count)
        return ReduceSumToShape(grad, wrt.shape)


class Mul(Node):
    def __init__(self, *elems, name="Mul"):
        if not elems:
            name = "1-" + name
        super().__init__(list(elems), name)
        self.shape = shape_from_elems(*self.children)

    def _ew_spec(self):
        return ("nary", _lib.EW_MUL, 1.0)

    def _eval_device(self):
        return _single(self)

    def _partial_derivative(self, wrt, previous_grad):
        add_list = []
        for loc, child in enumerate(self.children):
            if child == wrt:
                add_list.append(Mul(*[ch for i, ch in enumerate(self.children) if loc != i]))
        grad = previous_grad * Add(*add_list)
        return ReduceSumToShape(grad, wrt.shape)


# ------------------------------------------------------------------------------------------------ unary
class _Elementwise(Node):
    """One-input elementwise op: `program` is the bytecode applied to the loaded element."""
    default_name = "Elementwise"
    program = ()

    def __init__(self, node, name=None):
        super().__init__([node], self.default_name if name is None else name)
        self.node = self.children[0]
        self.shape = self.node.shape

    def _code(self):
        return list(self.program)

    def _ew_spec(self):
        return ("unary", self._code())

    def _eval_device(self):
        return _single(self)

    def _partial_derivative(self, wrt, previous_grad):
        if self.node == wrt:
            return self._pd(previous_grad)
        return 0


class Negate(_Elementwise):
    default_name = "Negate"
    program = ((_lib.EW_NEG, 0, 0.0),)

    def _pd(self, previous_grad):
        return -previous_grad


class Recipr(_Elementwise):
    """Elementwise 1 / (x + epsilon) — reference ops.py:105-121; every `/` in a graph routes here (node.py:91-97)."""
    default_name = "Reciprocal"

    def _code(self):
        return [(_lib.EW_RECIP_EPS, 0, Node.epsilon)]

    def _pd(self, previous_grad):
        return - previous_grad * self * self


class ReLU(_Elementwise):
    default_name = "ReLU"
    program = ((_lib.EW_RELU, 0, 0.0),)

    def _pd(self, previous_grad):
        return previous_grad * self * Recipr(self)


class Log(_Elementwise):
    default_name = "Log"

    def _code(self):
        return [(_lib.EW_LOG_EPS, 0, Node.epsilon)]

    def _pd(self, previous_grad):
        return previous_grad * Recipr(self.node)


class Identity(_Elementwise):
    default_name = "Identity"

    def _ew_spec(self):
        return ("alias",)              # the reference returns the child's array itself (ops.py:324)

    def _pd(self, previous_grad):
        return previous_grad


class Exp(_Elementwise):
    default_name = "Exp"
    program = ((_lib.EW_EXP, 0, 0.0),)

    def _pd(self, previous_grad):
        return previous_grad * self


class Sigmoid(_Elementwise):
    default_name = "Sigmoid"
    program = ((_lib.EW_SIGMOID, 0, 0.0),)

    def _pd(self, previous_grad):
        return previous_grad * self * (1 - self)


class NormalDistribution(Node):
    def __init__(self, node, mean=0, variance=1, name="Normal Distribution"):
        super().__init__([node], name=name)
        self.node = self.children[0]
        self.mean = mean
        self.variance = variance
        self.shape = self.node.shape

    def _eval_device(self):
        # reference ops.py:387-391: 1/sqrt(2*pi*v**2) * exp(-(x-m)**2 / (2*v**2)); scalar hyper-parameters are
        # folded on the host exactly as Python evaluates them there.
        m, v = self.mean, self.variance
        coef = float(1 / np.sqrt(2 * np.pi * (v ** 2)))
        denom = float(2 * v ** 2)
        x = dev(self.node)
        code = [(_lib.EW_LOAD | IN, 0, 0.0), (_lib.EW_ADD | IMM, 0, -float(m)), (_lib.EW_SQUARE, 0, 0.0), (_lib.EW_NEG, 0, 0.0),
                (_lib.EW_DIV | IMM, 0, denom), (_lib.EW_EXP, 0, 0.0), (_lib.EW_MUL | IMM, 0, coef), (_lib.EW_STORE_OUT, 0, 0.0)]
        return run_program(code, [x], x.shape)

    def _partial_derivative(self, wrt, previous_grad):
        if self.node == wrt:
            return -previous_grad * self.node * self
        return 0


class Pow(Node):
    def __init__(self, first, second, name="Pow"):
        super().__init__([first, second], name)
        self.first = self.children[0]
        self.second = self.children[1]
        self.shape = shape_from_elems(*self.children)

    def _ew_spec(self):
        return ("pow",)

    def _eval_device(self):
        return _single(self)

    def _partial_derivative(self, wrt, previous_grad):
        if self.first == self.second == wrt:
            return previous_grad * self * (Log(self.first) + 1)
        elif self.first == wrt:
            return previous_grad * self.second * Pow(self.first, self.second - 1)
        elif self.second == wrt:
            return previous_grad * Log(self.first) * self
        return 0


# ------------------------------------------------------------------------------------------------ einsum
_split_dots_cache = {}
_einsum_parse_cache = {}     # (op_str, operand shapes) -> (opnames, all_letters, letter_to_dim, shape); all read-only afterwards


class Einsum(Node):
    def __init__(self, op_str, *operands, name="EinSum"):
        super().__init__(list(operands), name + " " + op_str)
        # ellipsis can't be in the middle of op_letters (reference ops.py:127)
        self.op_str = op_str
        self.operands = self.children

        # parsing depends only on (op_str, operand shapes): training loops rebuild the same einsums every step
        key = (op_str, tuple(tuple(op.shape) for op in self.operands))
        hit = _einsum_parse_cache.get(key)
        if hit is not None:
            self.opnames, self.all_letters, self.letter_to_dim, self.shape = hit
            return

        self.opnames = re.split(",|->", self.op_str)
        self.all_letters = "".join(set("".join(self.opnames[:-1])))
        self.letter_to_dim = {}

        if len(self.operands) + 1 != len(self.opnames):
            raise ValueError("Number of operands doesn't match the einsum string!")

        for op, op_letters in zip(self.operands, self.opnames[:-1]):
            if len(op.shape) != 0 and len(op.shape) != len(op_letters) \
                    and "..." not in op_letters and op_letters != "":
                raise ValueError("Dimension of operand " + str(op) + " doesn't match the string! " +
                                 "Shape: " + str(op.shape) + " , string: '" + op_letters + "'")

            shp = tuple(op.shape)
            if op_letters[:3] == "...":
                op_letters = op_letters[::-1]
                shp = shp[::-1]
            for i, lett in enumerate(Einsum.split_dots(op_letters)):
                try:
                    if len(lett) == 1:
                        dim = [shp[i]]
                    else:
                        dim = list(shp[i:])
                    if self.letter_to_dim.get(lett, dim) != dim:
                        raise ValueError("Inconsistent dimension names!")
                    self.letter_to_dim[lett] = dim
                except IndexError:
                    pass  # letters that can't be matched are dimension 1

        self.shape = []
        for let in Einsum.split_dots(self.opnames[-1]):
            for l in self.letter_to_dim.get(let, [1]):
                self.shape.append(l)
        self.shape = tuple(self.shape)
        if len(_einsum_parse_cache) > 4096:
            _einsum_parse_cache.clear()
        _einsum_parse_cache[key] = (self.opnames, self.all_letters, self.letter_to_dim, self.shape)

    @staticmethod
    def split_dots(op_str):
        r = _split_dots_cache.get(op_str)
        if r is None:
            if len(_split_dots_cache) > 4096:
                _split_dots_cache.clear()
            r = _split_dots_cache[op_str] = tuple(re.findall(r"\.{3}|\S", op_str))
        return r

    def _ew_spec(self):
        """'ab->ab' (inserted by ReduceSumToShape, ops.py:31-39) is a pass-through for fusion."""
        a = self.opnames[0]
        if len(self.operands) == 1 and a == self.opnames[-1] and "." not in a and len(set(a)) == len(a) \
                and len(a) == len(self.operands[0].shape):
            return ("alias",)
        return None

    # ---- device evaluation
    def _eval_device(self):
        if len(self.children) == 2 and (self.children[0]._dev is None or self.children[1]._dev is None):
            from . import conv as _conv          # an operand without a value: an Im2Col the plan made virtual (core/conv.py)
            got = _conv.implicit_einsum(self)
            if got is not None:
                return got
        subs = list(self.opnames[:-1])
        out_sub = self.opnames[-1]
        arrs = []
        for i, op in enumerate(self.operands):
            a = dev(op)
            if a.ndim == 0 and subs[i] != "" and "..." not in subs[i]:
                # a number operand is broadcast to its letters' shape (reference ops.py:175-178)
                shp = [l for let in Einsum.split_dots(subs[i]) for l in self.letter_to_dim.get(let, [1])]
                a = a.broadcast_to(shp)
            arrs.append(a)
        subs, out_sub = _expand_ellipsis(subs, out_sub, [a.ndim for a in arrs])

        # extents (numpy rule: size-1 dims broadcast)
        extent = {}
        for sub, a in zip(subs, arrs):
            if len(sub) != a.ndim:
                raise ValueError("einsum: operand has rank %d but subscripts '%s'" % (a.ndim, sub))
            for l, d in zip(sub, a.shape):
                cur = extent.get(l)
                if cur is None or cur == 1:
                    extent[l] = d
                elif d != 1 and d != cur:
                    raise ValueError("einsum: inconsistent extent for '%s' (%d vs %d)" % (l, cur, d))
        for l in out_sub:
            if l not in extent:
                raise ValueError("einsum: output subscript '%s' never appeared in an input" % l)
        out_shape = tuple(extent[l] for l in out_sub)

        # all-ones operands (inserted by the gradient, ops.py:204-210) whose letters occur elsewhere multiply by 1
        keep = list(range(len(arrs)))
        if len(arrs) > 1:
            for i, op in enumerate(self.operands):
                if isinstance(op, ConstFill) and op.fill == 1.0 and op._host is None and len(keep) > 1:
                    others = set("".join(subs[j] for j in keep if j != i))
                    if all(l in others for l in subs[i]):
                        keep.remove(i)
        subs = [subs[i] for i in keep]
        arrs = [arrs[i] for i in keep]

        # a lone operand with no summed / repeated letter is a pure view (np.einsum returns a view too)
        if len(arrs) == 1 and len(set(subs[0])) == len(subs[0]) and set(subs[0]) == set(out_sub) \
                and tuple(arrs[0].shape) == tuple(extent[l] for l in subs[0]):
            return arrs[0].permute([subs[0].index(l) for l in out_sub])

        out = DeviceArray.empty(out_shape)
        n = len(arrs)
        tin = (_lib.Tensor * n)(*[a.tensor() for a in arrs])
        tout = out.tensor()
        spec = (",".join(subs) + "->" + out_sub).encode()
        if engine.kernel_timings is not None:
            t = torch()
            e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            e0.record()
            _lib.call("adb_einsum", spec, n, tin, C.byref(tout), C.c_void_p(stream_ptr()))
            e1.record()
            engine.kernel_timings.append((spec.decode() + " " + "|".join(str(op.name) for op in self.operands), e0, e1))
            return out
        _lib.call("adb_einsum", spec, n, tin, C.byref(tout), C.c_void_p(stream_ptr()))
        return out

    def _partial_derivative(self, wrt, previous_grad):
        """c = einsum("ij,jk->ik", a, b)  =>  dc/da = einsum("ik,jk->ij", g, b): the output subscript swaps places
        with the operand's (reference ops.py:182-213).  Letters of `wrt` that no remaining operand carries get an
        explicit ones-operand (never materialised here)."""
        order = list(range(len(self.opnames)))
        try:
            loc = self.operands.index(wrt)
        except ValueError:
            return 0
        order[loc], order[-1] = order[-1], order[loc]
        pool = self.operands + [previous_grad]
        operands_with_grad = [pool[i] for i in order]
        opnames = [self.opnames[i] for i in order]

        for i, letter in enumerate(Einsum.split_dots(self.opnames[loc])):
            if letter not in Einsum.split_dots("".join(opnames[:-1])):
                opnames.insert(0, letter)
                dim = wrt.shape[i]
                shape = (dim,) if isinstance(dim, numbers.Integral) else tuple(dim)
                operands_with_grad.insert(0, ConstFill(shape, 1.0, name="np.ones(" + str(dim) + ")"))
        op_str = Einsum.to_einsum_string(opnames)
        return Einsum(op_str, *operands_with_grad[:-1])

    @staticmethod
    def to_einsum_string(list_of_ops):
        return ",".join(list_of_ops[:-1]) + "->" + list_of_ops[-1]


def _expand_ellipsis(subs, out_sub, ranks):
    """numpy's "..." rule: the ellipsis of each operand stands for its unnamed leading/trailing dims, right-aligned
    across operands; the backend itself only takes single letters."""
    if not any("..." in s for s in subs) and "..." not in out_sub:
        return subs, out_sub
    used = set("".join(subs) + out_sub)
    pool = [c for c in ascii_uppercase + ascii_lowercase if c not in used]
    widths = []
    for s, r in zip(subs, ranks):
        widths.append(r - len(s.replace("...", "")) if "..." in s else 0)
    width = max(widths + [0])
    if width > len(pool):
        raise ValueError("einsum: too many dimensions under '...'")
    ell = "".join(pool[:width])
    new = []
    for s, w in zip(subs, widths):
        new.append(s.replace("...", ell[width - w:]) if "..." in s else s)
    return new, out_sub.replace("...", ell)


# ------------------------------------------------------------------------------------------------ losses
class SoftmaxCEWithLogits(Node):
    def __init__(self, labels, logits, name="SoftmaxCEWithLogits"):
        super().__init__([labels, logits], name=name)
        self.labels, self.logits = self.children
        self.shape = tuple(self.logits.shape[:-1])
        if self.labels.shape is None or len(self.labels.shape) != 2 or len(self.logits.shape) != 2:
            # the N-D path checks the labels on the HOST (axis-1 sums read back synchronously) and ignores the kernel's own
            # last-axis flag: a launch tape would replay neither decision, so a graph that holds such a node is never taped
            self._never_tape = True

    def _eval_device(self):
        lab, z = dev(self.labels), dev(self.logits)
        if z.ndim != 2 or lab.ndim != 2:
            return self._eval_device_nd(lab, z)
        lab = lab.broadcast_to(z.shape)
        out = DeviceArray.empty((z.shape[0],))
        t = torch()
        flag = t.zeros(1, dtype=t.int32, device="cuda")
        tl, tz, to = lab.tensor(), z.tensor(), out.tensor()
        _lib.call("adb_softmax_ce", C.byref(tl), C.byref(tz), C.byref(to), C.c_void_p(flag.data_ptr()), C.c_void_p(stream_ptr()))
        engine.add_pending_flag(flag, ValueError("Labels must be a valid probability distribution!"))
        return out

    def _eval_device_nd(self, lab, z):
        """Labels / logits that are not both 2-D, with the reference's exact semantics (ops.py:243-255): the distribution
        check sums the labels along axis ONE whatever their rank (so one-hot (B,T,V) labels are rejected, as upstream), the
        log-sum-exp and the final sum run along the LAST axis, and labels broadcast against logits.  Rare path: the check is
        read back synchronously; the arithmetic is the 2-D kernel over the flattened leading axes (its own last-axis label
        check is ignored)."""
        if lab.ndim < 2:
            raise np.exceptions.AxisError(1, lab.ndim)                               # np.sum(labels_val, axis=1)
        red_shape = tuple(1 if i == 1 else d for i, d in enumerate(lab.shape))
        red = DeviceArray.empty(red_shape)
        tl, tr = lab.tensor(), red.tensor()
        _lib.call("adb_reduce_sum", C.byref(tl), C.byref(tr), C.c_void_p(stream_ptr()))
        sums = to_host(red)
        if not np.allclose(sums, np.ones_like(sums)):
            raise ValueError("Labels must be a valid probability distribution!")
        full = tuple(np.broadcast_shapes(lab.shape, z.shape))
        rows, cols = int(np.prod(full[:-1], dtype=np.int64)), full[-1]

        def flat(a):
            a = a.broadcast_to(full)
            v = a.reshape_view((rows, cols)) if a.is_dense() else None
            if v is None:
                d = DeviceArray.empty(full)
                copy_device(a, d)
                v = d.reshape_view((rows, cols))
            return v
        out = DeviceArray.empty(full[:-1])
        if rows * cols == 0:
            if out.size:
                fill_device(out, 0.0)
            return out
        t = torch()
        flag = t.zeros(1, dtype=t.int32, device="cuda")
        tl, tz, to = flat(lab).tensor(), flat(z).tensor(), out.reshape_view((rows,)).tensor()
        _lib.call("adb_softmax_ce", C.byref(tl), C.byref(tz), C.byref(to), C.c_void_p(flag.data_ptr()), C.c_void_p(stream_ptr()))
        return out

    def _partial_derivative(self, wrt, previous_grad):
        if wrt == self.logits:
            return Einsum("...i,...ij->...ij", previous_grad, Softmax(self.logits) - self.labels)
        elif wrt == self.labels:
            return Variable(0)
        return 0


class SigmoidCEWithLogits(Node):
    def __init__(self, labels, logits, name="SigmoidCEWithLogits"):
        super().__init__([labels, logits], name)
        self.labels, self.logits = self.children
        self.shape = self.logits.shape

    def _eval_device(self):
        # reference ops.py:274: max(x, 0) - x*z + log(1 + exp(-|x|))   (inputs: 0 = x, 1 = z)
        z, x = dev(self.labels), dev(self.logits)
        code = [(_lib.EW_LOAD | IN, 0, 0.0), (_lib.EW_MAX | IMM, 0, 0.0), (_lib.EW_STORE_T, 0, 0.0),
                (_lib.EW_LOAD | IN, 0, 0.0), (_lib.EW_MUL | IN, 1, 0.0), (_lib.EW_NEG, 0, 0.0), (_lib.EW_ADD | TMP, 0, 0.0),
                (_lib.EW_STORE_T, 0, 0.0),
                (_lib.EW_LOAD | IN, 0, 0.0), (_lib.EW_ABS, 0, 0.0), (_lib.EW_NEG, 0, 0.0), (_lib.EW_EXP, 0, 0.0),
                (_lib.EW_ADD | IMM, 0, 1.0), (_lib.EW_LOG, 0, 0.0), (_lib.EW_ADD | TMP, 0, 0.0), (_lib.EW_STORE_OUT, 0, 0.0)]
        return run_program(code, [x, z], np.broadcast_shapes(x.shape, z.shape))

    def _partial_derivative(self, wrt, previous_grad):
        if wrt == self.logits:
            return Einsum("...ij,...ij->...ij", previous_grad, Sigmoid(self.logits) - self.labels)
        return 0


class FrobeniusNorm(Node):
    def __init__(self, *nodes, name="Frobenius Norm"):
        super().__init__(list(nodes), name=name)
        self.nodes = self.children
        self.shape = 1,

    def _eval_device(self):
        # reference ops.py:369: sqrt(sum([np.sum(np.square(x)) for x in nodes])) — one reduction per node, then
        # one tiny program adds the partials left to right and takes the root.
        if len(self.nodes) == 1:            # one tensor: the root is taken inside the reduction's last slice (one launch)
            x = dev(self.nodes[0])
            out = DeviceArray.empty(())
            tx = x.tensor()
            _lib.call("adb_norm2", C.byref(tx), C.c_void_p(out.ptr), C.c_void_p(stream_ptr()))
            return out
        parts = []
        for n in self.nodes:
            x = dev(n)
            part = DeviceArray.empty(())
            tx = x.tensor()
            _lib.call("adb_sum_squares", C.byref(tx), C.c_void_p(part.ptr), C.c_void_p(stream_ptr()))
            parts.append(part)
        acc = None
        for i in range(0, len(parts), _lib.EW_MAX_IN - 1):
            chunk = parts[i:i + _lib.EW_MAX_IN - 1]
            inputs = ([acc] if acc is not None else []) + chunk
            code = [((_lib.EW_LOAD if j == 0 else _lib.EW_ADD) | IN, j, 0.0) for j in range(len(inputs))]
            if i + _lib.EW_MAX_IN - 1 >= len(parts):
                code.append((_lib.EW_SQRT, 0, 0.0))
            code.append((_lib.EW_STORE_OUT, 0, 0.0))
            acc = run_program(code, inputs, ())
        if acc is None:
            acc = run_program([(_lib.EW_LOAD | IMM, 0, 0.0), (_lib.EW_STORE_OUT, 0, 0.0)], [], ())
        return acc

    def _partial_derivative(self, wrt, previous_grad):
        try:
            loc = self.nodes.index(wrt)
        except ValueError:
            return 0
        return previous_grad * self.nodes[loc] / self


# ------------------------------------------------------------------------------------------------ macros
@module_wrapper
def Tanh(x):
    val = Exp(-2 * x)
    return (1 - val) / (1 + val)


@module_wrapper
def SquaredDifference(x, y):
    diff = x - y
    return diff * diff


@module_wrapper
def MatMul(x, y):
    return Einsum("ij,jk->ik", x, y)


@module_wrapper
def Transpose(x):
    return Einsum("ij->ji", x)


@module_wrapper
def Softmax(x):
    # not max-shifted, exactly as the reference (ops.py:423); the denominator is an einsum against [1]
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


# ------------------------------------------------------------------------------------------------ launch tapes
# classes whose device evaluation depends only on (class, these attributes, children) - see core/tape.py
from . import tape as _tape  # noqa: E402
for _cls in (Add, Mul, Negate, Recipr, ReLU, Log, Identity, Exp, Sigmoid, Pow, SoftmaxCEWithLogits, SigmoidCEWithLogits, FrobeniusNorm):
    _tape.register(_cls)
_tape.register(NormalDistribution, "mean", "variance")
_tape.register(Einsum, "op_str")
_tape.register(ReduceSumKeepDims, "axes")
