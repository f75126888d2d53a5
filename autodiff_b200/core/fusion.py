"""Elementwise fusion: groups of Add / Mul / Negate / Recipr / ReLU / Log / Exp / Sigmoid / Pow / identity nodes
that feed one another become ONE bytecode program for `adb_ewise` (one kernel, one pass over HBM).

The program replays each member's reference operation in the reference's order (left-to-right n-ary folds from
ops.py:53,75; 1/(x+eps) from ops.py:116; x*(x>0) from ops.py:228 ...), so fusing changes no rounding: fp64 `+`/`*`
are commutative, and association order is preserved by evaluating nested operands into temporaries first.

A member may be absorbed into a group when every consumer of it (among the nodes being evaluated) is already in
the group and it has the group's shape; values consumed several times inside the group live in register temps
(Tanh's exp(-2x), reference ops.py:399-402, becomes a single kernel).  Absorbed nodes are simply never
materialised — if somebody asks for one later it is recomputed on demand, which is always correct because
nodes are pure.
"""
import numpy as np

from .. import _lib
from .node import ConstFill

IN, IMM, TMP = _lib.SRC_IN, _lib.SRC_IMM, _lib.SRC_TMP
MAX_MEMBERS = 24


_bshape_cache = {}


def broadcast_shapes(*shapes):
    """np.broadcast_shapes for plain tuples/lists (same result, same error), without numpy's dummy-array machinery:
    this runs once per Add/Mul node and the second-order char-rnn graph has 40 k of them - almost all with one of a
    handful of shape combinations, so results are memoised (tuples of tuples only: a list shape is not hashable and
    takes the computation below)."""
    try:
        hit = _bshape_cache.get(shapes)
    except TypeError:
        hit = None
    else:
        if hit is not None:
            return hit
        out = _broadcast_shapes(*shapes)
        if len(_bshape_cache) < 4096 and all(type(sh) is tuple for sh in shapes):
            _bshape_cache[shapes] = out
        return out
    return _broadcast_shapes(*shapes)


def _broadcast_shapes(*shapes):
    first = shapes[0]
    same = True
    for sh in shapes:
        if sh is not first and sh != first:          # (a list never equals a tuple: those take the general path below)
            same = False
            break
    if same:
        return first if type(first) is tuple else tuple(first)
    rank = max(len(sh) for sh in shapes)
    out = [1] * rank
    for sh in shapes:
        off = rank - len(sh)
        for i, d in enumerate(sh):
            d = int(d)
            if d != 1:
                cur = out[off + i]
                if cur == 1:
                    out[off + i] = d
                elif cur != d:
                    return tuple(np.broadcast_shapes(*[tuple(x) for x in shapes]))   # raises numpy's ValueError
    return tuple(out)


def norm_shape(shape):
    if shape is None:
        return None
    if isinstance(shape, (int, np.integer)):
        return (int(shape),)
    try:
        return tuple(int(s) for s in shape)
    except TypeError:
        return None


def ew_spec(node):
    """("nary", KIND, identity) | ("unary", code) | ("alias",) | ("pow",) | None."""
    fn = getattr(node, "_ew_spec", None)
    if fn is None or "_eval" in node.__dict__ or "_eval_device" in node.__dict__:
        return None
    return fn()


class FusionOverflow(Exception):
    pass


class Emitter:
    """Turns a group (root + absorbed members) into accumulator-machine code."""

    def __init__(self, root, members, dev_of, scalar_of):
        self.root = root
        self.members = members            # id -> node (includes root)
        self.dev_of = dev_of
        self.scalar_of = scalar_of
        self.code = []
        self.inputs = []
        self.input_ix = {}
        self.free_temps = list(range(_lib.EW_TEMPS - 1, -1, -1))
        self.temp_of = {}                 # id(member) -> temp holding its value
        self.max_temp = -1
        # how many times each member is consumed inside the group
        self.uses = {}
        for n in members.values():
            for ch in n.children:
                if id(ch) in members and ch is not root:
                    self.uses[id(ch)] = self.uses.get(id(ch), 0) + 1
        self.shapes = []

    # ---- helpers
    def _alloc(self):
        if not self.free_temps:
            raise FusionOverflow("temps")
        t = self.free_temps.pop()
        self.max_temp = max(self.max_temp, t)
        return t

    def _free(self, t):
        self.free_temps.append(t)

    def _emit(self, op, arg=0, imm=0.0):
        self.code.append((op, arg, imm))
        if len(self.code) > _lib.EW_MAX_INSTR - 1:
            raise FusionOverflow("instructions")

    def _input(self, node):
        k = id(node)
        if k not in self.input_ix:
            if len(self.inputs) >= _lib.EW_MAX_IN:
                raise FusionOverflow("inputs")
            a = self.dev_of(node)
            self.input_ix[k] = len(self.inputs)
            self.inputs.append(a)
            self.shapes.append(a.shape)
        return self.input_ix[k]

    def operand(self, ch):
        """('imm', v) | ('tmp', t) | ('sub', node) | ('in', i) | ('shape', shp)"""
        s = self.scalar_of(ch)
        if s is not None:
            return ("imm", s)
        k = id(ch)
        if k in self.members and ch is not self.root:
            if k in self.temp_of:
                return ("tmp", self.temp_of[k])
            return ("sub", ch)
        return ("in", self._input(ch))

    def _apply(self, kind, opd):
        if opd[0] == "imm":
            self._emit(kind | IMM, 0, opd[1])
        elif opd[0] == "tmp":
            self._emit(kind | TMP, opd[1])
        else:
            self._emit(kind | IN, opd[1])

    # ---- evaluation: leaves the node's value in acc
    def value(self, node):
        self.emit(node)
        if self.uses.get(id(node), 0) > 1:          # shared inside the group: keep it in a temp
            t = self._alloc()
            self._emit(_lib.EW_STORE_T, t)
            self.temp_of[id(node)] = t

    def into_temp(self, node):
        """Evaluates a sub-expression and parks it in a temp; returns (temp, transient?)."""
        if id(node) in self.temp_of:
            return self.temp_of[id(node)], False
        self.value(node)
        if id(node) in self.temp_of:
            return self.temp_of[id(node)], False
        t = self._alloc()
        self._emit(_lib.EW_STORE_T, t)
        return t, True

    def emit(self, node):
        spec = ew_spec(node)
        kind = spec[0]
        if kind == "nary":
            op, identity = spec[1], spec[2]
            kids = []
            for ch in node.children:
                if isinstance(ch, ConstFill) and ch.fill == identity and ch._host is None and not ch._dev_override:
                    if tuple(ch.shape) != ():
                        self.shapes.append(tuple(ch.shape))      # a block of ones/zeros only contributes its shape
                    continue
                s = self.scalar_of(ch)
                if s is not None and s == identity:
                    continue
                kids.append(ch)
            if not kids:
                self._emit(_lib.EW_LOAD | IMM, 0, float(identity))
                return
            # Nested operands other than the first are evaluated into temps BEFORE the fold starts, so the
            # left-to-right association ((k0 op k1) op k2) is kept; the first operand is evaluated last and
            # stays in acc (saves one temp and a STORE/LOAD pair).
            resolved, transient = [None] * len(kids), []
            for i in range(len(kids) - 1, 0, -1):
                o = self.operand(kids[i])
                if o[0] == "sub":
                    t, tr = self.into_temp(kids[i])
                    o = ("tmp", t)
                    if tr:
                        transient.append(t)
                resolved[i] = o
            resolved[0] = self.operand(kids[0])
            for i in range(1, len(kids)):            # sharing may have turned an input into a temp meanwhile
                if resolved[i][0] != "tmp":
                    resolved[i] = self.operand(kids[i])
            if resolved[0][0] == "sub":
                self.value(kids[0])
            else:
                self._apply(_lib.EW_LOAD, resolved[0])
            for o in resolved[1:]:
                self._apply(op, o)
            for t in transient:
                self._free(t)
        elif kind in ("unary", "alias"):
            o = self.operand(node.children[0])
            if o[0] == "sub":
                self.value(o[1])
            else:
                self._apply(_lib.EW_LOAD, o)
            if kind == "unary":
                for ins in spec[1]:
                    self._emit(*ins)
        elif kind == "pow":
            transient = None
            expo = self.operand(node.children[1])
            if expo[0] == "sub":
                t, tr = self.into_temp(node.children[1])
                expo = ("tmp", t)
                transient = t if tr else None
            base = self.operand(node.children[0])
            if base[0] == "sub":
                self.value(node.children[0])
            else:
                self._apply(_lib.EW_LOAD, base)
            self._apply(_lib.EW_POW, expo)
            if transient is not None:
                self._free(transient)
        else:
            raise FusionOverflow("unknown spec")

    def finish(self):
        self.emit(self.root)
        self._emit(_lib.EW_STORE_OUT, 0)
        out_shape = broadcast_shapes(*self.shapes) if self.shapes else ()
        return self.code, self.inputs, out_shape


def run_group(root, members, dev_of, scalar_of, run_program):
    """Executes one fusion group; returns the root's DeviceArray.  Raises FusionOverflow if it does not fit."""
    em = Emitter(root, members, dev_of, scalar_of)
    code, inputs, out_shape = em.finish()
    # pure pass-through of one operand of the right shape: alias, no kernel
    if len(code) == 2 and (code[0][0] == (_lib.EW_LOAD | IN)) and inputs[0].shape == out_shape:
        return inputs[0]
    return run_program(code, inputs, out_shape)
