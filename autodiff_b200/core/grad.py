"""`grad` - the graph -> graph reverse-mode transform of reference autodiff/core/grad.py:9-38 - plus a structural memo
that takes its Python cost off the critical path of training loops.

Loops written against the reference call `ad.grad` every step on an ISOMORPHIC graph (examples/feedforward_mnist.py:27,
examples/char_rnn.py:60); for the second-order char-rnn the two nested calls build 70 k Python objects (0.28 s) that the
launch tape then never looks at.  So, from the third call on a structure seen before, `grad` returns `DeferredGrad` nodes:
one node per requested gradient that knows (top, wrt list, index, shape) and EXPANDS into the reference's gradient graph
only if somebody walks into it (`.children`).  Evaluating deferred gradients whose structure has a launch tape
(core/tape.py) replays the tape straight from the leaves of the forward graph - same kernels, same operands, same order,
bit-identical - and never builds the gradient graph; nested calls (grad of a deferred grad) stay deferred.  Anything the
memo does not know (first sightings, interior `wrt` nodes, an explicit previous_grad, untapeable graphs) takes the plain path,
which is exactly the reference's algorithm.  `ADB_LAZY_GRAD=0` disables the memo."""
import collections
import os

from .node import ConstFill, Node, Variable
from .ops import Add  # noqa: F401  (grad accumulates with `+=`, i.e. Add nodes, as the reference does)
from .utils import gc_paused, reverse_topo_sort

LAZY = os.environ.get("ADB_LAZY_GRAD", "1") != "0"
LAZY_AFTER = 1          # plain (expanded) result for the first sighting of a structure; deferred nodes from then on
_known = {}             # request key -> [sightings, ((shape, name), ...)]
stats = {"deferred": 0, "expanded": 0, "plain": 0}


def grad(top_node, wrt_list, previous_grad=None):
    """Graph -> graph reverse-mode transform, semantics of reference autodiff/core/grad.py:9-38.

    Returns, for every node in wrt_list, a NEW graph whose evaluation is d(top_node)/d(node); because partial
    derivatives are themselves graphs, grad(grad(...)) works.  No arithmetic happens here (except the eager
    evaluation inside Slice's partial derivative, a reference quirk kept on purpose).

    :param previous_grad: incoming gradient of top_node; default ones(top_node.shape) (never materialised)
    """
    assert isinstance(wrt_list, list) or isinstance(wrt_list, tuple)
    if LAZY and previous_grad is None:
        req = _request(top_node, wrt_list)
        if req is not None:
            ent = _known.get(req.key)
            if ent is not None and ent[0] >= LAZY_AFTER:
                stats["deferred"] += 1
                return [DeferredGrad(req, i, shape, name) for i, (shape, name) in enumerate(ent[1])]
            out = _expand(top_node, wrt_list, None)
            if ent is None:
                if len(_known) > 256:
                    _known.clear()
                _known[req.key] = [1, tuple((g.shape, g.name) for g in out)]
            else:
                ent[0] += 1
            return out
    return _expand(top_node, wrt_list, previous_grad)


def _expand(top_node, wrt_list, previous_grad):
    stats["plain"] += 1
    if previous_grad is None:
        previous_grad = ConstFill(tuple(top_node.shape), 1.0, name=add_sum_name(top_node))

    with gc_paused():
        dct = collections.defaultdict(lambda: Variable(0))
        dct[top_node] += previous_grad

        for node in reverse_topo_sort(top_node):
            for child in set(node.children):
                dct[child] += node.partial_derivative(wrt=child, previous_grad=dct[node])

        return [dct[wrt] for wrt in wrt_list]


def add_sum_name(node):
    return "'" + node.name + "' grad_sum"


# ------------------------------------------------------------------------------------------------ deferred gradients
class _Request:
    """One `grad(top, wrt_list)` call: the key of its structure and, once somebody needed it, its expansion."""
    __slots__ = ("top", "wrt", "key", "real")

    def __init__(self, top, wrt, key):
        self.top, self.wrt, self.key, self.real = top, list(wrt), key, None

    def expand(self):
        if self.real is None:
            stats["expanded"] += 1
            self.real = _expand(self.top, self.wrt, None)
        return self.real


def _request(top, wrt_list):
    """The request for grad(top, wrt_list) if its structure can be keyed, else None."""
    from . import engine, tape
    if not tape.ENABLED:
        return None
    if type(top) is DeferredGrad and top._real is None:
        # grad of a deferred grad: keyed on the inner request; the wrt nodes must be leaves the inner request knows
        base = top._req
        pos = []
        for w in wrt_list:
            p = _leaf_position(base, w)
            if p is None:
                return None
            pos.append(p)
        return _Request(top, wrt_list, ("g", base.key, top._index, tuple(pos)))
    scanned = tape.scan([top] + list(wrt_list), engine.scalar_of, engine._has_value, engine._is_device_node)
    if scanned is None:
        return None
    order, leaves, sig = scanned
    where = {id(n): i for i, n in enumerate(order)}
    pos = []
    for w in wrt_list:
        i = where.get(id(w))
        if i is None or leaves[i] is not w:
            return None                           # an interior node as wrt: plain path
        pos.append(i)
    return _Request(top, wrt_list, ("n", Node.epsilon, sig, tuple(pos)))


def _leaf_position(req, node):
    """Stable description of a leaf relative to the forward graph at the bottom of a request chain: (depth, index in the wrt
    list of that request), or None.  (The wrt nodes are the only leaves a nested request is allowed to name.)"""
    depth = 0
    while True:
        for i, w in enumerate(req.wrt):
            if w is node:
                return (depth, i)
        if type(req.top) is DeferredGrad and req.top._real is None:
            req = req.top._req
            depth += 1
        else:
            return None


class DeferredGrad(Node):
    """d(top)/d(wrt[index]) as ONE node; behaves as Identity(the reference's gradient graph) once expanded."""
    _scalar_as_array = True        # (the reference's result is an Add node, ops.py:53)

    def __init__(self, req, index, shape, name):
        self._req, self._index, self._real = req, index, None
        self._name = name
        self._host = None
        self._dev = None
        self._dev_override = False
        self.shape = shape
        self.context_list = Node.context_list.copy()
        self.id = Node.id
        Node.id += 1

    @property
    def name(self):
        return self._name

    @name.setter
    def name(self, n):
        self._name = n

    @property
    def real(self):
        if self._real is None:
            self._real = self._req.expand()[self._index]
        return self._real

    @property
    def children(self):
        return [self.real]

    @children.setter
    def children(self, value):
        raise AttributeError("a deferred gradient has exactly one child: its expansion")

    def _ew_spec(self):
        return ("alias",)

    def _eval_device(self):
        from . import engine
        return engine.dev(self.real)

    def _partial_derivative(self, wrt, previous_grad):
        if self.real == wrt:
            return previous_grad
        return 0


from . import tape as _tape  # noqa: E402
_tape.register(DeferredGrad)          # once expanded it is a plain one-child alias op


# ------------------------------------------------------------------------------------------------ evaluation without expansion
# (called by core/engine.py evaluate_many)
_memo = {}              # evaluation key -> (tape, order_map, number of nodes of the expanded scan)


class _Ctx:
    __slots__ = ("key", "fwd_order", "tops")


def _top_desc(node, where):
    """Canonical description of one output relative to the forward scan: a forward node, or a (nested) deferred gradient."""
    if type(node) is DeferredGrad and node._real is None:
        req = node._req
        wpos = []
        for w in req.wrt:
            i = where.get(id(w))
            if i is None:
                return None
            wpos.append(i)
        inner = _top_desc(req.top, where)
        if inner is None:
            return None
        return ("g", node._index, tuple(wpos), inner)
    i = where.get(id(node))
    return None if i is None else ("n", i)


def lazy_context(tops, extra_key):
    """If `tops` contains unexpanded deferred gradients: the forward scan they hang off and the key of this evaluation."""
    from . import engine, tape
    roots, seen = [], set()

    def add(n):
        if id(n) not in seen:
            seen.add(id(n))
            roots.append(n)
    any_deferred = False
    for t in tops:
        n = t
        while type(n) is DeferredGrad and n._real is None:
            any_deferred = True
            for w in n._req.wrt:
                add(w)
            n = n._req.top
        add(n)
    if not any_deferred:
        return None
    # forward roots first (their order is the order the expanded graph's evaluation visits shared nodes in)
    scanned = tape.scan(roots, engine.scalar_of, engine._has_value, engine._is_device_node)
    if scanned is None:
        return None
    order, leaves, sig = scanned
    where = {id(n): i for i, n in enumerate(order)}
    descs = []
    for t in tops:
        d = _top_desc(t, where)
        if d is None:
            return None
        descs.append(d)
    ctx = _Ctx()
    ctx.key = (extra_key, sig, tuple(descs))
    ctx.fwd_order = order
    ctx.tops = tops
    return ctx


def lazy_replay(ctx, dev_of, add_pending_flag, on_ready):
    """Replays the launch tape of this evaluation straight from the forward graph's leaves.  False = not known (yet)."""
    from . import tape
    ent = _memo.get(ctx.key)
    if ent is None:
        return False
    tp, order_map, n = ent
    order, leaves = [None] * n, [None] * n
    fwd, tops = ctx.fwd_order, ctx.tops
    for i, kind, a, is_leaf in order_map:
        node = fwd[a] if kind == 0 else (tops[a] if kind == 1 else a)
        order[i] = node
        if is_leaf:
            leaves[i] = node
    if not tape.replay(tp, order, leaves, tops, dev_of, add_pending_flag, on_ready):
        return False
    stats["lazy_replays"] = stats.get("lazy_replays", 0) + 1
    return True


def lazy_learn(ctx, order, leaves, tp):
    """After the plain path evaluated (and taped) the EXPANDED graph of this evaluation: remember how the expanded scan maps
    onto the forward scan, so that the next evaluation of the same key can skip the expansion."""
    if ctx.key in _memo or tp is None:
        return
    import numpy as np
    where = {id(n): j for j, n in enumerate(ctx.fwd_order)}
    top_at = {id(t): k for k, t in enumerate(ctx.tops)}
    order_map, mapped = [], set()
    for i, node in enumerate(order):
        is_leaf = leaves[i] is not None
        j = where.get(id(node))
        if j is not None:
            order_map.append((i, 0, j, is_leaf))
        elif id(node) in top_at:
            order_map.append((i, 1, top_at[id(node)], is_leaf))
        elif is_leaf:
            # A leaf created by the expansion itself.  Only constants that the STRUCTURE determines may be reused from the memo:
            # never-materialised fills, scalars folded into kernels, tiny literal arrays (Softmax's np.array([1]), ops.py:428).
            # Anything else (the eagerly evaluated Slice gradient of reshape.py:96-102 is a data-dependent constant) keeps
            # this evaluation on the plain path.
            ok = (isinstance(node, ConstFill) and node.fill is not None and node._host is None) or \
                 (isinstance(node, Variable) and node.host_scalar() is not None) or \
                 (type(node) is Variable and isinstance(node._value, np.ndarray) and node._value.size <= 16 and node._host is None
                  and not node._dev_override)
            if not ok:
                return
            order_map.append((i, 2, node, True))
        else:
            continue      # interior node of the gradient graph: no object on the lazy path, its value stays anonymous in the arena
        mapped.add(i)
    if any(i not in mapped for i in tp.leaf_ext):
        return
    if len(_memo) > 64:
        _memo.clear()
    _memo[ctx.key] = (tp, order_map, len(order))
