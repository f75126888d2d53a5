"""Structure-keyed launch tapes: replaying a graph that is evaluated again and again.

Training loops written against the reference rebuild an ISOMORPHIC graph every step (examples/feedforward_mnist.py
builds `nn(inp)`, the loss and `ad.grad` inside the loop), and for the launch-bound configs the per-node Python of
planning, fusing and marshalling dominates (C1: 25 kernels of a few microseconds each; C5: 44 k nodes, 10 k kernels).
The tape removes that work for every evaluation after the second one of a given structure:

  scan     one pre-order walk of the not-yet-evaluated sub-DAG produces a structural signature - node classes, their
           static attributes (einsum subscripts, axes, slices ...), leaf kinds/shapes/strides, folded scalar values and
           the sharing pattern (back-references) - so equal signatures mean "same kernels, same operand wiring";
  record   the 2nd evaluation of a signature runs the normal planner/executor while every C-ABI compute call (with its
           ctypes argument structs) and every device allocation is logged; each device pointer in those arguments is
           resolved to (buffer, byte offset), where a buffer is a leaf's storage, a persistent constant or a slot of the
           step's arena;
  replay   later evaluations upload the leaves, allocate ONE arena, patch the pointers in the recorded argument structs
           and re-issue the calls (the C++ side re-plans tiles / access modes from the actual pointers, so alignment or
           workspace differences are handled there), then hand every materialised node a `DeviceArray` view of the arena.

  graph    from the 2nd replay on, a tape whose leaves are small (<= 64 MB) is replayed as ONE CUDA graph: the recorded calls
           are issued once more on a side stream in capture mode (C ABI adb_graph_begin / adb_graph_end) against a buffer
           the tape keeps - [copies of the leaf storages | arena] - and every later replay is: copy the leaves into their
           staging area, one `cudaGraphLaunch`.  The launch-bound configs (C1: 25 kernels, C5: 10 k kernels) then cost one
           driver call per step instead of one ctypes call + planner pass per kernel.  The buffer is reused only when no
           `DeviceArray` of the previous replay is still alive (refcount of the buffer tensor); otherwise, and whenever
           capture fails, the replay falls back to the call-by-call path above.  `ADB_TAPE_GRAPH=0` disables it.

Results are bit-identical to the normal path: the same kernels run on the same operands in the same order.  Graphs with
user-defined numpy nodes, unknown node classes or unresolvable pointers are simply never taped.  `ADB_TAPE=0` disables it.
"""
import bisect
import ctypes as C
from math import copysign as _copysign
import os
import sys

import numpy as np

from .. import _lib
from .. import device as _device
from ..device import DeviceArray, torch
from .node import ConstFill, Node, Variable

ENABLED = os.environ.get("ADB_TAPE", "1") != "0"
RECORD_AT = 2            # record on the 2nd sighting of a structure, replay from the 3rd
# ADB_TAPE_SELFCHECK=1 (test-suite stress mode): record on the FIRST sighting, then immediately replay the fresh tape onto
# the same graph and require every materialised node to come out bit-identical to the recorded run.
SELF_CHECK = os.environ.get("ADB_TAPE_SELFCHECK", "0") == "1"
if SELF_CHECK:
    RECORD_AT = 1
MAX_TAPES = 64
GRAPH = os.environ.get("ADB_TAPE_GRAPH", "1") != "0"
GRAPH_AFTER_REPLAYS = int(os.environ.get("ADB_TAPE_GRAPH_AFTER", "1"))   # call-by-call replays before a tape is captured (0 with
GRAPH_MIN_CALLS = int(os.environ.get("ADB_TAPE_GRAPH_MIN_CALLS", "4"))   # ADB_TAPE_SELFCHECK=1: every graph of the test-suite)
GRAPH_STREAMS = max(1, int(os.environ.get("ADB_TAPE_GRAPH_STREAMS", "4")))   # capture streams: independent calls of a tape become parallel
                                                                             # branches of the CUDA graph (1 = one linear chain)
# Which arguments of a recorded entry point are WRITTEN (include/adb200.h); every other tensor / pointer argument is read.  A tape
# that holds a call missing here is captured as one chain.
_OUT_ARGS = {"adb_ewise": (5,), "adb_einsum": (3,), "adb_reduce_sum": (1,), "adb_sum_squares": (1,), "adb_norm2": (1,),
             "adb_softmax_ce": (2,), "adb_im2col": (3,), "adb_col2im": (3,), "adb_conv_gemm": (5,), "adb_memcpy_d2d": (0,),
             "adb_memset_async": (0,)}
GRAPH_MAX_STAGE_BYTES = 64 << 20       # leaves are copied next to the arena on every replay: only worth it when they are small
# All graph buffers together (they outlive the step): 64 GB of the 180 GB.  The second-order char-rnn's arena is 26 GB; while
# its Python graph build dominated the evaluation its graph measured a wash (profiles/r01_graph_ab.log), but with deferred
# gradients (core/grad.py) the 9.5 k ctypes calls of a call-by-call replay ARE the host time, and one graph launch removes them.
GRAPH_MAX_TOTAL_BYTES = int(os.environ.get("ADB_TAPE_GRAPH_BYTES", str(64 << 30)))
SCE_ERROR = "Labels must be a valid probability distribution!"

_seen = {}               # signature -> sightings so far
_tapes = {}              # signature -> Tape | False (not tapeable)
_recorder = None         # active Recorder (set while a recording evaluation runs)
stats = {"replays": 0, "records": 0, "rejected": 0, "graphs": 0, "graph_launches": 0, "graph_rejected": 0, "graph_busy": 0, "graph_bytes": 0}
_cap_stream = None       # side stream the graphs are captured on (the legacy default stream cannot be captured)
_cap_side = []           # further capture streams: parallel branches of a graph (GRAPH_STREAMS)

# node classes whose device evaluation is a pure function of (class, these attributes, children): filled by register()
_ATTRS = {}


def register(cls, *attrs):
    _ATTRS[cls] = attrs


def _freeze(v):
    if isinstance(v, (list, tuple)):
        return tuple(_freeze(x) for x in v)
    if isinstance(v, slice):
        return ("slice", v.start, v.stop, v.step)
    if isinstance(v, np.ndarray):
        return ("nd", v.shape, v.tobytes()) if v.size <= 16 else None
    if isinstance(v, (np.integer, np.floating)):
        return v.item()
    return v


_CLASS_KIND = {}         # class -> (is a Variable subclass, tapeable device op, attribute names, is exactly Variable)   [filled on first sight]
_PLAIN_SIG = {}          # (class, number of children) -> the signature entry of an attribute-less op (one shared tuple)
_ATOMS = (str, int, float, bool, type(None))


def _class_kind(cls, is_device_node, probe):
    attrs = _ATTRS.get(cls)
    k = _CLASS_KIND[cls] = (issubclass(cls, Variable), attrs is not None and is_device_node(probe), attrs, cls is Variable)
    return k


def scan(tops, scalar_of, has_value, is_device_node):
    """Pre-order walk of the sub-DAG that an evaluation of `tops` would touch.
    Returns (order, leaves, signature) or None when the graph cannot be taped.
    (This loop runs over every node of every evaluation - 44 k nodes for the second-order char-rnn - so the helpers
    `has_value` / `scalar_of` / `is_device_node` of the engine are inlined for the common cases.)"""
    pos, order, leaves, sig = {}, [], [], []
    stack = list(reversed(tops))
    kinds, plain = _CLASS_KIND, _PLAIN_SIG
    ndarray = np.ndarray
    pop, push, get_pos = stack.pop, stack.extend, pos.get
    add_sig, add_leaf, add_order = sig.append, leaves.append, order.append
    while stack:
        n = pop()
        key = id(n)
        p = get_pos(key)
        if p is not None:
            add_sig(p)                           # sharing: a back-reference to an already visited node
            continue
        pos[key] = len(order)
        add_order(n)
        cls = type(n)
        kind = kinds.get(cls)
        if kind is None:
            kind = _class_kind(cls, is_device_node, n)
        if kind[0] or n._dev is not None or n._host is not None:
            # ---- a leaf of this evaluation: a Variable, or a node that already has a value
            if kind[0]:
                if kind[3] and n._host is None and not n._dev_override and type(n._value) is ndarray:
                    v = n._value
                    s = float(v) if v.ndim == 0 else None           # == Variable.host_scalar() for the common case
                else:
                    s = n.host_scalar()                             # (subclasses: ConstFill._value would MATERIALISE the constant)
                if s is not None:
                    # folded into programs as an immediate: the VALUE is structure (-0.0 == 0.0 and hashes alike, so the sign
                    # of a zero is keyed separately; NaN != NaN only costs a tape miss)
                    add_sig(("s", s) if s != 0.0 else ("s", s, _copysign(1.0, s)))
                    add_leaf(n)                  # (uploaded only if some kernel reads it as a tensor, see build_tape)
                    continue
            d = n._dev
            if cls is ConstFill and n.fill is not None and n._host is None:
                # keyed by the fill VALUE even when a one-element device copy already exists: Einsum drops operands that
                # are all ones (ops.py Einsum._eval_device), so the value decides which kernels run
                add_sig(("c", n.fill, n._fill_shape, d is not None))
            elif d is not None:
                if type(d) is not DeviceArray:
                    return None
                add_sig(("d", d.shape, d.strides))
            else:
                v = n._host if n._host is not None else getattr(n, "_value", None)
                if isinstance(v, DeviceArray):
                    add_sig(("d", v.shape, v.strides))
                elif isinstance(v, ndarray):
                    add_sig(("h", v.shape))
                else:
                    return None
            add_leaf(n)
            continue
        add_leaf(None)
        if not kind[1]:
            return None
        nd = n.__dict__
        if "_eval" in nd or "_eval_device" in nd or "_never_tape" in nd:
            return None                          # an instance-level override / a node that decides on the host: not a pure
                                                 # function of (class, attributes) that a replay could reproduce
        attrs = kind[2]
        children = n.children
        if attrs:
            key = []
            for a in attrs:
                v = getattr(n, a)
                if type(v) not in _ATOMS:
                    f = _freeze(v)
                    if f is None and v is not None:
                        return None
                    v = f
                key.append(v)
            add_sig((cls, len(children), tuple(key)))
        else:
            ck = (cls, len(children))
            e = plain.get(ck)
            if e is None:
                e = plain[ck] = ck
            add_sig(e)
        if children:
            push(children[::-1])
    return order, leaves, tuple(sig)


# ------------------------------------------------------------------------------------------------ recording
class Recorder:
    def __init__(self):
        self.allocs = []          # (ptr, nbytes, base tensor): every DeviceArray.empty during the recording (kept alive)
        self.calls = []           # (name, fn, args)
        self.top_ends = []        # number of calls logged when each top finished

    def on_alloc(self, base):
        self.allocs.append((base.data_ptr(), base.numel() * 8, base))

    def on_call(self, name, fn, args):
        # byref(DeviceArray.tensor()) points at a struct CACHED on a live DeviceArray: the tape must own a private copy,
        # or patching pointers at replay time would corrupt that array's descriptor
        own = []
        for a in args:
            obj = getattr(a, "_obj", None)
            if isinstance(obj, _lib.Tensor):
                cp = _lib.Tensor()
                C.memmove(C.byref(cp), C.byref(obj), C.sizeof(_lib.Tensor))
                own.append(C.byref(cp))
            else:
                own.append(a)
        self.calls.append((name, fn, own))


class _Call:
    __slots__ = ("fn", "args", "patches", "void_patches", "flag_arg", "reads", "writes")   # reads / writes: buffer indices, None = unknown


class Tape:
    __slots__ = ("n_nodes", "n_ext", "ext_leaf", "leaf_ext", "abs_keep", "slot_off", "arena_elems", "calls", "top_ends", "values",
                 "ext_size", "replays", "graph")


class _Graph:
    """One captured replay: the cudaGraphExec_t, the buffer its kernels address ([leaf staging | arena]) and its flags."""
    __slots__ = ("exec", "buf", "refs", "stage_off", "arena_off", "flags", "nflags", "last_stream")

    def __del__(self):
        buf = getattr(self, "buf", None)
        if buf is not None:
            stats["graph_bytes"] -= buf.numel() * 8
        ex = getattr(self, "exec", None)
        if ex:
            try:
                _lib.lib().adb_graph_destroy(ex)
            except Exception:
                pass


def _tensor_structs(arg):
    """The adb_tensor structs reachable from one ctypes argument."""
    if isinstance(arg, _lib.Tensor):
        return [arg]
    obj = getattr(arg, "_obj", None)           # C.byref(x)
    if isinstance(obj, _lib.Tensor):
        return [obj]
    if isinstance(arg, C.Array) and getattr(arg, "_type_", None) is _lib.Tensor:
        return [arg[i] for i in range(len(arg))]
    return []


def build_tape(rec, order, leaves, n_tops):
    """Resolves every device pointer of the recording; returns a Tape or None."""
    # --- buffers: [0, n_ext) leaf storages (deduplicated by base pointer), then arena slots
    ext_ptr, ext_leaf, leaf_ext, ext_size = {}, [], {}, []
    for i, leaf in enumerate(leaves):
        if leaf is None:
            continue
        d = leaf._dev
        if d is None:
            continue                             # a scalar that only ever appeared as an immediate: never uploaded
        b = d.base.data_ptr()
        j = ext_ptr.get(b)
        if j is None:
            j = ext_ptr[b] = len(ext_leaf)
            ext_leaf.append(i)
            ext_size.append(d.base.numel() * 8)
        leaf_ext[i] = (j, d.offset)
    n_ext = len(ext_leaf)
    ranges = []                                  # (start, end, buffer index)
    for b, j in ext_ptr.items():
        ranges.append((b, b + ext_size[j], j))
    slot_off, total = [], 0
    seen_alloc = set()
    for ptr, nbytes, _base in rec.allocs:
        if ptr in ext_ptr:
            continue                             # a leaf uploaded during the recording: its storage is the leaf's, not the arena's
        if ptr in seen_alloc:
            return None                          # address reuse inside one recording: cannot happen while bases are kept alive
        seen_alloc.add(ptr)
        slot_off.append(total)
        ranges.append((ptr, ptr + nbytes, n_ext + len(slot_off) - 1))
        total += (nbytes // 8 + 31) // 32 * 32   # 256-byte aligned slots
    abs_keep = []
    const_ptrs = {t.data_ptr(): t for t in _device._const_cache.values()}
    ranges.sort()
    starts = [r[0] for r in ranges]

    def resolve(ptr):
        k = bisect.bisect_right(starts, ptr) - 1
        if k >= 0 and ranges[k][0] <= ptr < ranges[k][1]:
            return ranges[k][2], ptr - ranges[k][0]
        for cp, t in const_ptrs.items():
            if cp <= ptr < cp + 16:
                abs_keep.append(t)               # persistent broadcast constant: absolute address, kept alive by the tape
                return -1, ptr
        return None

    tape = Tape()
    tape.calls = []
    for name, fn, args in rec.calls:
        c = _Call()
        c.fn, c.args, c.patches, c.void_patches, c.flag_arg = fn, list(args), [], [], -1
        outs = _OUT_ARGS.get(name)
        c.reads, c.writes = (set(), set()) if outs is not None else (None, None)
        for ai, a in enumerate(args):
            for ts in _tensor_structs(a):
                if not ts.ptr:
                    continue
                r = resolve(ts.ptr)
                if r is None:
                    return None
                if r[0] >= 0:
                    c.patches.append((ts, r[0], r[1]))
                    if outs is not None:
                        (c.writes if ai in outs else c.reads).add(r[0])
            if isinstance(a, C.c_void_p) and a.value:
                if name == "adb_softmax_ce" and ai == 3:
                    c.flag_arg = ai              # a fresh zeroed flag per replay
                elif ai == len(args) - 1:
                    pass                         # the stream
                else:
                    r = resolve(a.value)
                    if r is None:
                        return None
                    if r[0] >= 0:
                        c.void_patches.append((ai, r[0], r[1]))
                        if outs is not None:
                            (c.writes if ai in outs else c.reads).add(r[0])
        tape.calls.append(c)
    # --- where every materialised node's value lives
    values = []
    for i, n in enumerate(order):
        if leaves[i] is not None:
            continue
        d = n._dev
        if d is None:
            continue
        r = resolve(d.base.data_ptr() + 8 * d.offset) if d.size > 0 else None
        if r is None or r[0] < 0:
            if d.size > 0 and (r is None):
                return None
            if r is not None and r[0] < 0:
                values.append((i, -1, d, None, None))          # a view of a persistent constant: reuse the array object
                continue
            values.append((i, -2, d, None, None))              # empty array
            continue
        values.append((i, r[0], r[1] // 8, d.shape, d.strides))
    tape.n_nodes = len(order)
    tape.n_ext, tape.ext_leaf, tape.leaf_ext = n_ext, ext_leaf, leaf_ext
    tape.abs_keep = abs_keep
    tape.slot_off, tape.arena_elems = slot_off, total
    tape.top_ends = list(rec.top_ends)
    tape.values = values
    tape.ext_size, tape.replays, tape.graph = ext_size, 0, None
    return tape


# ------------------------------------------------------------------------------------------------ CUDA-graph replay
def plan_streams(calls, k):
    """List scheduling of a tape's calls onto k capture streams from their buffer read / write sets (RAW, WAW and WAR on whole
    buffers: two calls that touch different views of one buffer are ordered - conservative, never wrong).  Returns (stream of each
    call, calls on OTHER streams each call waits for, set of calls that must record an event), or None when some call's sets are
    unknown.  A call goes onto the stream whose tail is one of its predecessors (no event needed), else onto the stream idle longest."""
    if k <= 1 or any(c.reads is None for c in calls):
        return None
    last_writer, readers = {}, {}
    tail = [-1] * k
    assign, waits = [], []
    for i, c in enumerate(calls):
        deps = set()
        for b in c.reads:
            w = last_writer.get(b)
            if w is not None:
                deps.add(w)
        for b in c.writes:
            w = last_writer.get(b)
            if w is not None:
                deps.add(w)
            deps.update(readers.get(b, ()))
        st = -1
        for d in sorted(deps, reverse=True):
            if tail[assign[d]] == d:
                st = assign[d]
                break
        if st < 0:
            st = min(range(k), key=lambda j: tail[j])
        latest = {}
        for d in deps:
            if assign[d] != st and d > latest.get(assign[d], -1):
                latest[assign[d]] = d              # stream order implies the earlier calls of that stream
        waits.append(sorted(latest.values()))
        assign.append(st)
        tail[st] = i
        for b in c.writes:
            last_writer[b] = i
            readers[b] = []
        for b in c.reads:
            if b not in c.writes:
                readers.setdefault(b, []).append(i)
    return assign, waits, {d for w in waits for d in w}


def _capture(tape):
    """Issues the tape's calls against a private buffer: once eagerly on the capture stream (sizes every grow-only
    workspace of the library for that stream, so nothing allocates while recording), once in capture mode."""
    global _cap_stream, _cap_side
    t = torch()
    L = _lib.lib()
    g = _Graph()
    g.exec = None
    g.stage_off, total = [], 0
    for sz in tape.ext_size:
        g.stage_off.append(total)
        total += (sz // 8 + 31) // 32 * 32
    g.arena_off = total
    g.buf = None
    nbytes = 8 * (total + max(tape.arena_elems, 2))
    if stats["graph_bytes"] + nbytes > GRAPH_MAX_TOTAL_BYTES:
        stats["graph_rejected"] += 1              # the buffers of captured tapes stay allocated: bounded in total
        return False
    g.buf = t.empty(total + max(tape.arena_elems, 2), dtype=t.float64, device="cuda")
    stats["graph_bytes"] += nbytes
    g.nflags = sum(1 for c in tape.calls if c.flag_arg >= 0)
    g.flags = t.zeros(max(g.nflags, 1), dtype=t.int32, device="cuda")
    if _cap_stream is None:
        _cap_stream = t.cuda.Stream()
    _cap_stream.wait_stream(t.cuda.current_stream())
    b0, f0 = g.buf.data_ptr(), g.flags.data_ptr()
    bases = [b0 + 8 * o for o in g.stage_off] + [b0 + 8 * (total + o) for o in tape.slot_off]
    sp = C.c_void_p(_cap_stream.cuda_stream)
    check = _lib.check
    # independent calls go to different capture streams and become parallel branches of the graph (fork / join through events)
    plan = plan_streams(tape.calls, GRAPH_STREAMS) if len(tape.calls) >= 8 else None
    if plan is not None:
        while len(_cap_side) < GRAPH_STREAMS - 1:
            _cap_side.append(t.cuda.Stream())
        streams = [_cap_stream] + _cap_side[:GRAPH_STREAMS - 1]
        sps = [C.c_void_p(x.cuda_stream) for x in streams]
        assign, waits, need_event = plan
        if len(set(assign)) == 1:
            plan = None                           # a pure chain

    def issue():
        fi = 0
        events = {}
        if plan is not None:                      # fork: the side streams join the capture behind everything queued so far
            ev = t.cuda.Event()
            ev.record(_cap_stream)
            for x in streams[1:]:
                x.wait_event(ev)
        for i, c in enumerate(tape.calls):
            for ts, j, off in c.patches:
                ts.ptr = bases[j] + off
            args = list(c.args)
            for ai, j, off in c.void_patches:
                args[ai] = C.c_void_p(bases[j] + off)
            if plan is not None:
                k = assign[i]
                for d in waits[i]:
                    streams[k].wait_event(events[d])
                cs = sps[k]
            else:
                cs = sp
            args[-1] = cs                         # every compute entry point takes the stream last
            if c.flag_arg >= 0:
                fp = C.c_void_p(f0 + 4 * fi)
                fi += 1
                check(L.adb_memset_async(fp, 0, 4, cs))      # part of the graph: each launch starts from a clean flag
                args[c.flag_arg] = fp
            check(c.fn(*args))
            if plan is not None and i in need_event:
                ev = events[i] = t.cuda.Event()
                ev.record(streams[assign[i]])
        if plan is not None:                      # join: the origin stream ends behind every branch
            for x in streams[1:]:
                ev = t.cuda.Event()
                ev.record(x)
                _cap_stream.wait_event(ev)
    try:
        issue()
        check(L.adb_sync(sp))
        check(L.adb_graph_begin(sp))
        ex = C.c_void_p()
        try:
            issue()
        except Exception:
            L.adb_graph_end(sp, C.byref(ex))      # leave capture mode; the partial recording is dropped
            if ex.value:
                L.adb_graph_destroy(ex)
            raise
        check(L.adb_graph_end(sp, C.byref(ex)))
    except (_lib.BackendError, ValueError, RuntimeError):
        stats["graph_rejected"] += 1
        return False
    g.exec = ex
    g.last_stream = None
    g.refs = sys.getrefcount(g.buf)
    stats["graphs"] += 1
    if plan is not None:
        stats["graph_branches"] = stats.get("graph_branches", 0) + len(set(assign))
    return g


def _graph_for(tape, base_tensors):
    """The tape's graph if this replay may use it, else None (call-by-call replay)."""
    g = tape.graph
    if g is False or not GRAPH:
        return None
    for j, b in enumerate(base_tensors):
        if b.numel() * 8 != tape.ext_size[j]:
            return None                           # a leaf that is now a view of a differently sized array
    if g is None:
        if tape.replays <= GRAPH_AFTER_REPLAYS or len(tape.calls) < GRAPH_MIN_CALLS or sum(tape.ext_size) > GRAPH_MAX_STAGE_BYTES:
            return None
        g = tape.graph = _capture(tape)
        if g is False:
            return None
    if sys.getrefcount(g.buf) > g.refs:
        stats["graph_busy"] += 1                  # a value of the previous replay is still alive in the buffer
        return None
    return g


def _launch_graph(tape, g, order, bases, base_tensors, tops, add_pending_flag, on_ready):
    L = _lib.lib()
    check = _lib.check
    t = torch()
    if g.nflags:
        from . import engine
        if engine._pending_flags:
            engine.check_pending_flags()          # the flags of the previous launch are about to be re-zeroed
    sp = _device.stream_ptr()
    st = C.c_void_p(sp)
    if g.last_stream is not None and g.last_stream != sp:
        check(L.adb_sync(C.c_void_p(g.last_stream)))      # the buffer is shared by all launches: never two in flight on two streams
    g.last_stream = sp
    b0 = g.buf.data_ptr()
    for j, off in enumerate(g.stage_off):
        check(L.adb_memcpy_d2d(C.c_void_p(b0 + 8 * off), C.c_void_p(bases[j]), tape.ext_size[j], st))
    check(L.adb_graph_launch(g.exec, st))
    n_ext, slot_off, a0, buf = tape.n_ext, tape.slot_off, g.arena_off, g.buf
    make = DeviceArray._make
    for i, j, off, shape, strides in tape.values:
        n = order[i]
        if n is None:
            continue                              # (deferred-gradient replay: interior nodes have no object)
        if j >= n_ext:
            n._dev = make(buf, a0 + slot_off[j - n_ext] + off, shape, strides)
        elif j >= 0:
            n._dev = make(base_tensors[j], off, shape, strides)
        else:
            n._dev = off
    for k in range(g.nflags):
        add_pending_flag(g.flags[k:k + 1], ValueError(SCE_ERROR))
    if on_ready is not None:
        for top in tops:
            on_ready(top)
    stats["replays"] += 1
    stats["graph_launches"] += 1
    return True


# ------------------------------------------------------------------------------------------------ replay
def replay(tape, order, leaves, tops, dev_of, add_pending_flag, on_ready):
    """Re-issues the recorded calls for a new, isomorphic graph.  Returns False (nothing done) when the leaves of the
    new graph alias differently from the recorded ones."""
    t = torch()
    for i in tape.leaf_ext:
        if leaves[i]._dev is None:
            dev_of(leaves[i])                     # upload (host leaf) or materialise (constant / device value)
    bases, base_tensors = [0] * tape.n_ext, [None] * tape.n_ext
    for j, i in enumerate(tape.ext_leaf):
        d = leaves[i]._dev
        bases[j] = d.base.data_ptr()
        base_tensors[j] = d.base
    for i, (j, off) in tape.leaf_ext.items():
        d = leaves[i]._dev
        if d.base.data_ptr() != bases[j] or d.offset != off:
            return False                          # e.g. two leaves that were views of one array are now separate arrays
    tape.replays += 1
    g = _graph_for(tape, base_tensors)
    if g is not None:
        return _launch_graph(tape, g, order, bases, base_tensors, tops, add_pending_flag, on_ready)
    arena = t.empty(max(tape.arena_elems, 2), dtype=t.float64, device="cuda")
    a0 = arena.data_ptr()
    for off in tape.slot_off:
        bases.append(a0 + 8 * off)
    n_ext = tape.n_ext
    # values first: on_ready callbacks (gradient all-reduce) use node._dev as soon as a top's kernels are queued
    slot_off = tape.slot_off
    make = DeviceArray._make
    for i, j, off, shape, strides in tape.values:
        n = order[i]
        if n is None:
            continue                              # (deferred-gradient replay: interior nodes have no object)
        if j >= n_ext:
            n._dev = make(arena, slot_off[j - n_ext] + off, shape, strides)
        elif j >= 0:
            n._dev = make(base_tensors[j], off, shape, strides)
        else:
            n._dev = off                          # recorded DeviceArray (persistent constant / empty)
    check = _lib.check
    ends = tape.top_ends
    ti = 0
    st = C.c_void_p(_device.stream_ptr())         # the CURRENT stream (ad.use_stream), not the one of the recording
    while on_ready is not None and ti < len(ends) and ends[ti] == 0:
        on_ready(tops[ti])
        ti += 1
    for k, c in enumerate(tape.calls):
        for ts, j, off in c.patches:
            ts.ptr = bases[j] + off
        args = c.args
        for ai, j, off in c.void_patches:
            args[ai] = C.c_void_p(bases[j] + off)
        args[-1] = st                             # every compute entry point takes the stream last
        if c.flag_arg >= 0:
            flag = t.zeros(1, dtype=t.int32, device="cuda")
            args[c.flag_arg] = C.c_void_p(flag.data_ptr())
            check(c.fn(*args))
            add_pending_flag(flag, ValueError(SCE_ERROR))
        else:
            check(c.fn(*args))
        while on_ready is not None and ti < len(ends) and ends[ti] == k + 1:
            on_ready(tops[ti])
            ti += 1
    while on_ready is not None and ti < len(tops):
        on_ready(tops[ti])
        ti += 1
    stats["replays"] += 1
    return True


def self_check(tape, order, leaves, tops, dev_of, add_pending_flag):
    """Replays a freshly recorded tape onto the graph it was recorded from and compares every node value bitwise."""
    from ..device import to_host
    want = {}
    for i, n in enumerate(order):
        if leaves[i] is None and n._dev is not None:
            want[i] = (to_host(n._dev, pinned=False), n._dev)
    for i in want:
        order[i]._dev = None
    if not replay(tape, order, leaves, tops, dev_of, add_pending_flag, None):
        for i, (_, d) in want.items():
            order[i]._dev = d
        return
    for i, (host, d) in want.items():
        n = order[i]
        got = to_host(n._dev, pinned=False) if n._dev is not None else None
        if got is None or got.shape != host.shape or not np.array_equal(got, host, equal_nan=True):
            raise AssertionError("launch tape self-check: node %d (%s %s) differs between the recorded run and its replay"
                                 % (i, type(n).__name__, getattr(n, "name", "")))
    stats["self_checks"] = stats.get("self_checks", 0) + 1


def lookup(sig):
    """(tape | None, should_record)."""
    tp = _tapes.get(sig)
    if tp:
        return tp, False
    if tp is False:
        return None, False
    n = _seen.get(sig, 0) + 1
    if len(_seen) > 4 * MAX_TAPES:
        _seen.clear()
    _seen[sig] = n
    return None, n >= RECORD_AT


def store(sig, tape):
    if len(_tapes) >= MAX_TAPES:
        _tapes.clear()
    _tapes[sig] = tape if tape is not None else False
    stats["records" if tape is not None else "rejected"] += 1


def clear():
    _seen.clear()
    _tapes.clear()
    g = sys.modules.get(__package__ + ".grad")
    if g is not None:                             # the deferred-gradient memo holds tapes too
        g._known.clear()
        g._memo.clear()
