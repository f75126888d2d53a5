"""Iterative evaluator: walks the DAG post-order and runs each node's device kernel exactly once.

Replaces the recursive `child()` calls inside every reference `_eval` (e.g. autodiff/core/ops.py:173) — same
memoisation contract as reference node.py:42-46, no recursion limit, values stay in HBM between nodes."""
import ctypes as C

import numpy as np

from .. import _lib
from ..device import DeviceArray, stream_ptr, to_device, torch
from .node import Node, Variable

fusion_enabled = True   # set False to run one kernel per node (debugging / A-B measurements)
kernel_timings = None   # bench.py sets this to a list; Einsum then records (description, start, stop) CUDA events
_pending_flags = []   # (torch int32 tensor, exception) raised at the next host synchronisation


def _is_device_node(node):
    if "_eval" in node.__dict__:
        return False
    cls_eval = type(node)._eval
    return cls_eval is Node._eval or isinstance(node, Variable)


def _has_value(node):
    return node._dev is not None or node._host is not None


def scalar_of(node):
    """Host-known 0-d constant of a leaf, else None."""
    if isinstance(node, Variable):
        return node.host_scalar()
    return None


class _Plan:
    __slots__ = ("groups", "absorbed")

    def __init__(self):
        self.groups = {}      # id(root) -> {id(member): member}
        self.absorbed = {}    # id(member) -> root


def _plan(tops):
    """Finds the not-yet-evaluated sub-DAG under `tops`, counts consumers and forms elementwise fusion groups."""
    from . import fusion
    plan = _Plan()
    top_ids = {id(t) for t in tops}
    order, consumers, seen = [], {}, set()
    stack = [(t, False) for t in reversed(tops)]
    while stack:
        node, done = stack.pop()
        if done:
            order.append(node)
            continue
        if id(node) in seen or _has_value(node):
            continue
        seen.add(id(node))
        stack.append((node, True))
        if not _is_device_node(node):
            continue          # opaque user op: pulls its own inputs through child()
        for c in node.children:
            if _has_value(c) or scalar_of(c) is not None:
                continue
            consumers.setdefault(id(c), []).append(node)
            if id(c) not in seen:
                stack.append((c, False))
    if not fusion_enabled:
        return plan
    for node in reversed(order):          # consumers before producers
        if id(node) in plan.absorbed or fusion.ew_spec(node) is None:
            continue
        shape = fusion.norm_shape(node.shape)
        if shape is None:
            continue
        members = {id(node): node}
        frontier, deferred = list(node.children), []
        while frontier:
            x = frontier.pop()
            k = id(x)
            if k in members or k in plan.absorbed or k in top_ids or k not in consumers:
                continue
            if _has_value(x) or fusion.ew_spec(x) is None or fusion.norm_shape(x.shape) != shape:
                continue
            if len(members) >= fusion.MAX_MEMBERS:
                continue
            if not all(id(c) in members for c in consumers[k]):
                deferred.append(x)
                continue
            members[k] = x
            frontier.extend(x.children)
            frontier.extend(deferred)
            deferred = []
        if len(members) > 1:
            plan.groups[id(node)] = members
            for k, m in members.items():
                if m is not node:
                    plan.absorbed[k] = node
    return plan


def _group_inputs(members):
    deps = []
    for m in members.values():
        for c in m.children:
            if id(c) not in members and not _has_value(c) and scalar_of(c) is None:
                deps.append(c)
    return deps


def _execute(top, plan):
    from . import fusion
    stack = [top]
    while stack:
        node = stack[-1]
        if _has_value(node):
            stack.pop()
            continue
        if not _is_device_node(node):
            # user-defined numpy op (or ad.checkpoint's opaque node): its _eval pulls its own inputs through child()
            node._host = node._eval()
            stack.pop()
            continue
        members = plan.groups.get(id(node))
        if members is not None:
            pending = _group_inputs(members)
        else:
            pending = [c for c in node.children if not _has_value(c) and scalar_of(c) is None]
        if pending:
            stack.extend(reversed(pending))
            continue
        if members is not None:
            try:
                node._dev = fusion.run_group(node, members, dev, scalar_of, run_program)
            except fusion.FusionOverflow:
                del plan.groups[id(node)]          # does not fit one program: evaluate the members one by one
                for k in members:
                    plan.absorbed.pop(k, None)
                continue
        else:
            node._dev = node._eval_device()
        stack.pop()


def evaluate_many(tops, on_ready=None):
    """Evaluates several graph outputs with one fusion plan (consumer counts span all of them).
    `on_ready(node)` fires as soon as each output's kernels are queued (used to overlap all-reduces)."""
    tops = list(tops)
    plan = _plan([t for t in tops if not _has_value(t)])
    for t in tops:
        _execute(t, plan)
        if on_ready is not None:
            on_ready(t)


def evaluate(top):
    evaluate_many([top])


def dev(node):
    """Device value of an already-evaluated (or leaf) node, uploading a host value if that is all there is."""
    if node._dev is None:
        if node._host is not None:
            node._dev = to_device(np.asarray(node._host))
        else:
            evaluate(node)
            if node._dev is None:
                node._dev = to_device(np.asarray(node._host))
    return node._dev


def add_pending_flag(flag_tensor, exc):
    _pending_flags.append((flag_tensor, exc))


def check_pending_flags():
    """Called after every device->host synchronisation: surfaces deferred device-side validation errors
    (SoftmaxCEWithLogits label check, reference ops.py:246-248) as the reference's ValueError."""
    if not _pending_flags:
        return
    flags = list(_pending_flags)
    del _pending_flags[:]
    for flag, exc in flags:
        if int(flag.item()) != 0:
            raise exc


# ------------------------------------------------------------------------------------------- program helper
_prog_cache = {}


def _program(code):
    """ctypes instruction array for a program, cached by content (programs repeat every training step)."""
    key = tuple(code)
    prog = _prog_cache.get(key)
    if prog is None:
        n = len(key)
        if n > _lib.EW_MAX_INSTR:
            raise ValueError("elementwise program too long (%d instructions)" % n)
        prog = (_lib.EwInstr * n)()
        for i, (op, arg, imm) in enumerate(key):
            prog[i].op = op
            prog[i].arg = arg
            prog[i].imm = imm
        if len(_prog_cache) > 8192:
            _prog_cache.clear()
        _prog_cache[key] = prog
    return prog


def run_program(code, inputs, out_shape, n_out=1):
    """code: list of (op, arg, imm); inputs: list of DeviceArray; returns the output DeviceArray(s)."""
    outs = [DeviceArray.empty(out_shape) for _ in range(n_out)]
    run_program_into(code, inputs, outs)
    return outs[0] if n_out == 1 else outs


def run_program_into(code, inputs, outs):
    prog = _program(code)
    tin = (_lib.Tensor * max(len(inputs), 1))(*[a.tensor() for a in inputs])
    tout = (_lib.Tensor * len(outs))(*[o.tensor() for o in outs])
    _lib.check(_lib.lib().adb_ewise(prog, len(prog), len(inputs), tin, len(outs), tout, _STREAM()))


def _STREAM():
    return C.c_void_p(stream_ptr())
