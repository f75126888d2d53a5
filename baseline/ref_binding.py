"""INTEGRATION.md section B made executable: the ctypes stub a maintainer of bgavran/autodiff would add next to
autodiff/core/ops.py, binding libadb200.so INSIDE the reference's own classes.

`bind(ref)` replaces the numpy line of `_eval` in the reference's `Einsum`, `Add`, `Mul`, `Negate`, `Recipr`, `ReLU`, `Exp`,
`Log`, `Sigmoid` and `ReduceSumKeepDims` (ops.py:53,75,96,116,180,228,309,339,354; reshape.py:14) by one call through the C
ABI of include/adb200.h; everything else of the reference - graph construction, `grad`, memoisation, shapes, error
messages - runs unmodified.  Operands travel host -> device -> host around every node (this is the minimal binding, not
the fast one: the package `autodiff_b200` keeps values resident); `unbind(ref)` restores the stock bodies.

This file uses NOTHING of the autodiff_b200 Python package: only ctypes + numpy + the shared library, exactly what the
reference side would have.  tests/test_gpu_ref_binding.py runs the reference's graphs both ways and compares."""
import ctypes as C
import numbers
import os

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "autodiff_b200", "libadb200.so")
_lib = None


class Tensor(C.Structure):                       # include/adb200.h: adb_tensor
    _fields_ = [("ptr", C.c_void_p), ("rank", C.c_int32), ("_pad", C.c_int32),
                ("shape", C.c_int64 * 8), ("stride", C.c_int64 * 8)]


class EwInstr(C.Structure):                      # include/adb200.h: adb_ew_instr
    _fields_ = [("op", C.c_int32), ("arg", C.c_int32), ("imm", C.c_double)]


IN, IMM = 1, 2
LOAD, ADD, MUL, STORE_OUT, NEG, RECIP_EPS, LOG_EPS, EXP, RELU, SIGMOID = (k << 2 for k in (0, 1, 2, 8, 9, 10, 11, 12, 13, 14))


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(LIB_PATH)
        l.adb_last_error.restype = C.c_char_p
        for name in ("adb_malloc", "adb_free", "adb_memcpy_h2d", "adb_memcpy_d2h", "adb_sync", "adb_einsum", "adb_ewise", "adb_reduce_sum", "adb_init"):
            getattr(l, name).restype = C.c_int
        l.adb_malloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_void_p]
        l.adb_free.argtypes = [C.c_void_p, C.c_void_p]
        l.adb_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        l.adb_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        l.adb_sync.argtypes = [C.c_void_p]
        l.adb_einsum.argtypes = [C.c_char_p, C.c_int, C.POINTER(Tensor), C.POINTER(Tensor), C.c_void_p]
        l.adb_ewise.argtypes = [C.POINTER(EwInstr), C.c_int, C.c_int, C.POINTER(Tensor), C.c_int, C.POINTER(Tensor), C.c_void_p]
        l.adb_reduce_sum.argtypes = [C.POINTER(Tensor), C.POINTER(Tensor), C.c_void_p]
        check(l.adb_init(0), l)
        _lib = l
    return _lib


def check(rc, l=None):
    if rc:
        raise ValueError((l or _lib).adb_last_error().decode())


class Dev:
    """A dense device-resident fp64 array owned through adb_malloc / adb_free."""

    def __init__(self, shape):
        self.shape = tuple(int(s) for s in shape)
        self.ptr = C.c_void_p()
        check(lib().adb_malloc(C.byref(self.ptr), 8 * max(int(np.prod(self.shape, dtype=np.int64)), 1), None))

    def __del__(self):
        if _lib is not None and self.ptr:
            _lib.adb_free(self.ptr, None)

    def tensor(self, as_shape=None):
        """adb_tensor of this array; `as_shape` broadcasts it numpy-style (stride 0 on the added / stretched axes)."""
        shape = self.shape if as_shape is None else tuple(as_shape)
        t = Tensor()
        t.ptr, t.rank = self.ptr, len(shape)
        own = (1,) * (len(shape) - len(self.shape)) + self.shape
        acc = 1
        for i in range(len(shape) - 1, -1, -1):
            t.shape[i] = shape[i]
            t.stride[i] = 0 if (own[i] == 1 and shape[i] != 1) else acc
            acc *= own[i]
        return t

    @staticmethod
    def upload(a):
        a = np.asarray(a, dtype=np.float64, order="C")           # (np.ascontiguousarray would turn a 0-d value into shape (1,))
        d = Dev(a.shape)
        if a.size:
            check(lib().adb_memcpy_h2d(d.ptr, a.ctypes.data_as(C.c_void_p), a.nbytes, None))
            check(lib().adb_sync(None))
        return d

    def download(self):
        out = np.empty(self.shape)
        if out.size:
            check(lib().adb_memcpy_d2h(out.ctypes.data_as(C.c_void_p), self.ptr, out.nbytes, None))
        check(lib().adb_sync(None))
        return out


def _ewise(code, inputs, out_shape):
    prog = (EwInstr * len(code))(*[EwInstr(op, arg, imm) for op, arg, imm in code])
    devs = [Dev.upload(a) for a in inputs]
    out = Dev(out_shape)
    tin = (Tensor * len(devs))(*[d.tensor(out_shape) for d in devs])
    tout = (Tensor * 1)(out.tensor())
    check(lib().adb_ewise(prog, len(code), len(devs), tin, 1, tout, None))
    return out.download()


def _nary(kind, identity):
    def _eval(self):
        vals = [np.asarray(c(), dtype=np.float64) for c in self.children]
        if not vals:
            return np.array(identity)
        shape = np.broadcast_shapes(*[v.shape for v in vals])
        code = [(LOAD | IN, 0, 0.0)] + [(kind | IN, i, 0.0) for i in range(1, len(vals))] + [(STORE_OUT, 0, 0.0)]
        return _ewise(code, vals, shape)
    return _eval


def _unary(*instrs):
    def _eval(self):
        x = np.asarray(self.node(), dtype=np.float64)
        return _ewise([(LOAD | IN, 0, 0.0)] + list(instrs) + [(STORE_OUT, 0, 0.0)], [x], x.shape)
    return _eval


def _expand_ellipsis(op_str, ranks):
    """'...i,...ij->...ij' with explicit letters (adb_einsum takes letters only; the planner of the package does the same)."""
    if "..." not in op_str:
        return op_str
    lhs, rhs = op_str.split("->")
    subs = lhs.split(",")
    used = set(op_str) - set(".,->")
    pool = [c for c in "ABCDEFGHIJKLMNOPQRSTUVWXYZ" if c not in used]
    n_ell = max(r - (len(s) - 3) for s, r in zip(subs, ranks) if "..." in s)
    ell = "".join(pool[:n_ell])
    out = []
    for s, r in zip(subs, ranks):
        if "..." in s:
            k = r - (len(s) - 3)
            s = s.replace("...", ell[n_ell - k:])
        out.append(s)
    return ",".join(out) + "->" + rhs.replace("...", ell)


def _einsum_eval(ref):
    Einsum = ref.Einsum

    def _eval(self):
        arr = [op() for op in self.operands]                               # reference ops.py:173
        for i, val in enumerate(arr):                                      # reference ops.py:175-178
            if isinstance(val, numbers.Number):
                shp = [l for let in Einsum.split_dots(self.opnames[i]) for l in self.letter_to_dim.get(let, [1])]
                arr[i] = np.broadcast_to(val, shp)
        arr = [np.asarray(a, dtype=np.float64) for a in arr]
        op_str = _expand_ellipsis(self.op_str, [a.ndim for a in arr])
        lhs, rhs = op_str.split("->")
        ext = {}
        for sub, a in zip(lhs.split(","), arr):
            for letter, d in zip(sub, a.shape):
                ext[letter] = max(ext.get(letter, 1), d)
        out = Dev([ext[l] for l in rhs])
        devs = [Dev.upload(a) for a in arr]
        tin = (Tensor * len(devs))(*[d.tensor() for d in devs])
        to = out.tensor()
        check(lib().adb_einsum(op_str.encode(), len(devs), tin, C.byref(to), None))       # replaces np.einsum, ops.py:180
        return out.download()
    return _eval


def _reduce_eval(self):
    x = np.asarray(self.node(), dtype=np.float64)
    out = Dev([1 if i in self.axes else s for i, s in enumerate(x.shape)])
    d = Dev.upload(x)
    ti, to = d.tensor(), out.tensor()
    check(lib().adb_reduce_sum(C.byref(ti), C.byref(to), None))                            # replaces np.sum(keepdims), reshape.py:14
    return out.download()


_saved = {}


def bind(ref):
    """Patches the reference module `ref` (import autodiff as ref) in place."""
    eps = ref.Node.epsilon
    patches = {
        ref.Einsum: _einsum_eval(ref),
        ref.Add: _nary(ADD, 0),
        ref.Mul: _nary(MUL, 1),
        ref.Negate: _unary((NEG, 0, 0.0)),
        ref.Recipr: _unary((RECIP_EPS, 0, eps)),
        ref.ReLU: _unary((RELU, 0, 0.0)),
        ref.Exp: _unary((EXP, 0, 0.0)),
        ref.Log: _unary((LOG_EPS, 0, eps)),
        ref.Sigmoid: _unary((SIGMOID, 0, 0.0)),
        ref.ReduceSumKeepDims: _reduce_eval,
    }
    lib()
    for cls, fn in patches.items():
        if cls not in _saved:
            _saved[cls] = cls._eval
        cls._eval = fn
    return sorted(c.__name__ for c in patches)


def unbind(ref):
    for cls, fn in _saved.items():
        cls._eval = fn
    _saved.clear()
