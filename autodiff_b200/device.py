"""Device-resident arrays and the host<->device boundary.

`DeviceArray` replaces the numpy ndarray that the reference keeps in `Node.cached` / `Variable._value`
(reference autodiff/core/node.py:42-46,111-129).  Memory, streams and pinned host buffers come from PyTorch
(plumbing only); every byte of arithmetic and every strided copy goes through libadb200.so.
"""
import ctypes as C
import threading

import numpy as np

from . import _lib

_state = threading.local()
_torch = None
_initialised = False
_const_cache = {}


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def init(device=None):
    """Binds this process to one GPU.  Raises if there is no B200 — there is no CPU path."""
    global _initialised
    t = torch()
    if not t.cuda.is_available():
        raise _lib.BackendError("autodiff_b200 needs a CUDA device (B200, sm_100a); none is visible and there is no CPU fallback")
    if device is None:
        import os
        device = int(os.environ.get("LOCAL_RANK", "0"))
    t.cuda.set_device(device)
    _lib.check(_lib.lib().adb_init(int(device)))
    _initialised = True
    return device


def _ensure_init():
    if not _initialised:
        init()


_stream = None
_alloc_hook = None     # set by core/tape.py while a recording evaluation runs: hook(base tensor) for every DeviceArray.empty


def stream_ptr():
    """cudaStream_t every kernel of this process is queued on (torch's current stream at init; see use_stream)."""
    global _stream
    if _stream is None:
        _stream = torch().cuda.current_stream().cuda_stream
    return _stream


def use_stream(stream=None):
    """Re-reads (or sets) the stream after the caller switched torch's current stream."""
    global _stream
    _stream = (stream if stream is not None else torch().cuda.current_stream()).cuda_stream


def synchronize():
    _lib.check(_lib.lib().adb_sync(C.c_void_p(stream_ptr())))


def launch_count():
    return int(_lib.lib().adb_launch_count())


def _dense_strides(shape, pitch=None):
    strides = [0] * len(shape)
    acc = 1
    for i in range(len(shape) - 1, -1, -1):
        strides[i] = acc
        if i == len(shape) - 1 and pitch is not None:
            acc = pitch
        else:
            acc *= max(int(shape[i]), 1)
    return tuple(strides)


class DeviceArray:
    """fp64 strided view of device memory.  strides are in ELEMENTS and may be 0 (broadcast) or negative."""
    __slots__ = ("base", "offset", "shape", "strides", "_tensor")

    def __init__(self, base, offset, shape, strides):
        self.base = base            # torch.float64 cuda tensor that owns the storage
        self.offset = int(offset)   # elements from base.data_ptr()
        self.shape = tuple(int(s) for s in shape)
        self.strides = tuple(int(s) for s in strides)
        self._tensor = None         # cached adb_tensor (views are immutable)

    @staticmethod
    def _make(base, offset, shape, strides):
        """Constructor for callers that already hold int tuples (tape replay, internal views): skips the conversions."""
        d = object.__new__(DeviceArray)
        d.base, d.offset, d.shape, d.strides, d._tensor = base, offset, shape, strides, None
        return d

    # ---- construction
    @staticmethod
    def empty(shape):
        """Row pitch of >=2-D arrays is padded to a multiple of 4 elements (16 for rows of >= 256) so every row is 32-byte aligned
        (TMA global strides must be multiples of 16 B, the elementwise kernels use 256-bit accesses; the NN bias
        trick makes odd widths like 4097 common, reference high_level_ops.py:32-35)."""
        _ensure_init()
        shape = tuple(int(s) for s in shape)
        t = torch()
        if len(shape) >= 2:
            # short innermost extents (channels of an NHWC image) are never TMA rows: keep them dense.  Wide rows are padded to
            # whole 128-byte lines (16 elements): the blocked tensor map of an MN-contiguous GEMM operand (gemm_f64.cu:
            # make_tmap_blocked) reads 16-column blocks, and the last block of a 4097-wide row must stay inside the row
            w = shape[-1]
            pitch = w + (-w % 16) if w >= 256 else (w + (-w % 4) if w >= 16 else w)
            n = pitch
            for s in shape[:-1]:
                n *= s
            strides = _dense_strides(shape, pitch)
        else:
            n = shape[0] if shape else 1
            strides = _dense_strides(shape)
        base = t.empty(max(int(n), 2), dtype=t.float64, device="cuda")
        if _alloc_hook is not None:
            _alloc_hook(base)
        return DeviceArray(base, 0, shape, strides)

    @staticmethod
    def full(shape, value):
        """A broadcast constant: one element, all strides 0 (never materialised); buffers cached per value."""
        _ensure_init()
        value = float(value)
        base = _const_cache.get(value)
        if base is None:
            t = torch()
            base = t.empty(2, dtype=t.float64, device="cuda")
            host = np.array([value, value], dtype=np.float64)
            _lib.check(_lib.lib().adb_memcpy_h2d(C.c_void_p(base.data_ptr()), host.ctypes.data_as(C.c_void_p), 16, C.c_void_p(stream_ptr())))
            synchronize()   # host staging array dies at return
            if len(_const_cache) < 4096:
                _const_cache[value] = base
        return DeviceArray(base, 0, tuple(shape), (0,) * len(shape))

    # ---- properties
    @property
    def ptr(self):
        return self.base.data_ptr() + 8 * self.offset

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        n = 1
        for s in self.shape:
            n *= s
        return n

    def tensor(self):
        t = self._tensor
        if t is None:
            t = self._tensor = _lib.make_tensor(self.ptr, self.shape, self.strides)
        return t

    def is_dense(self):
        """C-contiguous with no padding."""
        acc = 1
        for d, s in zip(reversed(self.shape), reversed(self.strides)):
            if d != 1 and s != acc:
                return False
            acc *= d
        return True

    # ---- views (metadata only; mirrors numpy basic indexing, reference reshape.py:92-94)
    def view(self, shape, strides, offset=0):
        return DeviceArray(self.base, self.offset + offset, shape, strides)

    def permute(self, perm):
        return self.view([self.shape[p] for p in perm], [self.strides[p] for p in perm])

    def broadcast_to(self, shape):
        shape = tuple(shape)
        pad = len(shape) - self.ndim
        if pad < 0:
            raise ValueError("cannot broadcast %s to %s" % (self.shape, shape))
        st = []
        for i, d in enumerate(shape):
            j = i - pad
            if j < 0:
                st.append(0)
            elif self.shape[j] == d:
                st.append(self.strides[j] if d != 1 else 0)
            elif self.shape[j] == 1:
                st.append(0)
            else:
                raise ValueError("cannot broadcast %s to %s" % (self.shape, shape))
        return self.view(shape, st)

    def getitem(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        if any(i is Ellipsis for i in idx):
            n_real = sum(1 for i in idx if i is not Ellipsis and i is not None)
            k = idx.index(Ellipsis)
            idx = idx[:k] + (slice(None),) * (self.ndim - n_real) + idx[k + 1:]
        shape, strides, off, dim = [], [], 0, 0
        for i in idx:
            if i is None:
                shape.append(1); strides.append(0)
                continue
            if dim >= self.ndim:
                raise IndexError("too many indices for array")
            d, s = self.shape[dim], self.strides[dim]
            if isinstance(i, slice):
                start, stop, step = i.indices(d)
                n = len(range(start, stop, step))
                off += start * s if n > 0 else 0
                shape.append(n); strides.append(s * step)
            else:
                try:
                    ii = int(i)
                except TypeError:
                    raise IndexError("only integers, slices, None and Ellipsis are supported on device (got %r)" % (i,))
                if ii < 0:
                    ii += d
                if not 0 <= ii < d:
                    raise IndexError("index %d is out of bounds for axis %d with size %d" % (i, dim, d))
                off += ii * s
            dim += 1
        shape += list(self.shape[dim:]); strides += list(self.strides[dim:])
        return self.view(shape, strides, off)

    def reshape_view(self, shape):
        """Returns a view with the new shape if one exists without copying, else None (numpy rules, simplified:
        the array must be dense, or a broadcast scalar)."""
        shape = tuple(int(s) for s in shape)
        if all(s == 0 for s in self.strides) or self.size <= 1:
            if int(np.prod(shape, dtype=np.int64)) != self.size and self.size > 1:
                return None
            if self.size <= 1:
                return self.view(shape, (0,) * len(shape))
        if self.is_dense():
            return self.view(shape, _dense_strides(shape))
        return None


# --------------------------------------------------------------------------------------- host <-> device
class _PinnedBlock:
    """One cudaMallocHost allocation; goes back to the pool when the last numpy view of it dies."""
    __slots__ = ("ptr", "cap", "__weakref__")

    def __init__(self, ptr, cap):
        self.ptr, self.cap = ptr, cap

    def __del__(self):
        global _pinned_pooled
        try:
            if _pinned_pooled + self.cap <= _PINNED_POOL_MAX:
                _pinned_free.setdefault(self.cap, []).append(self.ptr)
                _pinned_pooled += self.cap
            else:
                _lib.lib().adb_host_free(C.c_void_p(self.ptr))
        except Exception:      # interpreter shutdown
            pass


_pinned_pooled = 0                 # bytes currently parked in the pool
_PINNED_POOL_MAX = 48 << 30        # beyond this, blocks go back to the driver instead of the pool
_pinned_free = {}      # capacity -> [host pointers]; page-locked memory is expensive to create (~0.3 s/GiB), so it is recycled
_PINNED_GRAIN = 1 << 21


def pinned_empty(shape, dtype=np.float64):
    """numpy array backed by page-locked memory from a size-class pool (cudaMallocHost through the C ABI); the block
    returns to the pool when the array and all its views are garbage collected."""
    n = int(np.prod(shape, dtype=np.int64))
    nbytes = max(n, 1) * np.dtype(dtype).itemsize
    if not torch().cuda.is_available():
        return np.empty(shape, dtype=dtype)
    cap = (nbytes + _PINNED_GRAIN - 1) // _PINNED_GRAIN * _PINNED_GRAIN
    global _pinned_pooled
    free = _pinned_free.get(cap)
    if free:
        ptr = free.pop()
        _pinned_pooled -= cap
    else:
        out = C.c_void_p()
        _lib.check(_lib.lib().adb_host_alloc(C.byref(out), cap))
        ptr = out.value
    block = _PinnedBlock(ptr, cap)
    raw = (C.c_char * nbytes).from_address(ptr)
    raw._adb_block = block                      # numpy keeps `raw` alive as the array's base
    return np.frombuffer(raw, dtype=dtype, count=n).reshape(shape)


def _enqueue_upload(a, out, st):
    """Queues the host->device copy of fp64 ndarray `a` into DeviceArray `out` on stream `st` (a c_void_p)."""
    L = _lib.lib()
    if a.ndim >= 2 and a.strides[-1] == 8 and _rows_regular(a):
        rows = a.size // a.shape[-1]
        src_pitch = a.strides[-2] if a.shape[-2] > 1 else a.shape[-1] * 8
        if all(d == 1 for d in a.shape[:-1]):
            src_pitch = a.shape[-1] * 8
        _lib.check(L.adb_memcpy2d_h2d(C.c_void_p(out.ptr), out.strides[-2] * 8 if out.ndim >= 2 else a.shape[-1] * 8,
                                      C.c_void_p(a.ctypes.data), src_pitch, a.shape[-1] * 8, rows, st))
        return a
    if not a.flags.c_contiguous:
        a = np.ascontiguousarray(a)
    if a.ndim >= 2:
        rows = a.size // a.shape[-1]
        _lib.check(L.adb_memcpy2d_h2d(C.c_void_p(out.ptr), out.strides[-2] * 8, C.c_void_p(a.ctypes.data), a.shape[-1] * 8,
                                      a.shape[-1] * 8, rows, st))
    else:
        _lib.check(L.adb_memcpy_h2d(C.c_void_p(out.ptr), C.c_void_p(a.ctypes.data), a.size * 8, st))
    return a


def to_device(array):
    """numpy (any dtype / strides) -> DeviceArray (fp64).  Mixed dtypes are promoted to float64 on upload
    (the reference computes in numpy's promoted dtype, float64 for every BASELINE config)."""
    _ensure_init()
    if isinstance(array, DeviceArray):
        return array
    a = np.asarray(array)
    if a.dtype != np.float64:
        a = a.astype(np.float64)
    if 0 < a.size <= _SMALL_CONST_ELEMS:
        # tiny host constants (the `np.array([1])` operand of every Softmax, reference ops.py:428-432; python numbers used as
        # einsum operands) are uploaded once per distinct content: a synchronous upload each time would drain the
        # stream hundreds of times per evaluation of an unrolled RNN.  Device arrays are never written in place.
        key = (a.shape, a.tobytes())
        hit = _small_const_cache.get(key)
        if hit is not None:
            return hit
    out = DeviceArray.empty(a.shape)
    if a.size == 0:
        return out
    _enqueue_upload(a, out, C.c_void_p(stream_ptr()))
    synchronize()   # the source may be pageable and may be mutated/freed by the caller right after
    if a.size <= _SMALL_CONST_ELEMS:
        if len(_small_const_cache) >= 4096:
            _small_const_cache.clear()
        _small_const_cache[(a.shape, a.tobytes())] = out
    return out


_SMALL_CONST_ELEMS = 16
_small_const_cache = {}


def to_device_async(array, stream):
    """Like to_device, but the copy is only QUEUED on `stream` (a torch.cuda.Stream).  Returns (DeviceArray, keepalive):
    the caller orders consumers after the copy with an event and keeps `keepalive` (the staged source) alive until then."""
    _ensure_init()
    a = np.asarray(array)
    if a.dtype != np.float64:
        a = a.astype(np.float64)
    out = DeviceArray.empty(a.shape)
    if a.size == 0:
        return out, a
    keep = _enqueue_upload(a, out, C.c_void_p(stream.cuda_stream))
    return out, keep


def enqueue_download(d, stream):
    """Queues the device->host copy of DeviceArray `d` on `stream` into a fresh pinned buffer; returns the ndarray
    (valid once the stream has been synchronised), or None when `d` needs a gather first (the caller falls back)."""
    L = _lib.lib()
    st = C.c_void_p(stream.cuda_stream)
    if d.size == 0:
        return np.empty(d.shape, dtype=np.float64)
    out = pinned_empty(d.shape) if d.size >= 4096 else np.empty(d.shape, dtype=np.float64)
    if d.is_dense():
        _lib.check(L.adb_memcpy_d2h(C.c_void_p(out.ctypes.data), C.c_void_p(d.ptr), d.size * 8, st))
        return out
    if d.ndim >= 2 and d.strides[-1] == 1 and _dev_rows_regular(d):
        rows = d.size // d.shape[-1]
        _lib.check(L.adb_memcpy2d_d2h(C.c_void_p(out.ctypes.data), d.shape[-1] * 8, C.c_void_p(d.ptr), d.strides[-2] * 8,
                                      d.shape[-1] * 8, rows, st))
        return out
    return None


def _rows_regular(a):
    """True when all leading dims of `a` collapse onto one constant row pitch (so a 2-D pitched copy applies)."""
    if a.ndim < 2:
        return False
    pitch = a.strides[-2]
    if pitch < a.shape[-1] * 8:
        return False
    acc = pitch * a.shape[-2]
    for d, s in zip(reversed(a.shape[:-2]), reversed(a.strides[:-2])):
        if d != 1 and s != acc:
            return False
        acc *= d
    return True


def copy_device(src, dst):
    """Strided device copy dst[...] = src (broadcasting src), via the ewise kernel's [LOAD; STORE] program."""
    code = (_lib.EwInstr * 2)()
    code[0].op = _lib.EW_LOAD | _lib.SRC_IN; code[0].arg = 0
    code[1].op = _lib.EW_STORE_OUT; code[1].arg = 0
    tin = (_lib.Tensor * 1)(src.tensor())
    tout = (_lib.Tensor * 1)(dst.tensor())
    _lib.call("adb_ewise", code, 2, 1, tin, 1, tout, C.c_void_p(stream_ptr()))


def fill_device(dst, value):
    code = (_lib.EwInstr * 2)()
    code[0].op = _lib.EW_LOAD | _lib.SRC_IMM; code[0].imm = float(value)
    code[1].op = _lib.EW_STORE_OUT; code[1].arg = 0
    tout = (_lib.Tensor * 1)(dst.tensor())
    _lib.call("adb_ewise", code, 2, 0, None, 1, tout, C.c_void_p(stream_ptr()))


def to_host(d, pinned=True):
    """DeviceArray -> C-contiguous numpy float64 array (synchronises the stream)."""
    L = _lib.lib()
    st = C.c_void_p(stream_ptr())
    out = pinned_empty(d.shape) if pinned and d.size >= 4096 else np.empty(d.shape, dtype=np.float64)
    if d.size == 0:
        return out
    if not d.is_dense():
        rows_ok = d.ndim >= 2 and d.strides[-1] == 1 and _dev_rows_regular(d)
        if rows_ok:
            rows = d.size // d.shape[-1]
            _lib.check(L.adb_memcpy2d_d2h(C.c_void_p(out.ctypes.data), d.shape[-1] * 8, C.c_void_p(d.ptr), d.strides[-2] * 8,
                                          d.shape[-1] * 8, rows, st))
            synchronize()
            return out
        dense = DeviceArray(torch().empty(max(d.size, 2), dtype=torch().float64, device="cuda"), 0, d.shape, _dense_strides(d.shape))
        copy_device(d, dense)
        d = dense
    _lib.check(L.adb_memcpy_d2h(C.c_void_p(out.ctypes.data), C.c_void_p(d.ptr), d.size * 8, st))
    synchronize()
    return out


def _dev_rows_regular(d):
    pitch = d.strides[-2]
    if pitch < d.shape[-1]:
        return False
    acc = pitch * d.shape[-2]
    for dim, s in zip(reversed(d.shape[:-2]), reversed(d.strides[:-2])):
        if dim != 1 and s != acc:
            return False
        acc *= dim
    return True


# --------------------------------------------------------------------------------------- precision mode
class precision:
    """`with ad.precision("tf32x3"): ...` — GEMM-shaped einsums run as 3xTF32 on the tcgen05 tensor cores (fp32
    accumulation in TMEM, ~1e-7 relative error, parity bar 1e-4) instead of FP64 DMMA.  Default is "fp64"."""
    MODES = {"fp64": 0, "tf32x3": 1}

    def __init__(self, mode):
        if mode not in self.MODES:
            raise ValueError("precision must be one of %s" % sorted(self.MODES))
        self.mode = self.MODES[mode]
        self.prev = None

    def __enter__(self):
        self.prev = _lib.lib().adb_get_precision()
        _set_precision_code(self.mode)
        return self

    def __exit__(self, *a):
        _set_precision_code(self.prev)


# Global kernel-selection state of the library, mirrored here so that launch tapes can key on it (a CUDA graph captured
# under one mode must never be replayed under another): [precision code, elementwise interpreter forced].
kernel_mode = [0, 0]


def _set_precision_code(code):
    _lib.check(_lib.lib().adb_set_precision(code))
    kernel_mode[0] = int(code)


def set_precision(mode):
    _set_precision_code(precision.MODES[mode])


# --------------------------------------------------------------------------------------- diagnostics
def ewise_stats():
    """One line per distinct elementwise program run so far: "<calls> <quads> <static|interp> op:arg ..."
    (input of tools/gen_ewise_catalog.py --merge)."""
    L = _lib.lib()
    n = L.adb_ewise_stats(None, 0)
    buf = C.create_string_buffer(n + 16)
    L.adb_ewise_stats(buf, n + 16)
    return buf.value.decode()


def force_interpreter(on=True):
    """Runs every elementwise program on the bytecode interpreter (A/B against the catalogued kernels)."""
    _lib.check(_lib.lib().adb_ewise_force_interpreter(1 if on else 0))
    kernel_mode[1] = 1 if on else 0
