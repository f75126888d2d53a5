"""ctypes binding of libadb200.so (include/adb200.h).  There is NO fallback: if the shared library is missing
or no B200 is visible, every compute entry point raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libadb200.so")

MAX_RANK = 8
MAX_OPERANDS = 8
EW_MAX_IN = 8
EW_MAX_OUT = 4
EW_MAX_INSTR = 64
EW_TEMPS = 8


class Tensor(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("rank", C.c_int32), ("_pad", C.c_int32),
                ("shape", C.c_int64 * MAX_RANK), ("stride", C.c_int64 * MAX_RANK)]


class GemmDesc(C.Structure):
    _fields_ = [("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64), ("batch", C.c_int64),
                ("A", C.c_void_p), ("a_sm", C.c_int64), ("a_sk", C.c_int64), ("a_sb", C.c_int64),
                ("B", C.c_void_p), ("b_sk", C.c_int64), ("b_sn", C.c_int64), ("b_sb", C.c_int64),
                ("C", C.c_void_p), ("c_sm", C.c_int64), ("c_sn", C.c_int64), ("c_sb", C.c_int64)]


class EwInstr(C.Structure):
    _fields_ = [("op", C.c_int32), ("arg", C.c_int32), ("imm", C.c_double)]


# opcode kinds (adb200.h); op = KIND | SRC
SRC_NONE, SRC_IN, SRC_IMM, SRC_TMP = 0, 1, 2, 3
(EW_LOAD, EW_ADD, EW_MUL, EW_DIV, EW_POW, EW_RPOW, EW_MAX, EW_STORE_T, EW_STORE_OUT, EW_NEG, EW_RECIP_EPS,
 EW_LOG_EPS, EW_EXP, EW_RELU, EW_SIGMOID, EW_SQRT, EW_ABS, EW_SQUARE, EW_LOG, EW_RECIP) = [k << 2 for k in range(20)]

# every symbol include/adb200.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "adb_init": (C.c_int, [C.c_int]),
    "adb_last_error": (C.c_char_p, []),
    "adb_version": (C.c_char_p, []),
    "adb_launch_count": (C.c_int64, []),
    "adb_sync": (C.c_int, [C.c_void_p]),
    "adb_malloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t, C.c_void_p]),
    "adb_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "adb_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "adb_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "adb_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "adb_host_free": (C.c_int, [C.c_void_p]),
    "adb_memcpy2d_h2d": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p]),
    "adb_memcpy2d_d2h": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p]),
    "adb_gemm_f64": (C.c_int, [C.POINTER(GemmDesc), C.c_void_p]),
    "adb_gemm_f64_path": (C.c_int, [C.POINTER(GemmDesc), C.c_void_p, C.c_int]),
    "adb_gemm_sk_plan": (C.c_int, [C.c_longlong] * 4 + [C.c_int, C.c_int, C.c_longlong, C.c_int, C.POINTER(C.c_longlong)]),
    "adb_gemm_tf32x3": (C.c_int, [C.POINTER(GemmDesc), C.c_void_p]),
    "adb_set_precision": (C.c_int, [C.c_int]),
    "adb_get_precision": (C.c_int, []),
    "adb_ewise": (C.c_int, [C.POINTER(EwInstr), C.c_int, C.c_int, C.POINTER(Tensor), C.c_int, C.POINTER(Tensor), C.c_void_p]),
    "adb_ewise_stats": (C.c_int, [C.c_char_p, C.c_size_t]),
    "adb_ewise_force_interpreter": (C.c_int, [C.c_int]),
    "adb_ewise_set_piece_limit": (C.c_int, [C.c_longlong]),
    "adb_einsum": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(Tensor), C.POINTER(Tensor), C.c_void_p]),
    "adb_einsum_describe": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(Tensor), C.POINTER(Tensor), C.c_char_p, C.c_size_t]),
    "adb_reduce_sum": (C.c_int, [C.POINTER(Tensor), C.POINTER(Tensor), C.c_void_p]),
    "adb_sum_squares": (C.c_int, [C.POINTER(Tensor), C.c_void_p, C.c_void_p]),
    "adb_norm2": (C.c_int, [C.POINTER(Tensor), C.c_void_p, C.c_void_p]),
    "adb_softmax_ce": (C.c_int, [C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(Tensor), C.c_void_p, C.c_void_p]),
    "adb_im2col": (C.c_int, [C.POINTER(Tensor), C.c_int, C.c_int, C.POINTER(Tensor), C.c_void_p]),
    "adb_col2im": (C.c_int, [C.POINTER(Tensor), C.c_int, C.c_int, C.POINTER(Tensor), C.c_void_p]),
    "adb_conv_gemm": (C.c_int, [C.c_int, C.POINTER(Tensor), C.c_int, C.c_int, C.POINTER(Tensor), C.POINTER(Tensor), C.c_void_p]),
    "adb_conv_gemm_ok": (C.c_int, [C.c_int, C.POINTER(Tensor), C.c_int, C.c_int, C.POINTER(Tensor), C.POINTER(Tensor)]),
    "adb_graph_begin": (C.c_int, [C.c_void_p]),
    "adb_graph_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "adb_graph_launch": (C.c_int, [C.c_void_p, C.c_void_p]),
    "adb_graph_destroy": (C.c_int, [C.c_void_p]),
    "adb_memset_async": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]),
    "adb_memcpy_d2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
}

_lib = None


class BackendError(RuntimeError):
    pass


def lib():
    """Loads libadb200.so (building nothing, falling back to nothing)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BackendError("libadb200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`; "
                               "autodiff_b200 has no CPU fallback." % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


_call_hook = None     # set by core/tape.py while a recording evaluation runs: hook(name, fn, args)


def call(name, *args):
    """One C-ABI compute call (everything that launches kernels goes through here so that tapes can log it)."""
    fn = getattr(lib(), name)
    rc = fn(*args)
    if rc != 0:
        check(rc)
    if _call_hook is not None:
        _call_hook(name, fn, args)


def check(rc):
    if rc != 0:
        msg = lib().adb_last_error().decode("utf-8", "replace")
        if msg.startswith("einsum:") or msg.startswith("Number of operands") or msg.startswith("ewise:") or msg.startswith("softmax_ce:") or msg.startswith("im2col:") or msg.startswith("col2im:"):
            raise ValueError(msg)
        raise BackendError(msg)


def make_tensor(ptr, shape, strides):
    t = Tensor()
    t.ptr = ptr
    r = len(shape)
    if r > MAX_RANK:
        raise ValueError("tensor rank %d exceeds backend limit %d" % (r, MAX_RANK))
    t.rank = r
    for i in range(r):
        t.shape[i] = shape[i]
        t.stride[i] = strides[i]
    return t
