"""ctypes binding of libbci.so (include/bci.h).  This is the only way Python reaches the
CUDA kernels; there is no CPU fallback behind it — if the library is missing or the GPU
cannot be initialised, callers get a `BciError` and must surface it."""
from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("BCI_LIB", PKG_DIR / "libbci.so"))

# ---- constants mirrored from include/bci.h ------------------------------------------
OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_OOM, ERR_NOT_INIT = 0, 1, 2, 3, 4, 5
F32, F64, I32, I64 = 0, 1, 2, 3
EW_ADD, EW_SUB, EW_MUL, EW_DIV, EW_MAXIMUM, EW_MINIMUM, EW_RSUB, EW_RDIV = range(8)
EW_SQUARE, EW_ABS, EW_NEG, EW_SQRT, EW_EXP, EW_LOG = 32, 33, 34, 35, 36, 37
EW_SIN, EW_COS, EW_TANH, EW_FLOOR, EW_CEIL, EW_LOG1P, EW_EXPM1 = 38, 39, 40, 41, 42, 43, 44
RED_SUM, RED_MEAN, RED_MAX, RED_MIN = 0, 1, 2, 3
PRE_NONE, PRE_SQUARE, PRE_ABS = 0, 1, 2
ARGMAX, ARGMIN = 0, 1
PCG_F64, PCG_F32, PCG_U32, PCG_BOUNDED32_I64, PCG_BOUNDED32_I32 = range(5)
FILL_CONST, FILL_ARANGE, FILL_LINSPACE = 0, 1, 2
ABI_VERSION = 1


class BciError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libbci error {code}: {message}")
        self.code = code


class Unsupported(BciError):
    """The call is valid numpy but outside what the kernels take: caller uses numpy."""


class Scalar(C.Union):
    _fields_ = [("f", C.c_double), ("i", C.c_int64)]


class DevInfo(C.Structure):
    _fields_ = [
        ("device", C.c_int),
        ("sm_count", C.c_int),
        ("cc_major", C.c_int),
        ("cc_minor", C.c_int),
        ("total_mem", C.c_size_t),
        ("name", C.c_char * 64),
    ]


class Pcg64Gen(C.Structure):
    _fields_ = [("state_lo", C.c_uint64), ("state_hi", C.c_uint64), ("inc_lo", C.c_uint64), ("inc_hi", C.c_uint64),
                ("low", C.c_int64), ("range", C.c_uint64)]


class Sha256State(C.Structure):
    _fields_ = [
        ("h", C.c_uint32 * 8),
        ("total_len", C.c_uint64),
        ("buf_len", C.c_uint32),
        ("buf", C.c_uint8 * 64),
    ]


_VPP = C.POINTER(C.c_void_p)
_SZP = C.POINTER(C.c_size_t)

# name -> (restype, argtypes); every symbol include/bci.h declares
SIGNATURES = {
    "bci_init": (C.c_int, [C.c_int]),
    "bci_shutdown": (None, []),
    "bci_abi_version": (C.c_int, []),
    "bci_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "bci_device_info": (C.c_int, [C.POINTER(DevInfo)]),
    "bci_mem_info": (C.c_int, [_SZP, _SZP]),
    "bci_last_error": (C.c_char_p, []),
    "bci_healthy": (C.c_int, []),
    "bci_alloc": (C.c_int, [C.c_size_t, _VPP]),
    "bci_free": (C.c_int, [C.c_void_p]),
    "bci_host_alloc": (C.c_int, [C.c_size_t, _VPP]),
    "bci_host_free": (C.c_int, [C.c_void_p]),
    "bci_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "bci_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "bci_d2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "bci_memset": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t]),
    "bci_h2d_fd": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_size_t]),
    "bci_d2h_fd": (C.c_int, [C.c_int, C.c_uint64, C.c_void_p, C.c_size_t]),
    "bci_sync": (C.c_int, []),
    "bci_timer_begin": (C.c_int, []),
    "bci_timer_end": (C.c_int, [C.POINTER(C.c_float)]),
    "bci_event_record": (C.c_int, [C.c_int]),
    "bci_event_elapsed": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "bci_launch_count": (C.c_uint64, []),
    "bci_flush_l2": (C.c_int, []),
    "bci_ewise": (C.c_int, [C.c_int, C.c_int, C.c_int, _VPP, _VPP, C.POINTER(Scalar), _VPP, _SZP]),
    "bci_reduce": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _VPP, _SZP, C.c_void_p]),
    "bci_reduce_fused": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _VPP, _VPP,
                                   C.POINTER(Scalar), C.POINTER(Scalar), _SZP, C.c_void_p]),
    "bci_argext": (C.c_int, [C.c_int, C.c_int, C.c_int, _VPP, _SZP, C.POINTER(C.c_int64)]),
    "bci_cumsum": (C.c_int, [C.c_int, C.c_int, _VPP, _VPP, _SZP]),
    "bci_sort": (C.c_int, [C.c_int, C.c_int, _VPP, _VPP, _SZP]),
    "bci_argsort": (C.c_int, [C.c_int, C.c_int, _VPP, _VPP, _SZP]),
    "bci_rand_f64": (C.c_int, [C.c_int, _VPP, _SZP, C.POINTER(C.c_uint64)]),
    "bci_rand_normal_f64": (C.c_int, [C.c_int, _VPP, _SZP, C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "bci_rand_mt19937_f64": (C.c_int, [C.c_int, _VPP, _SZP, C.POINTER(C.c_uint32)]),
    "bci_rand_pcg64": (C.c_int, [C.c_int, C.c_int, _VPP, _SZP, C.POINTER(Pcg64Gen), C.POINTER(C.c_uint32)]),
    "bci_reduce_rand_f64": (C.c_int, [C.c_int, C.c_int, C.c_int, _SZP, C.POINTER(C.c_uint64), C.c_void_p]),
    "bci_fill": (C.c_int, [C.c_int, C.c_int, C.c_int, _VPP, _SZP, C.POINTER(Scalar), C.POINTER(Scalar), C.POINTER(Scalar)]),
    "bci_cast": (C.c_int, [C.c_int, C.c_int, C.c_int, _VPP, _VPP, _SZP]),
    "bci_matmul_f32": (C.c_int, [C.c_int, _VPP, _VPP, _VPP, C.c_int, C.c_int, C.c_int]),
    "bci_sha256": (C.c_int, [C.c_int, _VPP, _SZP, C.c_int, C.c_void_p]),
    "bci_sha256_init": (C.c_int, [C.POINTER(Sha256State)]),
    "bci_sha256_update_batch": (C.c_int, [C.c_int, C.POINTER(C.POINTER(Sha256State)), _VPP, _SZP]),
    "bci_sha256_final_batch": (C.c_int, [C.c_int, C.POINTER(C.POINTER(Sha256State)), C.c_void_p]),
    # include/bci_broker.h
    "bci_broker_start": (C.c_int, [C.c_char_p, C.c_int]),
    "bci_broker_stop": (None, []),
    "bci_broker_stats": (C.c_int, [C.POINTER(C.c_uint64)]),
}

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Loads libbci.so (once) and binds every declared symbol.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise BciError(
                ERR_NOT_INIT,
                f"{LIB_PATH} is missing: build it with `python -m bee_code_interpreter_b200.build` "
                "(there is no CPU fallback for the GPU executor backend)",
            )
        lib = C.CDLL(str(LIB_PATH))
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export it
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.bci_abi_version() != ABI_VERSION:
            raise BciError(ERR_INVALID, f"libbci ABI {lib.bci_abi_version()} != binding {ABI_VERSION}")
        _lib = lib
        return lib


def last_error() -> str:
    msg = load().bci_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int) -> None:
    if rc == OK:
        return
    msg = last_error()
    if rc == ERR_UNSUPPORTED:
        raise Unsupported(rc, msg)
    raise BciError(rc, msg)


def ptr_array(ptrs) -> C.Array:
    arr = (C.c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


def size_array(ns) -> C.Array:
    arr = (C.c_size_t * len(ns))()
    for i, n in enumerate(ns):
        arr[i] = int(n)
    return arr
