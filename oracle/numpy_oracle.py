"""oracle/numpy_oracle.py — TEST INFRASTRUCTURE ONLY.

The reference's arithmetic for this path is stock numpy called by agent code
(examples/benchmark-numpy.py:18-22, examples/using_imports.py:20-22); this module states,
op class by op class, which numpy call the device result is compared with and under which
bar."""
from __future__ import annotations

import ctypes as C
import hashlib
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent
_LIB = None


def clib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = ORACLE_DIR / "liboracle.so"
        if not path.exists():
            raise RuntimeError(f"{path} missing: run __graft_entry__.build() (gcc oracle/*.c)")
        lib = C.CDLL(str(path))
        lib.oracle_sha256.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        lib.oracle_sha256.restype = None
        lib.oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        lib.oracle_philox4x32_10.restype = None
        lib.oracle_rand_f64.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
        lib.oracle_rand_f64.restype = None
        lib.oracle_rand_normal_f64.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_double, C.c_double]
        lib.oracle_rand_normal_f64.restype = None
        _LIB = lib
    return _LIB


# ---- K1 elementwise: numpy ufuncs, bit-exact bar for + - * / sqrt abs neg square max min
EWISE = {
    "add": np.add, "subtract": np.subtract, "multiply": np.multiply, "divide": np.divide,
    "maximum": np.maximum, "minimum": np.minimum, "square": np.square, "absolute": np.absolute,
    "negative": np.negative, "sqrt": np.sqrt, "exp": np.exp, "log": np.log,
    "sin": np.sin, "cos": np.cos, "tanh": np.tanh, "floor": np.floor, "ceil": np.ceil,
    "log1p": np.log1p, "expm1": np.expm1,
}


def ewise(name: str, x: np.ndarray, y=None) -> np.ndarray:
    f = EWISE[name]
    with np.errstate(all="ignore"):
        return f(x) if y is None else f(x, y)


# ---- K2 reductions (axis=None). Bar: <= 1e-5 rel float32, <= 1e-12 rel float64, exact ints
def reduce(name: str, x: np.ndarray):
    with np.errstate(all="ignore"):
        return {"sum": np.sum, "mean": np.mean, "max": np.max, "min": np.min}[name](x)


def benchmark_numpy_body(x: np.ndarray) -> np.float64:
    """examples/benchmark-numpy.py:21 on a given array."""
    return np.sum(np.square(x))


# ---- K6 argmax/argmin: exact
def argext(name: str, x: np.ndarray) -> int:
    return int({"argmax": np.argmax, "argmin": np.argmin}[name](x))


# ---- K7 cumsum. int32/int64/float64: numpy.cumsum (exact / 1e-12).  float32: numpy's own
# sequential float32 accumulation drifts ~1e-4 from the exact prefix sums on 1e7 elements
# (SURVEY.md hard part 2); the device accumulates in float64 and rounds once, so the bar is
# the float64 prefix sum rounded to float32 (this function), plus a looser check vs numpy.
def cumsum(x: np.ndarray) -> np.ndarray:
    if x.dtype == np.float32:
        return np.cumsum(x.astype(np.float64)).astype(np.float32)
    return np.cumsum(x)


# ---- K5 sort / stable argsort: exact (NaNs last; -0.0 == +0.0 compares value-equal)
def sort(x: np.ndarray) -> np.ndarray:
    return np.sort(x)


def argsort_stable(x: np.ndarray) -> np.ndarray:
    return np.argsort(x, kind="stable")


# ---- K3 matmul float32: bar max|dC| / max|C| <= 1e-5 against float64 matmul
def matmul_f64(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return a.astype(np.float64) @ b.astype(np.float64)


# ---- K8 RNG: bit-exact Philox4x32-10 stream + numpy's 53-bit double construction
def rand_f64(n: int, seed: int) -> np.ndarray:
    out = np.empty(int(n), dtype=np.float64)
    clib().oracle_rand_f64(out.ctypes.data, int(n), C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF))
    return out


def rand_normal_f64(n: int, seed: int, loc: float = 0.0, scale: float = 1.0) -> np.ndarray:
    """Box-Muller normals on the Philox stream (counter word 2 = 1): restates bci_rand.cu with libm;
    the device agrees to a few ulp of the transcendental functions, not bit for bit."""
    out = np.empty(int(n), dtype=np.float64)
    clib().oracle_rand_normal_f64(out.ctypes.data, int(n), C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF),
                                  C.c_double(loc), C.c_double(scale))
    return out


def philox4x32_10(ctr, key) -> list[int]:
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    clib().oracle_philox4x32_10(c, k, o)
    return list(o)


# ---- K4 SHA-256
def sha256_c(data: bytes) -> bytes:
    buf = np.frombuffer(data, dtype=np.uint8)
    out = np.empty(32, dtype=np.uint8)
    clib().oracle_sha256(buf.ctypes.data if buf.size else None, buf.size, out.ctypes.data)
    return out.tobytes()


def sha256_hashlib(data: bytes) -> bytes:
    return hashlib.sha256(data).digest()
