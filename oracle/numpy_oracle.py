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


# ---- K8d numpy.random.Generator(PCG64): restatement of numpy/random/src/pcg64/pcg64.h, pinned in
# tests/test_oracle.py against numpy's own `random_raw` and `advance` (the reference's arithmetic
# for the inputs of BASELINE configs C2/C3/C5 is `numpy.random.default_rng(seed)`).
PCG_MULT = (2549297995355413924 << 64) | 4865540595714422341
_M128, _M64 = (1 << 128) - 1, (1 << 64) - 1


def pcg64_output(state: int) -> int:
    """pcg_output_xsl_rr_128_64: rotr64(hi ^ lo, state >> 122)."""
    x, rot = ((state >> 64) ^ state) & _M64, state >> 122
    return ((x >> rot) | (x << ((64 - rot) & 63))) & _M64


def pcg64_raw(state: int, inc: int, n: int) -> list[int]:
    """n successive next64 outputs: step, then output of the NEW state."""
    out = []
    for _ in range(n):
        state = (state * PCG_MULT + inc) & _M128
        out.append(pcg64_output(state))
    return out


def pcg64_jump(k: int) -> tuple[int, int]:
    """(P, G) with s_{i+k} = P*s_i + inc*G: P = A^k, G = 1 + A + ... + A^(k-1) — the affine jump the device
    kernel's tables hold (csrc/bci_rand_pcg.cu jump_of)."""
    P, G, bp, bg = 1, 0, PCG_MULT, 1
    while k:
        if k & 1:
            G = (bp * G + bg) & _M128
            P = (bp * P) & _M128
        bg = (bp * bg + bg) & _M128
        bp = (bp * bp) & _M128
        k >>= 1
    return P, G


def pcg64_random(state: int, inc: int, n: int, kind: str, low: int = 0, span: int = 0) -> np.ndarray:
    """What numpy's Generator returns for n draws from (state, inc), per kind of the device kernel:
    'f64' (next64 >> 11) * 2^-53; 'f32' (next32 >> 8) * 2^-24; 'u32' the next32 stream;
    'bounded' low + ((next32 * span) >> 32) WITHOUT rejection (distributions.c buffered_bounded_lemire_uint32)."""
    draws = n if kind == "f64" else (n + 1) // 2
    raw = np.array(pcg64_raw(state, inc, draws), dtype=np.uint64)
    if kind == "f64":
        return (raw >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    words = np.empty(2 * draws, dtype=np.uint32)
    words[0::2] = (raw & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    words[1::2] = (raw >> np.uint64(32)).astype(np.uint32)
    words = words[:n]
    if kind == "f32":
        return (words >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    if kind == "u32":
        return words
    return (low + ((words.astype(np.uint64) * np.uint64(span)) >> np.uint64(32)).astype(np.int64))
