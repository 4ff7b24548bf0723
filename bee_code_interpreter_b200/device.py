"""Batch-first host API over the C ABI (include/bci.h): one call = one launch per op class
for B independent requests.  Everything here is thin marshalling; the work happens in the
sm_100a kernels.  No CPU fallback: errors propagate as `BciError`."""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, Sequence

import numpy as np

from . import _lib
from ._lib import BciError, Unsupported  # noqa: F401  (re-exported)

DTYPE_CODE = {
    np.dtype(np.float32): _lib.F32,
    np.dtype(np.float64): _lib.F64,
    np.dtype(np.int32): _lib.I32,
    np.dtype(np.int64): _lib.I64,
}
CODE_DTYPE = {v: k for k, v in DTYPE_CODE.items()}

_initialised_device: int | None = None


def dtype_code(dtype) -> int:
    try:
        return DTYPE_CODE[np.dtype(dtype)]
    except KeyError:
        raise Unsupported(_lib.ERR_UNSUPPORTED, f"dtype {dtype} is not handled on the device") from None


def default_device() -> int:
    for var in ("BCI_DEVICE", "LOCAL_RANK"):
        v = os.environ.get(var)
        if v not in (None, ""):
            return int(v)
    return 0


def init(device: int | None = None) -> int:
    """Binds this process to one GPU (one process per GPU).  Idempotent."""
    global _initialised_device
    lib = _lib.load()
    dev = default_device() if device is None else int(device)
    _lib.check(lib.bci_init(dev))
    _initialised_device = dev
    return dev


def is_initialised() -> bool:
    return _initialised_device is not None


def shutdown() -> None:
    global _initialised_device
    _lib.load().bci_shutdown()
    _initialised_device = None


def device_count() -> int:
    n = C.c_int(0)
    _lib.check(_lib.load().bci_device_count(C.byref(n)))
    return n.value


def device_info() -> dict:
    info = _lib.DevInfo()
    _lib.check(_lib.load().bci_device_info(C.byref(info)))
    return {
        "device": info.device,
        "sm_count": info.sm_count,
        "cc": (info.cc_major, info.cc_minor),
        "total_mem": info.total_mem,
        "name": info.name.decode(),
    }


def mem_info() -> tuple[int, int]:
    f, t = C.c_size_t(0), C.c_size_t(0)
    _lib.check(_lib.load().bci_mem_info(C.byref(f), C.byref(t)))
    return f.value, t.value


def alloc(nbytes: int) -> int:
    p = C.c_void_p(0)
    _lib.check(_lib.load().bci_alloc(int(nbytes), C.byref(p)))
    return int(p.value)


def free(ptr: int) -> None:
    if ptr:
        _lib.check(_lib.load().bci_free(C.c_void_p(ptr)))


def sync() -> None:
    _lib.check(_lib.load().bci_sync())


def flush_l2() -> None:
    _lib.check(_lib.load().bci_flush_l2())


def launch_count() -> int:
    return int(_lib.load().bci_launch_count())


def timer_begin() -> None:
    _lib.check(_lib.load().bci_timer_begin())


def timer_end() -> float:
    ms = C.c_float(0.0)
    _lib.check(_lib.load().bci_timer_end(C.byref(ms)))
    return float(ms.value)


def event_record(slot: int) -> None:
    _lib.check(_lib.load().bci_event_record(int(slot)))


def event_elapsed(slot_begin: int, slot_end: int) -> float:
    ms = C.c_float(0.0)
    _lib.check(_lib.load().bci_event_elapsed(int(slot_begin), int(slot_end), C.byref(ms)))
    return float(ms.value)


class PinnedBuffer:
    """Page-locked host memory exposed as a numpy uint8 array (for H2D/D2H staging)."""

    def __init__(self, nbytes: int):
        p = C.c_void_p(0)
        _lib.check(_lib.load().bci_host_alloc(int(nbytes), C.byref(p)))
        self.ptr = int(p.value)
        self.nbytes = int(nbytes)
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(1, self.nbytes)).from_address(self.ptr))[: self.nbytes]

    def close(self) -> None:
        if self.ptr:
            self.array = None
            _lib.load().bci_host_free(C.c_void_p(self.ptr))
            self.ptr = 0

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass


def h2d(dst_ptr: int, src: np.ndarray) -> None:
    src = np.ascontiguousarray(src)
    _lib.check(_lib.load().bci_h2d(C.c_void_p(dst_ptr), C.c_void_p(src.ctypes.data), src.nbytes))
    # pageable sources are staged by the runtime before the call returns; pinned ones must
    # stay alive until the next sync — callers that pass pinned memory sync themselves.


def d2h(dst: np.ndarray, src_ptr: int) -> None:
    assert dst.flags["C_CONTIGUOUS"]
    _lib.check(_lib.load().bci_d2h(C.c_void_p(dst.ctypes.data), C.c_void_p(src_ptr), dst.nbytes))


def d2d(dst_ptr: int, src_ptr: int, nbytes: int) -> None:
    _lib.check(_lib.load().bci_d2d(C.c_void_p(dst_ptr), C.c_void_p(src_ptr), int(nbytes)))


def upload(arr: np.ndarray) -> int:
    """Allocates a device buffer and copies `arr` into it; returns the device pointer."""
    arr = np.ascontiguousarray(arr)
    p = alloc(arr.nbytes)
    if arr.nbytes:
        h2d(p, arr)
    return p


def download(ptr: int, n: int, dtype, offset: int = 0) -> np.ndarray:
    """`n` elements of `dtype`, starting `offset` elements into the buffer."""
    out = np.empty(int(n), dtype=dtype)
    if out.nbytes:
        d2h(out, ptr + int(offset) * out.dtype.itemsize)
    return out


# ---- batched ops ---------------------------------------------------------------------
def _scalars(values: Sequence | None, code: int, B: int):
    if values is None:
        return None
    arr = (_lib.Scalar * B)()
    is_float = code in (_lib.F32, _lib.F64)
    for i, v in enumerate(values):
        if is_float:
            arr[i].f = float(v)
        else:
            arr[i].i = int(v)
    return arr


def ewise(op: int, dtype, in0: Sequence[int], in1: Sequence[int] | None, scalars: Sequence | None,
          out: Sequence[int], ns: Sequence[int]) -> None:
    code = dtype_code(dtype)
    B = len(in0)
    _lib.check(_lib.load().bci_ewise(
        op, code, B, _lib.ptr_array(in0), _lib.ptr_array(in1) if in1 is not None else None,
        _scalars(scalars, code, B), _lib.ptr_array(out), _lib.size_array(ns)))


def _decode_results(raw: np.ndarray, op: int, code: int) -> np.ndarray:
    if op == _lib.RED_MEAN:
        res_code = _lib.F32 if code == _lib.F32 else _lib.F64
    elif op == _lib.RED_SUM:
        res_code = _lib.I64 if code == _lib.I32 else code
    else:
        res_code = code
    dt = CODE_DTYPE[res_code]
    if dt.itemsize == 8:
        return raw.view(dt).copy()
    return raw.view(np.uint32)[::2].view(dt).copy()


def reduce(op: int, dtype, in_ptrs: Sequence[int], ns: Sequence[int], pre: int = _lib.PRE_NONE) -> np.ndarray:
    code = dtype_code(dtype)
    B = len(in_ptrs)
    raw = np.zeros(B, dtype=np.uint64)
    _lib.check(_lib.load().bci_reduce(op, pre, code, B, _lib.ptr_array(in_ptrs), _lib.size_array(ns),
                                      C.c_void_p(raw.ctypes.data)))
    return _decode_results(raw, op, code)


def reduce_fused(op: int, ew_op: int, dtype, in0: Sequence[int], in1: Sequence[int] | None,
                 scalars: Sequence | None, ns: Sequence[int], pre: int = _lib.PRE_NONE,
                 addends: Sequence | None = None) -> np.ndarray:
    code = dtype_code(dtype)
    B = len(in0)
    raw = np.zeros(B, dtype=np.uint64)
    _lib.check(_lib.load().bci_reduce_fused(
        op, pre, ew_op, code, B, _lib.ptr_array(in0), _lib.ptr_array(in1) if in1 is not None else None,
        _scalars(scalars, code, B), _scalars(addends, code, B), _lib.size_array(ns),
        C.c_void_p(raw.ctypes.data)))
    return _decode_results(raw, op, code)


def prepared_reduce(op: int, dtype, in_ptrs: Sequence[int], ns: Sequence[int], pre: int = _lib.PRE_NONE):
    """`reduce` with the pointer / size tables marshalled once: returns a callable that launches and returns the
    results.  The broker builds these tables in C++; a Python caller that launches the same batch repeatedly
    (bench_configs.py) would otherwise spend 0.1-0.3 ms per call converting 1024-entry lists (tools/prof_c3.py)."""
    code = dtype_code(dtype)
    B = len(in_ptrs)
    raw = np.zeros(B, dtype=np.uint64)
    fn, a0, n, out = _lib.load().bci_reduce, _lib.ptr_array(in_ptrs), _lib.size_array(ns), C.c_void_p(raw.ctypes.data)

    def call() -> np.ndarray:
        _lib.check(fn(op, pre, code, B, a0, n, out))
        return _decode_results(raw, op, code)
    return call


def prepared_reduce_fused(op: int, ew_op: int, dtype, in0: Sequence[int], in1: Sequence[int] | None,
                          scalars: Sequence | None, ns: Sequence[int], pre: int = _lib.PRE_NONE,
                          addends: Sequence | None = None):
    """`reduce_fused` with its tables marshalled once (see prepared_reduce)."""
    code = dtype_code(dtype)
    B = len(in0)
    raw = np.zeros(B, dtype=np.uint64)
    fn = _lib.load().bci_reduce_fused
    a0, a1 = _lib.ptr_array(in0), (_lib.ptr_array(in1) if in1 is not None else None)
    sc, ad, n, out = _scalars(scalars, code, B), _scalars(addends, code, B), _lib.size_array(ns), C.c_void_p(raw.ctypes.data)

    def call() -> np.ndarray:
        _lib.check(fn(op, pre, ew_op, code, B, a0, a1, sc, ad, n, out))
        return _decode_results(raw, op, code)
    return call


def argext(op: int, dtype, in_ptrs: Sequence[int], ns: Sequence[int]) -> np.ndarray:
    code = dtype_code(dtype)
    B = len(in_ptrs)
    out = np.zeros(B, dtype=np.int64)
    _lib.check(_lib.load().bci_argext(op, code, B, _lib.ptr_array(in_ptrs), _lib.size_array(ns),
                                      out.ctypes.data_as(C.POINTER(C.c_int64))))
    return out


def cumsum_out_dtype(dtype) -> np.dtype:
    dtype = np.dtype(dtype)
    return np.dtype(np.int64) if dtype == np.int32 else dtype


def cumsum(dtype, in_ptrs: Sequence[int], out_ptrs: Sequence[int], ns: Sequence[int]) -> None:
    code = dtype_code(dtype)
    _lib.check(_lib.load().bci_cumsum(code, len(in_ptrs), _lib.ptr_array(in_ptrs), _lib.ptr_array(out_ptrs),
                                      _lib.size_array(ns)))


def sort(dtype, in_ptrs: Sequence[int], out_ptrs: Sequence[int], ns: Sequence[int]) -> None:
    code = dtype_code(dtype)
    _lib.check(_lib.load().bci_sort(code, len(in_ptrs), _lib.ptr_array(in_ptrs), _lib.ptr_array(out_ptrs),
                                    _lib.size_array(ns)))


def argsort(dtype, in_ptrs: Sequence[int], out_idx_ptrs: Sequence[int], ns: Sequence[int]) -> None:
    code = dtype_code(dtype)
    _lib.check(_lib.load().bci_argsort(code, len(in_ptrs), _lib.ptr_array(in_ptrs),
                                       _lib.ptr_array(out_idx_ptrs), _lib.size_array(ns)))


def rand_f64(out_ptrs: Sequence[int], ns: Sequence[int], seeds: Iterable[int]) -> None:
    seeds = [int(s) & 0xFFFFFFFFFFFFFFFF for s in seeds]
    arr = (C.c_uint64 * len(seeds))(*seeds)
    _lib.check(_lib.load().bci_rand_f64(len(out_ptrs), _lib.ptr_array(out_ptrs), _lib.size_array(ns), arr))


def rand_normal_f64(out_ptrs: Sequence[int], ns: Sequence[int], seeds: Iterable[int], locs: Sequence[float],
                    scales: Sequence[float]) -> None:
    seeds = [int(s) & 0xFFFFFFFFFFFFFFFF for s in seeds]
    B = len(out_ptrs)
    _lib.check(_lib.load().bci_rand_normal_f64(B, _lib.ptr_array(out_ptrs), _lib.size_array(ns), (C.c_uint64 * B)(*seeds),
                                               (C.c_double * B)(*[float(v) for v in locs]),
                                               (C.c_double * B)(*[float(v) for v in scales])))


def rand_mt19937(out_ptrs: Sequence[int], ns: Sequence[int], states: np.ndarray) -> np.ndarray:
    """numpy's legacy generator: `states` is (B, 625) uint32 = key[624] + pos per request
    (numpy.random.get_state()); returns the states after drawing ns[i] doubles into out_ptrs[i]."""
    st = np.array(states, dtype=np.uint32, order="C").reshape(len(out_ptrs), 625)
    _lib.check(_lib.load().bci_rand_mt19937_f64(len(out_ptrs), _lib.ptr_array(out_ptrs), _lib.size_array(ns),
                                                st.ctypes.data_as(C.POINTER(C.c_uint32))))
    return st


PCG_OUT_DTYPE = {_lib.PCG_F64: np.dtype(np.float64), _lib.PCG_F32: np.dtype(np.float32), _lib.PCG_U32: np.dtype(np.uint32),
                 _lib.PCG_BOUNDED32_I64: np.dtype(np.int64), _lib.PCG_BOUNDED32_I32: np.dtype(np.int32)}


def rand_pcg64(kind: int, out_ptrs: Sequence[int], ns: Sequence[int], gens: Sequence[tuple]) -> np.ndarray:
    """numpy.random.Generator(PCG64) streams: gens[i] = (state, inc[, low, range]) with the 128-bit
    integers of `bit_generator.state["state"]`.  Returns the per-request count of draws numpy would
    have rejected (bounded kinds; 0 everywhere = the output IS numpy's stream)."""
    B = len(out_ptrs)
    arr = (_lib.Pcg64Gen * B)()
    m64 = (1 << 64) - 1
    for i, g in enumerate(gens):
        state, inc = int(g[0]), int(g[1])
        arr[i].state_lo, arr[i].state_hi = state & m64, (state >> 64) & m64
        arr[i].inc_lo, arr[i].inc_hi = inc & m64, (inc >> 64) & m64
        if len(g) > 2:
            arr[i].low, arr[i].range = int(g[2]), int(g[3])
    bounded = kind in (_lib.PCG_BOUNDED32_I64, _lib.PCG_BOUNDED32_I32)
    rej = np.zeros(B, dtype=np.uint32)
    _lib.check(_lib.load().bci_rand_pcg64(int(kind), B, _lib.ptr_array(out_ptrs), _lib.size_array(ns), arr,
                                          rej.ctypes.data_as(C.POINTER(C.c_uint32)) if bounded else None))
    return rej


def reduce_rand_f64(op: int, ns: Sequence[int], seeds: Iterable[int], pre: int = _lib.PRE_NONE) -> np.ndarray:
    """reduce_op(pre(rand(n))) with the operand generated in registers (same stream as rand_f64)."""
    seeds = [int(s) & 0xFFFFFFFFFFFFFFFF for s in seeds]
    B = len(seeds)
    raw = np.zeros(B, dtype=np.uint64)
    _lib.check(_lib.load().bci_reduce_rand_f64(op, pre, B, _lib.size_array(ns), (C.c_uint64 * B)(*seeds),
                                               C.c_void_p(raw.ctypes.data)))
    return raw.view(np.float64).copy()


def fill(kind: int, dtype, out_ptrs: Sequence[int], ns: Sequence[int], a: Sequence, b: Sequence | None = None,
         c: Sequence | None = None) -> None:
    code = dtype_code(dtype)
    B = len(out_ptrs)
    _lib.check(_lib.load().bci_fill(int(kind), code, B, _lib.ptr_array(out_ptrs), _lib.size_array(ns),
                                    _scalars(a, code, B), _scalars(b, code, B), _scalars(c, code, B)))


def cast(src_dtype, dst_dtype, in_ptrs: Sequence[int], out_ptrs: Sequence[int], ns: Sequence[int]) -> None:
    _lib.check(_lib.load().bci_cast(dtype_code(src_dtype), dtype_code(dst_dtype), len(in_ptrs), _lib.ptr_array(in_ptrs),
                                    _lib.ptr_array(out_ptrs), _lib.size_array(ns)))


def h2d_fd(dst_ptr: int, fd: int, offset: int, nbytes: int) -> None:
    _lib.check(_lib.load().bci_h2d_fd(C.c_void_p(dst_ptr), int(fd), int(offset), int(nbytes)))


def d2h_fd(fd: int, offset: int, src_ptr: int, nbytes: int) -> None:
    _lib.check(_lib.load().bci_d2h_fd(int(fd), int(offset), C.c_void_p(src_ptr), int(nbytes)))


def matmul_f32(a_ptrs: Sequence[int], b_ptrs: Sequence[int], c_ptrs: Sequence[int], M: int, N: int, K: int) -> None:
    _lib.check(_lib.load().bci_matmul_f32(len(a_ptrs), _lib.ptr_array(a_ptrs), _lib.ptr_array(b_ptrs),
                                          _lib.ptr_array(c_ptrs), int(M), int(N), int(K)))


def sha256_device(ptrs: Sequence[int], lens: Sequence[int]) -> list[bytes]:
    """SHA-256 of B device-resident messages, one launch."""
    B = len(ptrs)
    out = np.zeros(B * 32, dtype=np.uint8)
    _lib.check(_lib.load().bci_sha256(B, _lib.ptr_array(ptrs), _lib.size_array(lens), 1,
                                      C.c_void_p(out.ctypes.data)))
    return [out[32 * i: 32 * i + 32].tobytes() for i in range(B)]


def sha256_host(buffers: Sequence[bytes | bytearray | memoryview | np.ndarray]) -> list[bytes]:
    """SHA-256 of B host buffers: copied to the device and hashed in one launch."""
    B = len(buffers)
    arrs = [np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b.view(np.uint8).reshape(-1)
            for b in buffers]
    ptrs = [a.ctypes.data if a.size else 0 for a in arrs]
    out = np.zeros(B * 32, dtype=np.uint8)
    _lib.check(_lib.load().bci_sha256(B, _lib.ptr_array(ptrs), _lib.size_array([a.size for a in arrs]), 0,
                                      C.c_void_p(out.ctypes.data)))
    return [out[32 * i: 32 * i + 32].tobytes() for i in range(B)]


class Sha256Stream:
    """Incremental hasher state for one ObjectWriter; chunks from many writers are
    absorbed together with `sha256_update_many` (one launch per round of chunks)."""

    def __init__(self):
        self.state = _lib.Sha256State()
        _lib.check(_lib.load().bci_sha256_init(C.byref(self.state)))

    def to_bytes(self) -> bytes:
        return bytes(self.state)

    @classmethod
    def from_bytes(cls, raw: bytes) -> "Sha256Stream":
        obj = cls.__new__(cls)
        obj.state = _lib.Sha256State.from_buffer_copy(raw)
        return obj


def sha256_update_many(streams: Sequence[Sha256Stream], chunks: Sequence[bytes]) -> None:
    B = len(streams)
    st = (C.POINTER(_lib.Sha256State) * B)(*[C.pointer(s.state) for s in streams])
    arrs = [np.frombuffer(c, dtype=np.uint8) for c in chunks]
    ptrs = [a.ctypes.data if a.size else 0 for a in arrs]
    _lib.check(_lib.load().bci_sha256_update_batch(B, st, _lib.ptr_array(ptrs),
                                                   _lib.size_array([a.size for a in arrs])))


def sha256_final_many(streams: Sequence[Sha256Stream]) -> list[bytes]:
    B = len(streams)
    st = (C.POINTER(_lib.Sha256State) * B)(*[C.pointer(s.state) for s in streams])
    out = np.zeros(B * 32, dtype=np.uint8)
    _lib.check(_lib.load().bci_sha256_final_batch(B, st, C.c_void_p(out.ctypes.data)))
    return [out[32 * i: 32 * i + 32].tobytes() for i in range(B)]
