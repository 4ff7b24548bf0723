"""TEST DOUBLE for `bee_code_interpreter_b200.device` — lets the CPU-only suite exercise the
HOST logic (proxy fusion decisions, broker batching/ownership, executor process model,
storage protocol) without a GPU.  It lives under tests/ and is never importable from the
product; "device memory" is a dict of numpy byte buffers and every op is numpy (the oracle).
It records each batched call so tests can assert how requests were packed."""
from __future__ import annotations

import hashlib
import threading

import numpy as np

from bee_code_interpreter_b200 import _lib as L
from oracle import numpy_oracle as O

_mem: dict[int, np.ndarray] = {}
_next = [1 << 20]
_lock = threading.Lock()
calls: list[tuple] = []           # (name, batch_size, extra)
_initialised = [False]

BIN = {L.EW_ADD: np.add, L.EW_SUB: np.subtract, L.EW_MUL: np.multiply, L.EW_DIV: np.divide,
       L.EW_MAXIMUM: np.maximum, L.EW_MINIMUM: np.minimum,
       L.EW_RSUB: lambda a, s: np.subtract(s, a), L.EW_RDIV: lambda a, s: np.divide(s, a)}
UN = {L.EW_SQUARE: np.square, L.EW_ABS: np.absolute, L.EW_NEG: np.negative, L.EW_SQRT: np.sqrt,
      L.EW_EXP: np.exp, L.EW_LOG: np.log, L.EW_SIN: np.sin, L.EW_COS: np.cos, L.EW_TANH: np.tanh,
      L.EW_FLOOR: np.floor, L.EW_CEIL: np.ceil, L.EW_LOG1P: np.log1p, L.EW_EXPM1: np.expm1}
RED = {L.RED_SUM: np.sum, L.RED_MEAN: np.mean, L.RED_MAX: np.max, L.RED_MIN: np.min}
PRE = {L.PRE_NONE: lambda a: a, L.PRE_SQUARE: np.square, L.PRE_ABS: np.absolute}


_reset_floor = [0]


def reset():
    with _lock:
        _mem.clear()
        calls.clear()
        _reset_floor[0] = _next[0]


def live_buffers() -> int:
    return len(_mem)


def init(device=None):
    _initialised[0] = True
    return 0


def is_initialised():
    return _initialised[0]


def device_info():
    return {"device": 0, "sm_count": 148, "cc": (10, 0), "total_mem": 1 << 37, "name": "fake"}


def mem_info():
    return 1 << 36, 1 << 37


def sync():
    pass


def launch_count():
    return len(calls)


# timing / lifecycle entry points bench.py uses (wall clock here)
BciError, Unsupported = L.BciError, L.Unsupported
_events: dict[int, float] = {}
_t0 = [0.0]


def event_record(slot):
    import time

    _events[slot] = time.perf_counter()


def event_elapsed(a, b):
    return max(1e-6, (_events[b] - _events[a]) * 1e3)


def timer_begin():
    import time

    _t0[0] = time.perf_counter()


def timer_end():
    import time

    return max(1e-6, (time.perf_counter() - _t0[0]) * 1e3)


def flush_l2():
    pass


def shutdown():
    _initialised[0] = False


def alloc(nbytes):
    with _lock:
        p = _next[0]
        _next[0] += (int(nbytes) + 4095) // 4096 * 4096 + 4096
        _mem[p] = np.zeros(max(int(nbytes), 1), dtype=np.uint8)
        return p


def free(ptr):
    with _lock:
        if ptr and ptr >= _reset_floor[0]:  # buffers from before the last reset() are already gone
            _mem.pop(ptr)  # KeyError = double free / foreign pointer: a real bug


def _view(ptr, n, dtype):
    dtype = np.dtype(dtype)
    return _mem[ptr][: int(n) * dtype.itemsize].view(dtype)


def h2d(dst, src):
    src = np.ascontiguousarray(src)
    _mem[dst][: src.nbytes] = src.view(np.uint8).reshape(-1)


def d2h(dst, src):
    dst.view(np.uint8).reshape(-1)[:] = _mem[src][: dst.nbytes]


def d2d(dst, src, nbytes):
    _mem[dst][:nbytes] = _mem[src][:nbytes]


def h2d_fd(dst, fd, offset, nbytes):
    import os

    data = os.pread(fd, nbytes, offset)
    if len(data) < nbytes:
        raise L.BciError(L.ERR_INVALID, "file ends short")
    _mem[dst][:nbytes] = np.frombuffer(data, dtype=np.uint8)


def d2h_fd(fd, offset, src, nbytes):
    import os

    os.pwrite(fd, _mem[src][:nbytes].tobytes(), offset)


def upload(arr):
    arr = np.ascontiguousarray(arr)
    p = alloc(arr.nbytes)
    if arr.nbytes:
        h2d(p, arr)
    return p


def download(ptr, n, dtype, offset=0):
    return _view(ptr, int(offset) + int(n), dtype)[int(offset):].copy()


def ewise(op, dtype, in0, in1, scalars, out, ns):
    calls.append(("ewise", len(in0), op))
    dtype = np.dtype(dtype)
    with np.errstate(all="ignore"):
        for i in range(len(in0)):
            a = _view(in0[i], ns[i], dtype)
            if op in UN:
                r = UN[op](a)
            else:
                b = _view(in1[i], ns[i], dtype) if in1 is not None and in1[i] else dtype.type(scalars[i])
                r = BIN[op](a, b)
            _view(out[i], ns[i], dtype)[:] = r.astype(dtype)


def reduce(op, dtype, in_ptrs, ns, pre=L.PRE_NONE):
    calls.append(("reduce", len(in_ptrs), (op, pre)))
    with np.errstate(all="ignore"):
        return np.array([RED[op](PRE[pre](_view(p, n, dtype))) for p, n in zip(in_ptrs, ns)])


def prepared_reduce(op, dtype, in_ptrs, ns, pre=L.PRE_NONE):
    return lambda: reduce(op, dtype, in_ptrs, ns, pre)


def prepared_reduce_fused(op, ew_op, dtype, in0, in1, scalars, ns, pre=L.PRE_NONE, addends=None):
    return lambda: reduce_fused(op, ew_op, dtype, in0, in1, scalars, ns, pre, addends)


def reduce_fused(op, ew_op, dtype, in0, in1, scalars, ns, pre=L.PRE_NONE, addends=None):
    calls.append(("reduce_fused", len(in0), (op, ew_op, pre, addends is not None)))
    dtype = np.dtype(dtype)
    out = []
    with np.errstate(all="ignore"):
        for i in range(len(in0)):
            a = _view(in0[i], ns[i], dtype)
            b = _view(in1[i], ns[i], dtype) if in1 is not None else dtype.type(scalars[i])
            v = BIN[ew_op](a, b)
            if addends is not None:
                v = v + dtype.type(addends[i])
            out.append(RED[op](PRE[pre](v)))
    return np.array(out)


def argext(op, dtype, in_ptrs, ns):
    calls.append(("argext", len(in_ptrs), op))
    f = np.argmax if op == L.ARGMAX else np.argmin
    return np.array([f(_view(p, n, dtype)) for p, n in zip(in_ptrs, ns)], dtype=np.int64)


def cumsum_out_dtype(dtype):
    dtype = np.dtype(dtype)
    return np.dtype(np.int64) if dtype == np.int32 else dtype


def cumsum(dtype, in_ptrs, out_ptrs, ns):
    calls.append(("cumsum", len(in_ptrs), None))
    for p, o, n in zip(in_ptrs, out_ptrs, ns):
        r = O.cumsum(_view(p, n, dtype))
        _view(o, n, r.dtype)[:] = r


def sort(dtype, in_ptrs, out_ptrs, ns):
    calls.append(("sort", len(in_ptrs), None))
    for p, o, n in zip(in_ptrs, out_ptrs, ns):
        _view(o, n, dtype)[:] = np.sort(_view(p, n, dtype))


def argsort(dtype, in_ptrs, out_ptrs, ns):
    calls.append(("argsort", len(in_ptrs), None))
    for p, o, n in zip(in_ptrs, out_ptrs, ns):
        _view(o, n, np.int64)[:] = np.argsort(_view(p, n, dtype), kind="stable")


def rand_f64(out_ptrs, ns, seeds):
    calls.append(("rand", len(out_ptrs), None))
    for p, n, s in zip(out_ptrs, ns, seeds):
        _view(p, n, np.float64)[:] = O.rand_f64(n, s)


PCG_OUT_DTYPE = {L.PCG_F64: np.dtype(np.float64), L.PCG_F32: np.dtype(np.float32), L.PCG_U32: np.dtype(np.uint32),
                 L.PCG_BOUNDED32_I64: np.dtype(np.int64), L.PCG_BOUNDED32_I32: np.dtype(np.int32)}


def rand_pcg64(kind, out_ptrs, ns, gens):
    """numpy's own PCG64 Generator is the oracle: state in, n draws out."""
    calls.append(("rand_pcg", len(out_ptrs), kind))
    rej = np.zeros(len(out_ptrs), dtype=np.uint32)
    for i, (p, n, g) in enumerate(zip(out_ptrs, ns, gens)):
        bg = np.random.PCG64()
        bg.state = {"bit_generator": "PCG64", "state": {"state": int(g[0]), "inc": int(g[1])}, "has_uint32": 0, "uinteger": 0}
        gen = np.random.Generator(bg)
        if kind == L.PCG_F64:
            out = gen.random(n)
        elif kind == L.PCG_F32:
            out = gen.random(n, dtype=np.float32)
        elif kind == L.PCG_U32:
            out = gen.integers(0, 1 << 32, n, dtype=np.uint32)
        else:
            dt = np.int64 if kind == L.PCG_BOUNDED32_I64 else np.int32
            low, span = int(g[2]), int(g[3])
            out = gen.integers(low, low + span, n, dtype=dt)
            # the real kernel cannot follow rejections: report them like it does
            consumed = (bg.state["state"]["state"], bg.state["has_uint32"])
            ref = np.random.PCG64()
            ref.state = {"bit_generator": "PCG64", "state": {"state": int(g[0]), "inc": int(g[1])}, "has_uint32": 0, "uinteger": 0}
            ref.advance((n + 1) // 2)
            rej[i] = 0 if consumed[0] == ref.state["state"]["state"] else 1
        _view(p, n, PCG_OUT_DTYPE[kind])[:] = out
    return rej


def reduce_rand_f64(op, ns, seeds, pre=L.PRE_NONE):
    calls.append(("reduce_rand", len(ns), (op, pre)))
    return np.array([RED[op](PRE[pre](O.rand_f64(n, s))) for n, s in zip(ns, seeds)], dtype=np.float64)


def fill(kind, dtype, out_ptrs, ns, a, b=None, c=None):
    calls.append(("fill", len(out_ptrs), kind))
    dtype = np.dtype(dtype)
    for i, (p, n) in enumerate(zip(out_ptrs, ns)):
        v = _view(p, n, dtype)
        idx = np.arange(n).astype(dtype)
        if kind == L.FILL_CONST:
            v[:] = dtype.type(a[i])
        elif kind == L.FILL_ARANGE:
            v[:] = dtype.type(a[i]) + idx * dtype.type(b[i])
            if c is not None and n > 1:
                v[1] = dtype.type(c[i])
        else:
            v[:] = idx * dtype.type(b[i]) + dtype.type(a[i])
            if c is not None and n:
                v[-1] = dtype.type(c[i])


def cast(src_dtype, dst_dtype, in_ptrs, out_ptrs, ns):
    calls.append(("cast", len(in_ptrs), (np.dtype(src_dtype).name, np.dtype(dst_dtype).name)))
    for p, o, n in zip(in_ptrs, out_ptrs, ns):
        _view(o, n, dst_dtype)[:] = _view(p, n, src_dtype).astype(dst_dtype)


def rand_normal_f64(out_ptrs, ns, seeds, locs, scales):
    calls.append(("rand_normal", len(out_ptrs), None))
    for p, n, s, lo, sc in zip(out_ptrs, ns, seeds, locs, scales):
        _view(p, n, np.float64)[:] = O.rand_normal_f64(n, s, lo, sc)


def rand_mt19937(out_ptrs, ns, states):
    """numpy's own MT19937 is the oracle here: set_state -> random_sample -> get_state."""
    calls.append(("rand_mt", len(out_ptrs), None))
    states = np.array(states, dtype=np.uint32).reshape(len(out_ptrs), 625)
    for i, (p, n) in enumerate(zip(out_ptrs, ns)):
        rs = np.random.RandomState()
        rs.set_state(("MT19937", states[i, :624].copy(), int(states[i, 624]), 0, 0.0))
        _view(p, n, np.float64)[:] = rs.random_sample(n)
        st = rs.get_state()
        states[i, :624], states[i, 624] = st[1], st[2]
    return states


def matmul_f32(a_ptrs, b_ptrs, c_ptrs, M, N, K):
    calls.append(("matmul", len(a_ptrs), (M, N, K)))
    for a, b, c in zip(a_ptrs, b_ptrs, c_ptrs):
        _view(c, M * N, np.float32)[:] = (_view(a, M * K, np.float32).reshape(M, K) @
                                          _view(b, K * N, np.float32).reshape(K, N)).reshape(-1)


def sha256_device(ptrs, lens):
    calls.append(("sha256", len(ptrs), None))
    return [hashlib.sha256(_mem[p][:n].tobytes()).digest() for p, n in zip(ptrs, lens)]


def sha256_host(buffers):
    calls.append(("sha256", len(buffers), None))
    return [hashlib.sha256(bytes(b)).digest() for b in buffers]


_streams: dict[int, "hashlib._Hash"] = {}
_stream_ids = [0]


class Sha256Stream:
    """Serialises to a fixed-size blob like bci_sha256_state (an id into a registry here)."""

    def __init__(self):
        self.h = hashlib.sha256()

    def to_bytes(self):
        with _lock:
            _stream_ids[0] += 1
            sid = _stream_ids[0]
            _streams[sid] = self.h.copy()
        return sid.to_bytes(8, "little").ljust(112, b"\0")

    @classmethod
    def from_bytes(cls, raw):
        obj = cls()
        obj.h = _streams[int.from_bytes(raw[:8], "little")].copy()
        return obj


def sha256_update_many(streams, chunks):
    calls.append(("sha256_update", len(streams), None))
    for s, c in zip(streams, chunks):
        s.h.update(c)


def sha256_final_many(streams):
    calls.append(("sha256_final", len(streams), None))
    return [s.h.digest() for s in streams]
