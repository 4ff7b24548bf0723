"""Device-resident stand-ins for numpy arrays inside agent code (boundary ii).

`DeviceArray` is an ndarray duck (`__array__`, `__array_ufunc__`, `__array_function__`,
operators) whose data lives in HBM.  Calls numpy would run on the CPU — the ufuncs and
reductions of examples/benchmark-numpy.py:20-21, np.sort/argmax/cumsum/matmul of the other
BASELINE configs — are forwarded to a *backend* that speaks the batch-first device API
(`device.py` in-process, or `broker_client.RemoteBackend` from a sandboxed script process).
Anything the kernels do not take (unsupported dtype/shape/kwarg, or an `Unsupported` from
the library) is evaluated by numpy itself on a host copy: a script never fails where numpy
would not.

Single-use temporaries are not materialised: `np.square(x)`, `x*y+1.0`, `np.abs(x-y)` return
a *deferred* DeviceArray; if the next consumer is a full reduction the whole expression runs
as one fused kernel (bci_reduce / bci_reduce_fused), otherwise the elementwise kernel runs
first.  Every intermediate is rounded exactly as numpy's separate ufunc calls round it.
"""
from __future__ import annotations

import operator
import weakref

import numpy as np
from numpy.lib.mixins import NDArrayOperatorsMixin

from . import _lib as L

SUPPORTED = (np.dtype(np.float32), np.dtype(np.float64), np.dtype(np.int32), np.dtype(np.int64))

BINARY_UFUNCS = {np.add: L.EW_ADD, np.subtract: L.EW_SUB, np.multiply: L.EW_MUL, np.divide: L.EW_DIV,
                 np.true_divide: L.EW_DIV, np.maximum: L.EW_MAXIMUM, np.minimum: L.EW_MINIMUM}
REVERSED = {L.EW_SUB: L.EW_RSUB, L.EW_DIV: L.EW_RDIV}  # scalar on the left
COMMUTATIVE = {L.EW_ADD, L.EW_MUL, L.EW_MAXIMUM, L.EW_MINIMUM}
UNARY_UFUNCS = {np.square: L.EW_SQUARE, np.absolute: L.EW_ABS, np.fabs: L.EW_ABS, np.negative: L.EW_NEG,
                np.sqrt: L.EW_SQRT, np.exp: L.EW_EXP, np.log: L.EW_LOG, np.sin: L.EW_SIN, np.cos: L.EW_COS,
                np.tanh: L.EW_TANH, np.floor: L.EW_FLOOR, np.ceil: L.EW_CEIL, np.log1p: L.EW_LOG1P,
                np.expm1: L.EW_EXPM1}
FLOAT_ONLY = {L.EW_DIV, L.EW_RDIV, L.EW_SQRT, L.EW_EXP, L.EW_LOG, L.EW_SIN, L.EW_COS, L.EW_TANH, L.EW_FLOOR,
              L.EW_CEIL, L.EW_LOG1P, L.EW_EXPM1}
FUSABLE_PRODUCERS = {L.EW_ADD, L.EW_SUB, L.EW_MUL, L.EW_MAXIMUM, L.EW_MINIMUM, L.EW_RSUB}
PRE_OF = {L.EW_SQUARE: L.PRE_SQUARE, L.EW_ABS: L.PRE_ABS}
REDUCE_UFUNCS = {np.add: L.RED_SUM, np.maximum: L.RED_MAX, np.minimum: L.RED_MIN}


class _Fallback(Exception):
    """Raised internally when an operation must be evaluated by numpy on the host."""


def _weak_scalar_ok(value, dtype: np.dtype) -> bool:
    """numpy (NEP 50) keeps the array dtype for python scalars; numpy-typed scalars only
    when their dtype is the array's own."""
    if isinstance(value, (bool, np.bool_)):
        return False
    if isinstance(value, np.generic):
        return value.dtype == dtype
    if isinstance(value, int):
        if dtype.kind == "f":
            return True
        info = np.iinfo(dtype)
        return info.min <= value <= info.max
    if isinstance(value, float):
        return dtype.kind == "f"
    return False


class DeviceArray(NDArrayOperatorsMixin):
    __array_priority__ = 1000.0

    def __init__(self, backend, ptr, shape, dtype, *, deferred=None):
        self._backend = backend
        self._ptr = ptr                      # device pointer/handle, or None while deferred/host-only
        self._shape = tuple(int(s) for s in shape)
        self._dtype = np.dtype(dtype)
        self._host = None                    # cached host copy
        self._deferred = deferred            # ("unary", op, a) | ("binary", op, a, b_or_None, scalar) | ("rand", seed)
        self._dependents = {}                # id -> weakref of deferred arrays that read this one
        if deferred is not None:
            for src in deferred[2:4]:
                if isinstance(src, DeviceArray):
                    src._dependents[id(self)] = weakref.ref(self)

    # ---- construction -------------------------------------------------------------
    @classmethod
    def from_host(cls, backend, arr: np.ndarray) -> "DeviceArray":
        arr = np.ascontiguousarray(arr)
        if arr.dtype not in SUPPORTED:
            raise _Fallback(f"dtype {arr.dtype}")
        ptr = backend.upload(arr)
        out = cls(backend, ptr, arr.shape, arr.dtype)
        out._host = arr
        return out

    # ---- basic ndarray attributes ---------------------------------------------------
    @property
    def shape(self): return self._shape
    @property
    def dtype(self): return self._dtype
    @property
    def ndim(self): return len(self._shape)
    @property
    def size(self):
        n = 1
        for s in self._shape:
            n *= s
        return n
    @property
    def nbytes(self): return self.size * self._dtype.itemsize
    @property
    def itemsize(self): return self._dtype.itemsize
    @property
    def T(self): return self._expose_host().T

    def __len__(self):
        if not self._shape:
            raise TypeError("len() of unsized object")
        return self._shape[0]

    # ---- residency --------------------------------------------------------------------
    def _device_ptr(self):
        """Device pointer, running the deferred elementwise kernel or the upload if needed."""
        if self._ptr is None:
            if self._deferred is not None:
                self._materialise_deferred()
            elif self._host is not None:
                self._ptr = self._backend.upload(np.ascontiguousarray(self._host))
            else:  # pragma: no cover
                raise RuntimeError("DeviceArray has neither device nor host data")
        return self._ptr

    def _materialise_deferred(self):
        d = self._deferred
        n = self.size
        if d[0] == "rand":        # numpy.random.rand(n) nobody has read yet: now somebody does
            out = self._backend.rand(n, d[1])
        elif d[0] == "unary":
            _, op, a = d
            out = self._backend.ewise_new(op, self._dtype, a._device_ptr(), None, None, n)
        else:
            _, op, a, b, scalar = d
            out = self._backend.ewise_new(op, self._dtype, a._device_ptr(),
                                          b._device_ptr() if b is not None else None,
                                          scalar if b is None else None, n)
        self._ptr = out
        self._deferred = None

    def _host_view(self) -> np.ndarray:
        """The cached host copy for INTERNAL, read-only use (printing, fallbacks, indexing that
        copies).  Anything that hands the cache out where agent code could write to it goes
        through `_expose_host` instead."""
        if self._host is None:
            try:
                ptr = self._device_ptr()
                self._host = self._backend.download(ptr, self.size, self._dtype).reshape(self._shape)
            except L.BciError:
                self._host = self._eval_on_host()
        return self._host

    def _expose_host(self) -> np.ndarray:
        """The host copy is about to be handed out WRITABLE (np.asarray(x), a slice view, a
        mutating ndarray method): from here on it is the single source of truth — the device
        copy is dropped and uploaded again only if a later device op needs it."""
        host = self._host_view()
        if not host.flags.writeable:
            host = self._host = host.copy()
        self._before_mutation()
        if self._ptr is not None:
            self._backend.free(self._ptr)
            self._ptr = None
        self._deferred = None
        return host

    def __array__(self, dtype=None, copy=None):
        if copy or (dtype is not None and np.dtype(dtype) != self._dtype):
            arr = self._host_view()
            return arr.astype(dtype) if dtype is not None and np.dtype(dtype) != arr.dtype else arr.copy()
        return self._expose_host()

    def _eval_on_host(self) -> np.ndarray:
        """Deferred expression evaluated by numpy (device unavailable for it)."""
        d = self._deferred
        if d is None:
            raise L.BciError(L.ERR_CUDA, "device array lost: neither a host copy nor a producer to re-run")
        if d[0] == "rand":       # the unseeded legacy generator: any fresh stream is a valid answer
            return np.random.random_sample(self._shape)
        if d[0] == "unary":
            return _UNARY_NP[d[1]](_h(d[2]))
        _, op, a, b, scalar = d
        rhs = _h(b) if b is not None else self._dtype.type(scalar)
        return _BINARY_NP[op](_h(a), rhs)

    def _assign(self, value) -> None:
        """`self[...] = value` for a whole-array result (same shape): take its data over."""
        self._before_mutation()
        if isinstance(value, DeviceArray) and value.shape == self._shape and value.dtype == self._dtype:
            ptr = value._device_ptr()
            if self._ptr is not None and self._ptr != ptr:
                self._backend.free(self._ptr)
            self._ptr, value._ptr = ptr, None      # ownership moves; `value` is a dead temporary
            self._host, self._deferred = value._host, None
            return
        host = _h(value)
        if host.shape != self._shape:
            host = np.broadcast_to(host, self._shape)
        self._host = np.array(host, dtype=self._dtype)   # same-kind cast like numpy's out=
        if self._ptr is not None:
            self._backend.free(self._ptr)
            self._ptr = None
        self._deferred = None

    def _before_mutation(self):
        for ref in list(self._dependents.values()):
            dep = ref()
            if dep is not None and dep._deferred is not None:
                dep._device_ptr()
        self._dependents = {}

    def __del__(self):
        try:
            if self._ptr is not None and self._backend is not None:
                self._backend.free(self._ptr)
        except Exception:
            pass

    # ---- protocols --------------------------------------------------------------------
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        try:
            if method == "__call__" and not kwargs:
                if ufunc in UNARY_UFUNCS and len(inputs) == 1:
                    return _unary(UNARY_UFUNCS[ufunc], inputs[0])
                if ufunc in BINARY_UFUNCS and len(inputs) == 2:
                    return _binary(BINARY_UFUNCS[ufunc], inputs[0], inputs[1])
                if ufunc is np.matmul and len(inputs) == 2:
                    return matmul(inputs[0], inputs[1])
                if ufunc is np.power and len(inputs) == 2 and _is_two(inputs[1]) and isinstance(inputs[0], DeviceArray) \
                        and _weak_scalar_ok(inputs[1], inputs[0].dtype):
                    return _unary(L.EW_SQUARE, inputs[0])   # numpy's power loops square for a scalar exponent of 2
            if method == "reduce" and ufunc in REDUCE_UFUNCS and len(inputs) == 1:
                # ufunc.reduce defaults to axis=0 (NOT a full reduction for ndim > 1)
                axis = kwargs.get("axis", 0)
                full = axis is None or (getattr(inputs[0], "ndim", 0) == 1 and axis in (0, -1))
                if full and _plain_reduce_kwargs(kwargs):
                    return _reduce(REDUCE_UFUNCS[ufunc], inputs[0])
        except (_Fallback, L.Unsupported, L.BciError):
            pass
        if "out" in kwargs:
            # in-place forms (`x += y`, `np.add(x, y, out=x)`): evaluate without `out`, then the
            # proxy takes over the result's buffer — numpy semantics, no aliasing hazards
            outs = kwargs["out"]
            outs = outs if isinstance(outs, tuple) else (outs,)
            if len(outs) == 1 and isinstance(outs[0], DeviceArray) and method == "__call__":
                rest = {k: v for k, v in kwargs.items() if k != "out"}
                result = self.__array_ufunc__(ufunc, method, *inputs, **rest)
                if result is NotImplemented:
                    return NotImplemented
                outs[0]._assign(result)
                return outs[0]
            if any(isinstance(o, DeviceArray) for o in outs):
                return NotImplemented
        host_inputs = tuple(_h(x) if isinstance(x, DeviceArray) else x for x in inputs)
        return getattr(ufunc, method)(*host_inputs, **kwargs)

    def __array_function__(self, func, types, args, kwargs):
        handler = _FUNCTIONS.get(func)
        if handler is not None:
            try:
                return handler(*args, **kwargs)
            except (_Fallback, L.Unsupported, L.BciError, TypeError):
                pass
        if isinstance(kwargs.get("out"), DeviceArray):   # numpy writes into the host copy: it becomes the truth
            kwargs = dict(kwargs, out=kwargs["out"]._expose_host())
        return func(*_to_host(args), **_to_host(kwargs))

    # ---- methods numpy's own wrappers and agent code call --------------------------------
    def sum(self, axis=None, dtype=None, out=None, **kw): return _method_reduce(np.sum, L.RED_SUM, self, axis, dtype, out, kw)
    def mean(self, axis=None, dtype=None, out=None, **kw): return _method_reduce(np.mean, L.RED_MEAN, self, axis, dtype, out, kw)
    def max(self, axis=None, out=None, **kw): return _method_reduce(np.max, L.RED_MAX, self, axis, None, out, kw)
    def min(self, axis=None, out=None, **kw): return _method_reduce(np.min, L.RED_MIN, self, axis, None, out, kw)
    def argmax(self, axis=None, out=None, **kw): return argmax(self, axis=axis, out=out, **kw)
    def argmin(self, axis=None, out=None, **kw): return argmin(self, axis=axis, out=out, **kw)
    def cumsum(self, axis=None, dtype=None, out=None): return cumsum(self, axis=axis, dtype=dtype, out=out)
    def dot(self, other):
        try:
            return _dot(self, other)
        except (_Fallback, L.Unsupported, L.BciError):
            return np.dot(_h(self), _h(other))

    def var(self, axis=None, dtype=None, out=None, ddof=0, **kw):
        try:
            return _var(self, axis=axis, dtype=dtype, out=out, ddof=ddof, **kw)
        except (_Fallback, L.Unsupported, L.BciError):
            return _h(self).var(axis=axis, dtype=dtype, out=out, ddof=ddof, **kw)

    def std(self, axis=None, dtype=None, out=None, ddof=0, **kw):
        try:
            return _std(self, axis=axis, dtype=dtype, out=out, ddof=ddof, **kw)
        except (_Fallback, L.Unsupported, L.BciError):
            return _h(self).std(axis=axis, dtype=dtype, out=out, ddof=ddof, **kw)

    def astype(self, dtype, *a, **kw):
        if np.dtype(dtype) == self._dtype and not a and not kw:
            return self.copy()
        dst = np.dtype(dtype)
        if not a and not kw and dst in SUPPORTED and (self._dtype.kind == "i" or dst.kind == "f"):
            try:     # value-preserving / round-to-nearest casts run on the device (float -> int is numpy's)
                out = self._backend.cast_new(self._dtype, dst, self._device_ptr(), self.size)
                return DeviceArray(self._backend, out, self._shape, dst)
            except (L.BciError, AttributeError):
                pass
        out = _h(self).astype(dtype, *a, **kw)
        return out.copy() if out is self._host else out   # astype(copy=False) must not leak the cache

    def __pow__(self, other, mod=None):
        # ndarray.__pow__ has scalar fast paths (numpy/_core/src/multiarray/number.c: fast_scalar_power): `x ** 2` IS
        # np.square(x) and `x ** 0.5` IS np.sqrt(x) for float arrays, bit for bit
        if mod is None and isinstance(other, (int, float)) and not isinstance(other, bool):
            try:
                if other == 2 and _weak_scalar_ok(other, self._dtype):
                    return _unary(L.EW_SQUARE, self)
                if other == 0.5 and self._dtype.kind == "f":
                    return _unary(L.EW_SQRT, self)
            except (_Fallback, L.Unsupported, L.BciError):
                pass
        return NDArrayOperatorsMixin.__pow__(self, other)

    def clip(self, min=None, max=None, out=None, **kw):
        try:
            return _np_clip(self, min, max, out=out, **kw)
        except (_Fallback, L.Unsupported, L.BciError, TypeError):
            return _h(self).clip(min, max, out=out, **kw)

    def __copy__(self):
        return self.copy()

    def __deepcopy__(self, memo):
        return self.copy()

    def __reduce__(self):
        # pickles (multiprocessing, joblib, caches) as the ndarray it stands for
        return (np.asarray, (np.array(_h(self)),))

    def copy(self, order="C"):
        out = self._backend.copy_buffer(self._device_ptr(), self.nbytes)
        return DeviceArray(self._backend, out, self._shape, self._dtype)

    def tofile(self, fid, sep="", format="%s"):
        """ndarray.tofile: raw bytes straight from HBM into the file (pinned path in the broker)."""
        if sep == "" and isinstance(fid, (str, bytes)) or hasattr(fid, "__fspath__"):
            try:
                _export(self, fid, b"")
                return
            except (L.BciError, AttributeError, OSError):
                pass
        _h(self).tofile(fid, sep, format)

    def reshape(self, *shape, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return self._expose_host().reshape(*shape, **kw)   # a view of the host copy

    def ravel(self, *a, **kw): return self if self.ndim == 1 and not a and not kw else self._expose_host().ravel(*a, **kw)
    def flatten(self, *a, **kw): return _h(self).flatten(*a, **kw)
    def tolist(self): return _h(self).tolist()
    def item(self, *a): return _h(self).item(*a)

    def sort(self, axis=-1, kind=None, order=None):
        """In-place like ndarray.sort."""
        if order is not None or self.ndim != 1 or axis not in (-1, 0):
            raise _Fallback("sort kwargs")
        self._before_mutation()
        self._backend.sort_new(self._dtype, self._device_ptr(), self.size, inplace=True)
        self._host = None

    def __getitem__(self, key):
        if self._host is None and self.ndim == 1 and self.size:
            # x[i] of an array that lives only in HBM: fetch that element, not the array (`print(s[0], s[-1])`
            # after a sort of 10^7 rows moves 8 bytes, not 40 MB).  A scalar is a copy in numpy too; slices
            # are VIEWS there (writes go through), so they take the host-copy path below.
            n = self._shape[0]
            try:
                if isinstance(key, (int, np.integer)) and not isinstance(key, (bool, np.bool_)) and -n <= key < n:
                    return self._backend.download(self._device_ptr(), 1, self._dtype, int(key) % n)[0]
            except (L.BciError, TypeError):
                pass
        host = self._host_view()
        out = host[key]
        if isinstance(out, np.ndarray) and out.size and np.shares_memory(out, host):
            return self._expose_host()[key]     # a writable VIEW of the host copy leaves the proxy
        return out

    def __setitem__(self, key, value):
        self._before_mutation()
        host = _h(self)
        if not host.flags.writeable:
            host = host.copy()
        host[key] = _h(value) if isinstance(value, DeviceArray) else value
        self._host = host
        if self._ptr is not None:  # device copy is stale now
            self._backend.free(self._ptr)
            self._ptr = None
        self._deferred = None

    def __iter__(self): return iter(_h(self) if self.ndim <= 1 else self._expose_host())   # rows are views
    def __repr__(self): return repr(_h(self))
    def __str__(self): return str(_h(self))
    def __format__(self, spec): return format(_h(self), spec)
    def __float__(self): return float(_h(self))
    def __int__(self): return int(_h(self))
    def __bool__(self): return bool(_h(self))
    def __index__(self): return operator.index(_h(self))
    def __contains__(self, item): return item in _h(self)

    def __getattr__(self, name):
        # anything else an ndarray has (std, var, round, clip, ...): numpy on the host copy
        if name.startswith("_"):
            raise AttributeError(name)
        if name in _PURE_NDARRAY_ATTRS:
            return getattr(self._host_view(), name)
        # unknown ndarray attribute: it may mutate (fill, put, partition, resize, ...) or return a
        # view (squeeze, transpose, view, ...): the host copy becomes the single source of truth
        return getattr(self._expose_host(), name)


# ndarray attributes that neither mutate the array nor return a view of it
_PURE_NDARRAY_ATTRS = frozenset({
    "round", "clip", "any", "all", "prod", "nonzero", "searchsorted", "tobytes", "tofile", "dump", "dumps", "conj",
    "conjugate", "cumprod", "choose", "compress", "repeat", "take", "trace", "ptp", "strides", "flags"})


def _h(x):
    """Host ndarray of `x` for internal read-only use (never invalidates a device copy)."""
    return x._host_view() if isinstance(x, DeviceArray) else np.asarray(x)


_UNARY_NP = {L.EW_SQUARE: np.square, L.EW_ABS: np.absolute, L.EW_NEG: np.negative, L.EW_SQRT: np.sqrt,
             L.EW_EXP: np.exp, L.EW_LOG: np.log, L.EW_SIN: np.sin, L.EW_COS: np.cos, L.EW_TANH: np.tanh,
             L.EW_FLOOR: np.floor, L.EW_CEIL: np.ceil, L.EW_LOG1P: np.log1p, L.EW_EXPM1: np.expm1}
_BINARY_NP = {L.EW_ADD: np.add, L.EW_SUB: np.subtract, L.EW_MUL: np.multiply, L.EW_DIV: np.divide,
              L.EW_MAXIMUM: np.maximum, L.EW_MINIMUM: np.minimum,
              L.EW_RSUB: lambda a, s: np.subtract(s, a), L.EW_RDIV: lambda a, s: np.divide(s, a)}


def _to_host(obj):
    if isinstance(obj, DeviceArray):
        return obj._host_view()
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_host(o) for o in obj)
    if isinstance(obj, dict):
        return {k: _to_host(v) for k, v in obj.items()}
    return obj


def _plain_reduce_kwargs(kwargs) -> bool:
    """True when a ufunc.reduce / np.sum style call asks for the full reduction only."""
    for k, v in kwargs.items():
        if k == "axis":
            if v is not None and v != 0:
                return False
        elif k in ("dtype", "out", "where", "initial"):
            if v is not None and v is not np._NoValue and v is not True:
                return False
        elif k == "keepdims":
            if v not in (False, np._NoValue):
                return False
        else:
            return False
    return True


# ---- elementwise -----------------------------------------------------------------------------
def _unary(op, x):
    if not isinstance(x, DeviceArray):
        raise _Fallback("operand")
    if x.dtype not in SUPPORTED or (op in FLOAT_ONLY and x.dtype.kind != "f"):
        raise _Fallback("dtype")
    return DeviceArray(x._backend, None, x.shape, x.dtype, deferred=("unary", op, x))


def _binary(op, a, b):
    if isinstance(a, DeviceArray) and isinstance(b, DeviceArray):
        if a.dtype != b.dtype or a.shape != b.shape or a.dtype not in SUPPORTED or a._backend is not b._backend:
            raise _Fallback("operands")
        if op in FLOAT_ONLY and a.dtype.kind != "f":
            raise _Fallback("dtype")
        return DeviceArray(a._backend, None, a.shape, a.dtype, deferred=("binary", op, a, b, None))
    if isinstance(a, DeviceArray):
        arr, scalar, left = a, b, False
    elif isinstance(b, DeviceArray):
        arr, scalar, left = b, a, True
    else:
        raise _Fallback("operands")
    if arr.dtype not in SUPPORTED or not _weak_scalar_ok(scalar, arr.dtype):
        raise _Fallback("scalar")
    if left and op not in COMMUTATIVE:
        op = REVERSED[op]
    if op in FLOAT_ONLY and arr.dtype.kind != "f":
        raise _Fallback("dtype")
    value = float(scalar) if arr.dtype.kind == "f" else int(scalar)
    return DeviceArray(arr._backend, None, arr.shape, arr.dtype, deferred=("binary", op, arr, None, value))


# ---- reductions --------------------------------------------------------------------------------
def _result_scalar(value):
    return value.dtype.type(value) if isinstance(value, np.generic) else value


def _reduce(red_op, x):
    if not isinstance(x, DeviceArray) or x.dtype not in SUPPORTED:
        raise _Fallback("operand")
    be, n = x._backend, x.size
    if n == 0:
        raise _Fallback("empty")
    d = x._deferred
    if d is not None:
        # numpy.random.rand(n) that was never read: generate it in registers inside the reduction
        if d[0] == "rand":
            return _result_scalar(be.reduce_rand(red_op, n, d[1]))
        if d[0] == "unary" and d[1] in PRE_OF and d[2]._deferred is not None and d[2]._deferred[0] == "rand" \
                and d[2]._ptr is None:
            return _result_scalar(be.reduce_rand(red_op, n, d[2]._deferred[1], pre=PRE_OF[d[1]]))
        # fold the single-use producer into the reduction's load
        if d[0] == "unary" and d[1] in PRE_OF:
            pre, inner = PRE_OF[d[1]], d[2]
            di = inner._deferred
            if di is not None and di[0] == "binary" and di[1] in FUSABLE_PRODUCERS and _leaf(di[2]) and _leaf_or_none(di[3]):
                return be.reduce_fused(red_op, di[1], x.dtype, [di[2]._device_ptr()],
                                       [di[3]._device_ptr()] if di[3] is not None else None,
                                       [di[4]] if di[3] is None else None, [n], pre=pre)[0]
            return be.reduce(red_op, x.dtype, [inner._device_ptr()], [n], pre=pre)[0]
        if d[0] == "binary" and d[1] in FUSABLE_PRODUCERS:
            _, op, a, b, scalar = d
            if _leaf(a) and _leaf_or_none(b):
                return be.reduce_fused(red_op, op, x.dtype, [a._device_ptr()],
                                       [b._device_ptr()] if b is not None else None,
                                       [scalar] if b is None else None, [n])[0]
            # (u (op) v) + scalar  with an array-array producer underneath: mean(x*y+1.0)
            da = a._deferred
            if (op == L.EW_ADD and b is None and da is not None and da[0] == "binary" and da[3] is not None
                    and da[1] in FUSABLE_PRODUCERS and _leaf(da[2]) and _leaf(da[3])):
                return be.reduce_fused(red_op, da[1], x.dtype, [da[2]._device_ptr()], [da[3]._device_ptr()],
                                       None, [n], addends=[scalar])[0]
    return be.reduce(red_op, x.dtype, [x._device_ptr()], [n])[0]


def _leaf(a) -> bool:
    return isinstance(a, DeviceArray) and a._deferred is None


def _leaf_or_none(a) -> bool:
    return a is None or _leaf(a)


def _method_reduce(np_func, red_op, x, axis, dtype, out, kw):
    plain = (axis is None or (axis in (0, -1) and x.ndim == 1)) and dtype is None and out is None and \
        all(v in (False, np._NoValue, None, True) and k in ("keepdims", "where", "initial") and
            not (k == "keepdims" and v is True) and not (k == "initial" and v not in (None, np._NoValue))
            for k, v in kw.items())
    if plain:
        try:
            return _reduce(red_op, x)
        except (_Fallback, L.Unsupported, L.BciError):
            pass
    kw = {k: v for k, v in kw.items() if v is not np._NoValue}
    if dtype is not None:
        kw["dtype"] = dtype
    return np_func(_h(x), axis=axis, out=out, **kw)


def _np_sum(a, axis=None, dtype=None, out=None, **kw): return _method_reduce(np.sum, L.RED_SUM, _need_dev(a), axis, dtype, out, kw)
def _np_mean(a, axis=None, dtype=None, out=None, **kw): return _method_reduce(np.mean, L.RED_MEAN, _need_dev(a), axis, dtype, out, kw)
def _np_max(a, axis=None, out=None, **kw): return _method_reduce(np.max, L.RED_MAX, _need_dev(a), axis, None, out, kw)
def _np_min(a, axis=None, out=None, **kw): return _method_reduce(np.min, L.RED_MIN, _need_dev(a), axis, None, out, kw)


def _need_dev(a):
    if not isinstance(a, DeviceArray):
        raise _Fallback("operand")
    return a


def _arg(op, a, axis=None, out=None, **kw):
    a = _need_dev(a)
    if out is not None or kw.get("keepdims") or not (axis is None or (a.ndim == 1 and axis in (0, -1))):
        raise _Fallback("kwargs")
    if a.dtype not in SUPPORTED or a.size == 0:
        raise _Fallback("dtype")
    return np.int64(a._backend.argext(op, a.dtype, [a._device_ptr()], [a.size])[0])


def argmax(a, axis=None, out=None, **kw): return _arg(L.ARGMAX, a, axis, out, **kw)
def argmin(a, axis=None, out=None, **kw): return _arg(L.ARGMIN, a, axis, out, **kw)


def cumsum(a, axis=None, dtype=None, out=None):
    a = _need_dev(a)
    if dtype is not None or out is not None or a.ndim != 1 or axis not in (None, 0, -1) or a.dtype not in SUPPORTED:
        raise _Fallback("kwargs")
    be = a._backend
    odt = np.dtype(np.int64) if a.dtype == np.int32 else a.dtype
    return DeviceArray(be, be.cumsum_new(a.dtype, a._device_ptr(), a.size), a.shape, odt)


def sort(a, axis=-1, kind=None, order=None, **kw):
    a = _need_dev(a)
    if order is not None or kw or a.ndim != 1 or axis not in (-1, 0, None) or a.dtype not in SUPPORTED:
        raise _Fallback("kwargs")
    be = a._backend
    return DeviceArray(be, be.sort_new(a.dtype, a._device_ptr(), a.size), a.shape, a.dtype)


def argsort(a, axis=-1, kind=None, order=None, **kw):
    a = _need_dev(a)
    # the default kind (quicksort) is unstable in numpy: any valid permutation is acceptable,
    # the device always returns the stable one
    if order is not None or kw or a.ndim != 1 or axis not in (-1, 0, None) or a.dtype not in SUPPORTED:
        raise _Fallback("kwargs")
    be = a._backend
    return DeviceArray(be, be.argsort_new(a.dtype, a._device_ptr(), a.size), a.shape, np.int64)


# ---- order statistics: one device sort, then two or three elements cross the bus -------------------------------
# numpy partitions the array around the indices it needs (introselect) and interpolates between at most two elements.
# The same elements come out of the sorted array; the interpolation is done by numpy itself on a two-element stand-in
# (np.mean / np.quantile), so type rules and rounding are numpy's.
def _sorted_elem(s, i: int):
    return s._backend.download(s._device_ptr(), 1, s.dtype, int(i))[0]


def _order_stat_operand(a, axis, out, kw):
    a = _need_dev(a)
    if out is not None or any(kw.get(k) for k in ("keepdims", "overwrite_input")) or set(kw) - {"keepdims", "overwrite_input"}:
        raise _Fallback("kwargs")
    if a.ndim != 1 or axis not in (None, 0, -1) or a.dtype not in SUPPORTED or a.size < 2:
        raise _Fallback("shape")
    return a


def _np_median(a, axis=None, out=None, **kw):
    a = _order_stat_operand(a, axis, out, kw)
    n = a.size
    s = sort(a)
    if a.dtype.kind == "f" and np.isnan(_sorted_elem(s, n - 1)):      # NaNs sort last; numpy's median is NaN then
        return np.median(np.array([np.nan, np.nan], dtype=a.dtype))
    mid = np.array([_sorted_elem(s, (n - 1) // 2), _sorted_elem(s, n // 2)], dtype=a.dtype)
    return np.mean(mid[:1] if n % 2 else mid)       # numpy: mean of the middle element(s)


def _quantile_linear(a, q):
    """numpy/lib/_function_base_impl.py:_quantile for method="linear" on a 1-D array, restated step by step with the
    same numpy operations on the same dtypes; the partition is replaced by the device sort, and only the elements the
    interpolation reads (two per quantile, and the last one for the NaN check) are fetched."""
    n = a.size
    if q.ndim > 1 or q.size > 64 or q.dtype.kind not in "fiu":
        raise _Fallback("q")
    if q.size == 0 or not (q.min() >= 0 and q.max() <= 1):
        raise _Fallback("q range")                           # numpy raises / handles it
    s = sort(a)
    vi = np.asanyarray((n - 1) * q)                          # 'linear': virtual index = (n - 1) * q, in q's dtype
    integral = np.issubdtype(vi.dtype, np.integer)
    if integral:
        prev_i = next_i = vi.astype(np.intp)
    else:
        prev = np.asanyarray(np.floor(vi))
        nxt = np.asanyarray(prev + 1)
        above = vi >= n - 1
        if above.any():
            prev[above] = -1
            nxt[above] = -1
        prev_i, next_i = prev.astype(np.intp), nxt.astype(np.intp)
    # the elements numpy would read after its partition, in a small host array indexed the same way
    wanted = np.unique(np.concatenate(([n - 1], prev_i.ravel() % n, next_i.ravel() % n)))
    small = np.array([_sorted_elem(s, i) for i in wanted], dtype=a.dtype)
    at = lambda idx: small[np.searchsorted(wanted, idx % n)]    # noqa: E731  (0-d index -> numpy scalar, like arr[idx])
    last = small[-1]
    if integral:
        result = at(prev_i)                                  # no interpolation: the points themselves
    else:
        previous, nxt_v = at(prev_i), at(next_i)
        gamma = np.asanyarray(np.asanyarray(vi - prev_i), dtype=vi.dtype).reshape(vi.shape)
        # _lerp(previous, next, gamma)
        diff_b_a = np.subtract(nxt_v, previous)
        result = np.asanyarray(np.add(previous, diff_b_a * gamma))
        np.subtract(nxt_v, diff_b_a * (1 - gamma), out=result, where=gamma >= 0.5, casting="unsafe", dtype=type(result.dtype))
        if result.ndim == 0:
            result = result[()]
    if a.dtype.kind == "f" and np.isnan(last):               # NaNs sort last: every quantile is NaN then
        if np.ndim(result) == 0:
            return last
        np.copyto(result, last)
    return result


def _np_quantile(a, q, axis=None, out=None, method="linear", **kw):
    weights = kw.pop("weights", None)
    a = _order_stat_operand(a, axis, out, kw)
    if method != "linear" or weights is not None or isinstance(q, DeviceArray):
        raise _Fallback("method")
    # numpy.quantile: a python number takes the array's float type, anything else keeps its own
    qa = np.asanyarray(q, dtype=a.dtype) if isinstance(q, (int, float)) and a.dtype.kind == "f" else np.asanyarray(q)
    return _quantile_linear(a, qa)


def _np_percentile(a, q, axis=None, out=None, method="linear", **kw):
    weights = kw.pop("weights", None)
    a = _order_stat_operand(a, axis, out, kw)
    if method != "linear" or weights is not None or isinstance(q, DeviceArray):
        raise _Fallback("method")
    # numpy.percentile: the divisor takes the array's float type
    qa = np.true_divide(q, a.dtype.type(100) if a.dtype.kind == "f" else 100, out=...)
    return _quantile_linear(a, qa)


def _is_two(v) -> bool:
    return isinstance(v, (int, float, np.integer, np.floating)) and not isinstance(v, (bool, np.bool_)) and v == 2


def _np_clip(a, a_min=None, a_max=None, out=None, **kw):
    """np.clip = minimum(maximum(a, lo), hi) (numpy/_core/_methods.py:_clip), each bound a python scalar."""
    a = _need_dev(a)
    lo = kw.pop("min", a_min) if a_min is None else a_min
    hi = kw.pop("max", a_max) if a_max is None else a_max
    if out is not None or kw or (lo is None and hi is None) or a.dtype not in SUPPORTED:
        raise _Fallback("kwargs")
    for b in (lo, hi):
        if b is not None and (isinstance(b, DeviceArray) or not _weak_scalar_ok(b, a.dtype) or b != b):
            raise _Fallback("bound")
    r = a
    if lo is not None:
        r = _binary(L.EW_MAXIMUM, r, lo)
    if hi is not None:
        r = _binary(L.EW_MINIMUM, r, hi)
    return r


def _np_ptp(a, axis=None, out=None, **kw):
    a = _need_dev(a)
    if out is not None or any(v not in (False, np._NoValue) for v in kw.values()) or not (axis is None or (a.ndim == 1 and axis in (0, -1))):
        raise _Fallback("kwargs")
    with np.errstate(over="ignore"):
        return np.subtract(_reduce(L.RED_MAX, a), _reduce(L.RED_MIN, a))      # numpy: umr_maximum(a) - umr_minimum(a)


def _np_average(a, axis=None, weights=None, returned=False, **kw):
    a = _need_dev(a)
    if weights is not None or returned or any(v not in (False, np._NoValue) for v in kw.values()):
        raise _Fallback("weights")
    return _method_reduce(np.mean, L.RED_MEAN, a, axis, None, None, {})          # numpy: a.mean(axis)


def _np_norm(x, ord=None, axis=None, keepdims=False):
    """np.linalg.norm of a real 1-D float array: sqrt(x . x) (numpy/linalg/_linalg.py:norm, the ord=None branch)."""
    x = _need_dev(x)
    if axis is not None or keepdims or ord is not None or x.ndim != 1 or x.dtype.kind != "f":
        raise _Fallback("norm kwargs")      # (ord=2 takes numpy's general path: another operation order)
    return np.sqrt(_dot(x, x))


def _nan_free(a) -> bool:
    """Floats: np.max propagates NaN, so one reduction tells whether the array holds any.  Integers never do."""
    return a.dtype.kind != "f" or not np.isnan(_reduce(L.RED_MAX, a))


def _nan_reduce(red_op):
    def handler(a, axis=None, dtype=None, out=None, **kw):
        a = _need_dev(a)
        if dtype is not None or out is not None or any(v not in (False, np._NoValue, None) for v in kw.values()) \
                or not (axis is None or (a.ndim == 1 and axis in (0, -1))) or a.size == 0:
            raise _Fallback("kwargs")
        if not _nan_free(a):
            raise _Fallback("NaNs present: numpy's masked path")
        return _reduce(red_op, a)          # without NaNs nansum / nanmean / nanmax / nanmin are the plain reductions
    return handler


def _nan_free_then(handler):
    """nanmedian / nanvar / nanpercentile ... of NaN-free data are the plain functions (numpy removes the NaNs of a
    1-D array and calls them); with NaNs present numpy's own path runs on the host."""
    def wrapped(a, *args, **kw):
        a = _need_dev(a)
        if a.ndim != 1 or a.size == 0 or not _nan_free(a):
            raise _Fallback("NaNs present")
        return handler(a, *args, **kw)
    return wrapped


def _np_inner(a, b):
    if isinstance(a, DeviceArray) and isinstance(b, DeviceArray) and a.ndim == 1 and b.ndim == 1:
        return _dot(a, b)          # 1-D real inner / vdot = dot
    raise _Fallback("inner")


def matmul(a, b, out=None, **kw):
    if out is not None or kw:
        raise _Fallback("kwargs")
    if not (isinstance(a, DeviceArray) and isinstance(b, DeviceArray)):
        raise _Fallback("operands")
    if a.dtype != np.float32 or b.dtype != np.float32 or a.ndim != 2 or b.ndim != 2 or a.shape[1] != b.shape[0]:
        raise _Fallback("shape")
    be = a._backend
    M, K = a.shape
    N = b.shape[1]
    return DeviceArray(be, be.matmul_new(a._device_ptr(), b._device_ptr(), M, N, K), (M, N), np.float32)


def _dot(a, b, out=None):
    if isinstance(a, DeviceArray) and isinstance(b, DeviceArray) and a.ndim == 2 and b.ndim == 2:
        return matmul(a, b, out=out)
    # 1-D . 1-D = sum(a * b): one fused kernel, the product never reaches HBM
    if (out is None and isinstance(a, DeviceArray) and isinstance(b, DeviceArray) and a.ndim == 1 and b.ndim == 1
            and a.shape == b.shape and a.dtype == b.dtype and a.dtype in (np.float32, np.float64, np.int64) and a.size
            and a._backend is b._backend):
        return _result_scalar(a._backend.reduce_fused(L.RED_SUM, L.EW_MUL, a.dtype, [a._device_ptr()], [b._device_ptr()],
                                                      None, [a.size])[0])
    raise _Fallback("dot")


def _var(a, axis=None, dtype=None, out=None, ddof=0, **kw):
    """numpy.var for a float array, full reduction: mean(|a - mean(a)|^2) * n / (n - ddof) — numpy's
    own two-pass formula (numpy/_core/_methods.py:_var), each pass one kernel; `a - mean` and its
    square are never materialised."""
    a = _need_dev(a)
    plain = (axis is None or (a.ndim == 1 and axis in (0, -1))) and dtype is None and out is None and \
        all(v in (False, np._NoValue, None) or (k == "where" and v is True) for k, v in kw.items()) and \
        all(k in ("keepdims", "where", "mean", "correction") for k in kw)
    if not plain or a.dtype.kind != "f" or a.dtype not in SUPPORTED or a.size == 0:
        raise _Fallback("var kwargs")
    if isinstance(ddof, bool) or not isinstance(ddof, (int, float)) or a.size - ddof <= 0:
        raise _Fallback("ddof")
    be, n = a._backend, a.size
    ptr = a._device_ptr()                                    # read twice: materialise a deferred producer once
    mean = be.reduce(L.RED_MEAN, a.dtype, [ptr], [n])[0]
    msq = be.reduce_fused(L.RED_MEAN, L.EW_SUB, a.dtype, [ptr], None, [float(mean)], [n], pre=L.PRE_SQUARE)[0]
    return a.dtype.type(a.dtype.type(msq) * a.dtype.type(n) / a.dtype.type(n - ddof)) if ddof else a.dtype.type(msq)


def _std(a, axis=None, dtype=None, out=None, ddof=0, **kw):
    return np.sqrt(_var(a, axis=axis, dtype=dtype, out=out, ddof=ddof, **kw))


def _export(a: DeviceArray, path, header: bytes) -> None:
    """`a`'s bytes -> the file `path` behind `header`; the broker hashes header||array on the device while the
    bytes leave through the pinned path, so the object store knows the file's id without reading it back."""
    import os

    fd = os.open(os.fspath(path), os.O_RDWR | os.O_CREAT | os.O_TRUNC | getattr(os, "O_CLOEXEC", 0), 0o666)
    try:
        a._backend.export(a._device_ptr(), a.nbytes, fd, header)
    finally:
        os.close(fd)


def _np_save(file, arr, allow_pickle=True, fix_imports=np._NoValue):
    """numpy.save of a device array to a path: the .npy header is built here, the data never visits this process."""
    import io
    import os

    arr = _need_dev(arr)
    if not (isinstance(file, (str, bytes)) or hasattr(file, "__fspath__")):
        raise _Fallback("file object")
    path = os.fspath(file)
    if isinstance(path, bytes):
        path = os.fsdecode(path)
    if not path.endswith(".npy"):
        path += ".npy"
    head = io.BytesIO()
    np.lib.format.write_array_header_1_0(head, {"descr": np.lib.format.dtype_to_descr(arr.dtype), "fortran_order": False,
                                                "shape": arr.shape})
    try:
        _export(arr, path, head.getvalue())
    except (AttributeError, OSError) as e:
        raise _Fallback(str(e)) from None


_FUNCTIONS = {
    np.save: _np_save,
    np.sum: _np_sum, np.mean: _np_mean, np.max: _np_max, np.min: _np_min, np.amax: _np_max, np.amin: _np_min,
    np.argmax: argmax, np.argmin: argmin, np.cumsum: cumsum, np.sort: sort, np.argsort: argsort,
    np.matmul: matmul, np.dot: _dot, np.var: _var, np.std: _std,
    np.median: _np_median, np.quantile: _np_quantile, np.percentile: _np_percentile,
    np.clip: _np_clip, np.ptp: _np_ptp, np.average: _np_average, np.linalg.norm: _np_norm,
    np.nansum: _nan_reduce(L.RED_SUM), np.nanmean: _nan_reduce(L.RED_MEAN),
    np.nanmax: _nan_reduce(L.RED_MAX), np.nanmin: _nan_reduce(L.RED_MIN),
    np.nanmedian: _nan_free_then(_np_median), np.nanquantile: _nan_free_then(_np_quantile),
    np.nanpercentile: _nan_free_then(_np_percentile), np.nanvar: _nan_free_then(_var), np.nanstd: _nan_free_then(_std),
    np.inner: _np_inner, np.vdot: _np_inner,
}


# ---- entry points used by the import hook (numpy_hook.py) ---------------------------------------
OFFLOAD_MIN_ELEMS = 1 << 18   # host arrays smaller than this stay with numpy


def device_rand(backend, shape, seed: int) -> DeviceArray:
    """numpy.random.rand / random_sample for an unseeded legacy generator.  Nothing is generated yet:
    a reduction over the result (`numpy.sum(numpy.square(x))`, examples/benchmark-numpy.py:21) draws
    the stream in registers; any other use materialises it (same seed, same values)."""
    if not hasattr(backend, "reduce_rand"):
        n = 1
        for s in shape:
            n *= int(s)
        return DeviceArray(backend, backend.rand(n, seed), shape, np.float64)
    return DeviceArray(backend, None, shape, np.float64, deferred=("rand", int(seed) & 0xFFFFFFFFFFFFFFFF))


def device_normal(backend, shape, seed: int, loc: float = 0.0, scale: float = 1.0) -> DeviceArray:
    """numpy.random.normal / randn / standard_normal for an unseeded legacy generator."""
    n = 1
    for s in shape:
        n *= int(s)
    return DeviceArray(backend, backend.rand_normal(n, seed, float(loc), float(scale)), shape, np.float64)


def device_rand_legacy(backend, shape, random_module) -> DeviceArray:
    """numpy.random.rand / random_sample on the SEEDED global generator: the MT19937 state goes
    to the device, the same stream comes back bit for bit, and the advanced state is put back so
    that host draws before and after this call stay on numpy's sequence."""
    n = 1
    for s in shape:
        n *= int(s)
    name, key, pos, has_gauss, cached = random_module.get_state()
    if name != "MT19937":
        raise _Fallback("generator")
    state = np.empty(625, dtype=np.uint32)
    state[:624], state[624] = key, pos
    ptr, after = backend.rand_mt19937(n, state)
    random_module.set_state((name, np.asarray(after[:624], dtype=np.uint32), int(after[624]), has_gauss, cached))
    return DeviceArray(backend, ptr, shape, np.float64)


def offload(backend, arr):
    """Host ndarray -> DeviceArray when it is big enough to pay for the PCIe hop."""
    if isinstance(arr, DeviceArray):
        return arr
    if (isinstance(arr, np.ndarray) and arr.dtype in SUPPORTED and arr.size >= OFFLOAD_MIN_ELEMS
            and arr.flags["C_CONTIGUOUS"]):
        return DeviceArray.from_host(backend, arr)
    raise _Fallback("host array")


# ---- numpy.random.Generator(PCG64) on the device -----------------------------------------------------------
_M64 = (1 << 64) - 1
_PCG_MULT = (2549297995355413924 << 64) | 4865540595714422341


def _pcg_output(state: int) -> int:
    """XSL-RR output of a 128-bit PCG64 state (numpy/random/src/pcg64/pcg64.h)."""
    hi, lo = state >> 64, state & _M64
    x, rot = hi ^ lo, state >> 122
    return ((x >> rot) | (x << ((64 - rot) & 63))) & _M64


def device_pcg(backend, generator, shape, kind: int, low: int = 0, span: int = 0) -> DeviceArray:
    """`n` values of `generator` (a numpy Generator over PCG64) produced on the device, bit for bit what
    numpy would return; the host generator is advanced past them, so draws before and after stay on
    numpy's stream.  Raises _Fallback (host generator untouched) when the stream cannot be followed."""
    bg = generator.bit_generator
    st = bg.state
    if st.get("bit_generator") != "PCG64":
        raise _Fallback("bit generator")
    n = 1
    for s in shape:
        n *= int(s)
    wide = kind == L.PCG_F64
    if not wide and st["has_uint32"]:
        raise _Fallback("a buffered 32-bit word is pending")       # rare: an odd number of 32-bit draws came before
    state, inc = int(st["state"]["state"]), int(st["state"]["inc"])
    ptr, rejected = backend.rand_pcg(kind, n, state, inc, low, span)
    if rejected:
        backend.free(ptr)
        raise _Fallback("Lemire rejection in the stream")
    draws = n if wide else (n + 1) // 2
    bg.advance(draws)           # also clears numpy's buffered 32-bit word: put back what numpy would hold
    after = bg.state
    if wide:                    # 64-bit draws never touch the buffered word
        after["has_uint32"], after["uinteger"] = st["has_uint32"], st["uinteger"]
    else:                       # the high half of the last draw: pending when n is odd, spent (but still stored) otherwise
        after["has_uint32"] = n & 1
        after["uinteger"] = _pcg_output(int(after["state"]["state"])) >> 32 if n else st["uinteger"]
    bg.state = after
    dtype = {L.PCG_F64: np.float64, L.PCG_F32: np.float32, L.PCG_U32: np.uint32, L.PCG_BOUNDED32_I64: np.int64,
             L.PCG_BOUNDED32_I32: np.int32}[kind]
    return DeviceArray(backend, ptr, shape, dtype)


# ---- array creation ---------------------------------------------------------------------------------------
def device_full(backend, shape, dtype, value) -> DeviceArray:
    dtype = np.dtype(dtype)
    if dtype not in SUPPORTED:
        raise _Fallback("dtype")
    n = 1
    for s in shape:
        n *= int(s)
    v = dtype.type(value)            # numpy casts the fill value to the array dtype
    return DeviceArray(backend, backend.fill_new(L.FILL_CONST, dtype, n, v.item()), shape, dtype)


def device_arange(backend, start, stop, step, dtype) -> DeviceArray:
    """numpy.arange for python scalars: length = ceil((stop - start) / step); element i = start + i*delta with
    delta = dtype(start + step) - dtype(start), like numpy's @TYPE@_fill."""
    import math

    dtype = np.dtype(dtype)
    if dtype not in SUPPORTED:
        raise _Fallback("dtype")
    n = int(math.ceil((stop - start) / step))
    if n < OFFLOAD_MIN_ELEMS:
        raise _Fallback("small")
    first, second = dtype.type(start), dtype.type(start + step)
    with np.errstate(all="ignore"):
        delta = dtype.type(second - first)
    return DeviceArray(backend, backend.fill_new(L.FILL_ARANGE, dtype, n, first.item(), delta.item(), second.item()), (n,), dtype)


def device_linspace(backend, start: float, stop: float, num: int, endpoint: bool) -> DeviceArray:
    """numpy.linspace (float64, scalars): y = arange(num) * step + start, y[-1] = stop when endpoint."""
    div = (num - 1) if endpoint else num
    delta = float(stop) - float(start)
    step = delta / div
    if step == 0 or num < OFFLOAD_MIN_ELEMS:
        raise _Fallback("degenerate step")
    ptr = backend.fill_new(L.FILL_LINSPACE, np.float64, num, float(start), step, float(stop) if endpoint else None)
    return DeviceArray(backend, ptr, (num,), np.float64)


def device_load(backend, path) -> DeviceArray:
    """numpy.load of a plain .npy file: the header is parsed here, the data goes file -> pinned buffer -> HBM
    inside the broker (the descriptor is passed, this process never reads the array)."""
    import os

    with open(path, "rb") as f:
        version = np.lib.format.read_magic(f)
        if version == (1, 0):
            shape, fortran, dtype = np.lib.format.read_array_header_1_0(f)
        elif version == (2, 0):
            shape, fortran, dtype = np.lib.format.read_array_header_2_0(f)
        else:
            raise _Fallback("npy version")
        offset = f.tell()
        n = 1
        for s in shape:
            n *= int(s)
        if fortran and len(shape) > 1 or dtype not in SUPPORTED or n < OFFLOAD_MIN_ELEMS:
            raise _Fallback("layout")
        if os.fstat(f.fileno()).st_size - offset < n * dtype.itemsize:
            raise _Fallback("truncated file")
        ptr = backend.upload_fd(f.fileno(), offset, n * dtype.itemsize)
    return DeviceArray(backend, ptr, shape, dtype)
