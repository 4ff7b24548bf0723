"""Script-side backend: same call surface `numpy_proxy` uses on `device.py`, but every call is
a message to the GPU broker (broker.py).  The script process never loads CUDA.

One request = one connection; the broker frees everything the connection allocated when it
closes, so a crashed or killed script leaks nothing."""
from __future__ import annotations

import os
import tempfile
import threading
from multiprocessing.connection import Client

import numpy as np

from . import _lib as L
from .broker import INLINE_MAX

SOCKET_ENV = "BCI_BROKER_SOCKET"


class RemoteBackend:
    def __init__(self, socket_path: str | None = None):
        self.socket_path = socket_path or os.environ.get(SOCKET_ENV)
        if not self.socket_path:
            raise L.BciError(L.ERR_NOT_INIT, f"{SOCKET_ENV} is not set: no GPU broker to talk to")
        self._conn = None
        self._lock = threading.Lock()

    # ---- transport ---------------------------------------------------------------------
    def _connection(self):
        if self._conn is None:
            try:
                self._conn = Client(address=self.socket_path, family="AF_UNIX")
            except OSError as e:
                raise L.BciError(L.ERR_NOT_INIT, f"cannot reach the GPU broker at {self.socket_path}: {e}") from None
        return self._conn

    def _call(self, *msg):
        with self._lock:
            conn = self._connection()
            try:
                conn.send(msg)
                reply = conn.recv()
            except (EOFError, OSError) as e:
                self._conn = None
                raise L.BciError(L.ERR_CUDA, f"GPU broker connection lost: {e}") from None
        if reply[0] == "ok":
            return reply[1]
        _, code, text = reply
        if code == L.ERR_UNSUPPORTED:
            raise L.Unsupported(code, text)
        raise L.BciError(code, text)

    def _post(self, *msg):
        with self._lock:
            try:
                self._connection().send(msg)
            except (EOFError, OSError):
                self._conn = None

    def close(self):
        with self._lock:
            if self._conn is not None:
                try:
                    self._conn.close()
                finally:
                    self._conn = None

    def ping(self):
        return self._call("ping")

    # ---- device API subset used by numpy_proxy --------------------------------------------------
    def alloc(self, nbytes):
        return self._call("alloc", int(nbytes))

    def free(self, ptr):
        if ptr:
            self._post("free", ptr)

    def upload(self, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        if arr.nbytes <= INLINE_MAX:
            return self._call("upload", arr.dtype.str, arr.size, ("inline", arr.tobytes()))
        fd, path = tempfile.mkstemp(prefix="bci-up-", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            with os.fdopen(fd, "wb") as f:
                f.write(memoryview(arr).cast("B"))
            return self._call("upload", arr.dtype.str, arr.size, ("shm", path))
        finally:
            os.unlink(path)

    def download(self, ptr, n, dtype):
        dtype = np.dtype(dtype)
        nbytes = int(n) * dtype.itemsize
        if nbytes <= INLINE_MAX:
            return np.frombuffer(self._call("download", ptr, dtype.str, int(n), None), dtype=dtype).copy()
        fd, path = tempfile.mkstemp(prefix="bci-down-", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            os.ftruncate(fd, nbytes)
            os.close(fd)
            self._call("download", ptr, dtype.str, int(n), path)
            return np.fromfile(path, dtype=dtype, count=int(n))
        finally:
            os.unlink(path)

    def d2d(self, dst, src, nbytes):
        raise NotImplementedError  # proxies use copy_buffer() through the broker

    def copy_buffer(self, src, nbytes):
        return self._call("d2d", src, int(nbytes))

    def rand_f64(self, out_ptrs, ns, seeds):
        raise NotImplementedError  # allocation + fill is one message: rand()

    def rand(self, n, seed):
        return self._call("rand", int(n), int(seed) & 0xFFFFFFFFFFFFFFFF)

    def ewise_new(self, op, dtype, in0, in1, scalar, n):
        return self._call("ewise", op, np.dtype(dtype).str, in0, in1, scalar, int(n))

    def reduce(self, op, dtype, in_ptrs, ns, pre=L.PRE_NONE):
        return [self._call("reduce", op, pre, np.dtype(dtype).str, in_ptrs[0], int(ns[0]))]

    def reduce_fused(self, op, ew_op, dtype, in0, in1, scalars, ns, pre=L.PRE_NONE, addends=None):
        return [self._call("reduce_fused", op, pre, ew_op, np.dtype(dtype).str, in0[0],
                           in1[0] if in1 is not None else None, scalars[0] if scalars is not None else None,
                           addends[0] if addends is not None else None, int(ns[0]))]

    def argext(self, op, dtype, in_ptrs, ns):
        return [self._call("argext", op, np.dtype(dtype).str, in_ptrs[0], int(ns[0]))]

    def cumsum_new(self, dtype, ptr, n):
        return self._call("cumsum", np.dtype(dtype).str, ptr, int(n))

    def sort_new(self, dtype, ptr, n, inplace=False):
        return self._call("sort", np.dtype(dtype).str, ptr, int(n), bool(inplace))

    def argsort_new(self, dtype, ptr, n):
        return self._call("argsort", np.dtype(dtype).str, ptr, int(n), False)

    def matmul_new(self, a, b, M, N, K):
        return self._call("matmul", a, b, int(M), int(N), int(K))

    def sha256(self, ptr, nbytes):
        return self._call("sha256", ptr, int(nbytes))

    # ---- content hashing for the object store (storage.py) ------------------------------------
    def sha256_file(self, path: str) -> bytes:
        return self._call("sha256_file", path)

    def sha256_bytes(self, data: bytes) -> bytes:
        return self._call("sha256_bytes", data)

    def sha256_update(self, state: bytes | None, chunk: bytes) -> bytes:
        return self._call("sha256_update", state, chunk)

    def sha256_final(self, state: bytes) -> bytes:
        return self._call("sha256_final", state)
