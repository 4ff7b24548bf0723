"""In-process backend for numpy_proxy: the same "new buffer out" call surface as
broker_client.RemoteBackend, implemented directly on the C-ABI binding (device.py).
Used by the service process itself, the smoke test and the GPU parity tests."""
from __future__ import annotations

import numpy as np

from . import _lib as L


class LocalBackend:
    def __init__(self, device_module=None):
        if device_module is None:
            from . import device as device_module
            if not device_module.is_initialised():
                device_module.init()
        self.dev = device_module

    def alloc(self, nbytes): return self.dev.alloc(nbytes)
    def free(self, ptr): self.dev.free(ptr)
    def upload(self, arr): return self.dev.upload(arr)
    def download(self, ptr, n, dtype, offset=0): return self.dev.download(ptr, n, dtype, offset)

    def _new(self, nbytes, fill):
        out = self.dev.alloc(nbytes)
        try:
            fill(out)
        except Exception:
            self.dev.free(out)
            raise
        return out

    def copy_buffer(self, src, nbytes):
        return self._new(nbytes, lambda out: self.dev.d2d(out, src, nbytes))

    def rand(self, n, seed):
        return self._new(n * 8, lambda out: self.dev.rand_f64([out], [n], [seed]))

    def rand_pcg(self, kind, n, state, inc, low=0, span=0):
        """-> (buffer, draws numpy would have rejected)."""
        box = {}
        odt = self.dev.PCG_OUT_DTYPE[kind]

        def run(out):
            box["rej"] = int(self.dev.rand_pcg64(kind, [out], [n], [(state, inc, low, span)])[0])
        return self._new(n * odt.itemsize, run), box["rej"]

    def reduce_rand(self, op, n, seed, pre=L.PRE_NONE):
        return self.dev.reduce_rand_f64(op, [n], [seed], pre=pre)[0]

    def fill_new(self, kind, dtype, n, a, b=None, c=None):
        dtype = np.dtype(dtype)
        return self._new(n * dtype.itemsize, lambda out: self.dev.fill(
            kind, dtype, [out], [n], [a], [b] if b is not None else None, [c] if c is not None else None))

    def cast_new(self, src_dtype, dst_dtype, ptr, n):
        return self._new(n * np.dtype(dst_dtype).itemsize,
                         lambda out: self.dev.cast(src_dtype, dst_dtype, [ptr], [out], [n]))

    def upload_fd(self, fd, offset, nbytes):
        return self._new(nbytes, lambda out: self.dev.h2d_fd(out, fd, offset, nbytes))

    def export(self, ptr, nbytes, fd, header=b""):
        """Array -> file behind `header`, digest of header||array computed on the device."""
        import os

        hdr = np.frombuffer(header, dtype=np.uint8)
        if not len(hdr):
            digest = self.dev.sha256_device([ptr], [nbytes])[0]
            self.dev.d2h_fd(fd, 0, ptr, nbytes)
        else:   # one contiguous message on the device: header bytes, then the array
            host = np.empty(len(hdr) + nbytes, dtype=np.uint8)
            host[: len(hdr)] = hdr
            self.dev.d2h(host[len(hdr):], ptr)
            msg = self.dev.upload(host)
            try:
                digest = self.dev.sha256_device([msg], [len(host)])[0]
                self.dev.d2h_fd(fd, 0, msg, len(host))
            finally:
                self.dev.free(msg)
        os.ftruncate(fd, len(hdr) + nbytes)
        return digest

    def rand_normal(self, n, seed, loc, scale):
        return self._new(n * 8, lambda out: self.dev.rand_normal_f64([out], [n], [seed], [loc], [scale]))

    def rand_mt19937(self, n, state):
        """-> (buffer, state after); `state` = 625 uint32 (key + pos)."""
        box = {}

        def run(out):
            box["st"] = self.dev.rand_mt19937([out], [n], np.asarray(state, dtype=np.uint32).reshape(1, 625))[0]
        return self._new(n * 8, run), box["st"]

    def ewise_new(self, op, dtype, in0, in1, scalar, n):
        dtype = np.dtype(dtype)
        return self._new(n * dtype.itemsize, lambda out: self.dev.ewise(
            op, dtype, [in0], [in1] if in1 is not None else None,
            [scalar] if (in1 is None and scalar is not None) else None, [out], [n]))      # unary ops have neither

    def reduce(self, op, dtype, in_ptrs, ns, pre=L.PRE_NONE):
        return self.dev.reduce(op, dtype, in_ptrs, ns, pre=pre)

    def reduce_fused(self, op, ew_op, dtype, in0, in1, scalars, ns, pre=L.PRE_NONE, addends=None):
        return self.dev.reduce_fused(op, ew_op, dtype, in0, in1, scalars, ns, pre=pre, addends=addends)

    def argext(self, op, dtype, in_ptrs, ns):
        return self.dev.argext(op, dtype, in_ptrs, ns)

    def cumsum_new(self, dtype, ptr, n):
        odt = self.dev.cumsum_out_dtype(dtype)
        return self._new(n * odt.itemsize, lambda out: self.dev.cumsum(dtype, [ptr], [out], [n]))

    def sort_new(self, dtype, ptr, n, inplace=False):
        if inplace:
            self.dev.sort(dtype, [ptr], [ptr], [n])
            return ptr
        return self._new(n * np.dtype(dtype).itemsize, lambda out: self.dev.sort(dtype, [ptr], [out], [n]))

    def argsort_new(self, dtype, ptr, n):
        return self._new(n * 8, lambda out: self.dev.argsort(dtype, [ptr], [out], [n]))

    def matmul_new(self, a, b, M, N, K):
        return self._new(M * N * 4, lambda out: self.dev.matmul_f32([a], [b], [out], M, N, K))

    def sha256(self, ptr, nbytes):
        return self.dev.sha256_device([ptr], [nbytes])[0]
