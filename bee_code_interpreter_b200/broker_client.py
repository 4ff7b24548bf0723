"""Script-side backend: the call surface `numpy_proxy` uses, implemented as messages to the
per-GPU broker (include/bci_broker.h).  The script process never loads CUDA.

One request = one connection; the broker frees everything the connection allocated when it
closes, so a crashed or killed script leaks nothing.  Reduction results come back as the raw
8-byte result slot of bci_reduce and are decoded with the same rules as device.py."""
from __future__ import annotations

import os
import socket
import struct
import tempfile
import threading

import numpy as np

from . import _lib as L
from . import wire as W

SOCKET_ENV = "BCI_BROKER_SOCKET"
_CODE = {np.dtype(np.float32): L.F32, np.dtype(np.float64): L.F64, np.dtype(np.int32): L.I32, np.dtype(np.int64): L.I64}
_DTYPE = {v: k for k, v in _CODE.items()}


def _code(dtype) -> int:
    try:
        return _CODE[np.dtype(dtype)]
    except KeyError:
        raise L.Unsupported(L.ERR_UNSUPPORTED, f"dtype {dtype} is not handled on the device") from None


def _decode_reduce(bits: int, op: int, code: int):
    if op == L.RED_MEAN:
        res = L.F32 if code == L.F32 else L.F64
    elif op == L.RED_SUM:
        res = L.I64 if code == L.I32 else code
    else:
        res = code
    dt = _DTYPE[res]
    raw = struct.pack("<Q", bits)
    return np.frombuffer(raw[: dt.itemsize], dtype=dt)[0]


def _anon_file() -> int:
    """An anonymous file in memory (memfd; an unlinked tmpfs file where memfd_create is missing)."""
    try:
        return os.memfd_create("bci-xfer", os.MFD_CLOEXEC)
    except (AttributeError, OSError):
        fd, path = tempfile.mkstemp(prefix="bci-xfer-", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        os.unlink(path)
        return fd


class RemoteBackend:
    def __init__(self, socket_path: str | None = None):
        self.socket_path = socket_path or os.environ.get(SOCKET_ENV)
        if not self.socket_path:
            raise L.BciError(L.ERR_NOT_INIT, f"{SOCKET_ENV} is not set: no GPU broker to talk to")
        self._sock = None
        self._lock = threading.Lock()

    # ---- transport ---------------------------------------------------------------------
    def _connection(self):
        if self._sock is None:
            s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            try:
                s.connect(self.socket_path)
            except OSError as e:
                s.close()
                raise L.BciError(L.ERR_NOT_INIT, f"cannot reach the GPU broker at {self.socket_path}: {e}") from None
            self._sock = s
        return self._sock

    def _call(self, op, fd=None, **kw):
        """Returns (value, payload).  `fd`: a descriptor that rides on the request (SCM_RIGHTS)."""
        if fd is not None:
            kw["flags"] = kw.get("flags", 0) | W.F_FD
        msg = W.pack_req(op, **kw)
        with self._lock:
            try:
                s = self._connection()
                if fd is not None:
                    W.send_with_fd(s, msg, fd)
                else:
                    s.sendall(msg)
                status, plen, value = W.REP.unpack(W.recv_exact(s, W.REP.size))
                payload = W.recv_exact(s, plen) if plen else b""
            except (EOFError, OSError) as e:
                self._drop()
                raise L.BciError(L.ERR_CUDA, f"GPU broker connection lost: {e}") from None
        if status == L.OK:
            return value, payload
        text = payload.decode("utf-8", "replace")
        if status == L.ERR_UNSUPPORTED:
            raise L.Unsupported(status, text)
        raise L.BciError(status, text)

    def _post(self, op, **kw):
        with self._lock:
            try:
                self._connection().sendall(W.pack_req(op, **kw))
            except (OSError, L.BciError):
                self._drop()

    def _drop(self):
        if self._sock is not None:
            try:
                self._sock.close()
            except OSError:
                pass
            self._sock = None

    def close(self):
        with self._lock:
            self._drop()

    def ping(self) -> dict:
        _, payload = self._call(W.PING)
        jobs, batches, max_batch, conns = struct.unpack("<4Q", payload[:32])
        return {"stats": {"jobs": jobs, "batches": batches, "max_batch": max_batch, "connections": conns}}

    # ---- buffers ---------------------------------------------------------------------------
    def alloc(self, nbytes):
        return self._call(W.ALLOC, n=int(nbytes))[0]

    def free(self, ptr):
        if ptr:
            self._post(W.FREE, p0=ptr)

    def upload(self, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        if arr.nbytes <= W.INLINE_MAX:
            return self._call(W.UPLOAD, n=arr.nbytes, payload=arr.tobytes())[0]
        fd = _anon_file()                      # the bytes cross in an anonymous memory file, the broker preads it
        try:
            view = memoryview(arr).cast("B")
            done = 0
            while done < arr.nbytes:
                done += os.write(fd, view[done: done + (64 << 20)])
            return self._call(W.UPLOAD, fd=fd, n=arr.nbytes)[0]
        finally:
            os.close(fd)

    def upload_fd(self, fd: int, offset: int, nbytes: int):
        """Device buffer filled straight from an open file (an input of the `files` map): no host copy here."""
        return self._call(W.UPLOAD, fd=fd, n=int(nbytes), i0=int(offset))[0]

    def download(self, ptr, n, dtype, offset=0):
        """`n` elements starting `offset` elements into the buffer."""
        dtype = np.dtype(dtype)
        nbytes = int(n) * dtype.itemsize
        skip = int(offset) * dtype.itemsize
        if nbytes <= W.INLINE_MAX:
            return np.frombuffer(self._call(W.DOWNLOAD, p0=ptr, n=nbytes, i1=skip)[1], dtype=dtype).copy()
        fd = _anon_file()
        try:
            self._call(W.DOWNLOAD, fd=fd, p0=ptr, n=nbytes, i1=skip)
            out = np.empty(int(n), dtype=dtype)
            view = memoryview(out).cast("B")
            done = 0
            while done < nbytes:
                got = os.preadv(fd, [view[done: done + (64 << 20)]], done)
                if got <= 0:
                    raise L.BciError(L.ERR_CUDA, "GPU broker wrote a short download")
                done += got
            return out
        finally:
            os.close(fd)

    def export(self, ptr, nbytes: int, fd: int, header: bytes = b"") -> bytes:
        """Array -> the open file `fd` behind `header` (pinned path in the broker); returns the SHA-256 of
        header||array, computed on the device while the bytes are on their way out."""
        return self._call(W.EXPORT, fd=fd, p0=ptr, n=int(nbytes), payload=bytes(header))[1]

    def copy_buffer(self, src, nbytes):
        return self._call(W.COPY, p0=src, n=int(nbytes))[0]

    # ---- ops (one request each; the broker packs concurrent requests into one launch) -------
    def rand(self, n, seed):
        return self._call(W.RAND, n=int(n), s0=int(seed) & 0xFFFFFFFFFFFFFFFF)[0]

    def rand_pcg(self, kind, n, state, inc, low=0, span=0):
        m64 = (1 << 64) - 1
        payload = struct.pack("<4Q", state & m64, (state >> 64) & m64, inc & m64, (inc >> 64) & m64)
        ptr, rep = self._call(W.RAND_PCG, code=int(kind), n=int(n), i0=int(low), s0=int(span) & m64, payload=payload)
        return ptr, struct.unpack("<I", rep[:4])[0]

    def reduce_rand(self, op, n, seed, pre=L.PRE_NONE):
        bits = self._call(W.REDUCE_RAND, code=op, pre=pre, n=int(n), s0=int(seed) & 0xFFFFFFFFFFFFFFFF)[0]
        return _decode_reduce(bits, op, L.F64)

    def fill_new(self, kind, dtype, n, a, b=None, c=None):
        code = _code(dtype)
        isf = code in (L.F32, L.F64)
        cbits = W.scalar_bits(c, isf)
        if cbits >= 1 << 63:
            cbits -= 1 << 64          # i0 is a signed field
        return self._call(W.FILL, code=int(kind), dtype=code, n=int(n), s0=W.scalar_bits(a, isf), s1=W.scalar_bits(b, isf),
                          i0=cbits, flags=W.F_ADDEND if c is not None else 0)[0]

    def cast_new(self, src_dtype, dst_dtype, ptr, n):
        return self._call(W.CAST, dtype=_code(src_dtype), aux=_code(dst_dtype), p0=ptr, n=int(n))[0]

    def rand_normal(self, n, seed, loc, scale):
        return self._call(W.RAND_NORMAL, n=int(n), i0=int(seed) & 0x7FFFFFFFFFFFFFFF, s0=W.scalar_bits(loc, True),
                          s1=W.scalar_bits(scale, True))[0]

    def rand_mt19937(self, n, state):
        st = np.ascontiguousarray(state, dtype="<u4")
        if st.size != W.MT_STATE_WORDS:
            raise L.BciError(L.ERR_INVALID, "MT19937 state must be 625 words")
        ptr, payload = self._call(W.RAND_MT, n=int(n), payload=st.tobytes())
        return ptr, np.frombuffer(payload, dtype="<u4", count=W.MT_STATE_WORDS).copy()

    def ewise_new(self, op, dtype, in0, in1, scalar, n):
        code = _code(dtype)
        isf = code in (L.F32, L.F64)
        return self._call(W.EWISE, code=op, dtype=code, p0=in0, p1=in1 or 0, n=int(n),
                          flags=W.F_IN1 if in1 is not None else 0, s0=W.scalar_bits(scalar, isf))[0]

    def reduce(self, op, dtype, in_ptrs, ns, pre=L.PRE_NONE):
        code = _code(dtype)
        bits = self._call(W.REDUCE, code=op, dtype=code, pre=pre, p0=in_ptrs[0], n=int(ns[0]))[0]
        return [_decode_reduce(bits, op, code)]

    def reduce_fused(self, op, ew_op, dtype, in0, in1, scalars, ns, pre=L.PRE_NONE, addends=None):
        code = _code(dtype)
        isf = code in (L.F32, L.F64)
        flags = (W.F_IN1 if in1 is not None else 0) | (W.F_ADDEND if addends is not None else 0)
        bits = self._call(W.REDUCE_FUSED, code=op, dtype=code, pre=pre, aux=ew_op, flags=flags, p0=in0[0],
                          p1=in1[0] if in1 is not None else 0, n=int(ns[0]),
                          s0=W.scalar_bits(scalars[0] if scalars is not None else None, isf),
                          s1=W.scalar_bits(addends[0] if addends is not None else None, isf))[0]
        return [_decode_reduce(bits, op, code)]

    def argext(self, op, dtype, in_ptrs, ns):
        v = self._call(W.ARGEXT, code=op, dtype=_code(dtype), p0=in_ptrs[0], n=int(ns[0]))[0]
        return [np.int64(v)]

    def cumsum_new(self, dtype, ptr, n):
        return self._call(W.CUMSUM, dtype=_code(dtype), p0=ptr, n=int(n))[0]

    def sort_new(self, dtype, ptr, n, inplace=False):
        return self._call(W.SORT, dtype=_code(dtype), p0=ptr, n=int(n), flags=W.F_INPLACE if inplace else 0)[0]

    def argsort_new(self, dtype, ptr, n):
        return self._call(W.ARGSORT, dtype=_code(dtype), p0=ptr, n=int(n))[0]

    def matmul_new(self, a, b, M, N, K):
        return self._call(W.MATMUL, p0=a, p1=b, i0=int(M), i1=int(N), i2=int(K))[0]

    # ---- content hashing for the object store (storage.py) ------------------------------------
    def sha256(self, ptr, nbytes):
        return self._call(W.SHA256_DEV, p0=ptr, n=int(nbytes))[1]

    def sha256_fd(self, fd: int) -> bytes:
        """SHA-256 of an open regular file; the broker reads it through the descriptor."""
        return self._call(W.SHA256_FD, fd=fd)[1]

    def sha256_fd_ex(self, fd: int) -> tuple[bytes, bool]:
        """-> (digest, answered from the export table without reading the file)."""
        value, payload = self._call(W.SHA256_FD, fd=fd)
        return payload, bool(value)

    def sha256_bytes(self, data: bytes) -> bytes:
        return self._call(W.SHA256_BYTES, payload=bytes(data))[1]

    def sha256_update(self, state: bytes | None, chunk: bytes) -> bytes:
        if state is None:
            return self._call(W.SHA256_UPDATE, payload=bytes(chunk))[1]
        return self._call(W.SHA256_UPDATE, flags=W.F_STATE, payload=state + bytes(chunk))[1]

    def sha256_final(self, state: bytes) -> bytes:
        return self._call(W.SHA256_FINAL, payload=state)[1]
