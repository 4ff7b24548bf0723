"""GPU broker: the one process (per GPU) that owns the CUDA context and packs the device ops
of many concurrent `/v1/execute` requests into one launch per op class (protocol and
rationale: include/bci_broker.h).

  NativeBroker   the product: C++ threads inside libbci.so (csrc/bci_broker.cu) started with
                 bci_broker_start(); no Python on the per-op path.
  BrokerProcess  launches `python -m bee_code_interpreter_b200.broker` as the GPU-owning
                 process (it calls bci_init + bci_broker_start and then just sleeps).
  GpuBroker      a Python implementation of the SAME wire protocol over an injected device
                 module — used by the CPU test-suite with tests/fake_device.py to exercise
                 ownership, batching and the client without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
import queue
import socket
import struct
import tempfile
import threading
import time
from collections import defaultdict

import numpy as np

from . import _lib as L
from . import wire as W

STATE_SIZE = C.sizeof(L.Sha256State)
_DTYPE = {L.F32: np.dtype(np.float32), L.F64: np.dtype(np.float64), L.I32: np.dtype(np.int32), L.I64: np.dtype(np.int64)}


class NativeBroker:
    """In-process handle on the C++ broker of libbci.so (the calling process owns the GPU)."""

    def __init__(self, socket_path: str | None = None, coalesce_us: int = 150, device: int | None = None):
        from . import device as dev

        if not dev.is_initialised():
            dev.init(device)
        if socket_path is None:
            socket_path = os.path.join(tempfile.mkdtemp(prefix="bci-broker-"), "broker.sock")
        self.socket_path = socket_path
        L.check(L.load().bci_broker_start(socket_path.encode(), int(coalesce_us)))

    @property
    def stats(self) -> dict:
        arr = (C.c_uint64 * 4)()
        L.load().bci_broker_stats(arr)
        return {"jobs": arr[0], "batches": arr[1], "max_batch": arr[2], "connections": arr[3]}

    def healthy(self) -> bool:
        return bool(L.load().bci_healthy())

    def close(self):
        L.load().bci_broker_stop()


class _Job:
    __slots__ = ("key", "args", "result", "error", "done")

    def __init__(self, key, args):
        self.key, self.args = key, args
        self.result = None
        self.error = None
        self.done = threading.Event()


class GpuBroker:
    """Python stand-in for the native broker: same protocol, injected device module."""

    def __init__(self, device_module, socket_path: str | None = None, coalesce_us: int = 150):
        self.dev = device_module
        if not self.dev.is_initialised():
            self.dev.init()
        if socket_path is None:
            socket_path = os.path.join(tempfile.mkdtemp(prefix="bci-broker-"), "broker.sock")
        self.socket_path = socket_path
        self.coalesce_s = coalesce_us / 1e6
        self._srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        self._srv.bind(socket_path)
        self._srv.listen(1024)
        self._jobs: "queue.SimpleQueue[_Job | None]" = queue.SimpleQueue()
        self._running = True
        self.stats = {"batches": 0, "jobs": 0, "max_batch": 0, "connections": 0}
        for target, name in ((self._accept_loop, "bci-accept"), (self._dispatch_loop, "bci-dispatch")):
            threading.Thread(target=target, name=name, daemon=True).start()

    def close(self):
        if not self._running:
            return
        self._running = False
        self._jobs.put(None)
        try:
            self._srv.close()
        except OSError:
            pass
        try:
            os.unlink(self.socket_path)
        except OSError:
            pass

    # ---- batching dispatcher ---------------------------------------------------------------
    def submit(self, key, args):
        job = _Job(key, args)
        self._jobs.put(job)
        job.done.wait()
        if job.error is not None:
            raise job.error
        return job.result

    def _dispatch_loop(self):
        while self._running:
            job = self._jobs.get()
            if job is None:
                break
            jobs = [job]
            if self.coalesce_s > 0:
                time.sleep(self.coalesce_s)  # let concurrent requests' ops arrive
            while True:
                try:
                    j = self._jobs.get_nowait()
                except queue.Empty:
                    break
                if j is None:
                    self._running = False
                    break
                jobs.append(j)
            groups = defaultdict(list)
            for j in jobs:
                groups[j.key].append(j)
            for key, group in groups.items():
                self.stats["batches"] += 1
                self.stats["jobs"] += len(group)
                self.stats["max_batch"] = max(self.stats["max_batch"], len(group))
                try:
                    for j, r in zip(group, self._run_batch(key, [j.args for j in group])):
                        j.result = r
                except Exception as e:
                    if len(group) == 1:
                        group[0].error = e
                    else:  # isolate the offender: one bad request must not fail the others
                        for j in group:
                            try:
                                j.result = self._run_batch(key, [j.args])[0]
                            except Exception as e1:
                                j.error = e1
                for j in group:
                    j.done.set()

    def _run_batch(self, key, argl):
        dev, kind = self.dev, key[0]
        col = lambda i: [a[i] for a in argl]  # noqa: E731
        if kind == W.RAND:
            dev.rand_f64(col(0), col(1), col(2))
            return [None] * len(argl)
        if kind == W.RAND_NORMAL:
            dev.rand_normal_f64(col(0), col(1), col(2), col(3), col(4))
            return [None] * len(argl)
        if kind == W.RAND_MT:
            return list(dev.rand_mt19937(col(0), col(1), np.stack(col(2))))
        if kind == W.RAND_PCG:
            return [int(r) for r in dev.rand_pcg64(key[1], col(0), col(1), col(2))]
        if kind == W.REDUCE_RAND:
            return list(dev.reduce_rand_f64(key[1], col(0), col(1), pre=key[2]))
        if kind == W.FILL:
            _, fk, dt, has_c = key
            dev.fill(fk, dt, col(0), col(1), col(2), col(3), col(4) if has_c else None)
            return [None] * len(argl)
        if kind == W.CAST:
            dev.cast(key[1], key[2], col(0), col(1), col(2))
            return [None] * len(argl)
        if kind == W.EWISE:
            _, op, dt, has_in1 = key
            dev.ewise(op, dt, col(0), col(1) if has_in1 else None, None if has_in1 else col(2), col(3), col(4))
            return [None] * len(argl)
        if kind == W.REDUCE:
            _, op, pre, dt = key
            return list(dev.reduce(op, dt, col(0), col(1), pre=pre))
        if kind == W.REDUCE_FUSED:
            _, op, pre, ew, dt, has_in1, has_add = key
            return list(dev.reduce_fused(op, ew, dt, col(0), col(1) if has_in1 else None, None if has_in1 else col(2),
                                         col(4), pre=pre, addends=col(3) if has_add else None))
        if kind == W.ARGEXT:
            return list(dev.argext(key[1], key[2], col(0), col(1)))
        if kind in (W.CUMSUM, W.SORT, W.ARGSORT):
            fn = {W.CUMSUM: dev.cumsum, W.SORT: dev.sort, W.ARGSORT: dev.argsort}[kind]
            fn(key[1], col(0), col(1), col(2))
            return [None] * len(argl)
        if kind == W.MATMUL:
            dev.matmul_f32(col(0), col(1), col(2), key[1], key[2], key[3])
            return [None] * len(argl)
        if kind == W.SHA256_DEV:
            return dev.sha256_device(col(0), col(1))
        if kind == W.SHA256_BYTES:
            return dev.sha256_host(col(0))
        if kind == W.SHA256_UPDATE:
            dev.sha256_update_many(col(0), col(1))
            return col(0)
        if kind == W.SHA256_FINAL:
            return dev.sha256_final_many(col(0))
        raise ValueError(f"unknown op class {kind!r}")

    # ---- connections ------------------------------------------------------------------------
    def _accept_loop(self):
        while self._running:
            try:
                conn, _ = self._srv.accept()
            except OSError:
                break
            self.stats["connections"] += 1
            threading.Thread(target=self._serve, args=(conn,), name="bci-conn", daemon=True).start()

    def _serve(self, conn):
        owned: dict[int, int] = {}  # device pointer -> bytes: the buffers this script may name
        dev = self.dev

        def need(ptr, nbytes):
            have = owned.get(ptr)
            if have is None or nbytes > have:
                raise L.BciError(L.ERR_INVALID, "buffer does not belong to this request")
            return ptr

        def new(nbytes):
            p = dev.alloc(nbytes)
            owned[p] = max(int(nbytes), 1)
            return p

        try:
            while True:
                fds: list[int] = []
                try:
                    head = W.recv_exact_fds(conn, W.REQ.size, fds)
                    (op, code, dtc, pre, aux, flags, p0, p1, n, i0, i1, i2, s0, s1, plen, _r) = W.REQ.unpack(head)
                    payload = W.recv_exact_fds(conn, plen, fds) if plen else b""
                except (EOFError, OSError):
                    for f in fds:
                        os.close(f)
                    break
                if op == W.FREE:
                    if owned.pop(p0, None) is not None:
                        dev.free(p0)
                    continue
                try:
                    if flags & W.F_FD and not fds:
                        raise L.BciError(L.ERR_INVALID, "request announces a descriptor but none arrived")
                    value, out = self._handle(op, code, dtc, pre, aux, flags, p0, p1, n, (i0, i1, i2), s0, s1, payload, need, new,
                                              fds[0] if fds else None)
                    conn.sendall(W.REP.pack(L.OK, len(out), value & 0xFFFFFFFFFFFFFFFF) + out)
                except L.BciError as e:
                    text = str(e).encode()
                    conn.sendall(W.REP.pack(e.code, len(text), 0) + text)
                except Exception as e:  # protocol misuse etc.
                    text = repr(e).encode()
                    conn.sendall(W.REP.pack(L.ERR_INVALID, len(text), 0) + text)
                finally:
                    for f in fds:
                        os.close(f)
        except OSError:
            pass
        finally:
            for p in list(owned):
                try:
                    dev.free(p)
                except Exception:
                    pass
            owned.clear()
            conn.close()

    def _handle(self, op, code, dtc, pre, aux, flags, p0, p1, n, ints, s0, s1, payload, need, new, fd=None):
        dev = self.dev
        dt = _DTYPE.get(dtc)
        isf = dtc in (L.F32, L.F64)
        if op == W.ALLOC:
            return new(n), b""
        if op == W.PING:
            s = self.stats
            return 0, struct.pack("<4Q", s["jobs"], s["batches"], s["max_batch"], s["connections"])
        if op == W.RAND:
            p = new(n * 8)
            self.submit((W.RAND,), (p, n, s0))
            return p, b""
        if op == W.RAND_NORMAL:
            p = new(n * 8)
            self.submit((W.RAND_NORMAL,), (p, n, ints[0], W.scalar_from_bits(s0, True), W.scalar_from_bits(s1, True)))
            return p, b""
        if op == W.RAND_PCG:
            if len(payload) != 32 or code not in dev.PCG_OUT_DTYPE:
                raise L.BciError(L.ERR_INVALID, "PCG64 state must be 4 words")
            w = struct.unpack("<4Q", payload)
            p = new(n * dev.PCG_OUT_DTYPE[code].itemsize)
            rej = self.submit((W.RAND_PCG, code), (p, n, (w[0] | (w[1] << 64), w[2] | (w[3] << 64), ints[0], s0)))
            return p, struct.pack("<I", rej)
        if op == W.REDUCE_RAND:
            return _result_bits(np.float64(self.submit((W.REDUCE_RAND, code, pre), (n, s0)))), b""
        if op == W.FILL:
            has_c = bool(flags & W.F_ADDEND)
            out = new(n * dt.itemsize)
            self.submit((W.FILL, code, dt, has_c), (out, n, W.scalar_from_bits(s0, isf), W.scalar_from_bits(s1, isf),
                                                   W.scalar_from_bits(ints[0] & 0xFFFFFFFFFFFFFFFF, isf)))
            return out, b""
        if op == W.CAST:
            ddt = _DTYPE.get(aux)
            if dt is None or ddt is None:
                raise L.BciError(L.ERR_INVALID, "unknown dtype")
            out = new(n * ddt.itemsize)
            self.submit((W.CAST, dt, ddt), (need(p0, n * dt.itemsize), out, n))
            return out, b""
        if op == W.RAND_MT:
            if len(payload) != 4 * W.MT_STATE_WORDS:
                raise L.BciError(L.ERR_INVALID, "MT19937 state must be 625 words")
            p = new(n * 8)
            after = self.submit((W.RAND_MT,), (p, n, np.frombuffer(payload, dtype="<u4").copy()))
            return p, np.asarray(after, dtype="<u4").tobytes()
        if op == W.EWISE:
            nb = n * dt.itemsize
            has_in1 = bool(flags & W.F_IN1)
            out = new(nb)
            self.submit((W.EWISE, code, dt, has_in1),
                        (need(p0, nb), need(p1, nb) if has_in1 else None, W.scalar_from_bits(s0, isf), out, n))
            return out, b""
        if op == W.REDUCE:
            r = self.submit((W.REDUCE, code, pre, dt), (need(p0, n * dt.itemsize), n))
            return _result_bits(r), b""
        if op == W.REDUCE_FUSED:
            nb = n * dt.itemsize
            has_in1, has_add = bool(flags & W.F_IN1), bool(flags & W.F_ADDEND)
            r = self.submit((W.REDUCE_FUSED, code, pre, aux, dt, has_in1, has_add),
                            (need(p0, nb), need(p1, nb) if has_in1 else None, W.scalar_from_bits(s0, isf),
                             W.scalar_from_bits(s1, isf), n))
            return _result_bits(r), b""
        if op == W.ARGEXT:
            return int(self.submit((W.ARGEXT, code, dt), (need(p0, n * dt.itemsize), n))), b""
        if op == W.CUMSUM:
            out = new(n * dev.cumsum_out_dtype(dt).itemsize)
            self.submit((W.CUMSUM, dt), (need(p0, n * dt.itemsize), out, n))
            return out, b""
        if op in (W.SORT, W.ARGSORT):
            need(p0, n * dt.itemsize)
            inplace = op == W.SORT and bool(flags & W.F_INPLACE)
            out = p0 if inplace else new(n * (8 if op == W.ARGSORT else dt.itemsize))
            self.submit((op, dt), (p0, out, n))
            return out, b""
        if op == W.MATMUL:
            M, N, K = ints
            out = new(M * N * 4)
            self.submit((W.MATMUL, M, N, K), (need(p0, M * K * 4), need(p1, K * N * 4), out))
            return out, b""
        if op == W.COPY:
            out = new(n)
            dev.d2d(out, need(p0, n), n)
            return out, b""
        if op == W.UPLOAD:
            p = new(n)
            if n:
                if flags & W.F_FD:
                    dev.h2d_fd(p, fd, ints[0], n)
                else:
                    dev.h2d(p, np.frombuffer(payload, dtype=np.uint8, count=n))
                    dev.sync()
            return p, b""
        if op == W.DOWNLOAD:
            skip = ints[1]
            if skip < 0:
                raise L.BciError(L.ERR_INVALID, "negative offset")
            need(p0, skip + n)
            if flags & W.F_FD:
                if skip:        # (handles of an injected device module are opaque: go through the host)
                    os.pwrite(fd, dev.download(p0, n, np.uint8, skip).tobytes(), ints[0])
                else:
                    dev.d2h_fd(fd, ints[0], p0, n)
                return 0, b""
            return 0, dev.download(p0, n, np.uint8, skip).tobytes()
        if op == W.EXPORT:
            if fd is None:
                raise L.BciError(L.ERR_INVALID, "export needs a descriptor")
            need(p0, n)
            # (the native broker assembles header||array on the device; handles of an injected device
            # module are opaque, so this stand-in builds the message on the host)
            host = np.concatenate([np.frombuffer(payload, dtype=np.uint8), dev.download(p0, n, np.uint8)])
            msg = dev.upload(host)
            try:
                digest = self.submit((W.SHA256_DEV,), (msg, len(host)))
                dev.d2h_fd(fd, ints[0], msg, len(host))
                if ints[0] == 0:
                    os.ftruncate(fd, len(host))
            finally:
                dev.free(msg)
            return 0, digest
        if op == W.SHA256_DEV:
            return 0, self.submit((W.SHA256_DEV,), (need(p0, n), n))
        if op == W.SHA256_FD:
            if fd is None:
                raise L.BciError(L.ERR_INVALID, "hashing a file needs its descriptor")
            size = os.fstat(fd).st_size
            data = os.pread(fd, size, 0) if size else b""
            return 0, self.submit((W.SHA256_BYTES,), (data,))
        if op == W.SHA256_BYTES:
            return 0, self.submit((W.SHA256_BYTES,), (payload,))
        if op == W.SHA256_UPDATE:
            if flags & W.F_STATE:
                st, chunk = dev.Sha256Stream.from_bytes(payload[:STATE_SIZE]), payload[STATE_SIZE:]
            else:
                st, chunk = dev.Sha256Stream(), payload
            return 0, self.submit((W.SHA256_UPDATE,), (st, chunk)).to_bytes()
        if op == W.SHA256_FINAL:
            return 0, self.submit((W.SHA256_FINAL,), (dev.Sha256Stream.from_bytes(payload[:STATE_SIZE]),))
        raise ValueError(f"unknown message {op!r}")


def _result_bits(value) -> int:
    """The 8-byte result slot bci_reduce would have written for this value."""
    raw = np.asarray(value).tobytes()
    return int.from_bytes(raw.ljust(8, b"\0"), "little")


class BrokerProcess:
    """The GPU owner as its own OS process.  One per GPU; any number of service processes and
    script processes connect to its socket."""

    def __init__(self, device: int | None = None, coalesce_us: int = 150, fake_device: str | None = None,
                 extra_pythonpath: str | None = None):
        import sys

        self._dir = tempfile.mkdtemp(prefix="bci-broker-")
        self.socket_path = os.path.join(self._dir, "broker.sock")
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        env = dict(os.environ)
        paths = [root] + ([extra_pythonpath] if extra_pythonpath else []) + ([env["PYTHONPATH"]] if env.get("PYTHONPATH") else [])
        env["PYTHONPATH"] = os.pathsep.join(paths)
        cmd = [sys.executable, "-m", "bee_code_interpreter_b200.broker", "--socket", self.socket_path,
               "--coalesce-us", str(coalesce_us)]
        if device is not None:
            cmd += ["--device", str(device)]
        if fake_device:
            cmd += ["--fake-device", fake_device]
        self._cmd, self._env, self._root = cmd, env, root
        self.restarts = 0
        self._start()

    def _start(self):
        import subprocess

        self._proc = subprocess.Popen(self._cmd, stdout=subprocess.PIPE, env=self._env, cwd=self._root)
        line = self._proc.stdout.readline()
        if line.strip() != b"ready":
            raise L.BciError(L.ERR_NOT_INIT, "GPU broker process failed to start (is a B200 visible?)")

    def alive(self) -> bool:
        return self._proc.poll() is None

    def restart(self) -> None:
        """A broker that died (or gave up its poisoned CUDA context, see main()) is started again on the same socket."""
        if self.alive():
            return
        try:
            os.unlink(self.socket_path)
        except OSError:
            pass
        self.restarts += 1
        self._start()

    def close(self):
        if self._proc.poll() is None:
            self._proc.terminate()
            try:
                self._proc.wait(timeout=5)
            except Exception:
                self._proc.kill()
        try:
            os.rmdir(self._dir)
        except OSError:
            pass


def main():  # python -m bee_code_interpreter_b200.broker --socket PATH [--device N]
    import argparse
    import signal
    import sys

    ap = argparse.ArgumentParser()
    ap.add_argument("--socket", required=True)
    ap.add_argument("--device", type=int, default=None)
    ap.add_argument("--coalesce-us", type=int, default=150)
    ap.add_argument("--fake-device", default=None, help="(tests only) module name of a device double")
    args = ap.parse_args()
    if args.fake_device:
        import importlib

        broker = GpuBroker(importlib.import_module(args.fake_device), socket_path=args.socket, coalesce_us=args.coalesce_us)
    else:
        broker = NativeBroker(socket_path=args.socket, coalesce_us=args.coalesce_us, device=args.device)
    sys.stdout.write("ready\n")
    sys.stdout.flush()
    parent = os.getppid()
    signal.signal(signal.SIGTERM, lambda *_: sys.exit(0))
    try:
        while os.getppid() == parent:   # do not outlive the service that started us
            time.sleep(0.25)
            # a sticky CUDA error (one request's faulting launch) poisons the whole context: hand the GPU back
            # and let the executor start a fresh broker, instead of failing every later request
            if hasattr(broker, "healthy") and not broker.healthy():
                sys.stderr.write("bci broker: CUDA context is unusable, exiting for a restart\n")
                os._exit(17)
    finally:
        broker.close()


if __name__ == "__main__":
    main()
