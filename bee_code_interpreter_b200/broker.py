"""GPU broker: the one process (per GPU) that owns the CUDA context and packs the device ops
of many concurrent `/v1/execute` requests into one launch per op class.

Each reference request is its own OS process (executor/server.rs:152-157 spawns the script;
kubernetes_code_executor.py:248-264 gives it a fresh pod).  Here the sandboxed script
processes hold no CUDA context: their numpy proxies (numpy_proxy.DeviceArray) send small op
descriptors over a Unix socket; array bytes never cross it unless the script really reads
an array on the host.  The broker
  * owns every device buffer, scoped to the connection that created it (a script can only
    name its own buffers; everything it leaves behind is freed when it disconnects — the
    analogue of deleting the pod, kubernetes_code_executor.py:262-264);
  * queues compute ops and drains the queue in a dispatcher thread: all pending ops of the
    same class (op, dtype, pre-transform) become ONE batched C-ABI call = one launch.
Wire format: multiprocessing.connection (length-prefixed pickles) over AF_UNIX.
"""
from __future__ import annotations

import os
import queue
import tempfile
import threading
import time
from collections import defaultdict
from multiprocessing.connection import Listener

import numpy as np

from . import _lib as L

INLINE_MAX = 1 << 16  # arrays up to 64 KiB travel inline; larger ones through /dev/shm files


class _Job:
    __slots__ = ("key", "args", "result", "error", "done")

    def __init__(self, key, args):
        self.key, self.args = key, args
        self.result = None
        self.error = None
        self.done = threading.Event()


class GpuBroker:
    def __init__(self, device_module=None, socket_path: str | None = None, coalesce_us: int = 150):
        if device_module is None:
            from . import device as device_module  # the real C-ABI binding
            if not device_module.is_initialised():
                device_module.init()
        self.dev = device_module
        if socket_path is None:
            socket_path = os.path.join(tempfile.mkdtemp(prefix="bci-broker-"), "broker.sock")
        self.socket_path = socket_path
        self.coalesce_s = coalesce_us / 1e6
        self._listener = Listener(address=socket_path, family="AF_UNIX", backlog=1024)
        self._jobs: "queue.SimpleQueue[_Job | None]" = queue.SimpleQueue()
        self._running = True
        self.stats = {"batches": 0, "jobs": 0, "max_batch": 0, "connections": 0}
        self._threads = [threading.Thread(target=self._accept_loop, name="bci-accept", daemon=True),
                         threading.Thread(target=self._dispatch_loop, name="bci-dispatch", daemon=True)]
        for t in self._threads:
            t.start()

    # ---- lifecycle ---------------------------------------------------------------------
    def close(self):
        if not self._running:
            return
        self._running = False
        self._jobs.put(None)
        try:
            self._listener.close()
        except Exception:
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
                    results = self._run_batch(key, [j.args for j in group])
                    for j, r in zip(group, results):
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
        dev = self.dev
        kind = key[0]
        if kind == "rand":
            dev.rand_f64([a[0] for a in argl], [a[1] for a in argl], [a[2] for a in argl])
            return [None] * len(argl)
        if kind == "ewise":
            _, op, dt, has_in1 = key
            dev.ewise(op, dt, [a[0] for a in argl], [a[1] for a in argl] if has_in1 else None,
                      None if has_in1 else [a[2] for a in argl], [a[3] for a in argl], [a[4] for a in argl])
            return [None] * len(argl)
        if kind == "reduce":
            _, op, pre, dt = key
            res = dev.reduce(op, dt, [a[0] for a in argl], [a[1] for a in argl], pre=pre)
            return list(res)
        if kind == "reduce_fused":
            _, op, pre, ew, dt, has_in1, has_add = key
            res = dev.reduce_fused(op, ew, dt, [a[0] for a in argl], [a[1] for a in argl] if has_in1 else None,
                                   None if has_in1 else [a[2] for a in argl], [a[4] for a in argl], pre=pre,
                                   addends=[a[3] for a in argl] if has_add else None)
            return list(res)
        if kind == "argext":
            _, op, dt = key
            return list(dev.argext(op, dt, [a[0] for a in argl], [a[1] for a in argl]))
        if kind == "cumsum":
            dev.cumsum(key[1], [a[0] for a in argl], [a[1] for a in argl], [a[2] for a in argl])
            return [None] * len(argl)
        if kind == "sort":
            dev.sort(key[1], [a[0] for a in argl], [a[1] for a in argl], [a[2] for a in argl])
            return [None] * len(argl)
        if kind == "argsort":
            dev.argsort(key[1], [a[0] for a in argl], [a[1] for a in argl], [a[2] for a in argl])
            return [None] * len(argl)
        if kind == "matmul":
            _, M, N, K = key
            dev.matmul_f32([a[0] for a in argl], [a[1] for a in argl], [a[2] for a in argl], M, N, K)
            return [None] * len(argl)
        if kind == "sha256":
            return dev.sha256_device([a[0] for a in argl], [a[1] for a in argl])
        if kind == "sha256_host":      # whole payloads (bytes or mmapped files) of many requests
            return dev.sha256_host([a[0] for a in argl])
        if kind == "sha256_update":    # one chunk per streaming writer
            streams = [a[0] for a in argl]
            dev.sha256_update_many(streams, [a[1] for a in argl])
            return streams
        if kind == "sha256_final":
            return dev.sha256_final_many([a[0] for a in argl])
        raise ValueError(f"unknown op class {kind!r}")

    # ---- connections ------------------------------------------------------------------------
    def _accept_loop(self):
        while self._running:
            try:
                conn = self._listener.accept()
            except Exception:
                break
            self.stats["connections"] += 1
            threading.Thread(target=self._serve, args=(conn,), name="bci-conn", daemon=True).start()

    def _serve(self, conn):
        owned: dict[int, int] = {}  # device pointer -> bytes, buffers this script may name
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
                try:
                    msg = conn.recv()
                except (EOFError, OSError):
                    break
                op = msg[0]
                try:
                    if op == "free":  # fire and forget
                        p = msg[1]
                        if owned.pop(p, None) is not None:
                            dev.free(p)
                        continue
                    reply = self._handle(op, msg, need, new, owned)
                    conn.send(("ok", reply))
                except L.BciError as e:
                    conn.send(("err", e.code, str(e)))
                except Exception as e:  # protocol misuse etc.
                    conn.send(("err", L.ERR_INVALID, repr(e)))
        finally:
            for p in list(owned):
                try:
                    dev.free(p)
                except Exception:
                    pass
            owned.clear()
            try:
                conn.close()
            except Exception:
                pass

    def _handle(self, op, msg, need, new, owned):
        dev = self.dev
        if op == "alloc":
            return new(msg[1])
        if op == "rand":
            _, n, seed = msg
            p = new(n * 8)
            self.submit(("rand",), (p, n, seed))
            return p
        if op == "ewise":
            _, code, dt, in0, in1, scalar, n = msg
            dt = np.dtype(dt)
            nb = n * dt.itemsize
            out = new(nb)
            self.submit(("ewise", code, dt, in1 is not None),
                        (need(in0, nb), need(in1, nb) if in1 is not None else None, scalar, out, n))
            return out
        if op == "reduce":
            _, code, pre, dt, p, n = msg
            dt = np.dtype(dt)
            return self.submit(("reduce", code, pre, dt), (need(p, n * dt.itemsize), n))
        if op == "reduce_fused":
            _, code, pre, ew, dt, in0, in1, scalar, addend, n = msg
            dt = np.dtype(dt)
            nb = n * dt.itemsize
            return self.submit(("reduce_fused", code, pre, ew, dt, in1 is not None, addend is not None),
                               (need(in0, nb), need(in1, nb) if in1 is not None else None, scalar, addend, n))
        if op == "argext":
            _, code, dt, p, n = msg
            dt = np.dtype(dt)
            return int(self.submit(("argext", code, dt), (need(p, n * dt.itemsize), n)))
        if op == "cumsum":
            _, dt, p, n = msg
            dt = np.dtype(dt)
            odt = dev.cumsum_out_dtype(dt)
            out = new(n * odt.itemsize)
            self.submit(("cumsum", dt), (need(p, n * dt.itemsize), out, n))
            return out
        if op in ("sort", "argsort"):
            _, dt, p, n, inplace = msg
            dt = np.dtype(dt)
            need(p, n * dt.itemsize)
            out = p if (inplace and op == "sort") else new(n * (8 if op == "argsort" else dt.itemsize))
            self.submit((op, dt), (p, out, n))
            return out
        if op == "matmul":
            _, a, b, M, N, K = msg
            out = new(M * N * 4)
            self.submit(("matmul", M, N, K), (need(a, M * K * 4), need(b, K * N * 4), out))
            return out
        if op == "d2d":
            _, src, nbytes = msg
            out = new(nbytes)
            dev.d2d(out, need(src, nbytes), nbytes)
            return out
        if op == "upload":
            _, dt, n, payload = msg
            dt = np.dtype(dt)
            arr = _payload_array(payload, dt, n)
            p = new(arr.nbytes)
            if arr.nbytes:
                dev.h2d(p, arr)
                dev.sync()
            return p
        if op == "download":
            _, p, dt, n, shm_path = msg
            dt = np.dtype(dt)
            need(p, n * dt.itemsize)
            if shm_path is None:
                return dev.download(p, n, dt).tobytes()
            mm = np.memmap(shm_path, dtype=dt, mode="r+", shape=(n,))
            host = dev.download(p, n, dt)
            mm[:] = host
            mm.flush()
            del mm
            return None
        if op == "sha256":
            _, p, nbytes = msg
            return self.submit(("sha256",), (need(p, nbytes), nbytes))
        if op == "sha256_file":        # executor output files: hashed where they lie (page cache -> GPU)
            _, path = msg
            size = os.path.getsize(path)
            data = np.memmap(path, dtype=np.uint8, mode="r") if size else b""
            return self.submit(("sha256_host",), (data,))
        if op == "sha256_bytes":
            return self.submit(("sha256_host",), (msg[1],))
        if op == "sha256_update":      # streaming ObjectWriter: state travels with the message
            _, state, chunk = msg
            st = dev.Sha256Stream.from_bytes(state) if state is not None else dev.Sha256Stream()
            return self.submit(("sha256_update",), (st, chunk)).to_bytes()
        if op == "sha256_final":
            return self.submit(("sha256_final",), (dev.Sha256Stream.from_bytes(msg[1]),))
        if op == "ping":
            return {"device": dev.device_info() if hasattr(dev, "device_info") else None, "stats": dict(self.stats)}
        raise ValueError(f"unknown message {op!r}")


def _payload_array(payload, dt, n):
    kind, data = payload
    if kind == "inline":
        return np.frombuffer(data, dtype=dt, count=n)
    return np.array(np.memmap(data, dtype=dt, mode="r", shape=(n,)))


def main():  # standalone: python -m bee_code_interpreter_b200.broker --socket PATH [--device N]
    import argparse
    import signal
    import sys

    ap = argparse.ArgumentParser()
    ap.add_argument("--socket", required=True)
    ap.add_argument("--device", type=int, default=None)
    ap.add_argument("--coalesce-us", type=int, default=150)
    ap.add_argument("--fake-device", default=None, help="(tests only) python module path of a device double")
    args = ap.parse_args()
    if args.fake_device:
        import importlib

        device = importlib.import_module(args.fake_device)
        device.init(args.device)
    else:
        from . import device

        device.init(args.device)
    broker = GpuBroker(device, socket_path=args.socket, coalesce_us=args.coalesce_us)
    sys.stdout.write("ready\n")
    sys.stdout.flush()
    parent = os.getppid()
    signal.signal(signal.SIGTERM, lambda *_: sys.exit(0))
    try:
        while os.getppid() == parent:   # do not outlive the service that started us
            time.sleep(1.0)
    finally:
        broker.close()


class BrokerProcess:
    """The GPU owner as its own OS process (its Python threads do not share a GIL with the
    service's request handling).  One per GPU."""

    def __init__(self, device: int | None = None, coalesce_us: int = 150, fake_device: str | None = None,
                 extra_pythonpath: str | None = None):
        import subprocess
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
        self._proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, env=env, cwd=root)
        line = self._proc.stdout.readline()
        if line.strip() != b"ready":
            raise L.BciError(L.ERR_NOT_INIT, "GPU broker process failed to start (is a B200 visible?)")

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


if __name__ == "__main__":
    main()
