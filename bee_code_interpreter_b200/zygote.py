"""Warm-interpreter zygote: forks one child per `/v1/execute` request.

The reference keeps a queue of warm pods and spends one per request
(kubernetes_code_executor.py:151-189,248-264); inside the pod each request still pays a fresh
interpreter (`xonsh script.xsh`, executor/server.rs:152).  Here the warm resource is a Python
process with numpy, the proxy layer and the import hook already imported; a request costs one
fork() of it (~2 ms) and keeps the reference's isolation unit — one OS process per request.

The zygote is single-threaded and holds no CUDA context, so fork() is safe.  Protocol (Unix
stream socket, one connection per request):
    client -> zygote : 4-byte length + request dict (pickle from ZygoteClient, JSON from
                       csrc/bci_executor_server.cpp)
    child  -> client : 4-byte pid, first thing after the fork  (the zygote closes its copy)
    child  -> client : 4-byte exit code, just before os._exit  (absent if the child was killed
                       or crashed: the client maps EOF to -1, like server.rs:165)
Children are auto-reaped (SIGCHLD ignored).
"""
from __future__ import annotations

import json
import os
import pickle
import signal
import socket
import struct
import subprocess
import sys
import tempfile
import time


def _recv_exact(conn: socket.socket, n: int) -> bytes:
    buf = b""
    while len(buf) < n:
        chunk = conn.recv(n - len(buf))
        if not chunk:
            return buf
        buf += chunk
    return buf


def serve(socket_path: str) -> None:
    from . import script_runner  # imports numpy + proxy + hook once

    signal.signal(signal.SIGCHLD, signal.SIG_IGN)
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    try:
        os.unlink(socket_path)
    except OSError:
        pass
    srv.bind(socket_path)
    srv.listen(1024)
    sys.stdout.write("ready\n")
    sys.stdout.flush()
    parent = os.getppid()
    srv.settimeout(2.0)
    while True:
        try:
            conn, _ = srv.accept()
        except socket.timeout:
            if os.getppid() != parent:  # service died: do not linger
                break
            continue
        except OSError:
            break
        try:
            head = _recv_exact(conn, 4)
            if len(head) < 4:
                conn.close()
                continue
            (length,) = struct.unpack("<I", head)
            payload = _recv_exact(conn, length)
            # pickled dict from ZygoteClient, JSON object from the native executor server
            req = json.loads(payload) if payload[:1] == b"{" else pickle.loads(payload)
            if req.get("op") == "shutdown":
                conn.close()
                break
            pid = os.fork()
            if pid == 0:
                srv.close()
                code = 1
                try:
                    # the CHILD announces its pid: if the parent did, a short script could finish and
                    # write its exit code first, and the client would read the two words swapped
                    os.setsid()   # own process group: a timeout (and the end of the request) reaps what the script spawned
                    conn.sendall(struct.pack("<i", os.getpid()))
                    code = script_runner.run_in_child(req)
                finally:
                    try:
                        conn.sendall(struct.pack("<i", code))
                    except OSError:
                        pass
                    os._exit(code & 0xFF)
        except Exception:
            pass
        finally:
            conn.close()
    try:
        os.unlink(socket_path)
    except OSError:
        pass


def _kill_group(pid: int) -> None:
    """SIGKILL to the script's process group (the child called setsid, so pgid == pid)."""
    if pid <= 1:          # never signal 0 / -1 on a garbled reply
        return
    for fn in (os.killpg, os.kill):
        try:
            fn(pid, signal.SIGKILL)
        except (ProcessLookupError, PermissionError):
            pass


class ZygoteClient:
    """Service-side handle: starts the zygote process and submits requests to it."""

    def __init__(self):
        self._dir = tempfile.mkdtemp(prefix="bci-zygote-")
        self.socket_path = os.path.join(self._dir, "zygote.sock")
        env = dict(os.environ)
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        env["PYTHONPATH"] = root + (os.pathsep + env["PYTHONPATH"] if env.get("PYTHONPATH") else "")
        # no BLAS/OpenMP worker threads in the zygote: fork() must copy a single-threaded
        # process, and a sandboxed request gets one host thread (heavy array work goes to the GPU)
        for var in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
            env.setdefault(var, "1")
        self._proc = subprocess.Popen([sys.executable, "-m", "bee_code_interpreter_b200.zygote", self.socket_path],
                                      stdout=subprocess.PIPE, env=env, cwd=root)
        line = self._proc.stdout.readline()
        if line.strip() != b"ready":
            raise RuntimeError("zygote failed to start")

    def run(self, req: dict, timeout: float) -> int | None:
        """Forks a child for `req`; returns its exit code, -1 if it died without reporting,
        or None on timeout (the child is killed)."""
        payload = pickle.dumps(req, protocol=pickle.HIGHEST_PROTOCOL)
        conn = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        try:
            conn.connect(self.socket_path)
            conn.sendall(struct.pack("<I", len(payload)) + payload)
            head = _recv_exact(conn, 4)
            if len(head) < 4:
                raise RuntimeError("zygote closed the connection before forking")
            (pid,) = struct.unpack("<i", head)
            conn.settimeout(timeout)
            try:
                tail = _recv_exact(conn, 4)
            except socket.timeout:
                _kill_group(pid)
                return None
            _kill_group(pid)   # whatever the script left running goes with the request (the reference deletes the pod)
            if len(tail) < 4:
                return -1
            return struct.unpack("<i", tail)[0]
        finally:
            conn.close()

    def close(self) -> None:
        try:
            payload = pickle.dumps({"op": "shutdown"})
            conn = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            conn.settimeout(1.0)
            conn.connect(self.socket_path)
            conn.sendall(struct.pack("<I", len(payload)) + payload)
            conn.close()
        except OSError:
            pass
        try:
            self._proc.wait(timeout=3)
        except Exception:
            self._proc.kill()
        try:
            os.rmdir(self._dir)
        except OSError:
            pass


if __name__ == "__main__":
    serve(sys.argv[1])
