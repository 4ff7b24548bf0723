"""Warm-interpreter zygote: one fresh OS process per `/v1/execute` request, forked ahead of time.

The reference keeps a queue of warm pods and spends one per request
(kubernetes_code_executor.py:151-189,248-264); inside the pod each request still pays a fresh
interpreter (`xonsh script.xsh`, executor/server.rs:152).  Here the warm resource is a Python
process with numpy, the proxy layer and the import hook already imported, and — like the
reference's pod QUEUE — a few children of it that were forked BEFORE any request arrived and
sit blocked on a socket pair.  A request is handed to one of them (its connection and capture
descriptors travel as SCM_RIGHTS), and the replacement is forked after the hand-off, off the
request's critical path.  The isolation unit stays the reference's: one OS process per request,
never reused.

The zygote is single-threaded and holds no CUDA context, so fork() is safe.  Protocol (Unix
stream socket, one connection per request):
    client -> zygote : 4-byte length + request as a JSON object (ZygoteClient and
                       csrc/bci_executor_server.cpp speak the same form; nothing is unpickled:
                       any local process of the same user can reach this socket)
                       + optionally two descriptors (SCM_RIGHTS): where stdout / stderr go
    child  -> client : 4-byte pid, first thing              (the zygote closes its copy)
    child  -> client : 4-byte exit code, just before os._exit  (absent if the child was killed
                       or crashed: the client maps EOF to -1, like server.rs:165)
Children are auto-reaped (SIGCHLD ignored).
"""
from __future__ import annotations

import json
import os
import signal
import socket
import struct
import subprocess
import sys
import tempfile
import time

_ENV_ALLOW = {"PATH", "HOME", "USER", "LOGNAME", "SHELL", "LANG", "LANGUAGE", "TZ", "TMPDIR", "TERM", "PWD", "VIRTUAL_ENV",
              "LD_LIBRARY_PATH", "PYTHONPATH", "PYTHONHOME", "PYTHONHASHSEED", "PYTHONIOENCODING", "PYTHONUNBUFFERED",
              "PYTHONDONTWRITEBYTECODE", "MPLBACKEND", "MPLCONFIGDIR", "XDG_CACHE_HOME", "XDG_CONFIG_HOME"}
_ENV_ALLOW_PREFIXES = ("LC_", "BCI_", "OMP_", "OPENBLAS_", "MKL_", "NUMEXPR_")
_TRACE = bool(os.environ.get("BCI_TRACE"))
_YIELD = bool(os.environ.get("BCI_ZYGOTE_YIELD"))
PIN_ENV = "BCI_ZYGOTE_PIN"
POOL_ENV = "BCI_ZYGOTE_POOL"      # children forked ahead of requests (0 = fork on arrival)
MAX_REQUEST = 64 << 20


def _recv_exact(conn: socket.socket, n: int, fds: list | None = None) -> bytes:
    buf = b""
    while len(buf) < n:
        if fds is not None:
            chunk, got, _f, _a = socket.recv_fds(conn, n - len(buf), 4)
            fds.extend(got)
        else:
            chunk = conn.recv(n - len(buf))
        if not chunk:
            return buf
        buf += chunk
    return buf


def _send_with_fds(sock: socket.socket, data: bytes, fds: list[int]) -> None:
    sent = socket.send_fds(sock, [data], fds) if fds else sock.send(data)
    if sent < len(data):
        sock.sendall(data[sent:])


def _serve_request(conn: socket.socket, payload: bytes, out_fds: list[int]) -> None:
    """Body of the per-request process: never returns."""
    from . import script_runner

    code = 1
    try:
        # the CHILD announces its pid: if the parent did, a short script could finish and
        # write its exit code first, and the client would read the two words swapped
        t_child = time.time() if _TRACE else 0.0
        os.setsid()   # own process group: a timeout (and the end of the request) reaps what the script spawned
        conn.sendall(struct.pack("<i", os.getpid()))
        req = json.loads(payload)
        if not isinstance(req, dict):
            raise ValueError("request is not an object")
        if _TRACE:
            tz = req.get("_t_zygote") or [0.0, 0.0]
            conn.sendall(struct.pack("<3d", tz[0], tz[1], t_child))
        code = script_runner.run_in_child(req, out_fds[0] if len(out_fds) >= 2 else None,
                                          out_fds[1] if len(out_fds) >= 2 else None)
    finally:
        try:
            conn.sendall(struct.pack("<i", code))
        except OSError:
            pass
        os._exit(code & 0xFF)


class _Idle:
    """A child forked ahead of time, blocked on its end of a socket pair."""

    __slots__ = ("sock", "pid")

    def __init__(self, sock, pid):
        self.sock, self.pid = sock, pid


def _spawn_idle(srv: socket.socket, idle: list) -> "_Idle | None":
    a, b = socket.socketpair(socket.AF_UNIX, socket.SOCK_STREAM)
    try:
        pid = os.fork()
    except OSError:
        a.close()
        b.close()
        return None
    if pid:
        b.close()
        return _Idle(a, pid)
    # ---- pre-forked child: wait for a request (or for the zygote to go away) -------------------
    try:
        srv.close()
        a.close()
        for other in idle:            # siblings' hand-off sockets: keeping them open would hide the zygote's death
            other.sock.close()
        fds: list[int] = []
        head = _recv_exact(b, 4, fds)
        if len(head) < 4:
            os._exit(0)               # the zygote exited
        (length,) = struct.unpack("<I", head)
        payload = _recv_exact(b, length, fds)
        b.close()
        if len(payload) < length or not fds:
            os._exit(0)
        conn = socket.socket(fileno=fds[0])
        _serve_request(conn, payload, fds[1:])
    finally:
        os._exit(1)


def serve(socket_path: str) -> None:
    import gc

    from . import script_runner  # imports numpy + proxy + hook once

    # Everything imported so far lives as long as the zygote: move it to the permanent generation and stop
    # collecting here.  Without this (measured, profiles/r2f_*): the cyclic collector's generation counters are
    # inherited by every child, a few percent of the requests start with a FULL collection over ~10^5 module
    # objects (30-60 ms, p99 of the request latency) — and each such pass dirties the pages it walks, i.e.
    # copy-on-write faults on top.  Children re-enable the collector for the script's own objects.
    if not os.environ.get("BCI_ZYGOTE_NOFREEZE"):
        gc.collect()
        gc.freeze()
        gc.disable()
    # BCI_ZYGOTE_PIN=<cpu>: this zygote and the first milliseconds of every child run on ONE CPU — a forked child
    # placed on another core starts with cold caches for the whole interpreter (measured: 1.7 ms of user time for
    # a `pass` script); the child widens its affinity again before the agent's code runs (script_runner).
    pin = os.environ.get(PIN_ENV)
    if pin not in (None, ""):
        try:
            cpus = set(os.sched_getaffinity(0))
            os.sched_setaffinity(0, {sorted(cpus)[int(pin) % len(cpus)]})
            script_runner.ALL_CPUS = cpus
        except (OSError, ValueError, AttributeError):
            pass
    signal.signal(signal.SIGCHLD, signal.SIG_IGN)
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    try:
        os.unlink(socket_path)
    except OSError:
        pass
    srv.bind(socket_path)
    srv.listen(1024)
    pool_size = max(0, int(os.environ.get(POOL_ENV, "4") or 0))
    idle: list[_Idle] = []
    for _ in range(pool_size):
        ch = _spawn_idle(srv, idle)
        if ch is not None:
            idle.append(ch)
    sys.stdout.write("ready\n")
    sys.stdout.flush()
    script_runner.prepare_zygote_stdio()
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
        fds: list[int] = []
        t_accept = time.time() if _TRACE else 0.0
        try:
            conn.settimeout(30.0)
            head = _recv_exact(conn, 4, fds)
            if len(head) < 4:
                continue
            (length,) = struct.unpack("<I", head)
            if length > MAX_REQUEST:
                continue
            payload = _recv_exact(conn, length, fds)
            if len(payload) < length:
                continue
            if payload[:1] != b"{":
                continue                    # only JSON objects are requests
            if length < 64 and json.loads(payload).get("op") == "shutdown":
                break
            if _TRACE and payload.endswith(b"}"):      # tools/e2e_cost.py: when did the zygote see this request
                payload = payload[:-1] + (',"_t_zygote":[%r,%r]}' % (t_accept, time.time())).encode()
                length = len(payload)
            conn.settimeout(None)
            handed = False
            while idle and not handed:
                ch = idle.pop(0)
                try:
                    _send_with_fds(ch.sock, struct.pack("<I", length) + payload, [conn.fileno()] + fds)
                    handed = True
                except OSError:
                    pass                    # that child is gone: try the next one
                finally:
                    ch.sock.close()
            if not handed:
                pid = os.fork()
                if pid == 0:
                    srv.close()
                    for other in idle:
                        other.sock.close()
                    _serve_request(conn, payload, fds)
        except Exception:
            pass
        finally:
            conn.close()
            for f in fds:
                try:
                    os.close(f)
                except OSError:
                    pass
        if _YIELD:
            os.sched_yield()                # let the request's process run before this loop goes on forking
        while len(idle) < pool_size:        # the replacement is forked AFTER the hand-off
            ch = _spawn_idle(srv, idle)
            if ch is None:
                break
            idle.append(ch)
    for ch in idle:
        ch.sock.close()                     # EOF: the waiting children exit
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


class ZygoteDied(RuntimeError):
    """The warm interpreter is gone (RuntimeError: the class the reference's retry policy acts on,
    kubernetes_code_executor.py:75-79)."""


class ZygoteClient:
    """Service-side handle: starts the zygote process and submits requests to it."""

    def __init__(self, pool: int | None = None, pin_cpu: int | None = None):
        self._pin = pin_cpu
        self._dir = tempfile.mkdtemp(prefix="bci-zygote-")
        self.socket_path = os.path.join(self._dir, "zygote.sock")
        self._pool = pool
        self._proc = None
        self.trace = [] if os.environ.get("BCI_TRACE") else None      # tools/e2e_cost.py: per-request phase times
        self._start()

    def _start(self) -> None:
        # the scripts inherit the zygote's environment: start it from an allow-list, not from the service's
        # own environment (tokens, cloud credentials ... have no business in agent code; a pod of the
        # reference only sees the executor image's environment plus the request's `env`)
        env = {k: v for k, v in os.environ.items() if k in _ENV_ALLOW or k.startswith(_ENV_ALLOW_PREFIXES)}
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        env["PYTHONPATH"] = root + (os.pathsep + env["PYTHONPATH"] if env.get("PYTHONPATH") else "")
        # no BLAS/OpenMP worker threads in the zygote: fork() must copy a single-threaded
        # process, and a sandboxed request gets one host thread (heavy array work goes to the GPU)
        for var in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
            env.setdefault(var, "1")
        if self._pool is not None:
            env[POOL_ENV] = str(self._pool)
        if self._pin is not None:
            env[PIN_ENV] = str(self._pin)
        self._proc = subprocess.Popen([sys.executable, "-m", "bee_code_interpreter_b200.zygote", self.socket_path],
                                      stdout=subprocess.PIPE, env=env, cwd=root)
        line = self._proc.stdout.readline()
        if line.strip() != b"ready":
            raise ZygoteDied("zygote failed to start")

    def alive(self) -> bool:
        return self._proc is not None and self._proc.poll() is None

    def restart(self) -> None:
        """A dead zygote is replaced (the reference respawns pods that failed, kubernetes_code_executor.py:176-182)."""
        if self.alive():
            return
        try:
            os.unlink(self.socket_path)
        except OSError:
            pass
        self._start()

    def run(self, req: dict, timeout: float, stdout_fd: int | None = None, stderr_fd: int | None = None) -> int | None:
        """Hands `req` to a fresh child; returns its exit code, -1 if it died without reporting,
        or None on timeout (the child is killed).  Raises ZygoteDied when no child could be had."""
        payload = json.dumps(req).encode()
        conn = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        trace = self.trace
        t0 = time.perf_counter() if trace is not None else 0.0
        t0_wall = time.time() if trace is not None else 0.0
        try:
            try:
                conn.connect(self.socket_path)
                fds = [stdout_fd, stderr_fd] if stdout_fd is not None and stderr_fd is not None else []
                _send_with_fds(conn, struct.pack("<I", len(payload)) + payload, fds)
                head = _recv_exact(conn, 4)
            except OSError as e:
                raise ZygoteDied(f"zygote unreachable: {e}") from None
            if len(head) < 4:
                raise ZygoteDied("zygote closed the connection before a child took the request")
            (pid,) = struct.unpack("<i", head)
            t1 = time.perf_counter() if trace is not None else 0.0
            stamps = struct.unpack("<3d", _recv_exact(conn, 24)) if trace is not None else None
            conn.settimeout(timeout)
            try:
                tail = _recv_exact(conn, 4)
            except socket.timeout:
                _kill_group(pid)
                return None
            if trace is not None:
                # (submit -> child announced itself, child -> exit code, wall stamps: submit, zygote accept, zygote hand-off, child start)
                trace.append((t1 - t0, time.perf_counter() - t1, t0_wall) + stamps)
            _kill_group(pid)   # whatever the script left running goes with the request (the reference deletes the pod)
            if len(tail) < 4:
                return -1
            return struct.unpack("<i", tail)[0]
        finally:
            conn.close()

    def close(self) -> None:
        try:
            payload = json.dumps({"op": "shutdown"}).encode()
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
