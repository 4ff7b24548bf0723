"""LocalGpuExecutor — drop-in for the executor-backend duck type (boundary i).

Mirrors `KubernetesCodeExecutor.execute(source_code, files, env) -> Result`
(services/kubernetes_code_executor.py:47-52,75-149), the object `create_http_server`
(services/http_server.py:77-106) and `CustomToolExecutor` (custom_tool_executor.py:45-46)
consume, injected at `ApplicationContext.code_executor` (application_context.py:72-84).

Per request, in the reference's order (kubernetes_code_executor.py:95-149):
  1. fresh workspace (the "fresh pod"), files of the `files` map restored into it from the
     object store (upload_file, :100-113; PUT semantics of executor/server.rs:69-88);
  2. the script runs as its own OS process with stdout/stderr captured, the request env
     applied, a timeout (executor/server.rs:149-169: timeout -> ("", "Execution timed out",
     -1), killed by signal -> -1).  The process is a fork of a warm zygote
     (zygote.py, script_runner.py) whose numpy calls reach the GPU through the broker (broker.py);
  3. top-level regular files whose ctime is after the request start are the changed files
     (executor/server.rs:98-118), stored through the object store and returned as
     {"/workspace/<name>": <object id>} (download_file, :126-142).
Script failures are results (exit_code != 0 + stderr), not exceptions; backend failures
raise and become HTTP 500 in the reference's route (http_server.py:102-104).
"""
from __future__ import annotations

import asyncio
import os
import re
import shutil
import tempfile
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from pathlib import Path
from typing import Mapping

WORKSPACE_PREFIX = "/workspace/"
_SCRATCH_DIR = os.environ.get("BCI_TMPDIR") or ("/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None)
_ABS_PATH_RE = re.compile(r"^/[^/].*$")            # utils/validation.py:20-22
_HASH_RE = re.compile(r"^[0-9a-zA-Z_-]{1,255}$")   # utils/validation.py:19


@dataclass
class Result:
    """kubernetes_code_executor.py:47-52"""
    stdout: str
    stderr: str
    exit_code: int
    files: Mapping[str, str] = field(default_factory=dict)


class LocalGpuExecutor:
    Result = Result

    def __init__(self, file_storage=None, *, max_inflight: int = 16, timeout: float = 60.0,
                 broker=None, device_module=None, storage_path: str | None = None,
                 min_rand_elems: int | None = None, zygotes: int | None = None, device: int | None = None,
                 retry_attempts: int = 3, retry_wait: tuple[float, float] = (0.1, 2.0), zygote_pool: int | None = None,
                 pin_zygotes: bool | None = None):
        # max_inflight = requests admitted at once (the rest wait on a semaphore).  16 per GPU is the
        # measured optimum for short scripts: the broker needs ~2 dispatch groups of work to keep the
        # GPU busy, and every additional live script process lengthens the fork / set-up phase of
        # all the others (tools/e2e_trace.py: 16 admitted -> 3.4 k req/s, 128 -> 2.35 k on one B200).
        # Raise it for scripts that spend their time off the GPU.
        # The GPU is owned by a broker: an in-process GpuBroker when a device module is injected
        # (tests), a given broker (several executors can share one per GPU), or — the default —
        # a dedicated broker PROCESS, so its threads do not share this process's GIL.
        self._own_broker = broker is None
        if broker is None:
            if device_module is not None:
                from .broker import GpuBroker
                broker = GpuBroker(device_module)
            else:
                from .broker import BrokerProcess
                broker = BrokerProcess(device)
        self.broker = broker
        self._tmp_storage = None
        if file_storage is None:
            from .storage import ContentAddressedStorage, HashService, RemoteHashService
            if storage_path is None:
                self._tmp_storage = tempfile.mkdtemp(prefix="bci-objects-")
                storage_path = self._tmp_storage
            hasher = HashService(device_module) if device_module is not None else RemoteHashService(broker.socket_path)
            file_storage = ContentAddressedStorage(storage_path, hasher)
        self.file_storage = file_storage
        self.timeout = timeout
        self.max_inflight = max_inflight
        self.min_rand_elems = min_rand_elems
        self._pool = ThreadPoolExecutor(max_workers=max_inflight, thread_name_prefix="bci-exec")
        self._sem: asyncio.Semaphore | None = None
        self._sem_loop = None
        n_zyg = zygotes if zygotes is not None else max(1, min(8, (os.cpu_count() or 2) // 8))
        from .zygote import ZygoteClient
        if pin_zygotes is None:
            pin_zygotes = os.environ.get("BCI_PIN_ZYGOTES", "1") != "0"
        # Pinned (default): zygote k of this process and the first milliseconds of each of its children run on CPU
        # (pid * 7 + k) mod ncpus — several service processes spread out.  A forked child placed on another core starts
        # with cold caches; measured on the 16-CPU-quota box (profiles/r2k_*): scripts that do nothing 3.55 k -> 4.51 k
        # req/s, p99 30 -> 12 ms; the metric script 2.68 k -> 2.98 k.  The child widens its affinity again before the
        # agent's code runs (script_runner.py), so user code is never confined.
        self._zygotes = [ZygoteClient(pool=zygote_pool, pin_cpu=(os.getpid() * 7 + k) if pin_zygotes else None) for k in range(n_zyg)]
        self._rr = 0
        # backend failures (a warm interpreter or the broker died) are retried like the reference retries
        # RuntimeError (kubernetes_code_executor.py:75-79: 3 attempts, exponential wait) after the dead part
        # has been started again; a zygote comes back in ~0.3 s, so the waits are shorter than a pod's
        self.retry_attempts, self.retry_wait = max(1, retry_attempts), retry_wait
        self._heal_lock = threading.Lock()

    # ---- public API ---------------------------------------------------------------------
    async def execute(self, source_code: str, files: Mapping[str, str] = {}, env: Mapping[str, str] = {}) -> Result:
        _validate(source_code, files, env)
        attempt = 0
        while True:
            try:
                return await self._execute_once(source_code, files, env)
            except RuntimeError:
                attempt += 1
                if attempt >= self.retry_attempts:
                    raise
                await asyncio.sleep(min(self.retry_wait[1], self.retry_wait[0] * 2 ** (attempt - 1)))
                await asyncio.get_running_loop().run_in_executor(None, self.heal)

    def heal(self) -> dict:
        """Starts again whatever died: warm interpreters, and the broker process when this executor owns it."""
        done = {"zygotes": 0, "broker": 0}
        with self._heal_lock:
            for z in self._zygotes:
                if not z.alive():
                    z.restart()
                    done["zygotes"] += 1
            alive = getattr(self.broker, "alive", None)
            if self._own_broker and alive is not None and not alive():
                self.broker.restart()
                done["broker"] += 1
        return done

    async def _execute_once(self, source_code: str, files: Mapping[str, str], env: Mapping[str, str]) -> Result:
        loop = asyncio.get_running_loop()
        if self._sem is None or self._sem_loop is not loop:
            self._sem, self._sem_loop = asyncio.Semaphore(self.max_inflight), loop
        async with self._sem:  # admission: at most `max_inflight` requests resident on the GPU
            # one directory per request on tmpfs when there is one: ~12 metadata operations per
            # request at thousands of requests/s is what a journalled /tmp serialises on
            base = tempfile.mkdtemp(prefix="bci-req-", dir=_SCRATCH_DIR)
            workspace = Path(base, "workspace")
            os.mkdir(workspace)
            try:
                for file_path, object_id in files.items():
                    target = _workspace_target(workspace, file_path)
                    target.parent.mkdir(parents=True, exist_ok=True)
                    target.write_bytes(await self.file_storage.read(object_id))
                start_ns = time.time_ns()
                before = _snapshot(workspace) if files else {}
                stdout, stderr, exit_code = await loop.run_in_executor(
                    self._pool, self._run_process, source_code, str(workspace), dict(env), Path(base))
                stored = {}
                names = _changed_files(workspace, start_ns, before)
                if names:
                    ids = await asyncio.gather(*(self._store(workspace / n) for n in names))
                    stored = {WORKSPACE_PREFIX + n: i for n, i in zip(names, ids)}
                return Result(stdout=stdout, stderr=stderr, exit_code=exit_code, files=stored)
            finally:
                _remove_request_dir(base, str(workspace))

    def close(self) -> None:
        self._pool.shutdown(wait=False, cancel_futures=True)
        hasher = getattr(self.file_storage, "_hasher", None)
        if hasher is not None:
            hasher.close()
        for z in self._zygotes:
            z.close()
        if self._own_broker:
            self.broker.close()
        if self._tmp_storage:
            shutil.rmtree(self._tmp_storage, ignore_errors=True)

    # ---- internals -------------------------------------------------------------------------
    async def _store(self, path: Path) -> str:
        put_file = getattr(self.file_storage, "put_file", None)
        if put_file is not None:
            return await put_file(str(path))
        async with self.file_storage.writer() as w:  # any object store with the reference's protocol
            with open(path, "rb") as f:
                while True:
                    chunk = f.read(1 << 20)
                    if not chunk:
                        break
                    await w.write(chunk)
        return w.hash

    def _run_process(self, source_code: str, workspace: str, env: dict, capture: Path):
        # stdout / stderr are captured in two anonymous memory files whose descriptors travel with the
        # request; the script's source file is written by the request's own process
        out_fd, err_fd = _capture_fd("stdout"), _capture_fd("stderr")
        try:
            req = {"source_code": source_code, "workspace": workspace, "env": env,
                   "broker_socket": self.broker.socket_path, "min_rand_elems": self.min_rand_elems,
                   "script_path": str(capture / "script.py"), "write_script": True}
            alive = getattr(self.broker, "alive", None)
            if self._own_broker and alive is not None and not alive():
                raise RuntimeError("the GPU broker process is gone")    # retried after heal(): scripts must not run on the host silently
            self._rr = (self._rr + 1) % len(self._zygotes)
            code = self._zygotes[self._rr].run(req, self.timeout, out_fd, err_fd)
            if code is None:
                return "", "Execution timed out", -1   # executor/server.rs:169
            if code < 0:
                code = -1                               # died on a signal (executor/server.rs:165)
            return _read_fd(out_fd), _read_fd(err_fd), code
        finally:
            os.close(out_fd)
            os.close(err_fd)


def _remove_request_dir(base: str, workspace: str) -> None:
    """The common case by hand (a flat workspace + the script file): a generic tree walk costs several times the
    syscalls and this runs on the service's event-loop thread for every request."""
    try:
        with os.scandir(workspace) as it:
            for entry in it:
                if entry.is_dir(follow_symlinks=False):
                    raise OSError("nested")
                os.unlink(entry.path)
        os.rmdir(workspace)
        try:
            os.unlink(os.path.join(base, "script.py"))
        except FileNotFoundError:
            pass
        os.rmdir(base)
    except OSError:
        shutil.rmtree(base, ignore_errors=True)


def _capture_fd(name: str) -> int:
    try:
        return os.memfd_create("bci-" + name, os.MFD_CLOEXEC)
    except (AttributeError, OSError):
        fd, path = tempfile.mkstemp(prefix="bci-" + name + "-", dir=_SCRATCH_DIR)
        os.unlink(path)
        return fd


def _read_fd(fd: int) -> str:
    size = os.fstat(fd).st_size
    data = os.pread(fd, size, 0) if size else b""
    return data.decode("utf-8", "replace")               # lossy like server.rs:163-164


def _read_text(path: str) -> str:
    try:
        with open(path, "rb") as f:
            return f.read().decode("utf-8", "replace")  # lossy like server.rs:163-164
    except FileNotFoundError:
        return ""


def _workspace_target(workspace: Path, file_path: str) -> Path:
    """Where an entry of the request's `files` map is restored.  The reference builds the pod URL
    `/workspace/` + path-without-prefix and server.rs joins STRINGS (`{workspace}/{path}`,
    server.rs:73-74), so "/etc/passwd" lands in /workspace/etc/passwd.  Here the workspace is a
    directory of the host: the same mapping, computed without ever leaving it — `..` components are
    refused (the pod was the sandbox there)."""
    rel = file_path.removeprefix(WORKSPACE_PREFIX).lstrip("/")
    parts = [p for p in rel.split("/") if p not in ("", ".")]
    if not parts or ".." in parts:
        raise ValueError(f"files key {file_path!r} does not name a file inside the workspace")
    return workspace.joinpath(*parts)


def _snapshot(workspace: Path) -> dict:
    """Top-level regular files with what a change cannot leave untouched."""
    snap = {}
    for entry in os.scandir(workspace):
        if entry.is_file():                         # Path::is_file follows links (server.rs:103) ...
            st = entry.stat(follow_symlinks=False)   # ... DirEntry::metadata does not (server.rs:106)
            snap[entry.name] = (st.st_ino, st.st_size, st.st_mtime_ns, st.st_ctime_ns)
    return snap


def _changed_files(workspace: Path, since_ns: int, before: dict | None = None) -> list[str]:
    """server.rs:98-118: top-level regular files whose ctime is later than the start of the request.
    The reference's script takes >= 80 ms to start, ours ~2 ms — inside one tick of the coarse clock
    the kernel stamps ctime with — so a file written right away could carry a ctime just BEFORE
    `since_ns`.  Hence also: anything new or different from the snapshot taken at the start."""
    out = []
    for entry in os.scandir(workspace):
        if not entry.is_file():       # like the reference, a link to a regular file counts (its CONTENT is stored, never the link)
            continue
        st = entry.stat(follow_symlinks=False)
        if st.st_ctime_ns > since_ns or (before is not None and
                                         before.get(entry.name) != (st.st_ino, st.st_size, st.st_mtime_ns, st.st_ctime_ns)):
            out.append(entry.name)
    return sorted(out)


def _validate(source_code, files, env) -> None:
    """Same constraints pydantic enforces in the reference (@validate_call,
    kubernetes_code_executor.py:80; AbsolutePath / Hash, utils/validation.py:19-22)."""
    if not isinstance(source_code, str):
        raise ValueError("source_code must be a string")
    for path, object_id in files.items():
        if not isinstance(path, str) or not _ABS_PATH_RE.match(path):
            raise ValueError(f"files key {path!r} is not an absolute path")
        if not isinstance(object_id, str) or not _HASH_RE.match(object_id):
            raise ValueError(f"files value {object_id!r} is not a valid object id")
    for k, v in env.items():
        if not isinstance(k, str) or not isinstance(v, str):
            raise ValueError("env must map str to str")
