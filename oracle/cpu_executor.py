"""oracle/cpu_executor.py — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's executor for one `/v1/execute` request, i.e. what a
request costs and returns on the reference's own code path:

  services/kubernetes_code_executor.py:95-149  upload files -> POST /execute -> download
                                               changed files -> Result
  executor/server.rs:69-96                     PUT/GET /workspace/{path}
  executor/server.rs:98-118                    changed files = top-level regular files of
                                               the workspace with ctime > request start
  executor/server.rs:120-179                   script -> subprocess with timeout ->
                                               (stdout, stderr, exit_code); timeout =>
                                               ("", "Execution timed out", -1); lossy UTF-8
  services/storage.py:44-90                    object store with random 64-hex ids

Differences forced by this image (all favour the baseline's speed): no Kubernetes pod (a
fresh temp workspace per request instead), no `upm guess`/`pip install`, plain
`python script.py` instead of `xonsh script.xsh` (xonsh/upm/kubectl are absent).  The
script runs with stock numpy on the host CPU and with a restatement of the reference's
import hook (oracle/ref_hook/sitecustomize.py) on PYTHONPATH.
"""
from __future__ import annotations

import asyncio
import os
import secrets
import shutil
import subprocess
import sys
import tempfile
import time
from dataclasses import dataclass, field
from pathlib import Path
from typing import Mapping

REF_HOOK_DIR = Path(__file__).resolve().parent / "ref_hook"
WORKSPACE_PREFIX = "/workspace/"


class RefStorage:
    """services/storage.py:34-90 restated synchronously: flat directory, object id =
    secrets.token_hex(32) chosen before the first byte is written (storage.py:52)."""

    def __init__(self, storage_path: str | os.PathLike):
        self.storage_path = Path(storage_path)

    def write(self, data: bytes) -> str:
        self.storage_path.mkdir(parents=True, exist_ok=True)
        object_id = secrets.token_hex(32)
        (self.storage_path / object_id).write_bytes(data)
        return object_id

    def read(self, object_id: str) -> bytes:
        target = self.storage_path / object_id
        if not object_id or not target.exists():
            raise FileNotFoundError(f"File not found: {object_id}")
        return target.read_bytes()

    def exists(self, object_id: str) -> bool:
        return (self.storage_path / object_id).exists()


@dataclass
class Result:
    """kubernetes_code_executor.py:47-52"""
    stdout: str
    stderr: str
    exit_code: int
    files: Mapping[str, str] = field(default_factory=dict)


def changed_files(workspace: Path, since_ns: int) -> list[str]:
    """server.rs:98-118 — non-recursive, regular files, ctime strictly after `since`."""
    out = []
    for entry in os.scandir(workspace):
        if not entry.is_file():
            continue
        if entry.stat().st_ctime_ns > since_ns:
            out.append(entry.name)
    return out


def run_script(source_code: str, workspace: Path, env: Mapping[str, str] | None, timeout: float = 60.0,
               pythonpath: str | None = None) -> tuple[str, str, int, list[str]]:
    """server.rs:120-179 for one request against an existing workspace directory."""
    start_ns = time.time_ns()
    with tempfile.TemporaryDirectory() as source_dir:
        script = Path(source_dir) / "script.py"
        script.write_text(source_code)
        child_env = dict(os.environ)
        hook = str(REF_HOOK_DIR) if pythonpath is None else pythonpath
        child_env["PYTHONPATH"] = hook + (os.pathsep + child_env["PYTHONPATH"] if child_env.get("PYTHONPATH") else "")
        if env:
            child_env.update(env)
        try:
            proc = subprocess.run([sys.executable, str(script)], cwd=workspace, env=child_env,
                                  capture_output=True, timeout=timeout)
            stdout = proc.stdout.decode("utf-8", "replace")
            stderr = proc.stderr.decode("utf-8", "replace")
            exit_code = proc.returncode if proc.returncode >= 0 else -1  # killed by signal -> -1 (server.rs:165)
        except subprocess.TimeoutExpired:
            stdout, stderr, exit_code = "", "Execution timed out", -1
    return stdout, stderr, exit_code, changed_files(workspace, start_ns)


class LocalCpuExecutor:
    """Duck type of KubernetesCodeExecutor (kubernetes_code_executor.py:81-86): the CPU
    baseline backend and the stand-in the golden HTTP fixtures were recorded with."""

    Result = Result

    def __init__(self, file_storage, timeout: float = 60.0):
        self.file_storage = file_storage  # RefStorage, or the reference's async Storage
        self.timeout = timeout

    async def _read(self, object_id: str) -> bytes:
        r = self.file_storage.read(object_id)
        return await r if asyncio.iscoroutine(r) else r

    async def _write(self, data: bytes) -> str:
        r = self.file_storage.write(data)
        return await r if asyncio.iscoroutine(r) else r

    async def execute(self, source_code: str, files: Mapping[str, str] = {}, env: Mapping[str, str] = {}) -> Result:
        workspace = Path(tempfile.mkdtemp(prefix="bci-cpu-ws-"))
        try:
            for file_path, object_id in files.items():  # kubernetes_code_executor.py:100-113
                target = workspace / file_path.removeprefix(WORKSPACE_PREFIX)
                target.parent.mkdir(parents=True, exist_ok=True)  # server.rs:75-76
                target.write_bytes(await self._read(object_id))
            stdout, stderr, exit_code, changed = await asyncio.to_thread(
                run_script, source_code, workspace, env, self.timeout)
            stored = {}
            for name in changed:  # kubernetes_code_executor.py:126-142
                stored[WORKSPACE_PREFIX + name] = await self._write((workspace / name).read_bytes())
            return Result(stdout=stdout, stderr=stderr, exit_code=exit_code, files=stored)
        finally:
            shutil.rmtree(workspace, ignore_errors=True)
