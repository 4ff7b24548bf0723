"""Child-process body for one `/v1/execute` request: what `xonsh script.xsh` is in the
reference (executor/server.rs:149-169), minus the interpreter start-up.

Runs inside a fork of the warm zygote (zygote.py): cwd = request workspace, env = service env
+ request env, fds 1/2 -> capture files, builtins.__import__ = numpy_hook.patched_import
(boundary ii), then the agent source runs as `__main__`.  The exit status follows CPython's
own rules for scripts (uncaught exception -> traceback on stderr, status 1; sys.exit(n) -> n).
"""
from __future__ import annotations

import builtins
import os
import sys
import traceback

import numpy  # noqa: F401  (preloaded in the zygote: forked children share these pages)

from . import broker_client, numpy_hook, numpy_proxy  # noqa: F401  (everything a request touches is imported HERE,
#                                                                   in the zygote: a child must not pay imports)

numpy_hook.numpy_view()  # build the agent-facing numpy view once; children inherit it

ALL_CPUS: set = set()    # set by a zygote that pinned itself to one CPU: the mask children widen back to


def _finish_like_an_interpreter() -> None:
    """What Py_Finalize does before a script's process ends and os._exit() skips: wait for the
    script's non-daemon threads, then run the atexit callbacks.  Both are skipped when there is nothing
    to do (the usual case): every Python-level step of a freshly forked process runs on cold caches and
    copy-on-write pages, so the request pays ~0.4 ms for the empty versions."""
    import atexit
    import threading

    if threading.active_count() > 1:
        threading._shutdown()
    if atexit._ncallbacks():
        atexit._run_exitfuncs()


def prepare_zygote_stdio() -> None:
    """Called once by the zygote: fd 0 becomes /dev/null (the reference's `Command::output()` child has no
    stdin either, server.rs:153-156) and sys.stdin / stdout / stderr become line-buffered text wrappers over
    fds 0 / 1 / 2.  A child then only dup2()s its capture descriptors onto 1 and 2: building three
    TextIOWrappers per request cost ~0.3 ms of cold-start time."""
    null_fd = os.open(os.devnull, os.O_RDONLY)
    os.dup2(null_fd, 0)
    os.close(null_fd)
    sys.stdin = os.fdopen(0, "r", closefd=False)
    sys.stdout.flush()
    sys.stderr.flush()
    sys.stdout = os.fdopen(1, "w", buffering=1, closefd=False)
    sys.stderr = os.fdopen(2, "w", buffering=1, closefd=False)
    global _STDIO_READY
    _STDIO_READY = True


_STDIO_READY = False


def run_in_child(req: dict, stdout_fd: int | None = None, stderr_fd: int | None = None) -> int:
    """`stdout_fd` / `stderr_fd`: capture descriptors the service passed along with the request (anonymous
    memory files); without them the request names capture FILES (`stdout_path`, `stderr_path`: the native
    executor server's form)."""
    os.chdir(req["workspace"])
    env = req.get("env")
    if env:
        os.environ.update(env)
    if req.get("broker_socket"):
        os.environ["BCI_BROKER_SOCKET"] = req["broker_socket"]
    # stdin is /dev/null, as for the reference's `Command::output()` child (server.rs:153-156): a script
    # that calls input() gets EOFError, it must never read the service's terminal
    if not _STDIO_READY:
        null_fd = os.open(os.devnull, os.O_RDONLY)
        os.dup2(null_fd, 0)
        os.close(null_fd)
        sys.stdin = os.fdopen(0, "r", closefd=False)
    out_fd = stdout_fd if stdout_fd is not None else os.open(req["stdout_path"], os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o600)
    err_fd = stderr_fd if stderr_fd is not None else os.open(req["stderr_path"], os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o600)
    os.dup2(out_fd, 1)
    os.dup2(err_fd, 2)
    os.close(out_fd)
    os.close(err_fd)
    if not _STDIO_READY:
        sys.stdout = os.fdopen(1, "w", buffering=1, closefd=False)
        sys.stderr = os.fdopen(2, "w", buffering=1, closefd=False)
    # the service wrote the source next to the capture files (server.rs:124 writes script.py into a
    # temp dir): tracebacks quote the failing line and inspect / multiprocessing find a real file
    script = req.get("script_path") or "script.py"
    if req.get("write_script"):       # the request's own process writes its source file (one less job for the service)
        try:
            with open(script, "w", encoding="utf-8", errors="surrogatepass") as f:
                f.write(req["source_code"])
        except OSError:
            pass
    sys.argv = [script]
    numpy_hook.reset_request_state()
    numpy_hook.install(min_rand_elems=req.get("min_rand_elems"))
    import gc

    gc.enable()      # the zygote runs with the collector off (its objects are frozen); the script gets a normal one
    if ALL_CPUS:                # a pinned zygote (zygote.py): the agent's code may use every CPU the service could
        try:
            os.sched_setaffinity(0, ALL_CPUS)
        except OSError:
            pass
    code = 0
    try:
        compiled = compile(req["source_code"], script, "exec")
        exec(compiled, {"__name__": "__main__", "__builtins__": builtins, "__file__": script})
    except SystemExit as e:
        if e.code is None:
            code = 0
        elif isinstance(e.code, int):
            code = e.code & 0xFF
        else:
            print(e.code, file=sys.stderr)
            code = 1
    except BaseException:
        traceback.print_exc()
        code = 1
    finally:
        try:
            _finish_like_an_interpreter()      # also after sys.exit() / an uncaught exception, like CPython
        except BaseException:
            pass
        try:
            sys.stdout.flush()
            sys.stderr.flush()
        except Exception:
            pass
    return code
