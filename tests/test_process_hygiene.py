"""What a request may not leave behind or reach: stdin is /dev/null (server.rs runs the script
through Command::output(), which does not inherit stdin), and processes the script started die
with the request — on timeout and on normal completion — where the reference deletes the pod."""
import os
import sys
import time
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))


def _alive(pid):
    try:
        os.kill(pid, 0)
    except ProcessLookupError:
        return False
    except PermissionError:
        return True
    try:  # a zombie still "exists": look at its state
        return open(f"/proc/{pid}/stat").read().split(")")[1].split()[0] != "Z"
    except OSError:
        return False


@pytest.fixture
def zyg(built, tmp_path):
    from bee_code_interpreter_b200.zygote import ZygoteClient

    z = ZygoteClient()
    yield z, tmp_path
    z.close()


def _req(tmp_path, src):
    return {"source_code": src, "workspace": str(tmp_path), "env": {}, "stdout_path": str(tmp_path / "o"),
            "stderr_path": str(tmp_path / "e"), "broker_socket": "/nonexistent", "min_rand_elems": None}


def test_stdin_is_dev_null(zyg):
    z, d = zyg
    assert z.run(_req(d, "import sys\nprint(repr(sys.stdin.read()))"), 30) == 0
    assert (d / "o").read_text() == "''\n"
    assert z.run(_req(d, "input()"), 30) == 1
    assert "EOFError" in (d / "e").read_text()


def test_background_processes_die_with_the_request(zyg):
    z, d = zyg
    spawn = ("import subprocess, sys\n"
             "p = subprocess.Popen([sys.executable, '-c', 'import time; time.sleep(300)'])\n"
             "open('child.pid', 'w').write(str(p.pid))\n")
    assert z.run(_req(d, spawn), 30) == 0                                   # normal completion
    pid = int((d / "child.pid").read_text())
    for _ in range(100):
        if not _alive(pid):
            break
        time.sleep(0.02)
    assert not _alive(pid)
    assert z.run(_req(d, spawn + "import time\ntime.sleep(300)\n"), 1.0) is None   # timeout
    pid = int((d / "child.pid").read_text())
    for _ in range(100):
        if not _alive(pid):
            break
        time.sleep(0.02)
    assert not _alive(pid)


def test_script_ends_like_an_interpreter(zyg):
    """Non-daemon threads are joined and atexit callbacks run before the process ends (os._exit
    alone would cut both short); tracebacks quote the failing source line."""
    z, d = zyg
    src = ("import atexit, threading, time\n"
           "def late():\n"
           "    time.sleep(0.3)\n"
           "    open('from_thread.txt', 'w').write('t')\n"
           "threading.Thread(target=late).start()\n"
           "atexit.register(lambda: open('from_atexit.txt', 'w').write('a'))\n"
           "print('main done')\n")
    script = d / "script.py"
    script.write_text(src)
    req = dict(_req(d, src), script_path=str(script))
    assert z.run(req, 30) == 0
    assert (d / "from_thread.txt").read_text() == "t" and (d / "from_atexit.txt").read_text() == "a"
    src = "import sys\nx = 1\nsys.exit(5)\n"
    assert z.run(_req(d, "import atexit\natexit.register(lambda: open('after_exit.txt', 'w').write('e'))\n" + src), 30) == 5
    assert (d / "after_exit.txt").read_text() == "e"                       # also after sys.exit()
    bad = "def f():\n    return 1 / 0\nf()\n"
    script.write_text(bad)
    assert z.run(dict(_req(d, bad), script_path=str(script)), 30) == 1
    err = (d / "e").read_text()
    assert "return 1 / 0" in err and "ZeroDivisionError" in err and str(script) in err


def test_files_map_cannot_leave_the_workspace(tmp_path):
    from bee_code_interpreter_b200 import executor as E

    ws = tmp_path / "workspace"
    ws.mkdir()
    assert E._workspace_target(ws, "/workspace/a/b.txt") == ws / "a" / "b.txt"
    assert E._workspace_target(ws, "/etc/passwd") == ws / "etc" / "passwd"        # server.rs joins strings: stays inside
    assert E._workspace_target(ws, "/workspace//x") == ws / "x"
    for bad in ("/workspace/../evil", "/workspace/a/../../evil", "/..", "/workspace/"):
        with pytest.raises(ValueError):
            E._workspace_target(ws, bad)


def test_dead_zygote_and_dead_broker_are_restarted_and_the_request_retried(built):
    """kubernetes_code_executor.py:75-79 retries `execute` on RuntimeError; here the dead part (warm interpreter,
    broker process) is started again between the attempts, so the caller sees a normal result."""
    import asyncio
    import os
    import signal
    import sys
    from pathlib import Path

    from bee_code_interpreter_b200.broker import BrokerProcess
    from bee_code_interpreter_b200.executor import LocalGpuExecutor

    tests_dir = str(Path(__file__).resolve().parent)
    broker = BrokerProcess(fake_device="fake_device", extra_pythonpath=tests_dir)
    ex = LocalGpuExecutor(broker=broker, max_inflight=4, timeout=20, zygotes=1, min_rand_elems=1000, retry_wait=(0.05, 0.2))
    ex._own_broker = True      # this executor is responsible for the broker it was given
    try:
        src = "import numpy as np\nx = np.random.rand(5000)\nprint(type(x).__name__, float(np.sum(x)) > 0)\n"
        assert asyncio.run(ex.execute(src)).stdout == "DeviceArray True\n"
        os.kill(ex._zygotes[0]._proc.pid, signal.SIGKILL)
        ex._zygotes[0]._proc.wait()
        r = asyncio.run(ex.execute(src))
        assert r.exit_code == 0 and r.stdout == "DeviceArray True\n"
        os.kill(broker._proc.pid, signal.SIGKILL)
        broker._proc.wait()
        r = asyncio.run(ex.execute(src))
        assert r.exit_code == 0 and r.stdout == "DeviceArray True\n" and broker.restarts == 1
    finally:
        ex.close()


def test_symlinks_in_the_workspace_are_stored_by_content_never_as_links(built, tmp_path):
    """`latest.csv -> data.csv` must not put a (dangling) link into the content-addressed store.
    The reference reports links to regular files as changed files (Path::is_file follows them, server.rs:103)
    and stores what they point to; so do we — as a copy of the bytes that were hashed."""
    import asyncio
    import hashlib
    import os
    import sys
    from pathlib import Path

    from bee_code_interpreter_b200.executor import LocalGpuExecutor

    sys.path.insert(0, str(Path(__file__).resolve().parent))
    import fake_device

    ex = LocalGpuExecutor(device_module=fake_device, max_inflight=2, timeout=20, zygotes=1, storage_path=str(tmp_path / "store"))
    try:
        src = ("import os\nopen('data.csv', 'w').write('a,b\\n1,2\\n')\nos.symlink('data.csv', 'latest.csv')\n"
               "os.symlink('/nonexistent/target', 'dangling')\nos.mkdir('sub')\n")
        r = asyncio.run(ex.execute(src))
        assert r.exit_code == 0, r.stderr
        assert set(r.files) == {"/workspace/data.csv", "/workspace/latest.csv"}
        digest = hashlib.sha256(b"a,b\n1,2\n").hexdigest()
        assert r.files["/workspace/data.csv"] == r.files["/workspace/latest.csv"] == digest
        stored = list((tmp_path / "store").iterdir())
        assert [p.name for p in stored] == [digest] and not stored[0].is_symlink() and stored[0].read_bytes() == b"a,b\n1,2\n"
    finally:
        ex.close()
