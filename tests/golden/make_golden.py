"""Regenerates the committed golden fixtures.  Runs ONLY where the reference checkout exists
(/root/reference, the authoring container); tests and the GPU box use the committed outputs.

  http_replay.json  the reference's own FastAPI app (services/http_server.py:77-106), imported
                    unmodified from /root/reference, driven in-process with the bodies of the
                    reference's e2e tests (test/e2e/test_http.py:23-99) and example scripts
                    (examples/{crash,escaping,files,ls}.py-style), with oracle.LocalCpuExecutor
                    as the executor backend and the reference's own Storage
                    (services/storage.py) as the object store.  Pins: response model shape,
                    status codes (200 / 422 / 500), file-map round trip, env propagation,
                    exit-code and stderr behaviour of failing scripts.
  pod_replay.json   the reference's FULL service stack — FastAPI app + the unmodified
                    KubernetesCodeExecutor.execute (kubernetes_code_executor.py:80-149: httpx PUT
                    uploads, POST /execute, streamed GET downloads) + the reference's Storage —
                    with bci-executor-server standing where the Rust executor pod stands (only
                    `executor_pod()` is overridden, to hand out 127.0.0.1 instead of asking
                    kubectl for a pod; one fresh server + workspace per request, like a fresh
                    pod).  Same request bodies as http_replay.json.
  examples_replay.json  the reference's example scripts under plain CPython (stdout / exit code /
                    last stderr line / files written): replayed through the native executor
                    server in both of its modes (tests/test_executor_server.py).
  numeric.npz       seeded inputs and the results stock numpy (the reference's arithmetic)
                    gives for every op class of SURVEY.md §8a; numpy version recorded.
"""
import json
import sys
import tempfile
from contextvars import ContextVar
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference/src")
OUT = Path(__file__).resolve().parent


def http_replay():
    from fastapi.testclient import TestClient

    from code_interpreter.services.http_server import create_http_server
    from code_interpreter.services.storage import Storage
    from oracle.cpu_executor import LocalCpuExecutor

    storage_dir = tempfile.mkdtemp(prefix="bci-golden-store-")
    executor = LocalCpuExecutor(Storage(storage_dir), timeout=5)

    class NoTools:  # custom tools are out of scope
        def parse(self, **kw):
            raise NotImplementedError

        async def execute(self, **kw):
            raise NotImplementedError

    app = create_http_server(executor, NoTools(), ContextVar("request_id", default=""))
    client = TestClient(app, raise_server_exceptions=False)
    cases = []

    def post(name, body, note):
        r = client.post("/v1/execute", json=body)
        try:
            payload = r.json()
        except Exception:
            payload = r.text
        cases.append({"name": name, "request": body, "status": r.status_code, "response": payload, "pins": note})
        return payload

    post("hello", {"source_code": "print(21 * 2)"}, "health_check.py:45-53 expects '42\\n'")
    created = post("create_file", {"source_code": "\nwith open('file.txt', 'w') as f:\n    f.write(\"Hello, World!\")\n", "files": {}},
                   "test_http.py:47-65: exit 0, files keys == {/workspace/file.txt}")
    fid = created["files"]["/workspace/file.txt"]
    post("read_file_back", {"source_code": "\nwith open('file.txt', 'r') as f:\n    print(f.read())\n",
                            "files": {"/workspace/file.txt": fid}},
         "test_http.py:67-85: stdout == content + newline, no files")
    post("env", {"source_code": "import os\nprint('Hello ' + os.environ['MY_NAME'])", "files": {}, "env": {"MY_NAME": "John Doe"}},
         "test_http.py:88-99")
    post("crash", {"source_code": "0/0"}, "examples/crash.py:15: non-zero exit, traceback on stderr, HTTP 200")
    post("sys_exit_3", {"source_code": "import sys\nprint('x')\nsys.exit(3)"}, "exit code passthrough (server.rs:165)")
    post("stderr_and_unicode", {"source_code": "import sys\nprint('caf\\u00e9 \\u2713')\nprint('warn', file=sys.stderr)"}, "utf-8 text on both streams")
    post("subdir_not_listed", {"source_code": "import os\nos.makedirs('d', exist_ok=True)\nopen('d/inner.txt','w').write('x')\nopen('top.bin','wb').write(bytes(range(256)))"},
         "server.rs:98-118: only top-level regular files are reported")
    post("timeout", {"source_code": "import time\ntime.sleep(30)"}, "server.rs:169: ('', 'Execution timed out', -1) [fixture timeout = 5 s]")
    post("bad_path", {"source_code": "pass", "files": {"relative/path": "abc"}}, "validation.py:20-22: AbsolutePath -> 422")
    post("bad_hash", {"source_code": "pass", "files": {"/workspace/x": "not a hash!"}}, "validation.py:19: Hash -> 422")
    post("unknown_object", {"source_code": "pass", "files": {"/workspace/x": "deadbeef"}}, "storage.py:72-73 FileNotFoundError -> http_server.py:102-104 HTTP 500")
    post("numpy_small", {"source_code": "import numpy as np\nnp.random.seed(42)\nx = np.random.normal(loc=10, scale=2, size=100)\nprint(repr(float(x.mean())))"},
         "examples/using_imports.py:20-22: seeded legacy generator stays on numpy's MT19937 stream")
    (OUT / "http_replay.json").write_text(json.dumps({"generated_with": {"numpy": np.__version__}, "cases": cases}, indent=1))
    print("http_replay.json:", len(cases), "cases")


def pod_replay():
    import contextlib
    import os
    import subprocess

    from fastapi.testclient import TestClient

    from bee_code_interpreter_b200 import build
    from code_interpreter.services.http_server import create_http_server
    from code_interpreter.services.kubernetes_code_executor import KubernetesCodeExecutor
    from code_interpreter.services.storage import Storage

    server_bin = str(build.build_server())
    storage = Storage(tempfile.mkdtemp(prefix="bci-golden-store-"))

    class PodIsLocalServer(KubernetesCodeExecutor):
        """Everything of the reference class except pod acquisition."""

        @contextlib.asynccontextmanager
        async def executor_pod(self):
            ws = tempfile.mkdtemp(prefix="bci-golden-pod-")
            # the pod's /workspace cannot exist in the authoring container: files live in `ws` and are
            # reported under the name they have in the pod
            env = dict(os.environ, APP_LISTEN_ADDR="127.0.0.1:8000", APP_WORKSPACE=ws, BCI_WORKSPACE_REPORT_AS="/workspace", BCI_PYTHON=sys.executable,
                       PYTHONPATH=str(ROOT / "bee_code_interpreter_b200" / "site"))
            proc = subprocess.Popen([server_bin], env=env, stdout=subprocess.PIPE, text=True)
            try:
                assert proc.stdout.readline().startswith("listening on")
                yield {"status": {"podIP": "127.0.0.1"}}
            finally:
                proc.terminate()
                proc.wait()

    executor = PodIsLocalServer(kubectl=None, executor_image="", container_resources={}, file_storage=storage,
                                executor_pod_spec_extra={}, executor_pod_queue_target_length=0, executor_pod_name_prefix="x")

    class NoTools:
        def parse(self, **kw):
            raise NotImplementedError

        async def execute(self, **kw):
            raise NotImplementedError

    client = TestClient(create_http_server(executor, NoTools(), ContextVar("request_id", default="")), raise_server_exceptions=False)
    base = json.loads((OUT / "http_replay.json").read_text())["cases"]
    cases, ids = [], {}
    for case in base:
        if case["name"] in ("timeout", "unknown_object"):
            continue   # 60 s default timeout of the pod contract / retry back-off of the reference class: not replayed
        body = json.loads(json.dumps(case["request"]))
        for path, old in list(body.get("files", {}).items()):
            body["files"][path] = ids.get(old, old)           # object ids are random per Storage instance
        r = client.post("/v1/execute", json=body)
        try:
            payload = r.json()
        except Exception:
            payload = r.text
        if isinstance(payload, dict) and isinstance(case["response"], dict):
            for path, new in payload.get("files", {}).items():
                if path in case["response"].get("files", {}):
                    ids[case["response"]["files"][path]] = new
        cases.append({"name": case["name"], "request": case["request"], "status": r.status_code, "response": payload})
    (OUT / "pod_replay.json").write_text(json.dumps({"cases": cases}, indent=1))
    print("pod_replay.json:", len(cases), "cases")


def examples_replay():
    """The reference's example scripts (examples/*.py, the inputs of its e2e tests and health check)
    run by PLAIN CPython in an empty directory: stdout, exit code and the last stderr line are what
    any executor of this path must reproduce.  Skipped: the benchmarks (minutes of CPU), fib.py
    (a minute of recursion), cowsay.py (module absent here), tcp.py (needs a peer)."""
    import os
    import subprocess

    cases = []
    prev_ws = None
    for name in ["crash", "escaping", "files", "hello_world_write_file", "hello_world_read_file", "ls", "using_imports"]:
        src = Path(f"/root/reference/examples/{name}.py").read_text()
        ws = prev_ws if name == "hello_world_read_file" else tempfile.mkdtemp(prefix="bci-golden-ex-")
        r = subprocess.run([sys.executable, "-c", src], cwd=ws, capture_output=True, text=True,
                           env=dict(os.environ, HELLO="h", WORLD="w"))
        cases.append({"name": name, "source_code": src, "env": {"HELLO": "h", "WORLD": "w"},
                      "needs_files_of": "hello_world_write_file" if name == "hello_world_read_file" else None,
                      "stdout": r.stdout.replace(ws, "{WORKSPACE}"), "exit_code": r.returncode,
                      "stderr_last_line": r.stderr.strip().splitlines()[-1:] or [],
                      "files": sorted(f for f in os.listdir(ws) if os.path.isfile(os.path.join(ws, f)))})
        prev_ws = ws
    (OUT / "examples_replay.json").write_text(json.dumps({"generated_with": {"python": sys.version.split()[0], "numpy": np.__version__},
                                                            "cases": cases}, indent=1))
    print("examples_replay.json:", len(cases), "cases")


def numeric():
    data = {}
    for name, dt in [("f32", np.float32), ("f64", np.float64), ("i32", np.int32), ("i64", np.int64)]:
        rng = np.random.default_rng(9000 + len(name) + ord(name[0]) + int(name[1:]))
        if np.issubdtype(dt, np.floating):
            x = ((rng.random(5003) - 0.5) * 200).astype(dt)
            y = ((rng.random(5003) - 0.5) * 200).astype(dt)
        else:
            x = rng.integers(-1000, 1000, 5003, dtype=dt)
            y = rng.integers(-1000, 1000, 5003, dtype=dt)
        data[f"{name}_x"], data[f"{name}_y"] = x, y
        data[f"{name}_add"] = x + y
        data[f"{name}_mul"] = x * y
        data[f"{name}_square"] = np.square(x)
        data[f"{name}_abs_sub"] = np.abs(x - y)
        data[f"{name}_sum"] = np.asarray(np.sum(x))
        data[f"{name}_mean"] = np.asarray(np.mean(x))
        data[f"{name}_max"] = np.asarray(np.max(x))
        data[f"{name}_min"] = np.asarray(np.min(x))
        data[f"{name}_sumsq"] = np.asarray(np.sum(np.square(x)))
        data[f"{name}_argmax"] = np.asarray(np.argmax(x))
        data[f"{name}_argmin"] = np.asarray(np.argmin(x))
        data[f"{name}_cumsum"] = np.cumsum(x)
        data[f"{name}_sort"] = np.sort(x)
        data[f"{name}_argsort"] = np.argsort(x, kind="stable")
    a = np.random.default_rng(1000).random((96, 160), dtype=np.float32)
    b = np.random.default_rng(1001).random((160, 64), dtype=np.float32)
    data["mm_a"], data["mm_b"], data["mm_c"] = a, b, a @ b
    data["numpy_version"] = np.asarray(np.__version__)
    np.savez_compressed(OUT / "numeric.npz", **data)
    print("numeric.npz:", len(data), "arrays")


if __name__ == "__main__":
    which = sys.argv[1:] or ["http", "pod", "examples", "numeric"]
    if "http" in which:
        http_replay()
    if "pod" in which:
        pod_replay()
    if "examples" in which:
        examples_replay()
    if "numeric" in which:
        numeric()
