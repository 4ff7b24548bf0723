"""bci-executor-server against the reference's in-pod wire contract (executor/server.rs:69-192):
PUT/GET /workspace/{path}, POST /execute.  The request bodies are the ones
KubernetesCodeExecutor sends (kubernetes_code_executor.py:100-131) and the scripts those of the
reference's e2e tests (test/e2e/test_http.py:23-99, examples/).  CPU only: exec mode runs
`python script.py` with the product's sitecustomize on PYTHONPATH; zygote mode forks the warm
interpreter.  Neither has a GPU broker here, so array calls take the hook's numpy fallback."""
import http.client
import json
import os
import socket
import subprocess
import sys
import time
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "bee_code_interpreter_b200"


class Server:
    def __init__(self, tmp_path, zygote_socket=None):
        from bee_code_interpreter_b200 import build

        self.workspace = tmp_path / "workspace"
        self.workspace.mkdir()
        env = dict(os.environ)
        env.update({"APP_LISTEN_ADDR": "127.0.0.1:0", "APP_WORKSPACE": str(self.workspace), "BCI_PYTHON": sys.executable,
                    "PYTHONPATH": str(PKG / "site"), "BCI_SCRATCH": str(tmp_path)})
        env.pop("BCI_BROKER_SOCKET", None)
        if zygote_socket:
            env["BCI_ZYGOTE_SOCKET"] = zygote_socket
        self.proc = subprocess.Popen([str(build.build_server())], env=env, stdout=subprocess.PIPE, text=True)
        line = self.proc.stdout.readline()
        assert line.startswith("listening on"), line
        self.port = int(line.split()[2].rsplit(":", 1)[1])
        self.mode = line.strip().rsplit("=", 1)[1]

    def conn(self):
        return http.client.HTTPConnection("127.0.0.1", self.port, timeout=30)

    def request(self, method, path, body=None, headers=None):
        c = self.conn()
        c.request(method, path, body=body, headers=headers or {})
        r = c.getresponse()
        data = r.read()
        c.close()
        return r.status, dict((k.lower(), v) for k, v in r.getheaders()), data

    def execute(self, source_code, **extra):
        status, headers, data = self.request("POST", "/execute", json.dumps({"source_code": source_code, **extra}),
                                             {"content-type": "application/json"})
        assert status == 200, (status, data)
        assert headers["content-type"] == "application/json"
        return json.loads(data)

    def close(self):
        self.proc.terminate()
        self.proc.wait(timeout=5)


@pytest.fixture
def server(tmp_path, built):
    s = Server(tmp_path)
    yield s
    s.close()


def test_execute_stdout_stderr_exit_code(server):
    assert server.mode == "exec"
    r = server.execute("print(21 * 2)")
    assert r == {"stdout": "42\n", "stderr": "", "exit_code": 0, "files": []}
    r = server.execute("import sys\nprint('to err', file=sys.stderr)\nsys.exit(3)")
    assert (r["stdout"], r["stderr"], r["exit_code"]) == ("", "to err\n", 3)
    r = server.execute("raise ValueError('boom')")                     # examples/crash.py shape
    assert r["exit_code"] == 1 and "ValueError: boom" in r["stderr"] and r["stdout"] == ""
    r = server.execute("import os, signal\nos.kill(os.getpid(), signal.SIGKILL)")
    assert r["exit_code"] == -1                                         # server.rs:165: no exit code -> -1


def test_file_round_trip_like_the_reference_e2e_test(server):
    # test_http.py:47-75: write a file in one request, read it back in the next one
    r = server.execute("with open('file.txt', 'w') as f:\n    f.write('Hello, World!')")
    assert r["exit_code"] == 0 and r["files"] == [str(server.workspace / "file.txt")]
    status, headers, data = server.request("GET", "/workspace/file.txt")
    assert status == 200 and data == b"Hello, World!" and headers["content-type"] == "application/octet-stream"
    r = server.execute("print(open('file.txt').read())")
    assert r["stdout"] == "Hello, World!\n" and r["files"] == []        # unchanged file is not reported


def test_put_creates_parents_and_streams_chunked_bodies(server):
    payload = os.urandom(3 * 2**20 + 17)
    status, _, _ = server.request("PUT", "/workspace/a/b/c.bin", payload)
    assert status == 204 and (server.workspace / "a" / "b" / "c.bin").read_bytes() == payload

    def chunks():                                                         # httpx streams uploads chunked
        for i in range(0, len(payload), 65536 + 7):
            yield payload[i:i + 65536 + 7]

    c = server.conn()
    c.request("PUT", "/workspace/streamed%20file.bin", body=chunks(), encode_chunked=True,
              headers={"transfer-encoding": "chunked"})
    r = c.getresponse()
    r.read()
    assert r.status == 204
    c.request("GET", "/workspace/streamed%20file.bin")                   # same connection: keep-alive
    r = c.getresponse()
    assert r.status == 200 and r.read() == payload
    c.close()
    status, _, _ = server.request("PUT", "/workspace/over.bin", b"first version, longer")
    status, _, _ = server.request("PUT", "/workspace/over.bin", b"second")
    assert (server.workspace / "over.bin").read_bytes() == b"second"    # truncate like server.rs:80
    # only top-level regular files are listed, nested ones are not (server.rs:100-115)
    r = server.execute("import os\nos.makedirs('d', exist_ok=True)\nopen('d/inner.txt','w').write('x')\nopen('top.txt','w').write('y')")
    assert r["files"] == [str(server.workspace / "top.txt")]


def test_env_timeout_and_lossy_output(server):
    r = server.execute("import os\nprint(os.environ['GREETING'])", env={"GREETING": "hi there"})   # test_http.py:78-99
    assert r["stdout"] == "hi there\n"
    r = server.execute("print('x')", env=None, timeout=None)
    assert r["stdout"] == "x\n"
    t0 = time.time()
    r = server.execute("import time\nprint('before', flush=True)\ntime.sleep(30)", timeout=1)
    assert r == {"stdout": "", "stderr": "Execution timed out", "exit_code": -1, "files": []}      # server.rs:169
    assert time.time() - t0 < 10
    r = server.execute("import sys\nsys.stdout.buffer.write(b'ok \\xff\\xfe \\xe2\\x82 end \\xe2\\x82\\xac')")
    assert r["stdout"] == "ok \ufffd\ufffd \ufffd end \u20ac"                                       # from_utf8_lossy
    r = server.execute("print('snow \u2603 \U0001F600 \"quoted\" \\\\ tab\\t')")
    assert r["stdout"] == 'snow \u2603 \U0001F600 "quoted" \\ tab\t\n'


def test_errors_match_the_contract(server):
    assert server.request("GET", "/workspace/missing.txt")[0] == 404
    assert server.request("GET", "/workspace/../etc/passwd")[0] == 400
    assert server.request("PUT", "/workspace/%2e%2e/x", b"1")[0] == 400
    assert server.request("POST", "/execute", b"{not json", {"content-type": "application/json"})[0] == 400
    assert server.request("POST", "/execute", json.dumps({"env": {}}), {"content-type": "application/json"})[0] == 400
    assert server.request("POST", "/execute", json.dumps({"source_code": "1", "timeout": -1}))[0] == 400
    assert server.request("POST", "/execute", json.dumps({"source_code": "1", "env": {"A": 1}}))[0] == 400
    assert server.request("GET", "/nope")[0] == 404
    assert server.request("DELETE", "/execute")[0] == 405
    assert server.execute("print(1)")["stdout"] == "1\n"               # still serving


def test_hook_is_active_in_scripts(server):
    src = ("import numpy as np\nimport json as j\n"
           "x = np.random.rand(1000)\n"
           "print(type(np).__name__, float(np.sum(np.square(x))) > 0, np.pi > 3)\n")
    r = server.execute(src)
    assert r["exit_code"] == 0 and r["stdout"] == "_ModuleView True True\n", r


def test_concurrent_requests(server):
    import concurrent.futures as cf

    def one(i):
        return server.execute(f"import time\ntime.sleep(0.2)\nprint({i} * {i})")["stdout"]

    t0 = time.time()
    with cf.ThreadPoolExecutor(16) as pool:
        outs = list(pool.map(one, range(16)))
    assert outs == [f"{i * i}\n" for i in range(16)]
    assert time.time() - t0 < 3.0                                       # not serialised


def test_zygote_mode_forks_the_warm_interpreter(tmp_path, built):
    from bee_code_interpreter_b200.zygote import ZygoteClient

    z = ZygoteClient()
    s = Server(tmp_path, zygote_socket=z.socket_path)
    try:
        assert s.mode == "zygote"
        r = s.execute("import os\nprint(os.getcwd())\nopen('made.txt','w').write('z')", env={"K": "v"})
        assert r["stdout"].strip() == str(s.workspace) and r["files"] == [str(s.workspace / "made.txt")]
        r = s.execute("import os, sys\nprint(os.environ.get('K'))\nsys.exit(7)", env={"K": "v2"})
        assert (r["stdout"], r["exit_code"]) == ("v2\n", 7)
        r = s.execute("import os\nprint(os.environ.get('K'))")         # env of one request does not leak into the next
        assert r["stdout"] == "None\n"
        r = s.execute("import time\ntime.sleep(30)", timeout=1)
        assert (r["stderr"], r["exit_code"]) == ("Execution timed out", -1)
        r = s.execute("import os, signal\nos.kill(os.getpid(), signal.SIGKILL)")
        assert r["exit_code"] == -1
        times = []
        for _ in range(20):
            t0 = time.perf_counter()
            assert s.execute("print('w')")["stdout"] == "w\n"
            times.append(time.perf_counter() - t0)
        assert sorted(times)[len(times) // 2] < 0.05                    # a fork, not an interpreter start
    finally:
        s.close()
        z.close()


# ---- golden replays -----------------------------------------------------------------------------
GOLDEN = Path(__file__).resolve().parent / "golden"


def _cases(name):
    return {c["name"]: c for c in json.loads((GOLDEN / name).read_text())["cases"]}


def test_reference_stack_with_this_server_as_the_pod_matches_the_oracle_path():
    """pod_replay.json was recorded by the reference's own app + unmodified KubernetesCodeExecutor
    (httpx PUT / POST / streamed GET) talking to bci-executor-server; http_replay.json by the same
    app over the oracle executor.  Same requests, same answers (object ids are random per store)."""
    api, pod = _cases("http_replay.json"), _cases("pod_replay.json")
    assert len(pod) >= 11
    for name, c in pod.items():
        want, got = api[name], c
        assert got["status"] == want["status"], name
        if isinstance(want["response"], dict) and "stdout" in want["response"]:
            for key in ("stdout", "exit_code"):
                assert got["response"][key] == want["response"][key], (name, key)
            assert sorted(got["response"]["files"]) == sorted(want["response"]["files"]), name
            assert got["response"]["stderr"].splitlines()[-1:] == want["response"]["stderr"].splitlines()[-1:], name
        else:
            assert got["response"] == want["response"], name            # 422 bodies from pydantic, verbatim


def test_live_replay_of_the_api_fixtures_through_the_wire_contract(tmp_path, built):
    """The three calls KubernetesCodeExecutor makes per request (kubernetes_code_executor.py:100-142),
    issued here with http.client against a fresh server per request (= fresh pod)."""
    api = _cases("http_replay.json")
    store = {}                                                       # object id -> bytes (the Storage role)
    for name in ("hello", "create_file", "read_file_back", "env", "crash", "sys_exit_3", "stderr_and_unicode",
                 "subdir_not_listed", "numpy_small"):
        case = api[name]
        d = tmp_path / name
        d.mkdir()
        s = Server(d)
        try:
            for path, object_id in case["request"].get("files", {}).items():
                status, _, _ = s.request("PUT", "/workspace/" + path.removeprefix("/workspace/"), store[object_id])
                assert status == 204
            r = s.execute(case["request"]["source_code"], env=case["request"].get("env", {}))
            want = case["response"]
            assert r["stdout"] == want["stdout"] and r["exit_code"] == want["exit_code"], name
            assert r["stderr"].splitlines()[-1:] == want["stderr"].splitlines()[-1:], name
            got_files = {}
            for p in r["files"]:
                rel = p.removeprefix(str(s.workspace) + "/")
                status, _, data = s.request("GET", "/workspace/" + rel)
                assert status == 200
                got_files["/workspace/" + rel] = data
            assert sorted(got_files) == sorted(want["files"]), name
            for path, data in got_files.items():
                store[want["files"][path]] = data                      # next request refers to it by the recorded id
        finally:
            s.close()
    assert store and b"Hello, World!" in store.values()


def test_zygote_pid_and_exit_code_never_swap(built, tmp_path):
    """The child announces its own pid before it runs the script: with the parent announcing it, a
    script that finishes first could get its exit code read as the pid (and the pid as the code)."""
    from bee_code_interpreter_b200.zygote import ZygoteClient

    z = ZygoteClient()
    try:
        req = {"source_code": "import sys\nsys.exit(3)", "workspace": str(tmp_path), "env": {}, "stdout_path": str(tmp_path / "o"),
               "stderr_path": str(tmp_path / "e"), "broker_socket": "/nonexistent", "min_rand_elems": None}
        codes = {z.run(req, 30) for _ in range(400)}
        assert codes == {3}
    finally:
        z.close()


def test_files_written_in_the_first_millisecond_are_reported(server):
    """Many immediate writers: none may be lost to the coarse-clock ctime of the kernel."""
    for i in range(60):
        r = server.execute(f"open('early{i}.txt', 'w').write('x')")
        assert r["files"] == [str(server.workspace / f"early{i}.txt")], (i, r)
    r = server.execute("open('early0.txt', 'w').write('y')")              # same size, rewritten: still reported
    assert r["files"] == [str(server.workspace / "early0.txt")]
    r = server.execute("pass")
    assert r["files"] == []


@pytest.mark.parametrize("mode", ["exec", "zygote"])
def test_reference_example_scripts_behave_as_under_plain_python(tmp_path, built, mode):
    """examples_replay.json = the reference's examples/*.py under plain CPython.  Through the native
    server (import hook active, seeded numpy stream, pandas / scipy importing the REAL numpy) the
    stdout, exit code, last stderr line and written files must be the same."""
    cases = json.loads((GOLDEN / "examples_replay.json").read_text())["cases"]
    z = None
    if mode == "zygote":
        from bee_code_interpreter_b200.zygote import ZygoteClient

        z = ZygoteClient()
    try:
        kept = None
        for case in cases:
            d = kept if case["needs_files_of"] else tmp_path / case["name"]
            if not case["needs_files_of"]:
                d.mkdir()
                s = Server(d, zygote_socket=z.socket_path if z else None)
            try:
                if case["needs_files_of"]:
                    s = Server.__new__(Server)          # same workspace as the previous case: a new server on it
                    s.workspace = kept / "workspace"
                    env = dict(os.environ, APP_LISTEN_ADDR="127.0.0.1:0", APP_WORKSPACE=str(s.workspace), BCI_PYTHON=sys.executable,
                               PYTHONPATH=str(PKG / "site"), BCI_SCRATCH=str(kept))
                    if z:
                        env["BCI_ZYGOTE_SOCKET"] = z.socket_path
                    from bee_code_interpreter_b200 import build

                    s.proc = subprocess.Popen([str(build.build_server())], env=env, stdout=subprocess.PIPE, text=True)
                    s.port = int(s.proc.stdout.readline().split()[2].rsplit(":", 1)[1])
                r = s.execute(case["source_code"], env=case["env"])
                assert r["stdout"] == case["stdout"].replace("{WORKSPACE}", str(s.workspace)), case["name"]
                assert r["exit_code"] == case["exit_code"], case["name"]
                assert r["stderr"].strip().splitlines()[-1:] == case["stderr_last_line"], case["name"]
                on_disk = sorted(f.name for f in s.workspace.iterdir() if f.is_file())
                assert on_disk == case["files"], case["name"]
            finally:
                s.close()
            kept = d
    finally:
        if z:
            z.close()


def test_json_round_trip_of_arbitrary_text(server):
    """Random unicode through both directions of the server's own JSON code: request strings
    arrive \\u-escaped (surrogate pairs included) or as raw UTF-8, stdout goes back escaped by the
    server; what the script prints must be what was sent."""
    from hypothesis import given, settings, strategies as st

    alphabet = st.characters(blacklist_categories=("Cs",))            # no lone surrogates: not encodable as UTF-8

    @settings(max_examples=40, deadline=None)
    @given(text=st.text(alphabet=alphabet, max_size=200), ascii_escapes=st.booleans())
    def run(text, ascii_escapes):
        src = "import sys\nsys.stdout.write(" + repr(text) + ")\nsys.stderr.write(" + repr(text[::-1]) + ")"
        body = json.dumps({"source_code": src, "env": {"T": text.replace("\x00", "")}}, ensure_ascii=ascii_escapes)
        status, headers, data = server.request("POST", "/execute", body.encode("utf-8"), {"content-type": "application/json"})
        assert status == 200, data
        r = json.loads(data)
        # text mode translates "\r" and "\r\n" on neither side here: the script writes through sys.stdout in
        # universal-newlines-off mode for output, so the bytes are the text's UTF-8
        assert r["stdout"] == text and r["stderr"] == text[::-1] and r["exit_code"] == 0

    run()
