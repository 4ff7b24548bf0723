"""End-to-end on the real device: agent source text -> zygote fork -> import hook -> broker ->
batched CUDA launches -> stdout; output files hashed by the SHA-256 kernel.  Checked against
numpy (the reference's arithmetic) and hashlib."""
import asyncio
import hashlib

import numpy as np
import pytest

from bee_code_interpreter_b200 import numpy_proxy as P
from bee_code_interpreter_b200 import workloads as W

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def executor(dev):
    from bee_code_interpreter_b200.executor import LocalGpuExecutor

    ex = LocalGpuExecutor(max_inflight=16, timeout=60, zygotes=2)
    yield ex
    ex.close()


def test_metric_script_through_executor(executor, dev):
    n = 10_000_000
    res = asyncio.run(_many(executor, W.benchmark_numpy_source(n), 24))
    for r in res:
        assert r.exit_code == 0, r.stderr
        v = float(r.stdout.splitlines()[0].split()[1])
        assert abs(v / n - 1 / 3) < 6 * np.sqrt(4 / 45 / n)
    # the GPU work happened in the broker process (this process launched nothing for it)
    from bee_code_interpreter_b200.broker_client import RemoteBackend

    stats = RemoteBackend(executor.broker.socket_path).ping()["stats"]
    assert stats["jobs"] >= 24 and stats["batches"] <= stats["jobs"]      # 24 x ONE job: rand generated inside the fused reduce


async def _many(ex, src, k):
    return await asyncio.gather(*(ex.execute(src) for _ in range(k)))


C3_SCRIPT = '''
import numpy as np
x = np.load("x.npy"); y = np.load("y.npy")
import numpy
bx = numpy.sort(x)                      # large host array: offloaded by the hook
print(repr(float(bx[0])), repr(float(bx[-1])))
r = numpy.random.rand(2_000_000)
print(type(r).__name__, r.shape, float(numpy.min(r)) >= 0.0, float(numpy.max(r)) < 1.0)
s = r * 2.0 + 1.0
print(repr(float(numpy.mean(s))))
numpy.save("out.npy", numpy.asarray(numpy.sqrt(r)))
'''


def test_script_with_files_and_hook_offload(executor):
    from bee_code_interpreter_b200.storage import ContentAddressedStorage  # noqa: F401
    import io

    x = np.random.default_rng(1).random(P.OFFLOAD_MIN_ELEMS * 2).astype(np.float32)
    y = np.random.default_rng(2).random(1000).astype(np.float32)

    def npy(a):
        b = io.BytesIO()
        np.save(b, a)
        return b.getvalue()

    async def go():
        hx = await executor.file_storage.write(npy(x))
        hy = await executor.file_storage.write(npy(y))
        assert hx == hashlib.sha256(npy(x)).hexdigest()
        r = await executor.execute(C3_SCRIPT, files={"/workspace/x.npy": hx, "/workspace/y.npy": hy})
        assert r.exit_code == 0, r.stderr
        lines = r.stdout.splitlines()
        assert lines[0] == f"{float(x.min())!r} {float(x.max())!r}"
        assert lines[1] == "DeviceArray (2000000,) True True"
        assert abs(float(lines[2]) - 2.0) < 0.01
        assert set(r.files) == {"/workspace/out.npy"}
        blob = await executor.file_storage.read(r.files["/workspace/out.npy"])
        assert hashlib.sha256(blob).hexdigest() == r.files["/workspace/out.npy"]
        out = np.load(io.BytesIO(blob))
        assert out.shape == (2_000_000,) and 0.0 <= out.min() and out.max() < 1.0

    asyncio.run(go())


def test_proxy_on_device_matches_numpy(dev):
    from bee_code_interpreter_b200.backend import LocalBackend

    be = LocalBackend(dev)
    rng = np.random.default_rng(5)
    xh, yh = rng.random(1 << 20, dtype=np.float32), rng.random(1 << 20, dtype=np.float32)
    x, y = P.DeviceArray.from_host(be, xh), P.DeviceArray.from_host(be, yh)
    assert abs(np.sum(np.square(x)) - np.sum(np.square(xh))) <= 1e-5 * np.sum(np.square(xh))
    assert abs(np.mean(x * y + 1.0) - np.mean(xh * yh + 1.0)) <= 1e-5
    assert np.max(np.abs(x - y)) == np.max(np.abs(xh - yh))
    assert np.array_equal(np.asarray(x * y + 1.0), xh * yh + 1.0)          # unfused path, bit-exact
    assert np.argmax(x) == np.argmax(xh)
    assert np.array_equal(np.asarray(np.sort(x)), np.sort(xh))
    k = P.DeviceArray.from_host(be, rng.integers(0, 2**31, 1 << 20, dtype=np.int64))
    assert np.array_equal(np.asarray(np.cumsum(k)), np.cumsum(np.asarray(k)))
    # powers, clip, ptp, NaN-aware reductions built on the same kernels
    assert np.array_equal(np.asarray(x ** 2), xh ** 2) and np.array_equal(np.asarray(np.power(x, 2)), np.power(xh, 2))
    assert np.array_equal(np.asarray(x ** 0.5), xh ** 0.5)
    ih = rng.integers(-2**20, 2**20, 1 << 20, dtype=np.int32)
    i32 = P.DeviceArray.from_host(be, ih)
    assert np.array_equal(np.asarray(i32 ** 2), ih ** 2)                      # wraps like numpy
    assert np.array_equal(np.asarray(np.clip(x, 0.25, 0.75)), np.clip(xh, 0.25, 0.75))
    assert np.array_equal(np.asarray(np.clip(i32, -5, None)), np.clip(ih, -5, None))
    nh = xh.copy(); nh[12345] = np.nan
    nx = P.DeviceArray.from_host(be, nh)
    assert np.array_equal(np.asarray(np.clip(nx, 0.25, 0.75)), np.clip(nh, 0.25, 0.75), equal_nan=True)
    assert np.ptp(x) == np.ptp(xh) and np.ptp(i32) == np.ptp(ih) and type(np.ptp(x)) is type(np.ptp(xh))
    assert np.nanmax(x) == np.nanmax(xh) and np.nanmin(i32) == np.nanmin(ih)
    assert abs(np.nansum(x) - np.nansum(xh)) <= 1e-5 * np.nansum(xh) and np.nansum(i32) == np.nansum(ih)
    assert np.nanmax(nx) == np.nanmax(nh) and abs(np.nanmean(nx) - np.nanmean(nh)) <= 1e-6      # NaN present: numpy's path
    assert abs(np.linalg.norm(x) - np.linalg.norm(xh)) <= 1e-5 * np.linalg.norm(xh)
    assert abs(np.sum(x ** 2) - np.sum(xh ** 2)) <= 1e-5 * np.sum(xh ** 2)
    # order statistics: one device sort + a few elements; values and types are numpy's
    kh = np.asarray(k).copy()
    for f, args in ((np.median, ()), (np.percentile, (95,)), (np.quantile, (0.001,)), (np.percentile, ([25, 50, 75],))):
        for h, d in ((xh, x), (kh, k)):
            want, got = f(h, *args), f(d, *args)
            assert type(want) is type(got) and np.array_equal(np.asarray(want), np.asarray(got)), (f.__name__, args)


def test_native_executor_server_runs_scripts_on_the_gpu(executor, tmp_path):
    """bci-executor-server (the server.rs stand-in) in zygote mode, sharing the executor's broker:
    POST /execute with the metric script -> the array work shows up as broker jobs."""
    import http.client
    import json
    import os
    import subprocess

    from bee_code_interpreter_b200 import build
    from bee_code_interpreter_b200.broker_client import RemoteBackend
    from bee_code_interpreter_b200.zygote import ZygoteClient

    ws = tmp_path / "workspace"
    ws.mkdir()
    z = ZygoteClient()
    env = dict(os.environ, APP_LISTEN_ADDR="127.0.0.1:0", APP_WORKSPACE=str(ws), BCI_ZYGOTE_SOCKET=z.socket_path,
               BCI_BROKER_SOCKET=executor.broker.socket_path)
    proc = subprocess.Popen([str(build.build_server())], env=env, stdout=subprocess.PIPE, text=True)
    try:
        port = int(proc.stdout.readline().split()[2].rsplit(":", 1)[1])
        before = RemoteBackend(executor.broker.socket_path).ping()["stats"]["jobs"]
        n = 10_000_000
        c = http.client.HTTPConnection("127.0.0.1", port, timeout=120)
        for _ in range(4):
            c.request("POST", "/execute", json.dumps({"source_code": W.benchmark_numpy_source(n), "env": {"BCI_DEBUG_FALLBACK": "1"}}),
                      {"content-type": "application/json"})
            r = json.loads(c.getresponse().read())
            assert r["exit_code"] == 0 and r["stderr"] == "", r
            v = float(r["stdout"].splitlines()[0].split()[1])
            assert abs(v / n - 1 / 3) < 6 * np.sqrt(4 / 45 / n)
        after = RemoteBackend(executor.broker.socket_path).ping()["stats"]["jobs"]
        assert after - before >= 4                                      # 4 x the fused generate+reduce job, none on the CPU
    finally:
        proc.terminate()
        proc.wait(timeout=5)
        z.close()


def test_seeded_script_gets_numpys_stream_from_the_gpu(executor):
    """np.random.seed(...) + np.random.rand(big): the device continues MT19937 bit for bit and
    hands the state back, so the script prints exactly what it prints on a CPU."""
    src = ("import numpy as np\n"
           "np.random.seed(20240921)\n"
           "a = np.random.rand(5)\n"
           "x = np.random.rand(3_000_001)\n"
           "b = np.random.rand(2)\n"
           "print(type(x).__name__)\n"
           "print(repr(float(x[0])), repr(float(x[1_500_000])), repr(float(x[-1])))\n"
           "print(repr(a.tolist()), repr(b.tolist()))\n"
           "print(repr(float(np.max(x))), int(np.argmax(x)))\n")
    r = asyncio.run(executor.execute(src, env={"BCI_DEBUG_FALLBACK": "1"}))
    assert r.exit_code == 0 and r.stderr == "", r.stderr
    np.random.seed(20240921)
    a = np.random.rand(5); x = np.random.rand(3_000_001); b = np.random.rand(2)
    lines = r.stdout.splitlines()
    assert lines[0] == "DeviceArray"
    assert lines[1] == f"{float(x[0])!r} {float(x[1_500_000])!r} {float(x[-1])!r}"
    assert lines[2] == f"{a.tolist()!r} {b.tolist()!r}"
    assert lines[3] == f"{float(np.max(x))!r} {int(np.argmax(x))}"


def test_native_broker_survives_hostile_requests(executor):
    """Same hostile sequence as the CPU test of the Python stand-in, against libbci.so's broker."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent))
    from test_broker_fuzz import hostile_requests

    hostile_requests(executor.broker.socket_path)
    r = asyncio.run(executor.execute(W.benchmark_numpy_source(1_000_000)))
    assert r.exit_code == 0 and r.stdout.startswith("Result:")
