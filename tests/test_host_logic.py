"""Host-side logic on CPU, with tests/fake_device.py standing in for the GPU binding:
numpy proxy semantics and fusion decisions, import hook, broker batching and buffer
ownership, executor process model replayed against the reference-app fixtures, and the
content-addressed storage protocol."""
import asyncio
import hashlib
import json
import sys
import threading
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
import fake_device  # noqa: E402

from bee_code_interpreter_b200 import _lib as L  # noqa: E402
from bee_code_interpreter_b200 import numpy_proxy as P  # noqa: E402
from bee_code_interpreter_b200.backend import LocalBackend  # noqa: E402

GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.fixture
def be(built):
    fake_device.reset()
    return LocalBackend(fake_device)


def names(calls):
    return [c[0] for c in calls]


# ---- proxy ------------------------------------------------------------------------------
def test_benchmark_numpy_expression_is_two_launches(be):
    x = P.device_rand(be, (100_000,), seed=7)
    r = np.sum(np.square(x))                       # examples/benchmark-numpy.py:21
    # ONE launch: the generator runs inside the reduction, neither rand(n) nor its square reaches memory
    assert names(fake_device.calls) == ["reduce_rand"] and x._ptr is None
    assert fake_device.calls[0][2] == (L.RED_SUM, L.PRE_SQUARE)
    assert r == np.sum(np.square(np.asarray(x)))   # reading x materialises the SAME stream (same seed)
    assert names(fake_device.calls)[1] == "rand" and isinstance(r, np.float64)
    assert float(x.max()) == np.asarray(x).max() and names(fake_device.calls)[-1] == "reduce"   # materialised now: plain reduce


def test_c3_request_mix_fuses_into_three_launches(be):
    rng = np.random.default_rng(0)
    xh, yh = rng.random(5000, dtype=np.float32), rng.random(5000, dtype=np.float32)
    x, y = P.DeviceArray.from_host(be, xh), P.DeviceArray.from_host(be, yh)
    fake_device.calls.clear()
    r1 = np.sum(np.square(x))
    r2 = np.mean(x * y + 1.0)
    r3 = np.max(np.abs(x - y))
    assert names(fake_device.calls) == ["reduce", "reduce_fused", "reduce_fused"]
    assert fake_device.calls[1][2] == (L.RED_MEAN, L.EW_MUL, L.PRE_NONE, True)
    assert fake_device.calls[2][2] == (L.RED_MAX, L.EW_SUB, L.PRE_ABS, False)
    assert np.isclose(r1, np.sum(np.square(xh)), rtol=1e-6)
    assert np.isclose(r2, np.mean(xh * yh + 1.0), rtol=1e-6)
    assert r3 == np.max(np.abs(xh - yh))
    assert r2.dtype == np.float32 and r3.dtype == np.float32


def test_deferred_value_used_twice_is_materialised_once(be):
    x = P.DeviceArray.from_host(be, np.arange(10, dtype=np.float64))
    fake_device.calls.clear()
    sq = np.square(x)
    total = sq.sum()
    first = sq[0:3]                                 # needs the array itself
    assert total == 285.0 and list(first) == [0.0, 1.0, 4.0]
    assert names(fake_device.calls).count("ewise") == 1


def test_mutation_forces_dependents_first(be):
    x = P.DeviceArray.from_host(be, np.arange(8, dtype=np.float64))
    sq = np.square(x)          # deferred on x
    x[0] = 100.0               # numpy semantics: sq was computed from the OLD x
    assert np.array_equal(np.asarray(sq), np.square(np.arange(8, dtype=np.float64)))
    assert np.asarray(x)[0] == 100.0
    assert np.sum(x) == 100.0 + sum(range(1, 8))


def test_scalar_typing_follows_nep50(be):
    x32 = P.DeviceArray.from_host(be, np.ones(4, dtype=np.float32))
    assert (x32 * 2.5).dtype == np.float32                       # python float is weak
    r = x32 * np.float64(2.5)                                    # numpy float64 scalar promotes: host path
    assert isinstance(r, np.ndarray) and r.dtype == np.float64
    xi = P.DeviceArray.from_host(be, np.arange(4, dtype=np.int32))
    assert isinstance(xi + 1, P.DeviceArray) and (xi + 1).dtype == np.int32
    r = xi * 0.5                                                 # int array * python float -> float64 on host
    assert isinstance(r, np.ndarray) and r.dtype == np.float64
    assert np.array_equal(np.asarray(2 - xi), 2 - np.arange(4, dtype=np.int32))   # reflected subtract
    assert np.array_equal(np.asarray(1.0 / x32), np.ones(4, dtype=np.float32))


def test_unsupported_calls_fall_back_to_numpy_results(be):
    h = np.random.default_rng(3).random((6, 5))
    x = P.DeviceArray.from_host(be, h)
    assert np.array_equal(np.sum(x, axis=0), h.sum(axis=0))
    assert np.isclose(np.std(x), h.std())
    assert np.array_equal(np.asarray(x.T), h.T)
    assert np.array_equal(np.where(x > 0.5, 1, 0), np.where(h > 0.5, 1, 0))
    assert x.tolist() == h.tolist()
    assert np.array_equal(np.concatenate([x, x]), np.concatenate([h, h]))
    b = P.DeviceArray.from_host(be, np.array([1, 2, 3], dtype=np.int64))
    assert np.array_equal(np.asarray(b + np.array([1, 1, 1])), np.array([2, 3, 4]))  # mixed device/host operands
    assert str(b) == str(np.array([1, 2, 3])) and len(b) == 3 and float(b[1]) == 2.0


def test_sort_argmax_cumsum_matmul_entry_points(be):
    h = np.random.default_rng(4).random(1000).astype(np.float32)
    x = P.DeviceArray.from_host(be, h)
    assert np.array_equal(np.asarray(np.sort(x)), np.sort(h))
    assert np.argmax(x) == np.argmax(h) and x.argmin() == np.argmin(h)
    k = P.DeviceArray.from_host(be, np.arange(100, dtype=np.int32))
    cs = np.cumsum(k)
    assert cs.dtype == np.int64 and np.array_equal(np.asarray(cs), np.cumsum(np.arange(100, dtype=np.int32)))
    idx = np.argsort(x, kind="stable")
    assert np.array_equal(np.asarray(idx), np.argsort(h, kind="stable"))
    a = P.DeviceArray.from_host(be, np.random.default_rng(5).random((8, 6), dtype=np.float32))
    b = P.DeviceArray.from_host(be, np.random.default_rng(6).random((6, 4), dtype=np.float32))
    c = a @ b
    assert isinstance(c, P.DeviceArray) and c.shape == (8, 4)
    assert np.allclose(np.asarray(c), np.asarray(a) @ np.asarray(b), rtol=1e-6)
    x.sort()
    assert np.array_equal(np.asarray(x), np.sort(h))


def test_median_quantile_percentile_are_one_sort_and_a_few_elements(be):
    """np.median / np.quantile / np.percentile of a DeviceArray: one device sort, the two neighbouring order statistics
    (and the last element, for NaN) come back, numpy interpolates — values, dtypes and result types are numpy's."""
    import warnings
    rng = np.random.default_rng(0)
    calls = ((np.median, ()), (np.quantile, (0.3,)), (np.quantile, (np.float32(0.25),)), (np.quantile, (1,)), (np.quantile, (0.0,)),
             (np.percentile, (37,)), (np.percentile, (100,)), (np.percentile, (0,)), (np.percentile, (np.float32(12.5),)),
             (np.quantile, ([0.1, 0.5, 0.9],)), (np.percentile, (np.array([5.0, 95.0]),)), (np.percentile, ([25, 50, 75],)))
    for dt in (np.float32, np.float64, np.int32, np.int64):
        for n in (2, 3, 10, 11, 1000, 1001):
            a = (rng.standard_normal(n) * 100).astype(dt)
            for f, args in calls:
                x = P.DeviceArray.from_host(be, a)
                fake_device.calls.clear()
                want, got = f(a, *args), f(x, *args)
                assert names(fake_device.calls).count("sort") == 1, (f.__name__, args)
                assert type(want) is type(got) and np.asarray(want).dtype == np.asarray(got).dtype and np.shape(want) == np.shape(got)
                assert np.array_equal(np.asarray(want), np.asarray(got)), (dt, n, f.__name__, args)
    a = np.array([1.0, np.nan, 3.0, 2.0], dtype=np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for f, args in ((np.median, ()), (np.quantile, (0.5,)), (np.quantile, ([0.5, 0.2],))):
            x = P.DeviceArray.from_host(be, a)
            want, got = f(a, *args), f(x, *args)
            assert np.array_equal(np.asarray(want), np.asarray(got), equal_nan=True) and np.asarray(want).dtype == np.asarray(got).dtype
    # what the device path does not take falls back to numpy: other methods, 2-D, weights
    x = P.DeviceArray.from_host(be, a)
    assert np.quantile(x, 0.5, method="nearest") is not None and np.isnan(np.quantile(x, 0.5, method="nearest"))


def test_power_clip_ptp_average_norm_and_nan_reductions_stay_on_the_device(be):
    """`x ** 2` / np.power(x, 2) are np.square (numpy's own fast paths), `x ** 0.5` is np.sqrt; np.clip is maximum + minimum;
    np.ptp, np.average, np.linalg.norm and the nan-reductions of NaN-free data run as the plain reductions."""
    rng = np.random.default_rng(1)

    def same(want, got, exact=True):
        w, g = np.asarray(want), np.asarray(got)
        assert w.dtype == g.dtype and w.shape == g.shape
        assert np.array_equal(w, g, equal_nan=True) if exact else np.allclose(w, g, rtol=1e-6, equal_nan=True)

    for dt in (np.float32, np.float64, np.int32, np.int64):
        a = (rng.standard_normal(5000) * 100).astype(dt)
        x = P.DeviceArray.from_host(be, a)
        is_int = np.dtype(dt).kind == "i"
        fake_device.calls.clear()
        same(a ** 2, x ** 2); same(np.power(a, 2), np.power(x, 2)); same(np.sum(a ** 2), np.sum(x ** 2), exact=is_int)
        same(np.clip(a, -10, 20), np.clip(x, -10, 20)); same(a.clip(-10, 20), x.clip(-10, 20))
        same(np.clip(a, 0, None), np.clip(x, 0, None)); same(np.clip(a, min=-5, max=5), np.clip(x, min=-5, max=5))
        same(np.ptp(a), np.ptp(x)); same(np.average(a), np.average(x), exact=False)
        same(np.nansum(a), np.nansum(x), exact=is_int); same(np.nanmax(a), np.nanmax(x)); same(np.nanmin(a), np.nanmin(x))
        same(np.nanmean(a), np.nanmean(x), exact=False)
        assert isinstance(x ** 2, P.DeviceArray) and isinstance(np.clip(x, -1, 1), P.DeviceArray)
        if not is_int:
            same(np.abs(a) ** 0.5, np.abs(x) ** 0.5); same(np.power(a, 2.0), np.power(x, 2.0)); same(a ** 3, x ** 3)
            same(np.linalg.norm(a), np.linalg.norm(x), exact=False)
            assert isinstance(np.abs(x) ** 0.5, P.DeviceArray)
        else:
            same(a ** 2.0, x ** 2.0)          # int ** float promotes to float64: numpy evaluates it
        kinds = set(names(fake_device.calls))
        assert {"reduce", "ewise"} <= kinds and (is_int or "reduce_fused" in kinds)
    a = np.array([1.0, np.nan, 3.0, 2.0] * 1000)
    x = P.DeviceArray.from_host(be, a)
    same(np.nansum(a), np.nansum(x)); same(np.nanmax(a), np.nanmax(x)); same(np.nanmean(a), np.nanmean(x))   # NaNs: numpy's masked path
    same(np.clip(a, 0, 2), np.clip(x, 0, 2))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        same(np.nanmedian(a), np.nanmedian(x)); same(np.nanvar(a), np.nanvar(x)); same(np.nanpercentile(a, 50), np.nanpercentile(x, 50))
    # NaN-free data: the nan- order statistics / moments and inner / vdot are the plain device entry points
    for dt in (np.float32, np.float64, np.int64):
        a, b = (rng.standard_normal(4001) * 50).astype(dt), (rng.standard_normal(4001) * 50).astype(dt)
        x, y = P.DeviceArray.from_host(be, a), P.DeviceArray.from_host(be, b)
        fake_device.calls.clear()
        same(np.nanmedian(a), np.nanmedian(x)); same(np.nanpercentile(a, 30), np.nanpercentile(x, 30))
        same(np.nanquantile(a, [0.2, 0.8]), np.nanquantile(x, [0.2, 0.8]))
        if dt is np.int64:
            same(np.inner(a, b), np.inner(x, y)); same(np.vdot(a, b), np.vdot(x, y))
        else:
            tol = 1e-5 if dt is np.float32 else 1e-12
            assert np.isclose(np.nanvar(x), np.nanvar(a), rtol=tol) and np.isclose(np.nanstd(x, ddof=1), np.nanstd(a, ddof=1), rtol=tol)
            assert np.isclose(np.inner(x, y), np.inner(a, b), rtol=tol, atol=tol * 1e4) and type(np.vdot(x, y)) is type(np.vdot(a, b))
        assert {"sort", "reduce_fused"} <= set(names(fake_device.calls))


def test_buffers_are_released(be):
    x = P.device_rand(be, (1000,), seed=1)
    y = np.square(x) + 1.0
    _ = np.asarray(y)
    del x, y
    import gc

    gc.collect()
    assert fake_device.live_buffers() == 0


# ---- import hook ----------------------------------------------------------------------------
def test_hook_gives_agent_code_the_view_and_libraries_the_real_module(be):
    from bee_code_interpreter_b200 import numpy_hook as H

    H._state.update(backend=None, backend_factory=lambda: be, view=None, seeded=False, min_rand=100)
    agent = {"__name__": "__main__"}
    lib = {"__name__": "pandas.core.frame"}
    view = H.patched_import("numpy", agent, None, None, 0)
    assert view is not np and view.pi == np.pi and view.ndarray is np.ndarray
    assert H.patched_import("numpy", lib, None, None, 0) is np
    assert H.patched_import("numpy", agent, None, None, 0) is view                     # idempotent
    assert H.patched_import("numpy.random", agent, None, ("rand",), 0) is view.random
    assert H.patched_import("json", agent, None, None, 0).__name__ == "json"
    x = view.random.rand(500)
    assert isinstance(x, P.DeviceArray) and x.shape == (500,)
    assert isinstance(view.random.rand(10), np.ndarray)                                   # below the offload floor
    assert isinstance(view.random.rand(), float)
    # seeded => numpy's MT19937 stream, continued on the device and handed back: draws before,
    # on and after the device stay on ONE sequence (odd number of host draws first: pairs straddle)
    view.random.seed(42)
    h0 = view.random.rand(7)                                                              # host (small)
    a = view.random.rand(500)                                                             # device
    h1 = view.random.rand(3)                                                              # host again
    b = view.random.random_sample((10, 60))                                               # device, 2-D
    np.random.seed(42)
    want = np.random.rand(7 + 500 + 3 + 600)
    assert isinstance(h0, np.ndarray) and isinstance(a, P.DeviceArray) and isinstance(b, P.DeviceArray)
    assert np.array_equal(np.concatenate([h0, np.asarray(a), h1, np.asarray(b).ravel()]), want)
    assert ("rand_mt", 1, None) in fake_device.calls
    big = np.random.default_rng(0).random(P.OFFLOAD_MIN_ELEMS).astype(np.float32)
    s = view.sort(big)
    assert isinstance(s, P.DeviceArray) and np.array_equal(np.asarray(s), np.sort(big))
    small = big[:100]
    assert isinstance(view.sort(small), np.ndarray)
    assert np.array_equal(view.cumsum([1, 2, 3]), [1, 3, 6])                             # lists untouched
    H._state.update(backend=None, backend_factory=None, view=None, seeded=False, min_rand=1 << 16)


# ---- broker -----------------------------------------------------------------------------------
def test_broker_batches_concurrent_clients_and_scopes_buffers(built):
    from bee_code_interpreter_b200.broker import GpuBroker
    from bee_code_interpreter_b200.broker_client import RemoteBackend

    fake_device.reset()
    broker = GpuBroker(fake_device, coalesce_us=20_000)
    try:
        n_clients = 6
        results = [None] * n_clients
        barrier = threading.Barrier(n_clients)

        def client(i):
            rb = RemoteBackend(broker.socket_path)
            x = P.device_rand(rb, (2000 + i,), seed=100 + i)
            barrier.wait()
            results[i] = float(np.sum(np.square(x)))
            del x
            rb.close()

        threads = [threading.Thread(target=client, args=(i,)) for i in range(n_clients)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(30)
        from oracle import numpy_oracle as O

        for i in range(n_clients):
            assert np.isclose(results[i], np.sum(np.square(O.rand_f64(2000 + i, 100 + i))), rtol=1e-12)
        red = [c for c in fake_device.calls if c[0] == "reduce_rand"]
        assert sum(c[1] for c in red) == n_clients and max(c[1] for c in red) >= 2   # packed into shared launches
        # ownership: a client cannot name another client's buffer
        a, b = RemoteBackend(broker.socket_path), RemoteBackend(broker.socket_path)
        pa = a.rand(100, 1)
        with pytest.raises(L.BciError):
            b.reduce(L.RED_SUM, np.float64, [pa], [100])
        with pytest.raises(L.BciError):
            a.reduce(L.RED_SUM, np.float64, [pa], [101])                              # past the end of its own buffer
        assert np.isclose(a.reduce(L.RED_SUM, np.float64, [pa], [100])[0], O.rand_f64(100, 1).sum())
        # big arrays cross behind a passed descriptor (SCM_RIGHTS), small ones inline
        big = np.random.default_rng(1).random(50_000)
        pb = a.upload(big)
        assert np.array_equal(a.download(pb, big.size, big.dtype), big)
        # Generator(PCG64) draws, creation, casts over the wire
        st64 = np.random.PCG64(9).state["state"]
        pg, rej = a.rand_pcg(L.PCG_F32, 3001, st64["state"], st64["inc"])
        assert rej == 0 and np.array_equal(a.download(pg, 3001, np.float32), np.random.default_rng(9).random(3001, dtype=np.float32))
        pi, rej = a.rand_pcg(L.PCG_BOUNDED32_I64, 500, st64["state"], st64["inc"], -5, 2**31)
        assert rej == 0 and np.array_equal(a.download(pi, 500, np.int64), np.random.default_rng(9).integers(-5, 2**31 - 5, 500))
        pf = a.fill_new(L.FILL_ARANGE, np.float64, 1000, 0.5, 0.25, 0.75)
        assert np.array_equal(a.download(pf, 1000, np.float64), np.arange(0.5, 250.5, 0.25))
        pc = a.cast_new(np.float64, np.float32, pf, 1000)
        assert np.array_equal(a.download(pc, 1000, np.float32), np.arange(0.5, 250.5, 0.25).astype(np.float32))
        with pytest.raises(L.BciError):
            b.cast_new(np.float64, np.float32, pf, 1000)                               # not b's buffer
        # export: array -> file behind a header, digest of the file computed on the "device"
        import os, tempfile
        fd, path = tempfile.mkstemp()
        digest = a.export(pf, 8000, fd, b"HEADER")
        os.close(fd)
        blob = open(path, "rb").read()
        os.unlink(path)
        assert blob == b"HEADER" + np.arange(0.5, 250.5, 0.25).tobytes() and digest == hashlib.sha256(blob).digest()
        fd, path = tempfile.mkstemp()
        os.write(fd, b"x" * 100_000)
        assert a.sha256_fd(fd) == hashlib.sha256(b"x" * 100_000).digest()
        os.close(fd); os.unlink(path)
        # seeded legacy generator over the wire: state in, same stream + advanced state out
        rs = np.random.RandomState(1234)
        rs.random_sample(5)                                                           # odd position in the block
        st = rs.get_state()
        state = np.concatenate([st[1], [st[2]]]).astype(np.uint32)
        pm, after = a.rand_mt19937(1001, state)
        assert np.array_equal(a.download(pm, 1001, np.float64), rs.random_sample(1001))
        st2 = rs.get_state()
        assert np.array_equal(after[:624], st2[1]) and int(after[624]) == st2[2]
        with pytest.raises(L.BciError):
            a.rand_mt19937(10, state[:100])
        a.close(); b.close()
        import time

        for _ in range(100):
            if fake_device.live_buffers() == 0:
                break
            time.sleep(0.02)
        assert fake_device.live_buffers() == 0                                          # disconnect frees everything
    finally:
        broker.close()


# ---- executor -------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def executor(built):
    from bee_code_interpreter_b200.executor import LocalGpuExecutor

    fake_device.reset()
    ex = LocalGpuExecutor(device_module=fake_device, max_inflight=8, timeout=5, zygotes=1, min_rand_elems=1000)
    yield ex
    ex.close()


def run(ex, case):
    b = case["request"]
    return asyncio.run(ex.execute(b["source_code"], files=b.get("files", {}), env=b.get("env", {})))


def test_executor_matches_reference_app_fixtures(executor):
    cases = {c["name"]: c for c in json.loads((GOLDEN / "http_replay.json").read_text())["cases"]}
    for name in ("hello", "env", "sys_exit_3", "stderr_and_unicode", "numpy_small"):
        r, want = run(executor, cases[name]), cases[name]["response"]
        assert (r.stdout, r.stderr, r.exit_code, dict(r.files)) == (want["stdout"], want["stderr"], want["exit_code"], {}), name
    r = run(executor, cases["create_file"])
    assert r.exit_code == 0 and set(r.files) == {"/workspace/file.txt"}
    assert r.files["/workspace/file.txt"] == hashlib.sha256(b"Hello, World!").hexdigest()   # content addressed
    back = asyncio.run(executor.execute(cases["read_file_back"]["request"]["source_code"], files=dict(r.files)))
    assert back.stdout == cases["read_file_back"]["response"]["stdout"] and back.exit_code == 0 and not back.files
    r = run(executor, cases["crash"])
    assert r.exit_code == 1 and r.stdout == "" and "ZeroDivisionError: division by zero" in r.stderr
    r = run(executor, cases["subdir_not_listed"])
    assert set(r.files) == {"/workspace/top.bin"}
    assert r.files["/workspace/top.bin"] == hashlib.sha256(bytes(range(256))).hexdigest()
    r = run(executor, cases["timeout"])
    assert (r.stdout, r.stderr, r.exit_code) == ("", "Execution timed out", -1)
    for bad in ("bad_path", "bad_hash"):
        with pytest.raises(ValueError):
            run(executor, cases[bad])
    with pytest.raises(FileNotFoundError):
        run(executor, cases["unknown_object"])
    r = asyncio.run(executor.execute("import os, signal\nos.kill(os.getpid(), signal.SIGKILL)"))
    assert r.exit_code == -1                                                            # server.rs:165


def test_executor_runs_the_metric_script_on_the_device_backend(executor):
    from bee_code_interpreter_b200 import workloads as W

    fake_device.calls.clear()
    n = 50_000
    res = asyncio.run(_many(executor, W.benchmark_numpy_source(n), 12))
    for r in res:
        assert r.exit_code == 0, r.stderr
        lines = r.stdout.splitlines()
        assert lines[0].startswith("Result: ") and lines[1].startswith("Execution Time: ")
        assert abs(float(lines[0].split()[1]) / n - 1 / 3) < 6 * np.sqrt(4 / 45 / n)
    kinds = names(fake_device.calls)
    assert set(kinds) == {"reduce_rand"}                      # one fused launch class on the "device", nothing else
    assert sum(c[1] for c in fake_device.calls if c[0] == "reduce_rand") == 12


async def _many(ex, src, k):
    return await asyncio.gather(*(ex.execute(src) for _ in range(k)))


# ---- storage -----------------------------------------------------------------------------------------
def test_content_addressed_storage_protocol(built, tmp_path):
    from bee_code_interpreter_b200.storage import ContentAddressedStorage, HashService

    fake_device.reset()
    st = ContentAddressedStorage(str(tmp_path / "objects"), HashService(fake_device, coalesce_us=5000))

    async def scenario():
        payloads = [np.random.default_rng(i).bytes(n) for i, n in enumerate([0, 1, 63, 64, 65, 70_000, 9_000_000])]

        async def put(p):
            async with st.writer() as w:
                for off in range(0, len(p), 65536):                                     # httpx-sized chunks
                    await w.write(p[off:off + 65536])
            return w.hash                                                                # read after exit, like the reference

        ids = await asyncio.gather(*(put(p) for p in payloads))
        for p, i in zip(payloads, ids):
            assert i == hashlib.sha256(p).hexdigest()
            assert await st.exists(i) and await st.read(i) == p
        assert await st.write(payloads[3]) == ids[3]                                    # dedupe, same id
        assert len(list((tmp_path / "objects").iterdir())) == len(set(ids))           # no temp files left
        with pytest.raises(FileNotFoundError):
            await st.read("0" * 64)
        with pytest.raises(ValueError):
            await st.read("../etc/passwd")
        async with st.reader(ids[2]) as f:
            assert await f.read() == payloads[2]

    asyncio.run(scenario())
    assert max(c[1] for c in fake_device.calls if c[0].startswith("sha256")) >= 2        # writers hashed together
    st.hasher.close()


def test_in_place_operators_and_out_argument(be):
    h = np.arange(10, dtype=np.float64)
    x = P.DeviceArray.from_host(be, h.copy())
    y = P.DeviceArray.from_host(be, np.ones(10))
    sq = np.square(x)                      # deferred on the OLD x
    x += y
    x *= 2.0
    assert isinstance(x, P.DeviceArray)
    assert np.array_equal(np.asarray(x), (h + 1) * 2)
    assert np.array_equal(np.asarray(sq), h * h)            # not affected by the in-place update
    np.subtract(x, y, out=x)
    assert np.array_equal(np.asarray(x), (h + 1) * 2 - 1)
    x /= 4                                   # true divide of float array by python int stays float64
    assert np.allclose(np.asarray(x), ((h + 1) * 2 - 1) / 4)
    z = np.zeros(10)
    np.add(x, y, out=z)                      # host `out` with device inputs: numpy's own path
    assert np.allclose(z, np.asarray(x) + 1)


def test_random_expressions_match_numpy(be):
    """Fuzz: random trees of supported operations on random arrays, evaluated through the proxy
    (deferred / fused / materialised as the proxy decides) and by plain numpy."""
    from hypothesis import given, settings, strategies as st

    dtypes = [np.float32, np.float64, np.int32, np.int64]
    unary = [np.square, np.abs, np.negative, lambda v: v ** 2, lambda v: np.power(v, 2), lambda v: np.clip(v, -3, 4),
             lambda v: v.clip(min=-1), lambda v: np.clip(v, None, 2)]
    binary = [np.add, np.subtract, np.multiply, np.maximum, np.minimum]
    reducers = [np.sum, np.max, np.min, np.mean, np.ptp, np.nansum, np.nanmax, np.median, lambda v: np.percentile(v, 30), None, None]

    @settings(max_examples=60, deadline=None)
    @given(seed=st.integers(0, 2**31 - 1), depth=st.integers(1, 4), di=st.integers(0, 3), n=st.integers(1, 3000),
           ri=st.integers(0, len(reducers) - 1))
    def run(seed, depth, di, n, ri):
        rng = np.random.default_rng(seed)
        dt = dtypes[di]
        mk = (lambda: rng.integers(-50, 50, n, dtype=dt)) if np.issubdtype(dt, np.integer) else (lambda: ((rng.random(n) - 0.5) * 20).astype(dt))
        hosts = [mk() for _ in range(3)]
        devs = [P.DeviceArray.from_host(be, h.copy()) for h in hosts]
        hv, dv = hosts[0], devs[0]
        for _ in range(depth):
            c = int(rng.integers(0, 3))
            if c == 0:
                f = unary[int(rng.integers(0, len(unary)))]
                hv, dv = f(hv), f(dv)
            elif c == 1:
                f = binary[int(rng.integers(0, len(binary)))]
                j = int(rng.integers(0, 3))
                hv, dv = f(hv, hosts[j]), f(dv, devs[j])
            else:
                f = binary[int(rng.integers(0, 3))]
                sc = int(rng.integers(-3, 4)) if np.issubdtype(dt, np.integer) else float(rng.integers(-3, 4)) + 0.5
                hv, dv = f(hv, sc), f(dv, sc)
        red = reducers[ri]
        with np.errstate(all="ignore"):
            if red is None:
                got, want = np.asarray(dv), hv
                assert got.dtype == want.dtype and np.array_equal(got, want)
            else:
                got, want = red(dv), red(hv)
                assert np.asarray(got).dtype == np.asarray(want).dtype
                if np.issubdtype(dt, np.integer) and red in (np.sum, np.max, np.min, np.ptp, np.nansum, np.nanmax):
                    assert got == want
                else:
                    assert np.isclose(float(got), float(want), rtol=1e-5 if dt == np.float32 else 1e-11, atol=1e-6)

    run()


def test_var_std_and_vector_dot_run_as_fused_reductions(be):
    rng = np.random.default_rng(11)
    for dt, rtol in ((np.float64, 1e-12), (np.float32, 1e-5)):
        h = (rng.standard_normal(20_001) * 3 + 5).astype(dt)
        g = rng.standard_normal(20_001).astype(dt)
        x, y = P.DeviceArray.from_host(be, h), P.DeviceArray.from_host(be, g)
        fake_device.calls.clear()
        v, s, v1 = np.var(x), np.std(x), x.var(ddof=1)
        assert names(fake_device.calls) == ["reduce", "reduce_fused"] * 3          # two kernels each, nothing materialised
        assert fake_device.calls[1][2] == (L.RED_MEAN, L.EW_SUB, L.PRE_SQUARE, False)
        assert isinstance(v, dt) and isinstance(s, dt) and isinstance(v1, dt)
        assert np.isclose(v, np.var(h), rtol=rtol) and np.isclose(s, np.std(h), rtol=rtol) and np.isclose(v1, h.var(ddof=1), rtol=rtol)
        assert np.isclose(x.std(ddof=1), h.std(ddof=1), rtol=rtol)
        fake_device.calls.clear()
        d = np.dot(x, y)
        assert names(fake_device.calls) == ["reduce_fused"] and isinstance(d, dt)
        assert np.isclose(d, np.dot(h, g), rtol=1e-4 if dt == np.float32 else 1e-10)
        assert np.isclose(x.dot(y), np.dot(h, g), rtol=1e-4 if dt == np.float32 else 1e-10)
        assert np.isclose(np.var(x * 2.0), np.var(h * dt(2)), rtol=rtol)             # deferred input: materialised once
    # what the kernels do not take falls back to numpy with numpy's result
    m = P.DeviceArray.from_host(be, rng.standard_normal((50, 40)))
    assert np.allclose(np.var(m, axis=0), np.var(np.asarray(m), axis=0))
    k = P.DeviceArray.from_host(be, np.arange(100, dtype=np.int64))
    assert np.var(k) == np.var(np.arange(100)) and np.dot(k, k) == np.dot(np.arange(100), np.arange(100))
    k32 = P.DeviceArray.from_host(be, np.arange(100, dtype=np.int32))
    assert np.dot(k32, k32) == np.dot(np.arange(100, dtype=np.int32), np.arange(100, dtype=np.int32))


def test_changed_files_do_not_depend_on_the_coarse_clock(tmp_path):
    """_changed_files: a file written within the same kernel tick as the request start is still
    reported (snapshot comparison), untouched inputs are not."""
    import os
    import time

    from bee_code_interpreter_b200 import executor as E

    ws = tmp_path
    (ws / "input.txt").write_text("in")
    before = E._snapshot(ws)
    start = time.time_ns() + 50_000_000          # pretend the clock comparison alone would miss everything
    (ws / "new.txt").write_text("n")
    (ws / "input2.txt").write_text("x")
    assert E._changed_files(ws, start, before) == ["input2.txt", "new.txt"]
    os.utime(ws / "input.txt")                   # touched: ctime / mtime move
    assert "input.txt" in E._changed_files(ws, start, before)
    assert E._changed_files(ws, start, E._snapshot(ws)) == []


def test_ufunc_reduce_of_a_2d_array_keeps_numpys_axis_default(be):
    """ufunc.reduce defaults to axis=0: only 1-D arrays (or axis=None) are full reductions."""
    h = np.arange(12, dtype=np.float64).reshape(3, 4)
    x = P.DeviceArray.from_host(be, h.copy())
    assert np.array_equal(np.add.reduce(x), np.add.reduce(h))
    assert np.array_equal(np.maximum.reduce(x, axis=0), np.maximum.reduce(h, axis=0))
    assert np.add.reduce(x, axis=None) == h.sum()
    v = P.DeviceArray.from_host(be, np.arange(5, dtype=np.float64))
    fake_device.calls.clear()
    assert np.add.reduce(v) == 10.0 and names(fake_device.calls) == ["reduce"]


def test_writes_through_an_exposed_host_copy_reach_later_device_ops(be):
    """A view / np.asarray / mutating ndarray method hands out the host cache: it becomes the single
    source of truth, later device ops re-upload it (stale device copy)."""
    mk = lambda: P.DeviceArray.from_host(be, np.arange(8, dtype=np.float64))  # noqa: E731
    x = mk(); v = x[:4]; v[:] = 100; assert x.sum() == 422.0
    x = mk(); x.fill(0.5); assert x.sum() == 4.0
    x = mk(); a = np.asarray(x); a[0] = 50; assert x.max() == 50.0 and np.sum(np.square(x)) == float((a * a).sum())
    x = mk(); assert x[3] == 3.0 and x[[1, 2]].tolist() == [1.0, 2.0] and x._ptr is not None    # copies keep the device buffer
    x = mk(); np.cumsum(np.arange(8.0), out=x); assert x.sum() == float(np.cumsum(np.arange(8.0)).sum())
    x = mk(); y = np.square(x); x[0] = 7.0; assert np.asarray(y)[0] == 0.0                       # deferred reader saw the old value


# ---- round 2: Generator / creation / load / save through the hook, whole request path on the fake device ---------
CONFIG_SCRIPT = '''
import numpy as np
rng = np.random.default_rng(2005)
x = rng.random(2**20, dtype=np.float32)
y = rng.random(2**20, dtype=np.float32)
print(type(x).__name__, repr(float(np.sum(np.square(x)))), repr(float(np.mean(x * y + 1.0))), repr(float(np.max(np.abs(x - y)))))
k = np.random.default_rng(5001).integers(0, 2**31, 300_001, dtype=np.int64)
c = np.cumsum(k)
print(type(k).__name__, int(np.argmax(k)), int(c[-1]), repr(float(rng.random())), repr(rng.standard_normal(2).tolist()))
z = np.load("in.npy")
t = np.arange(z.size, dtype=np.float64)
print(type(z).__name__, type(t).__name__, repr(float(np.dot(z, t))), repr(float(np.linspace(0.0, 1.0, 400_001).sum())))
np.save("out.npy", z * 2.0)
(z + 1.0).tofile("raw.bin")
'''


def test_generator_creation_load_save_through_the_executor(executor):
    """A C3/C5-shaped request end to end on the fake device: default_rng draws, creation functions and np.load
    produce device arrays (descriptor passing included), np.save / tofile export them, results equal numpy's."""
    import io

    zin = np.random.default_rng(3).random(300_000)
    buf = io.BytesIO()
    np.save(buf, zin)

    async def go():
        h = await executor.file_storage.write(buf.getvalue())
        return await executor.execute(CONFIG_SCRIPT, files={"/workspace/in.npy": h}, env={"BCI_DEBUG_FALLBACK": "1"})

    fake_device.calls.clear()
    r = asyncio.run(go())
    assert r.exit_code == 0 and r.stderr == "", r.stderr
    rng = np.random.default_rng(2005)
    x, y = rng.random(2**20, dtype=np.float32), rng.random(2**20, dtype=np.float32)
    k = np.random.default_rng(5001).integers(0, 2**31, 300_001, dtype=np.int64)
    lines = r.stdout.splitlines()
    assert lines[0] == f"DeviceArray {float(np.sum(np.square(x)))!r} {float(np.mean(x * y + 1.0))!r} {float(np.max(np.abs(x - y)))!r}"
    assert lines[1] == f"DeviceArray {int(np.argmax(k))} {int(np.cumsum(k)[-1])} {float(rng.random())!r} {rng.standard_normal(2).tolist()!r}"
    t = np.arange(zin.size, dtype=np.float64)
    assert lines[2] == f"DeviceArray DeviceArray {float(np.dot(zin, t))!r} {float(np.linspace(0.0, 1.0, 400_001).sum())!r}"
    assert set(r.files) == {"/workspace/out.npy", "/workspace/raw.bin"}

    async def blobs():
        return [await executor.file_storage.read(r.files[f"/workspace/{n}"]) for n in ("out.npy", "raw.bin")]

    out_npy, raw = asyncio.run(blobs())
    assert np.array_equal(np.load(io.BytesIO(out_npy)), zin * 2.0) and raw == (zin + 1.0).tobytes()
    assert r.files["/workspace/out.npy"] == hashlib.sha256(out_npy).hexdigest()
    kinds = names(fake_device.calls)
    assert kinds.count("rand_pcg") == 3 and "fill" in kinds and "reduce_fused" in kinds and "cumsum" in kinds


PANDAS_SCIPY_SCRIPT = '''
import numpy as np
import pandas as pd
from scipy.stats import ttest_ind
import scipy.stats as st
rng = np.random.default_rng(42)
control = rng.random(300_000) * 2 + 10
experimental = rng.random(300_000) * 2 + 10.01
s = pd.Series(control)
print(type(control).__name__, type(s).__name__, isinstance(s, pd.Series), isinstance(pd.Series([1.0]), pd.Series))
print(repr(s.mean()), repr(pd.Series(experimental).mean()), repr(s.std()), repr(pd.Series(experimental).std()))
t, p = ttest_ind(control, experimental)
print(repr(float(t)), repr(float(p)), repr(float(st.ttest_ind(control, experimental, equal_var=False).statistic)))
print(repr(float(s.median())), repr(float(s.quantile(0.95))), repr(float(s.quantile(0.001))), repr(float(s.mode().iloc[0])))
'''


def test_pandas_and_scipy_entry_points_run_as_device_reductions(executor):
    """examples/using_imports.py:25-36 at a size where it matters: pd.Series(x).mean()/.std() and scipy's
    ttest_ind on device arrays run as fused reductions (no host copy), equal to pandas / scipy on the host
    within float64 summation-order tolerance; median / quantile are one device sort plus two elements and equal
    pandas bit for bit; anything else (mode) is pandas itself."""
    import pandas as pd
    import scipy.stats as st

    fake_device.calls.clear()
    r = asyncio.run(executor.execute(PANDAS_SCIPY_SCRIPT, env={"BCI_DEBUG_FALLBACK": "1"}))
    assert r.exit_code == 0 and r.stderr == "", r.stderr
    rng = np.random.default_rng(42)
    a, b = rng.random(300_000) * 2 + 10, rng.random(300_000) * 2 + 10.01
    lines = r.stdout.splitlines()
    assert lines[0] == "DeviceArray DeviceSeries True True"
    got = [float(t) for t in lines[1].split()]
    want = [pd.Series(a).mean(), pd.Series(b).mean(), pd.Series(a).std(), pd.Series(b).std()]
    assert np.allclose(got, want, rtol=1e-12, atol=0)
    t, p = st.ttest_ind(a, b)
    got = [float(x) for x in lines[2].split()]
    assert np.allclose(got, [t, p, st.ttest_ind(a, b, equal_var=False).statistic], rtol=1e-9, atol=0)
    sa = pd.Series(a)
    assert [float(v) for v in lines[3].split()] == [float(sa.median()), float(sa.quantile(0.95)), float(sa.quantile(0.001)), float(sa.mode().iloc[0])]
    kinds = names(fake_device.calls)
    assert kinds.count("rand_pcg") == 2 and "reduce_fused" in kinds and kinds.count("sort") == 3


# ---- creation functions and Generator draws through the view: host-side parameter derivation vs numpy -------------
def _view_with_fake(min_elems=64):
    from bee_code_interpreter_b200 import numpy_hook as H

    H._state["view"] = None
    H._state["backend"] = None
    H.install(backend_factory=lambda: LocalBackend(fake_device), min_rand_elems=min_elems)
    return H.numpy_view()


def test_arange_linspace_full_match_numpy_for_random_parameters(be, monkeypatch):
    from hypothesis import given, settings, strategies as st

    monkeypatch.setattr(P, "OFFLOAD_MIN_ELEMS", 64)
    vnp = _view_with_fake()
    finite = st.floats(-1e6, 1e6, allow_nan=False, allow_infinity=False)

    @settings(max_examples=150, deadline=None)
    @given(start=finite, span=st.floats(1.0, 5e3), step=st.floats(1e-3, 7.0), neg=st.booleans(),
           dt=st.sampled_from([None, np.float32, np.float64]))
    def floats(start, span, step, neg, dt):
        stop = start + span
        args = (stop, start, -step) if neg else (start, stop, step)
        kw = {} if dt is None else {"dtype": dt}
        want = np.arange(*args, **kw)
        got = vnp.arange(*args, **kw)
        assert got.dtype == want.dtype and got.shape == want.shape
        assert np.array_equal(np.asarray(got), want)

    @settings(max_examples=100, deadline=None)
    @given(start=st.integers(-10**6, 10**6), n=st.integers(1, 5000), step=st.integers(1, 9), dt=st.sampled_from([None, np.int32, np.int64, np.float32]))
    def ints(start, n, step, dt):
        kw = {} if dt is None else {"dtype": dt}
        want = np.arange(start, start + n * step, step, **kw)
        got = vnp.arange(start, start + n * step, step, **kw)
        assert got.dtype == want.dtype and np.array_equal(np.asarray(got), want)

    @settings(max_examples=100, deadline=None)
    @given(start=finite, stop=finite, num=st.integers(2, 4000), endpoint=st.booleans())
    def lin(start, stop, num, endpoint):
        want = np.linspace(start, stop, num, endpoint=endpoint)
        got = vnp.linspace(start, stop, num, endpoint=endpoint)
        assert np.array_equal(np.asarray(got), want)

    floats(); ints(); lin()
    for shape, val, dt in [((100,), 2.5, None), ((8, 16), 7, np.int32), (300, -1.0, np.float32), ((70,), 3, None)]:
        kw = {} if dt is None else {"dtype": dt}
        got, want = vnp.full(shape, val, **kw), np.full(shape, val, **kw)
        assert type(got) is P.DeviceArray and got.dtype == want.dtype and got.shape == want.shape and np.array_equal(np.asarray(got), want)
    assert type(vnp.zeros(10)) is np.ndarray and type(vnp.arange(5)) is np.ndarray          # small: numpy itself


def test_generator_draws_track_numpys_stream_for_random_call_sequences(be):
    """Any interleaving of device-sized and scalar draws leaves the Generator exactly where numpy's would be
    (values and bit-generator state, including the buffered 32-bit word)."""
    from hypothesis import given, settings, strategies as st

    vnp = _view_with_fake(min_elems=32)
    call = st.one_of(
        st.tuples(st.just("random64"), st.integers(1, 300)), st.tuples(st.just("random32"), st.integers(1, 301)),
        st.tuples(st.just("integers"), st.integers(1, 300), st.integers(-50, 50), st.sampled_from([2, 3, 100, 2**16, 2**31, 2**32 - 1])),
        st.tuples(st.just("int32"), st.integers(1, 300)), st.tuples(st.just("uniform"), st.integers(1, 300)),
        st.tuples(st.just("normal"), st.integers(1, 5)), st.tuples(st.just("scalar")), st.tuples(st.just("bytes"), st.integers(1, 50)))

    @settings(max_examples=120, deadline=None)
    @given(seed=st.integers(0, 2**32), calls=st.lists(call, min_size=1, max_size=8))
    def run(seed, calls):
        g, r = vnp.random.default_rng(seed), np.random.default_rng(seed)
        assert isinstance(g, np.random.Generator)
        for c in calls:
            if c[0] == "random64":
                a, b = g.random(c[1]), r.random(c[1])
            elif c[0] == "random32":
                a, b = g.random(c[1], dtype=np.float32), r.random(c[1], dtype=np.float32)
            elif c[0] == "integers":
                a, b = g.integers(c[2], c[2] + c[3], c[1]), r.integers(c[2], c[2] + c[3], c[1])
            elif c[0] == "int32":
                a, b = g.integers(0, 1000, c[1], dtype=np.int32), r.integers(0, 1000, c[1], dtype=np.int32)
            elif c[0] == "uniform":
                a, b = g.uniform(-2.0, 3.5, c[1]), r.uniform(-2.0, 3.5, c[1])
            elif c[0] == "normal":
                a, b = g.standard_normal(c[1]), r.standard_normal(c[1])
            elif c[0] == "bytes":
                a, b = np.frombuffer(g.bytes(c[1]), np.uint8), np.frombuffer(r.bytes(c[1]), np.uint8)
            else:
                a, b = np.array(g.random()), np.array(r.random())
            a = np.asarray(a)
            assert a.dtype == b.dtype and np.array_equal(a, b), c
            assert g.bit_generator.state == r.bit_generator.state, c

    run()
