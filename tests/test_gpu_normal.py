"""K8c on the device: same stream as oracle/philox.c up to the ulps of log / sincos, right
distribution, loc/scale applied with numpy's two roundings, batched ragged requests."""
import numpy as np
import pytest

from oracle import numpy_oracle as O

pytestmark = pytest.mark.gpu


def test_device_normals_match_the_oracle_stream_and_distribution(dev):
    from scipy import stats

    sizes = [1, 2, 3, 4095, 4096, 4097, 100_003, 2_000_001]
    params = [(0.0, 1.0), (10.0, 2.0), (-3.5, 0.25), (0.0, 1.0), (1e6, 1e-3), (0.0, 1.0), (0.0, 0.0), (0.0, 1.0)]
    outs = [dev.alloc(8 * n) for n in sizes]
    seeds = [1000 + i for i in range(len(sizes))]
    dev.rand_normal_f64(outs, sizes, seeds, [p[0] for p in params], [p[1] for p in params])
    for n, p, s, (loc, scale) in zip(sizes, outs, seeds, params):
        got = dev.download(p, n, np.float64)
        want = O.rand_normal_f64(n, s, loc, scale)
        # |z| <= ~8.6 for 53-bit uniforms; log/sincos of libdevice and glibc differ by a few ulp
        assert np.all(np.abs(got - want) <= 1e-12 * (abs(loc) + abs(scale) * 10 + 1e-300) + 1e-300), (n, loc, scale)
    z = dev.download(outs[-1], sizes[-1], np.float64)
    assert abs(z.mean()) < 5 / np.sqrt(z.size) and abs(z.var() - 1) < 5 * np.sqrt(2 / z.size)
    assert abs(stats.skew(z)) < 0.01 and abs(stats.kurtosis(z)) < 0.02
    assert stats.kstest(z[:200_000], "norm").pvalue > 1e-4
    assert np.isfinite(z).all() and np.abs(z).max() < 9
    for p in outs:
        dev.free(p)


def test_script_normals_come_from_the_gpu(dev):
    """examples/using_imports.py:20-22 shape (seed + normal) at a size worth offloading: unseeded
    draws are DeviceArrays with the right moments, seeded ones are numpy's own stream."""
    import asyncio

    from bee_code_interpreter_b200.executor import LocalGpuExecutor

    src = ("import numpy as np\n"
           "x = np.random.normal(10, 2, 5_000_000)\n"
           "y = np.random.randn(1000, 3000)\n"
           "print(type(x).__name__, type(y).__name__, y.shape)\n"
           "print(abs(float(np.mean(x)) - 10) < 0.01, abs(float(np.std(x)) - 2) < 0.01, abs(float(np.mean(y))) < 0.01)\n"
           "np.random.seed(42)\n"
           "z = np.random.normal(loc=10, scale=2, size=300_000)\n"
           "print(type(z).__name__, repr(float(z[0])), repr(float(z.mean())))\n")
    ex = LocalGpuExecutor(max_inflight=4, timeout=60, zygotes=1)
    try:
        r = asyncio.run(ex.execute(src, env={"BCI_DEBUG_FALLBACK": "1"}))
    finally:
        ex.close()
    assert r.exit_code == 0 and r.stderr == "", r.stderr
    lines = r.stdout.splitlines()
    assert lines[0] == "DeviceArray DeviceArray (1000, 3000)"
    assert lines[1] == "True True True"
    np.random.seed(42)
    z = np.random.normal(loc=10, scale=2, size=300_000)
    assert lines[2] == f"ndarray {float(z[0])!r} {float(z.mean())!r}"
