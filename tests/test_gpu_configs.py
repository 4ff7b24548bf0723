"""BASELINE configs C2-C5 as request SCRIPTS through LocalGpuExecutor.execute(): source text in ->
zygote fork -> import hook -> broker -> batched CUDA launches -> stdout / stored files out, compared
with the oracle executor (the reference's CPU path, oracle/cpu_executor.py) running the SAME source.
Every test also asserts the broker's job counter moved: a script that silently ran on host numpy fails."""
import asyncio
import hashlib
import io
import tempfile

import numpy as np
import pytest

from bee_code_interpreter_b200 import workloads as W

pytestmark = pytest.mark.gpu

DEBUG_ENV = {"BCI_DEBUG_FALLBACK": "1"}      # a fallback to host numpy shows up on stderr (asserted empty)


@pytest.fixture(scope="module")
def executor(dev):
    from bee_code_interpreter_b200.executor import LocalGpuExecutor

    ex = LocalGpuExecutor(max_inflight=16, timeout=120, zygotes=2)
    yield ex
    ex.close()


@pytest.fixture(scope="module")
def oracle_executor():
    from oracle.cpu_executor import LocalCpuExecutor, RefStorage

    with tempfile.TemporaryDirectory() as td:
        yield LocalCpuExecutor(RefStorage(td), timeout=600)


def _jobs(executor) -> int:
    from bee_code_interpreter_b200.broker_client import RemoteBackend

    be = RemoteBackend(executor.broker.socket_path)
    try:
        return be.ping()["stats"]["jobs"]
    finally:
        be.close()


def _run_both(executor, oracle_executor, sources, files=None):
    async def go():
        ours = await asyncio.gather(*(executor.execute(s, files=files or {}, env=DEBUG_ENV) for s in sources))
        return ours
    before = _jobs(executor)
    ours = asyncio.run(go())
    jobs = _jobs(executor) - before
    ref = [asyncio.run(oracle_executor.execute(s)) for s in sources] if files is None else None
    for r in ours:
        assert r.exit_code == 0 and r.stderr == "", r.stderr
    return ours, ref, jobs


def _floats(stdout):
    return [float(t) for t in stdout.split()]


def test_c2_matmul_scripts(executor, oracle_executor):
    """C2: default_rng(1000+i).random((m,m), f32) x2 -> A @ B.  max|dC|/max|C| <= 1e-5 (north star bar)
    through printed max/min/mean, on 4 concurrent requests (m = 1024) and one at the config's 4096."""
    srcs = [W.c2_source(i, 1024) for i in range(4)] + [W.c2_source(7, 4096)]
    ours, ref, jobs = _run_both(executor, oracle_executor, srcs)
    assert jobs >= 5 * (2 + 1 + 3)            # per request: 2 generator draws, 1 matmul, 3 reductions
    for a, b in zip(ours, ref):
        got, want = _floats(a.stdout), _floats(b.stdout)
        scale = abs(want[0])
        assert all(abs(g - w) <= 1e-5 * scale for g, w in zip(got, want)), (got, want)


def test_c3_mixed_scripts(executor, oracle_executor):
    """C3: sum(square(x)), mean(x*y+1), max(abs(x-y)) on 2^20-element f32 arrays, 32 concurrent requests."""
    srcs = [W.c3_source(i) for i in range(32)]
    ours, ref, jobs = _run_both(executor, oracle_executor, srcs)
    assert jobs >= 32 * 5                      # 2 draws + 3 fused reductions each, nothing materialised in between
    for a, b in zip(ours, ref):
        got, want = _floats(a.stdout), _floats(b.stdout)
        assert all(abs(g - w) <= 1e-5 * abs(w) for g, w in zip(got, want)), (got, want)
        assert got[2] == want[2]               # max(abs(x-y)) is exact


def test_c3_inputs_from_the_files_map(executor):
    """The same program on .npy inputs restored from the object store: np.load hands the descriptor to the
    broker (pinned path), the arrays never pass through the script process."""
    n = 2**20
    x, y = W.c3_inputs(5, n)

    def npy(a):
        b = io.BytesIO()
        np.save(b, a)
        return b.getvalue()

    async def go():
        hx, hy = await executor.file_storage.write(npy(x)), await executor.file_storage.write(npy(y))
        assert hx == hashlib.sha256(npy(x)).hexdigest()
        return await executor.execute(W.C3_FILES_SOURCE, files={"/workspace/x.npy": hx, "/workspace/y.npy": hy}, env=DEBUG_ENV)

    before = _jobs(executor)
    r = asyncio.run(go())
    assert r.exit_code == 0 and r.stderr == "", r.stderr
    assert _jobs(executor) - before >= 3
    got = _floats(r.stdout)
    want = [float(np.sum(np.square(x))), float(np.mean(x * y + 1.0)), float(np.max(np.abs(x - y)))]
    assert all(abs(g - w) <= 1e-5 * abs(w) for g, w in zip(got, want)) and got[2] == want[2]


@pytest.mark.parametrize("nbytes", [8 * 2**20, 64 * 2**20 + 1])
def test_c4_output_files_are_named_by_their_sha256(executor, nbytes):
    """C4: each request writes one output file; its object id is the SHA-256 the K4 kernel computed
    (bit-exact vs hashlib), whether the bytes were born on the host or in HBM."""
    host = [W.c4_source(i, nbytes) for i in range(4)]
    device = [W.c4_source(i, nbytes - nbytes % 8, device=True) for i in range(4)]

    async def go():
        res = await asyncio.gather(*(executor.execute(s, env=DEBUG_ENV) for s in host + device))
        blobs = [await executor.file_storage.read(r.files["/workspace/out.bin"]) for r in res]
        return res, blobs

    before = _jobs(executor)
    res, blobs = asyncio.run(go())
    assert _jobs(executor) - before >= 8 + 4      # 8 file hashes + 4 device draws
    for i, (r, blob) in enumerate(zip(res, blobs)):
        assert r.exit_code == 0 and r.stderr == "", r.stderr
        assert set(r.files) == {"/workspace/out.bin"}
        assert hashlib.sha256(blob).hexdigest() == r.files["/workspace/out.bin"]
        if i < 4:
            assert blob == W.c4_payload(i, nbytes)
        else:
            n = (nbytes - nbytes % 8) // 8
            assert blob == np.random.default_rng(3000 + i - 4).random(n).tobytes()


def test_c5_analysis_scripts(executor, oracle_executor):
    """C5: sort / argmax / cumsum on 10^7-row arrays: stdout identical to the CPU executor's, byte for byte
    (sorted values, first-occurrence argmax, int64 prefix sums)."""
    srcs = [W.c5_source(i) for i in range(4)]
    ours, ref, jobs = _run_both(executor, oracle_executor, srcs)
    assert jobs >= 4 * 5                       # 2 draws, sort, argmax, cumsum
    for a, b in zip(ours, ref):
        assert a.stdout == b.stdout, (a.stdout, b.stdout)
