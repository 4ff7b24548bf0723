"""Round-2 kernels through the C ABI against numpy (the reference's arithmetic): PCG64 Generator
streams, reductions over a generated operand, creation / cast kernels, the pinned descriptor
transfer path and device-side hashing of exported arrays."""
import hashlib
import os
import tempfile

import numpy as np
import pytest

from bee_code_interpreter_b200 import _lib as L
from oracle import numpy_oracle as O

pytestmark = pytest.mark.gpu

SIZES = [0, 1, 7, 16, 16383, 16384, 16385, 32769, 1_000_003, (1 << 22) + 5]


def _gen(seed):
    st = np.random.PCG64(seed).state["state"]
    return st["state"], st["inc"]


def _ref(seed):
    return np.random.default_rng(seed)


@pytest.mark.parametrize("kind,dtype,draw", [
    (L.PCG_F64, np.float64, lambda g, n: g.random(n)),
    (L.PCG_F32, np.float32, lambda g, n: g.random(n, dtype=np.float32)),
    (L.PCG_U32, np.uint32, lambda g, n: g.integers(0, 1 << 32, n, dtype=np.uint32)),
])
def test_pcg64_streams_are_numpys(dev, kind, dtype, draw):
    """One ragged batch per kind: every request's output is bit for bit what numpy's Generator returns."""
    seeds = [1000 + i for i in range(len(SIZES))]
    ptrs = [dev.alloc(max(n, 1) * np.dtype(dtype).itemsize) for n in SIZES]
    rej = dev.rand_pcg64(kind, ptrs, SIZES, [_gen(s) for s in seeds])
    assert not rej.any()
    for p, n, s in zip(ptrs, SIZES, seeds):
        got = dev.download(p, n, dtype)
        assert np.array_equal(got, draw(_ref(s), n)), (kind, n)
        dev.free(p)


@pytest.mark.parametrize("out_kind,dt", [(L.PCG_BOUNDED32_I64, np.int64), (L.PCG_BOUNDED32_I32, np.int32)])
def test_pcg64_bounded_integers(dev, out_kind, dt):
    """Lemire 32-bit path: exact for every range when no draw is rejected; rejections are COUNTED (the
    caller then falls back to numpy) and the count matches how far numpy's generator really went."""
    cases = [(0, 2**31, 10**6 + 1), (5, 2**20, 70001), (-7, 3, 16385), (0, 100, 2 * 10**6), (0, 2**31 + 12345, 5 * 10**5),
             (10, 2**32 - 1, 33333)]
    if dt == np.int32:
        cases = [(lo, min(span, 2**31 - 1 - lo), n) for lo, span, n in cases]
    for i, (low, span, n) in enumerate(cases):
        seed = 5000 + i
        p = dev.alloc(n * np.dtype(dt).itemsize)
        state, inc = _gen(seed)
        rej = int(dev.rand_pcg64(out_kind, [p], [n], [(state, inc, low, span)])[0])
        got = dev.download(p, n, dt)
        dev.free(p)
        want = _ref(seed).integers(low, low + span, n, dtype=dt)
        # numpy's stream equals the unrejected one up to its first rejection, which lies among the first n
        # words: the kernel's count is zero exactly when the two outputs agree everywhere
        assert (rej == 0) == np.array_equal(got, want), (low, span, rej)
        assert got.min() >= low and got.max() < low + span
        assert np.array_equal(got, O.pcg64_random(state, inc, n, "bounded", low, span).astype(dt))   # the unrejected stream itself


def test_reduction_over_generated_operand_is_bit_identical(dev):
    """bci_reduce_rand_f64 == bci_rand_f64 + bci_reduce for every op / pre-transform and ragged sizes."""
    ns = [1, 5, 16383, 16384, 16385, 100_003, 3_000_017]
    seeds = [77 + i for i in range(len(ns))]
    ptrs = [dev.alloc(n * 8) for n in ns]
    dev.rand_f64(ptrs, ns, seeds)
    for op in (L.RED_SUM, L.RED_MEAN, L.RED_MAX, L.RED_MIN):
        for pre in (L.PRE_NONE, L.PRE_SQUARE, L.PRE_ABS):
            two = dev.reduce(op, np.float64, ptrs, ns, pre=pre)
            one = dev.reduce_rand_f64(op, ns, seeds, pre=pre)
            assert np.array_equal(two, one), (op, pre)
    x = dev.download(ptrs[-1], ns[-1], np.float64)
    assert abs(dev.reduce_rand_f64(L.RED_SUM, [ns[-1]], [seeds[-1]], pre=L.PRE_SQUARE)[0] - np.sum(np.square(x))) <= 1e-12 * np.sum(np.square(x))
    for p in ptrs:
        dev.free(p)


def test_creation_kernels_match_numpy(dev):
    for dt in (np.float32, np.float64, np.int32, np.int64):
        for n in (0, 1, 4095, 4096, 4097, 100_003):
            p = dev.alloc(max(n, 1) * np.dtype(dt).itemsize)
            dev.fill(L.FILL_CONST, dt, [p], [n], [3])
            assert np.array_equal(dev.download(p, n, dt), np.full(n, 3, dtype=dt))
            dev.free(p)
    for args, dt in [((0, 300_001, 1), np.int64), ((5, 900_000, 3), np.int32), ((0.5, 30000.25, 0.37), np.float64),
                     ((0, 200_001, 1), np.float32), ((1.5, 70000.0, 0.1), np.float32), ((10.0, -50000.0, -0.7), np.float64),
                     ((0, 2**25 + 7, 1), np.float32)]:
        want = np.arange(*args, dtype=dt)
        n = len(want)
        first, second = np.dtype(dt).type(args[0]), np.dtype(dt).type(args[0] + args[2])
        delta = np.dtype(dt).type(second - first)
        p = dev.alloc(n * np.dtype(dt).itemsize)
        dev.fill(L.FILL_ARANGE, dt, [p], [n], [first.item()], [delta.item()], [second.item()])
        assert np.array_equal(dev.download(p, n, dt), want), (args, dt)
        dev.free(p)
    for start, stop, num, endpoint in [(0.0, 1.0, 100_001, True), (-3.3, 17.1, 65537, False), (1e-3, 1e9, 4099, True)]:
        want = np.linspace(start, stop, num, endpoint=endpoint)
        step = (stop - start) / ((num - 1) if endpoint else num)
        p = dev.alloc(num * 8)
        dev.fill(L.FILL_LINSPACE, np.float64, [p], [num], [start], [step], [stop] if endpoint else None)
        assert np.array_equal(dev.download(p, num, np.float64), want), (start, stop, num)
        dev.free(p)


def test_casts_match_numpy_astype(dev):
    rng = np.random.default_rng(3)
    src = {np.float32: rng.standard_normal(100_003).astype(np.float32) * 1e3,
           np.float64: rng.standard_normal(100_003) * 1e10,
           np.int32: rng.integers(-2**31, 2**31, 100_003, dtype=np.int32),
           np.int64: rng.integers(-2**62, 2**62, 100_003, dtype=np.int64)}
    pairs = [(np.float32, np.float64), (np.float64, np.float32), (np.int32, np.int64), (np.int64, np.int32),
             (np.int32, np.float32), (np.int32, np.float64), (np.int64, np.float32), (np.int64, np.float64)]
    for s, d in pairs:
        for n in (0, 5, 4096, 100_003):
            x = src[s][:n]
            p = dev.upload(x) if n else dev.alloc(1)
            q = dev.alloc(max(n, 1) * np.dtype(d).itemsize)
            dev.cast(s, d, [p], [q], [n])
            with np.errstate(all="ignore"):
                assert np.array_equal(dev.download(q, n, d), x.astype(d)), (s, d, n)
            dev.free(p)
            dev.free(q)
    with pytest.raises(dev.Unsupported):
        dev.cast(np.float64, np.int64, [dev.alloc(8)], [dev.alloc(8)], [1])


def test_descriptor_transfers_round_trip(dev, tmp_path):
    """bci_h2d_fd / bci_d2h_fd: sizes around the 8 MiB bounce chunk, non-zero file offsets, short files."""
    rng = np.random.default_rng(11)
    for nbytes, off in [(1, 0), (4097, 3), ((8 << 20) - 1, 0), (8 << 20, 128), ((8 << 20) + 1, 0), ((40 << 20) + 12345, 77)]:
        data = rng.integers(0, 256, nbytes, dtype=np.uint8)
        src = tmp_path / "in.bin"
        with open(src, "wb") as f:
            f.write(b"\xee" * off)
            f.write(data.tobytes())
        p = dev.alloc(nbytes)
        fd = os.open(src, os.O_RDONLY)
        dev.h2d_fd(p, fd, off, nbytes)
        os.close(fd)
        assert np.array_equal(dev.download(p, nbytes, np.uint8), data)
        assert dev.sha256_device([p], [nbytes])[0] == hashlib.sha256(data.tobytes()).digest()
        dst = tmp_path / "out.bin"
        fd = os.open(dst, os.O_RDWR | os.O_CREAT | os.O_TRUNC)
        dev.d2h_fd(fd, off, p, nbytes)
        os.close(fd)
        assert dst.read_bytes()[off:] == data.tobytes()
        dev.free(p)
    p = dev.alloc(1 << 20)
    fd = os.open(tmp_path / "in.bin", os.O_RDONLY)
    with pytest.raises(dev.BciError):
        dev.h2d_fd(p, fd, 0, 1 << 30)      # the file ends first: an error, not a fault
    os.close(fd)
    dev.free(p)


def test_c2_matmul_on_the_configs_own_inputs(dev):
    """BASELINE C2 names its inputs: A, B = default_rng(1000 + i).random((4096, 4096), float32).  Generated ON the
    device (bit-exact with numpy, asserted), multiplied by the tcgen05 kernel, compared with numpy's own sgemm on
    the whole matrix (bar 1e-5 of max|C|) and with a float64 product on sampled rows (bar 2e-6)."""
    m = 4096
    n = m * m
    for i in (0, 3):
        rng = np.random.default_rng(1000 + i)
        a_h, b_h = rng.random((m, m), dtype=np.float32), rng.random((m, m), dtype=np.float32)
        st = np.random.PCG64(1000 + i)
        s0 = st.state["state"]
        st.advance(n // 2)
        s1 = st.state["state"]
        A, B, C = dev.alloc(n * 4), dev.alloc(n * 4), dev.alloc(n * 4)
        dev.rand_pcg64(L.PCG_F32, [A, B], [n, n], [(s0["state"], s0["inc"]), (s1["state"], s1["inc"])])
        assert np.array_equal(dev.download(A, n, np.float32), a_h.ravel()) and np.array_equal(dev.download(B, n, np.float32), b_h.ravel())
        dev.matmul_f32([A], [B], [C], m, m, m)
        c = dev.download(C, n, np.float32).reshape(m, m)
        ref32 = a_h @ b_h
        assert np.abs(c - ref32).max() <= 1e-5 * np.abs(ref32).max()
        rows = [0, 17, m // 2, m - 1]
        ref64 = a_h[rows].astype(np.float64) @ b_h.astype(np.float64)
        assert np.abs(c[rows] - ref64).max() <= 2e-6 * np.abs(ref64).max()
        for p in (A, B, C):
            dev.free(p)


def test_exported_array_is_hashed_on_the_device_and_not_read_back(dev, tmp_path):
    """BRK_EXPORT: a device array goes to a file behind an .npy header through the pinned path while the device
    hashes header||array; a later BRK_SHA256_FD of the untouched file is answered from the export table, and
    any write to the file invalidates the entry."""
    import io

    from bee_code_interpreter_b200.broker import BrokerProcess
    from bee_code_interpreter_b200.broker_client import RemoteBackend

    broker = BrokerProcess(device=0)     # its own CUDA context beside this process's
    try:
        be = RemoteBackend(broker.socket_path)
        st = np.random.PCG64(77).state["state"]
        n = (8 << 20) + 3
        ptr, rej = be.rand_pcg(L.PCG_F64, n, st["state"], st["inc"])
        want = np.random.default_rng(77).random(n)
        head = io.BytesIO()
        np.lib.format.write_array_header_1_0(head, {"descr": "<f8", "fortran_order": False, "shape": (n,)})
        path = tmp_path / "out.npy"
        fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC)
        digest = be.export(ptr, n * 8, fd, head.getvalue())
        os.close(fd)
        blob = path.read_bytes()
        assert digest == hashlib.sha256(blob).digest()
        assert np.array_equal(np.load(path), want)
        fd = os.open(path, os.O_RDONLY)
        d2, cached = be.sha256_fd_ex(fd)
        assert d2 == digest and cached
        d3, cached = be.sha256_fd_ex(fd)            # one consumer: the entry is gone, the file is read and hashed
        assert d3 == digest and not cached
        os.close(fd)
        fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC)
        be.export(ptr, n * 8, fd, b"")
        os.pwrite(fd, b"\x01", 5)                    # the script touches the file afterwards
        os.close(fd)
        fd = os.open(path, os.O_RDONLY)
        d4, cached = be.sha256_fd_ex(fd)
        os.close(fd)
        assert not cached and d4 == hashlib.sha256(path.read_bytes()).digest()
        be.close()
    finally:
        broker.close()


def test_sort_passes_whose_digit_is_constant_are_copied_not_ranked(dev):
    """One batch mixing requests whose passes are trivial (all keys share the digit: small non-negative int64, a
    constant array, floats of one binade) with requests whose passes are not: sort and stable argsort stay bit-exact."""
    rng = np.random.default_rng(12)
    n = 300_007
    cases = {
        np.int64: [rng.integers(0, 2**31, n, dtype=np.int64), np.full(n, 7, dtype=np.int64), rng.integers(-2**62, 2**62, n, dtype=np.int64),
                   rng.integers(0, 200, 5000, dtype=np.int64), rng.integers(-3, 3, n, dtype=np.int64)],
        np.float32: [(1.0 + rng.random(n, dtype=np.float32)), np.zeros(n, dtype=np.float32), rng.standard_normal(n).astype(np.float32),
                     np.full(777, np.nan, dtype=np.float32)],
        np.float64: [rng.random(n) + 4.0, rng.standard_normal(n), np.full(n, -1.5)],
        np.int32: [rng.integers(0, 255, n, dtype=np.int32), rng.integers(-2**31, 2**31 - 1, n, dtype=np.int32)],
    }
    for dt, arrs in cases.items():
        ins = [dev.upload(a) for a in arrs]
        outs = [dev.alloc(a.nbytes) for a in arrs]
        idx = [dev.alloc(a.size * 8) for a in arrs]
        ns = [a.size for a in arrs]
        dev.sort(dt, ins, outs, ns)
        dev.argsort(dt, ins, idx, ns)
        for a, o, ix in zip(arrs, outs, idx):
            got = dev.download(o, a.size, dt)
            assert np.array_equal(got, np.sort(a), equal_nan=True), dt
            perm = dev.download(ix, a.size, np.int64)
            want = np.argsort(a, kind="stable")
            assert np.array_equal(a[perm], a[want], equal_nan=True) and (np.array_equal(perm, want) or np.isnan(a).any() or (a == 0).any()), dt
        for p in ins + outs + idx:
            dev.free(p)


@pytest.mark.gpu
def test_sort_leaves_out_passes_that_are_constant_for_the_whole_batch(dev):
    """Batches above 2^20 keys in which every request shares the constant digits run fewer sweeps (odd and even counts,
    so both ping-pong parities), also when sorting in place: sort and stable argsort stay bit-exact."""
    rng = np.random.default_rng(13)
    n = 700_001
    batches = [
        [rng.integers(0, 2**31, n, dtype=np.int64), rng.integers(0, 2**24, n, dtype=np.int64)],        # 8 -> 4 sweeps
        [rng.integers(0, 2**16, n, dtype=np.int64), rng.integers(0, 2**9, n, dtype=np.int64)],         # 8 -> 2
        [rng.integers(0, 2**24, n, dtype=np.int64)] * 2,                                                # 8 -> 3
        [rng.integers(0, 256, n, dtype=np.int32), rng.integers(0, 3, n, dtype=np.int32)],               # 4 -> 1
        [rng.integers(0, 2**16, n, dtype=np.int32), rng.integers(0, 2**16, n, dtype=np.int32)],         # 4 -> 2
        [rng.integers(0, 2**8, n, dtype=np.int64) << 40, (rng.integers(0, 2**8, n, dtype=np.int64) << 40) + 5],  # middle digit only
        [1.0 + rng.random(n), 1.5 + rng.random(n) / 4],                                                 # f64 in one binade
    ]
    for arrs in batches:
        dt = arrs[0].dtype.type
        ns = [a.size for a in arrs]
        ins = [dev.upload(a) for a in arrs]
        outs = [dev.alloc(a.nbytes) for a in arrs]
        idx = [dev.alloc(a.size * 8) for a in arrs]
        dev.sort(dt, ins, outs, ns)
        dev.argsort(dt, ins, idx, ns)
        for a, o, ix in zip(arrs, outs, idx):
            assert np.array_equal(dev.download(o, a.size, dt), np.sort(a))
            assert np.array_equal(dev.download(ix, a.size, np.int64), np.argsort(a, kind="stable"))
        dev.sort(dt, ins, ins, ns)  # in place
        for a, o in zip(arrs, ins):
            assert np.array_equal(dev.download(o, a.size, dt), np.sort(a))
        for p in ins + outs + idx:
            dev.free(p)


@pytest.mark.gpu
def test_sort_sizes_around_the_tile_and_ranking_unit_boundaries(dev):
    """Every size class of the sweep: below one 16-key row, one ranking unit (16 lanes x 16 or 20 rows) +-1, one tile +-1
    (4096 / 5120 keys; 10240 for 32-bit keys alone), several tiles with a ragged tail; sorted, reversed and two-valued
    inputs among them."""
    rng = np.random.default_rng(14)
    sizes = [1, 2, 15, 16, 17, 255, 256, 257, 319, 320, 321, 4095, 4096, 4097, 5119, 5120, 5121, 8192, 10239, 10240, 10241,
             12289, 3 * 10240 - 1, 5 * 10240 + 300]
    for dt in (np.float32, np.int32, np.int64, np.float64):
        arrs = []
        for j, n in enumerate(sizes):
            if np.issubdtype(dt, np.floating):
                a = rng.standard_normal(n).astype(dt)
            else:
                a = rng.integers(np.iinfo(dt).min, np.iinfo(dt).max, n, dtype=dt)
            if j % 4 == 1:
                a = np.sort(a)
            elif j % 4 == 2:
                a = np.sort(a)[::-1].copy()
            elif j % 4 == 3:
                a = np.where(rng.random(n) < 0.5, a[0], a[-1]).astype(dt)
            arrs.append(a)
        ns = [a.size for a in arrs]
        ins = [dev.upload(a) for a in arrs]
        outs = [dev.alloc(a.nbytes) for a in arrs]
        idx = [dev.alloc(a.size * 8) for a in arrs]
        dev.sort(dt, ins, outs, ns)
        dev.argsort(dt, ins, idx, ns)
        for a, o, ix in zip(arrs, outs, idx):
            assert np.array_equal(dev.download(o, a.size, dt), np.sort(a)), (dt, a.size)
            assert np.array_equal(dev.download(ix, a.size, np.int64), np.argsort(a, kind="stable")), (dt, a.size)
        for p_ in ins + outs + idx:
            dev.free(p_)


@pytest.mark.gpu
def test_prepared_reductions_return_what_the_plain_calls_return(dev):
    """device.prepared_reduce / prepared_reduce_fused marshal the tables once; repeated calls give the plain calls' bits."""
    from bee_code_interpreter_b200 import _lib as L
    rng = np.random.default_rng(15)
    xs = [rng.random(n, dtype=np.float32) for n in (1, 1000, 2**20, 2**20 + 3)]
    ys = [rng.random(a.size, dtype=np.float32) for a in xs]
    X, Y, ns = [dev.upload(a) for a in xs], [dev.upload(a) for a in ys], [a.size for a in xs]
    k1 = dev.prepared_reduce(L.RED_SUM, np.float32, X, ns, pre=L.PRE_SQUARE)
    k2 = dev.prepared_reduce_fused(L.RED_MEAN, L.EW_MUL, np.float32, X, Y, None, ns, addends=[1.0] * len(ns))
    k3 = dev.prepared_reduce_fused(L.RED_MAX, L.EW_SUB, np.float32, X, Y, None, ns, pre=L.PRE_ABS)
    for _ in range(2):
        assert np.array_equal(k1(), dev.reduce(L.RED_SUM, np.float32, X, ns, pre=L.PRE_SQUARE))
        assert np.array_equal(k2(), dev.reduce_fused(L.RED_MEAN, L.EW_MUL, np.float32, X, Y, None, ns, addends=[1.0] * len(ns)))
        assert np.array_equal(k3(), dev.reduce_fused(L.RED_MAX, L.EW_SUB, np.float32, X, Y, None, ns, pre=L.PRE_ABS))
    want = [np.max(np.abs(a - b)) for a, b in zip(xs, ys)]
    assert np.array_equal(k3(), np.array(want, dtype=np.float32))
    for p_ in X + Y:
        dev.free(p_)
