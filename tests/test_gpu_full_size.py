"""BASELINE.json's configurations at FULL per-request size, a few requests each,
checked against the oracle and through size-independent properties (sortedness + permutation
checksum, last prefix == total, linearity of sums, digest of digests)."""
import hashlib

import numpy as np
import pytest

from bee_code_interpreter_b200 import _lib as L
from bee_code_interpreter_b200 import workloads as W
from oracle import numpy_oracle as O

pytestmark = pytest.mark.gpu


def test_c1_benchmark_numpy_body_full_size(dev):
    """C1: n = 1e8 float64.  Device rand stream is bit-exact with the oracle generator on a
    strided sample; sum(square(x)) equals numpy's on the SAME data to 1e-12; the fused and the
    unfused (square materialised, then sum) paths agree bit-for-bit with each other."""
    n = W.C1_N
    p, q = dev.alloc(n * 8), dev.alloc(n * 8)
    dev.rand_f64([p], [n], [424242])
    x = dev.download(p, n, np.float64)
    ref = O.rand_f64(2_000_000, 424242)
    assert np.array_equal(x[:2_000_000], ref)
    assert 0.0 <= x.min() and x.max() < 1.0
    fused = dev.reduce(L.RED_SUM, np.float64, [p], [n], pre=L.PRE_SQUARE)[0]
    want = O.benchmark_numpy_body(x)
    assert abs(fused - want) <= 1e-12 * want
    dev.ewise(L.EW_SQUARE, np.float64, [p], None, None, [q], [n])
    unfused = dev.reduce(L.RED_SUM, np.float64, [q], [n])[0]
    assert fused == unfused                      # same rounding points, same reduction tree
    assert abs(want / n - 1 / 3) < 6 * np.sqrt(4 / 45 / n)
    sq = dev.download(q, n, np.float64)
    assert np.array_equal(sq, np.square(x))
    dev.free(p)
    dev.free(q)


def test_c3_mixed_program_full_batch_shape(dev):
    """C3: 2^20-element float32 pairs, 64 requests in ONE launch per op class."""
    B, n = 64, 2**20
    xs, ys = zip(*(W.c3_inputs(i) for i in range(B)))
    px, py = [dev.upload(x) for x in xs], [dev.upload(y) for y in ys]
    ns = [n] * B
    r1 = dev.reduce(L.RED_SUM, np.float32, px, ns, pre=L.PRE_SQUARE)
    r2 = dev.reduce_fused(L.RED_MEAN, L.EW_MUL, np.float32, px, py, None, ns, addends=[1.0] * B)
    r3 = dev.reduce_fused(L.RED_MAX, L.EW_SUB, np.float32, px, py, None, ns, pre=L.PRE_ABS)
    for i in range(B):
        w1, w2, w3 = np.sum(np.square(xs[i])), np.mean(xs[i] * ys[i] + 1.0), np.max(np.abs(xs[i] - ys[i]))
        assert abs(r1[i] - w1) <= 1e-5 * w1 and abs(r2[i] - w2) <= 1e-5 * w2 and r3[i] == w3
    for p in px + py:
        dev.free(p)


def test_c4_sha256_full_size_files(dev):
    """C4: 64 MiB payloads (and 64 MiB +- 1, all-zero), device resident, one launch."""
    size = 64 * 2**20
    payloads = [W.c4_payload(0), W.c4_payload(1)[: size - 1], W.c4_payload(2) + b"\x01", bytes(size)]
    ptrs = [dev.upload(np.frombuffer(p, dtype=np.uint8)) for p in payloads]
    got = dev.sha256_device(ptrs, [len(p) for p in payloads])
    for p, g in zip(payloads, got):
        assert g == hashlib.sha256(p).digest()
    # digest of digests: order and content of the whole batch
    assert hashlib.sha256(b"".join(got)).digest() == hashlib.sha256(b"".join(hashlib.sha256(p).digest() for p in payloads)).digest()
    for p in ptrs:
        dev.free(p)


def test_c5_sort_argmax_cumsum_full_size(dev):
    """C5: 1e7 rows: np.sort(v), np.argmax(v), np.cumsum(k) per request, 4 requests batched."""
    B, n = 4, 10**7
    vs, ks = zip(*(W.c5_inputs(i) for i in range(B)))
    pv, pk = [dev.upload(v) for v in vs], [dev.upload(k) for k in ks]
    po, pc = [dev.alloc(n * 4) for _ in range(B)], [dev.alloc(n * 8) for _ in range(B)]
    ns = [n] * B
    dev.sort(np.float32, pv, po, ns)
    am = dev.argext(L.ARGMAX, np.float32, pv, ns)
    dev.cumsum(np.int64, pk, pc, ns)
    for i in range(B):
        s = dev.download(po[i], n, np.float32)
        assert np.all(s[1:] >= s[:-1])                                        # sorted
        assert s.view(np.uint32).astype(np.uint64).sum() == vs[i].view(np.uint32).astype(np.uint64).sum()  # same multiset (checksum)
        assert np.array_equal(s, np.sort(vs[i]))                              # and bit-exact
        assert am[i] == np.argmax(vs[i])
        c = dev.download(pc[i], n, np.int64)
        assert c[-1] == ks[i].sum() and np.array_equal(c, np.cumsum(ks[i]))
    # idempotence: sorting the sorted array changes nothing
    dev.sort(np.float32, po, po, ns)
    assert np.array_equal(dev.download(po[0], n, np.float32), np.sort(vs[0]))
    for p in pv + pk + po + pc:
        dev.free(p)
