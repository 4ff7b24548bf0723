"""Requests past 2^31 elements: tile counts, element offsets and result indices must all be
64-bit.  One int32 array of 2^31 + 12345 elements (8.6 GB) built on the device; every check is a
closed form (numpy's answer for an array of ones with a few planted values)."""
import numpy as np
import pytest

from bee_code_interpreter_b200 import _lib as L

pytestmark = pytest.mark.gpu

N = 2**31 + 12345


def _fill_ones(dev, ptr, n):
    chunk = 1 << 24
    dev.h2d(ptr, np.ones(chunk, dtype=np.int32))
    done = chunk
    while done < n:
        step = min(done, n - done)
        dev.d2d(ptr + 4 * done, ptr, 4 * step)
        done += step
    dev.sync()


def test_int32_ops_past_two_to_the_31_elements(dev):
    free_b, _ = dev.mem_info()
    if free_b < 60 * 2**30:
        pytest.skip("needs ~45 GB of free device memory")
    x = dev.alloc(4 * N)
    out = dev.alloc(8 * N)
    try:
        _fill_ones(dev, x, N)
        assert int(dev.reduce(L.RED_SUM, np.int32, [x], [N])[0]) == N                     # int64 accumulator
        big, small = 2**31 + 7, 2**31 + 9000
        dev.h2d(x + 4 * big, np.array([5], dtype=np.int32))
        dev.h2d(x + 4 * small, np.array([-3], dtype=np.int32))
        dev.h2d(x + 4 * 17, np.array([5], dtype=np.int32))                                  # earlier tie: first occurrence wins
        dev.sync()
        assert int(dev.argext(L.ARGMAX, np.int32, [x], [N])[0]) == 17
        assert int(dev.argext(L.ARGMIN, np.int32, [x], [N])[0]) == small
        dev.h2d(x + 4 * 17, np.array([1], dtype=np.int32))
        dev.sync()
        assert int(dev.argext(L.ARGMAX, np.int32, [x], [N])[0]) == big                     # index needs 32+ bits
        assert int(dev.reduce(L.RED_MAX, np.int32, [x], [N])[0]) == 5
        # cumsum: int32 -> int64, single-pass kernel, ~524 k tiles in one request
        dev.cumsum(np.int32, [x], [out], [N])
        total = N + 4 - 4                                                                   # +4 (the 5) and -4 (the -3)
        tail = dev.download(out + 8 * (N - 4), 4, np.int64)
        assert tail.tolist() == [total - 3, total - 2, total - 1, total]
        assert dev.download(out + 8 * (big - 1), 3, np.int64).tolist() == [big, big + 5, big + 6]
        assert dev.download(out, 3, np.int64).tolist() == [1, 2, 3]
        # elementwise: x * 3 - 1 in place over the whole array, sampled at both ends and the plants
        dev.ewise(L.EW_MUL, np.int32, [x], None, [3], [x], [N])
        dev.ewise(L.EW_SUB, np.int32, [x], None, [1], [x], [N])
        dev.sync()
        assert dev.download(x + 4 * big, 2, np.int32).tolist() == [14, 2]
        assert dev.download(x + 4 * small, 1, np.int32).tolist() == [-10]
        assert dev.download(x + 4 * (N - 2), 2, np.int32).tolist() == [2, 2]
        assert int(dev.reduce(L.RED_SUM, np.int32, [x], [N])[0]) == 2 * (N - 2) + 14 - 10
        # sort: a request of >= 2^31 keys is refused loudly (32-bit positions inside a request; the
        # proxy then falls back to numpy), 2^30 <= n < 2^31 takes the three-kernel form (the one-sweep
        # status words hold 30-bit counts)
        with pytest.raises(L.Unsupported):
            dev.sort(np.int32, [x], [out], [N])
        M = 2**30 + 5
        dev.h2d(x + 4 * 100, np.array([-7], dtype=np.int32))
        dev.h2d(x + 4 * (M - 1), np.array([9], dtype=np.int32))
        dev.sync()
        dev.sort(np.int32, [x], [out], [M])
        assert dev.download(out, 2, np.int32).tolist() == [-7, 2]
        assert dev.download(out + 4 * (M - 2), 2, np.int32).tolist() == [2, 9]
        assert int(dev.reduce(L.RED_SUM, np.int32, [out], [M])[0]) == 2 * (M - 2) - 7 + 9
    finally:
        dev.free(x)
        dev.free(out)
