"""numpy.random.normal / randn / standard_normal on the Philox stream (K8c): the oracle's libm
restatement has the right distribution, and the hook routes large unseeded draws to the device
and everything else (seeded, small, array-valued loc/scale) to numpy."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
import fake_device  # noqa: E402

from bee_code_interpreter_b200 import numpy_proxy as P  # noqa: E402
from bee_code_interpreter_b200.backend import LocalBackend  # noqa: E402
from oracle import numpy_oracle as O  # noqa: E402


def test_oracle_normals_are_standard_normal(built):
    from scipy import stats

    z = O.rand_normal_f64(1_000_001, 123)
    assert abs(z.mean()) < 5 / np.sqrt(z.size) and abs(z.var() - 1) < 5 * np.sqrt(2 / z.size)
    assert abs(stats.skew(z)) < 0.02 and abs(stats.kurtosis(z)) < 0.03
    assert stats.kstest(z[:200_000], "norm").pvalue > 1e-4
    m = z.size // 2
    assert abs(np.corrcoef(z[0:2 * m:2], z[1:2 * m:2])[0, 1]) < 0.01                            # the pair from one counter
    y = O.rand_normal_f64(1001, 123, loc=10.0, scale=2.0)
    assert np.allclose(y, 10.0 + 2.0 * z[:1001], rtol=0, atol=1e-12)
    assert not np.array_equal(O.rand_normal_f64(100, 124), z[:100])                           # keyed by the seed
    u = O.rand_f64(1000, 123)
    assert abs(np.corrcoef(u, z[:1000])[0, 1]) < 0.15                                           # not the uniform stream reshaped


def test_hook_routes_normals(built):
    from bee_code_interpreter_b200 import numpy_hook as H

    fake_device.reset()
    be = LocalBackend(fake_device)
    H._state.update(backend=None, backend_factory=lambda: be, view=None, seeded=False, min_rand=100)
    view = H.numpy_view()
    a = view.random.randn(500)
    b = view.random.normal(3.0, 0.5, size=(20, 30))
    c = view.random.standard_normal(700)
    assert all(isinstance(v, P.DeviceArray) for v in (a, b, c)) and b.shape == (20, 30)
    assert abs(float(np.mean(b)) - 3.0) < 0.1 and abs(float(np.std(b)) - 0.5) < 0.1
    assert [x[0] for x in fake_device.calls].count("rand_normal") == 3
    assert isinstance(view.random.randn(5), np.ndarray) and isinstance(view.random.normal(), float)      # small / scalar
    assert isinstance(view.random.normal(np.zeros(500), 1.0, 500), np.ndarray)                             # array-valued loc
    view.random.seed(7)                                                                                     # seeded: numpy's stream
    s = view.random.normal(0, 1, 500)
    np.random.seed(7)
    assert isinstance(s, np.ndarray) and np.array_equal(s, np.random.normal(0, 1, 500))
    H._state.update(seeded=False)
