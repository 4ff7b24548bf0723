"""Float unary ufuncs the elementwise kernel takes (sin, cos, tanh, floor, ceil, log1p, expm1 next
to sqrt/exp/log): deferred on the proxy, evaluated by one batched call, numpy's result; integer
inputs (numpy promotes them to float64) stay on the host path."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
import fake_device  # noqa: E402

from bee_code_interpreter_b200 import numpy_proxy as P  # noqa: E402
from bee_code_interpreter_b200.backend import LocalBackend  # noqa: E402


@pytest.fixture
def be(built):
    fake_device.reset()
    return LocalBackend(fake_device)


@pytest.mark.parametrize("fn", [np.sin, np.cos, np.tanh, np.floor, np.ceil, np.log1p, np.expm1, np.sqrt, np.exp, np.log])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_float_unary_ufuncs_go_through_the_device(be, fn, dt):
    h = (np.random.default_rng(5).random(5000) * 6 + 0.01).astype(dt)
    x = P.DeviceArray.from_host(be, h)
    fake_device.calls.clear()
    y = fn(x)
    assert isinstance(y, P.DeviceArray) and y.dtype == dt
    got = np.asarray(y)
    assert [c[0] for c in fake_device.calls] == ["ewise"]
    assert np.array_equal(got, fn(h))
    assert np.isclose(float(np.sum(fn(x))), float(np.sum(fn(h))), rtol=1e-5)      # materialised, then reduced


def test_integer_input_keeps_numpy_promotion(be):
    k = P.DeviceArray.from_host(be, np.arange(1, 50, dtype=np.int64))
    r = np.sin(k)
    assert isinstance(r, np.ndarray) and r.dtype == np.float64
    assert np.array_equal(r, np.sin(np.arange(1, 50, dtype=np.int64)))
    assert np.array_equal(np.floor(k), np.floor(np.arange(1, 50, dtype=np.int64)))
