"""Protocol robustness of the broker (Python stand-in on the test double; the native broker
implements the same 96-byte request framing, include/bci_broker.h): garbage, truncated frames,
absurd sizes and foreign buffer handles must produce an error reply or a closed connection —
never a wedged broker or a leaked device buffer."""
import socket
import struct
import sys
import time
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
import fake_device  # noqa: E402

from bee_code_interpreter_b200 import _lib as L  # noqa: E402
from bee_code_interpreter_b200 import wire as W  # noqa: E402


@pytest.fixture
def broker(built):
    from bee_code_interpreter_b200.broker import GpuBroker

    fake_device.reset()
    b = GpuBroker(fake_device, coalesce_us=100)
    yield b
    b.close()


def _raw(path):
    s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    s.settimeout(5)
    s.connect(path)
    return s


def _healthy(broker):
    from bee_code_interpreter_b200.broker_client import RemoteBackend

    rb = RemoteBackend(broker.socket_path)
    p = rb.upload(np.arange(1000, dtype=np.float64))
    ok = rb.reduce(L.RED_SUM, np.float64, [p], [1000])[0] == 499500.0
    rb.free(p)
    rb.close()
    return ok


def test_garbage_and_truncated_frames(broker):
    rng = np.random.default_rng(0)
    for i in range(40):
        s = _raw(broker.socket_path)
        blob = rng.integers(0, 256, int(rng.integers(1, 400)), dtype=np.uint8).tobytes()
        try:
            s.sendall(blob)
            s.shutdown(socket.SHUT_WR)
            while s.recv(65536):
                pass
        except OSError:
            pass
        s.close()
    assert _healthy(broker)


def hostile_requests(socket_path):
    """The well-framed but hostile sequence; also run against the native broker on the GPU box
    (tests/test_gpu_executor.py)."""
    s = _raw(socket_path)

    def call(op, **kw):
        s.sendall(W.pack_req(op, **kw))
        status, plen, value = W.REP.unpack(W.recv_exact(s, W.REP.size))
        payload = W.recv_exact(s, plen) if plen else b""
        return status, value, payload

    assert call(9999)[0] != L.OK                                             # unknown op
    assert call(W.REDUCE, code=L.RED_SUM, dtype=L.F64, p0=0xdeadbeef, n=10)[0] != L.OK   # foreign handle
    assert call(W.EWISE, code=L.EW_ADD, dtype=77, p0=1, n=10)[0] != L.OK      # unknown dtype
    st, ptr, _ = call(W.ALLOC, n=800)
    assert st == L.OK and ptr
    assert call(W.REDUCE, code=L.RED_SUM, dtype=L.F64, p0=ptr, n=101)[0] != L.OK        # past the end of its buffer
    assert call(W.REDUCE, code=L.RED_SUM, dtype=L.F64, p0=ptr, n=2**61)[0] != L.OK      # size overflow
    assert call(W.MATMUL, p0=ptr, p1=ptr, i0=-5, i1=3, i2=3)[0] != L.OK                 # negative dimension
    assert call(W.RAND_MT, n=10, payload=b"\0" * 17)[0] != L.OK                         # short generator state
    assert call(W.UPLOAD, n=10**9, payload=b"abc")[0] != L.OK                           # declared size > payload
    assert call(W.SHA256_FINAL, payload=b"\1" * 5)[0] != L.OK                           # short hash state
    assert call(W.PING)[0] == L.OK                                                       # the connection still works
    s.close()


def test_hostile_but_well_framed_requests(broker):
    hostile_requests(broker.socket_path)
    for _ in range(100):
        if fake_device.live_buffers() == 0:
            break
        time.sleep(0.02)
    assert fake_device.live_buffers() == 0                                               # nothing leaked
    assert _healthy(broker)


def test_client_vanishing_mid_request(broker):
    for cut in (10, 50, 95):
        s = _raw(broker.socket_path)
        s.sendall(W.pack_req(W.UPLOAD, n=4096, payload=b"x" * 4096)[:cut])
        s.close()
    s = _raw(broker.socket_path)
    s.sendall(W.pack_req(W.UPLOAD, n=4096, payload=b"x" * 4096)[:96 + 100])   # header + part of the payload
    s.close()
    assert _healthy(broker)
    for _ in range(100):
        if fake_device.live_buffers() == 0:
            break
        time.sleep(0.02)
    assert fake_device.live_buffers() == 0
