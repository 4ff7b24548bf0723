"""Wire protocol of the per-GPU broker (include/bci_broker.h), shared by the script-side
client (broker_client.py), the native broker inside libbci.so and its Python stand-in used
with the test double (broker.py)."""
from __future__ import annotations

import struct

ALLOC, FREE, RAND, EWISE, REDUCE, REDUCE_FUSED, ARGEXT, CUMSUM, SORT, ARGSORT, MATMUL, COPY, UPLOAD, DOWNLOAD, \
    SHA256_DEV, SHA256_FD, SHA256_BYTES, SHA256_UPDATE, SHA256_FINAL, PING, RAND_MT, RAND_NORMAL, RAND_PCG, \
    REDUCE_RAND, FILL, CAST, EXPORT = range(1, 28)
MT_STATE_WORDS = 625   # MT19937 key[624] + pos, little-endian uint32

F_IN1, F_ADDEND, F_INPLACE, F_FD, F_STATE = 1, 2, 4, 8, 16

# bci_wire_req: op, code, dtype, pre, aux, flags | p0, p1, n | i0, i1, i2 | s0, s1 | payload_len, reserved
REQ = struct.Struct("<6I3Q3q2QII")
REP = struct.Struct("<iIQ")
assert REQ.size == 96 and REP.size == 16

INLINE_MAX = 1 << 16   # array payloads up to 64 KiB travel inline; larger ones through a descriptor (SCM_RIGHTS)


def scalar_bits(value, is_float: bool) -> int:
    """bci_scalar bit pattern: double for float dtypes, int64 otherwise."""
    if value is None:
        return 0
    if is_float:
        return struct.unpack("<Q", struct.pack("<d", float(value)))[0]
    return int(value) & 0xFFFFFFFFFFFFFFFF


def scalar_from_bits(bits: int, is_float: bool):
    if is_float:
        return struct.unpack("<d", struct.pack("<Q", bits))[0]
    return struct.unpack("<q", struct.pack("<Q", bits))[0]


def pack_req(op, code=0, dtype=0, pre=0, aux=0, flags=0, p0=0, p1=0, n=0, i0=0, i1=0, i2=0, s0=0, s1=0, payload=b""):
    return REQ.pack(op, code, dtype, pre, aux, flags, p0 or 0, p1 or 0, n, i0, i1, i2, s0, s1, len(payload), 0) + payload


def recv_exact(sock, n: int) -> bytes:
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            raise EOFError("connection closed")
        buf += chunk
    return bytes(buf)


def send_with_fd(sock, msg: bytes, fd: int) -> None:
    """One request whose bulk bytes live behind `fd`: the descriptor rides on the first byte of the
    header as SCM_RIGHTS ancillary data (the broker never opens a path a client names)."""
    import socket as _socket

    sent = _socket.send_fds(sock, [msg], [fd])
    if sent < len(msg):
        sock.sendall(msg[sent:])


def recv_exact_fds(sock, n: int, fds: list) -> bytes:
    """recv_exact that also collects descriptors arriving as ancillary data (broker stand-in)."""
    import socket as _socket

    buf = bytearray()
    while len(buf) < n:
        chunk, got, _flags, _addr = _socket.recv_fds(sock, n - len(buf), 8)
        fds.extend(got)
        if not chunk:
            raise EOFError("connection closed")
        buf += chunk
    return bytes(buf)
